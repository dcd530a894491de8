"""Micro-benchmarks of the sm_100a kernels at the SD-v1.4 shapes (CUDA events, L2 flushed between launches).
    python tools/bench_kernels.py [attn] [gemm] [conv] [norm]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch
from e4t_b200 import ops

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
B = int(os.environ.get("B", "16"))


def timeit(fn, iters=6):
    fn(); fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def bf(*s, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).to(torch.bfloat16)


what = sys.argv[1:] or ["attn", "gemm", "conv", "norm"]
if "attn" in what:
    print(f"== attention (B={B}, 8 heads)  E4T_ATTN_CG={os.environ.get('E4T_ATTN_CG')}")
    for (N, M, C) in [(4096, 4096, 320), (1024, 1024, 640), (256, 256, 1280), (64, 64, 1280), (4096, 77, 320), (1024, 77, 640), (256, 77, 1280)]:
        q = bf(B, N, C, scale=0.5); k = bf(B, M, C, scale=0.5); v = bf(B, M, C, scale=0.5)
        o, lse = ops.attn_fwd(q, k, v, 8)
        do = bf(B, N, C)
        tf = timeit(lambda: ops.attn_fwd(q, k, v, 8))
        tb = timeit(lambda: ops.attn_bwd(q, k, v, o, do, lse, 8))
        tb2 = timeit(lambda: ops.attn_bwd(q, k, v, o, do, lse, 8, fused=False))
        fl = 4.0 * N * M * C * B
        print(f"N={N:5d} M={M:5d} dh={C//8:3d}: fwd {tf:7.3f} ms {fl/tf/1e9:7.1f} TF/s | bwd fused {tb:7.3f} ms {2.5*fl/tb/1e9:7.1f} TF/s | bwd 2-kernel {tb2:7.3f} ms")
if "gemm" in what:
    print("== gemm  (M,N,K) [a_mn,b_mn]")
    shapes = [(B * 4096, 960, 320, 0, 0), (B * 4096, 320, 320, 0, 0), (B * 4096, 2560, 320, 0, 0), (B * 4096, 320, 1280, 0, 0),
              (B * 1024, 1920, 640, 0, 0), (B * 1024, 5120, 640, 0, 0), (B * 1024, 640, 2560, 0, 0),
              (B * 256, 3840, 1280, 0, 0), (B * 256, 10240, 1280, 0, 0), (B * 256, 1280, 5120, 0, 0),
              (B * 64, 10240, 1280, 0, 0), (B * 77, 640, 768, 0, 0),
              (B * 4096, 320, 960, 0, 1), (B * 4096, 320, 2560, 0, 1), (B * 1024, 640, 5120, 0, 1)]
    for (M, N, K, amn, bmn) in shapes:
        A = bf(M, K); Bm = bf(K, N) if bmn else bf(N, K)
        t = timeit(lambda: ops.gemm(A, Bm, b_mn=bool(bmn)))
        print(f"M={M:6d} N={N:5d} K={K:5d} bmn={bmn}: {t:7.3f} ms {2.0*M*N*K/t/1e9:7.1f} TF/s")
    # weight-gradient (split-K, both MN-major)
    for (m, C, R) in [(B * 4096, 960, 320), (B * 1024, 1920, 640), (B * 256, 3840, 1280), (B * 77, 640, 768)]:
        dY = bf(m, C); X = bf(m, R); acc = torch.zeros(C, R, device=dev)
        tiles = ((C + 127) // 128) * ((R + 127) // 128)
        splits = max(1, min((m + 63) // 64, (2 * 148 + tiles - 1) // tiles))
        t = timeit(lambda: ops.gemm(dY, X, a_mn=True, b_mn=True, out=acc, accumulate=True, splits=splits))
        print(f"dW m={m:6d} C={C:5d} R={R:5d} splits={splits}: {t:7.3f} ms {2.0*m*C*R/t/1e9:7.1f} TF/s")
if "conv" in what:
    print("== conv3x3 (B,H,W,Cin->Cout)")
    for (H, Cin, Cout) in [(64, 320, 320), (64, 640, 320), (64, 960, 320), (32, 640, 640), (32, 1280, 640), (32, 1920, 640),
                           (16, 1280, 1280), (16, 2560, 1280), (8, 1280, 1280), (8, 2560, 1280), (64, 320, 640)]:
        x = bf(B, H, H, Cin); w = bf(9, Cout, Cin, scale=0.05)
        t = timeit(lambda: ops.conv3x3(x, w))
        print(f"H={H:3d} {Cin:5d}->{Cout:5d}: {t:7.3f} ms {2.0*B*H*H*9*Cin*Cout/t/1e9:7.1f} TF/s")
if "norm" in what:
    print("== norms (GB/s of algorithmic traffic)")
    for (HW, C) in [(4096, 320), (4096, 960), (1024, 640), (1024, 1920), (256, 1280), (256, 2560), (64, 2560)]:
        x = bf(B, HW, C); g = torch.ones(C, device=dev); b_ = torch.zeros(C, device=dev)
        y, st = ops.groupnorm_fwd(x, g, b_, 32, 1e-5, True)
        t = timeit(lambda: ops.groupnorm_fwd(x, g, b_, 32, 1e-5, True))
        tb = timeit(lambda: ops.groupnorm_bwd(x, y, g, b_, st, 32, 1e-5, True))
        n = x.numel() * 2
        print(f"GN HW={HW:5d} C={C:5d}: fwd {t:6.3f} ms {3*n/t/1e6:7.0f} GB/s | bwd {tb:6.3f} ms {5*n/tb/1e6:7.0f} GB/s")
    for (rows, C) in [(B * 4096, 320), (B * 1024, 640), (B * 256, 1280)]:
        x = bf(rows, C); g = torch.ones(C, device=dev); b_ = torch.zeros(C, device=dev)
        y, st = ops.layernorm_fwd(x, g, b_, 1e-5)
        t = timeit(lambda: ops.layernorm_fwd(x, g, b_, 1e-5))
        tb = timeit(lambda: ops.layernorm_bwd(x, y, g, st, 1e-5))
        n = x.numel() * 2
        print(f"LN rows={rows:6d} C={C:5d}: fwd {t:6.3f} ms {2*n/t/1e6:7.0f} GB/s | bwd {tb:6.3f} ms {3*n/tb/1e6:7.0f} GB/s")
    h = bf(B * 4096, 2560)
    t = timeit(lambda: ops.geglu_fwd(h))
    print(f"GEGLU fwd {t:6.3f} ms {h.numel()*2*1.5/t/1e6:7.0f} GB/s")
