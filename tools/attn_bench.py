"""Attention-core timing of the runtime-selectable variants at the UNet / ViT shapes (CUDA events, L2 flushed between
iterations by rotating over buffers larger than the 126 MB L2).  Not a bench value: it ranks variants.

    python tools/attn_bench.py [fwd] [bwd]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch  # noqa: E402

from e4t_b200 import ops  # noqa: E402

PEAK = 1698.1
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
except Exception:
    pass

SHAPES = [  # (name, B, H, N, M, dh)
    ("unet L0 self", 16, 8, 4096, 4096, 40), ("unet L1 self", 16, 8, 1024, 1024, 80),
    ("unet L2 self", 16, 8, 256, 256, 160), ("unet L0 cross", 16, 8, 4096, 77, 40),
    ("unet L1 cross", 16, 8, 1024, 77, 80), ("vit-h/14", 16, 16, 257, 257, 80), ("clip-l text", 16, 12, 77, 77, 64)]


def timeit(fn, sets, iters=8, warm=2):
    for i in range(warm):
        fn(*sets[i % len(sets)])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(iters):
        fn(*sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def main():
    which = set(sys.argv[1:]) or {"fwd", "bwd"}
    out = []
    for name, B, H, N, M, dh in (SHAPES[:1] if "one" in which else SHAPES):
        C = H * dh
        nbytes = (2 * B * N * C + 2 * B * M * C) * 2
        nsets = max(2, min(6, int(160e6 // nbytes) + 1))
        sets = [tuple((torch.randn(B, n, C, device="cuda") * 0.5).to(torch.bfloat16) for n in (N, M, M)) for _ in range(nsets)]
        flops = 4.0 * N * M * C * B
        row = dict(shape=name, B=B, H=H, N=N, M=M, dh=dh)
        if "fwd" in which:
            os.environ["E4T_ATTN_FWD2"] = "0"
            o_ref, lse_ref = ops.attn_fwd(*sets[0], H)
            for tag in ("0", "s", "d", "dp1", "dp2", "dp3"):
                os.environ["E4T_ATTN_FWD2"] = tag
                o, lse = ops.attn_fwd(*sets[0], H)
                torch.cuda.synchronize()
                ms = timeit(lambda q, k, v: ops.attn_fwd(q, k, v, H), sets)
                row[f"fwd[{tag}]"] = dict(ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1), frac=round(flops / ms / 1e9 / PEAK, 3),
                                          err_vs_legacy=rel(o, o_ref), dlse=(lse - lse_ref).abs().max().item())
            os.environ.pop("E4T_ATTN_FWD2", None)
        if "bwd" in which:
            q, k, v = sets[0]
            o, lse = ops.attn_fwd(q, k, v, H)
            do = torch.randn_like(o)
            ms = timeit(lambda q, k, v: ops.attn_bwd(q, k, v, o, do, lse, H), sets)
            row["bwd"] = dict(ms=round(ms, 4), tflops_2x=round(2 * flops / ms / 1e9, 1), frac_2x=round(2 * flops / ms / 1e9 / PEAK, 3))
        print(json.dumps(row), flush=True)
        out.append(row)
        del sets
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "attn_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
