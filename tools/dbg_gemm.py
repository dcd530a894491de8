import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch
from e4t_b200 import ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=6):
    fn(); fn(); ts=[]
    for _ in range(iters):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]
for (M,N,K) in [(65536,960,320),(65536,320,320),(16384,5120,640),(4096,10240,1280)]:
    A=(torch.randn(M,K,device="cuda")).to(torch.bfloat16); B=(torch.randn(N,K,device="cuda")*0.05).to(torch.bfloat16)
    for bn in (0, 64, 128, 256):
        t=timeit(lambda: ops.gemm(A,B,force_bn=bn))
        print(f"dbg={os.environ.get('E4T_GEMM_DEBUG')} tma={os.environ.get('E4T_GEMM_TMA_STORE')} M={M} N={N} K={K} bn={bn}: {t*1e3:7.1f} us {2.0*M*N*K/t/1e9:7.1f} TF/s")
