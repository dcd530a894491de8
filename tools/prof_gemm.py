"""A few GEMM shapes for `ncu --set full -k regex:e4t_gemm`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch
from e4t_b200 import ops
B = 16
shapes = [(B * 4096, 960, 320), (B * 4096, 320, 320), (B * 1024, 5120, 640), (B * 256, 10240, 1280)]
ts = [((torch.randn(M, K, device="cuda")).to(torch.bfloat16), (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)) for M, N, K in shapes]
for _ in range(2):
    for A, Bm in ts:
        ops.gemm(A, Bm)
torch.cuda.synchronize()
