"""One level-0 self-attention fwd+bwd (B=16, N=4096, 8x40) — target for `ncu --set full -k regex:attn`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch
from e4t_b200 import ops
B, N, C = int(os.environ.get("B", "16")), 4096, 320
q = (torch.randn(B, N, C, device="cuda") * 0.5).to(torch.bfloat16)
k = (torch.randn(B, N, C, device="cuda") * 0.5).to(torch.bfloat16)
v = (torch.randn(B, N, C, device="cuda") * 0.5).to(torch.bfloat16)
do = torch.randn(B, N, C, device="cuda").to(torch.bfloat16)
for _ in range(2):
    o, lse = ops.attn_fwd(q, k, v, 8)
    ops.attn_bwd(q, k, v, o, do, lse, 8)
torch.cuda.synchronize()
