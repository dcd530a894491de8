"""GroupNorm(+SiLU) fwd+bwd at the two largest UNet shapes — target for `ncu --set full -k regex:gn_`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch
from e4t_b200 import ops
for HW, C in ((4096, 320), (1024, 1280)):
    x = torch.randn(16, HW, C, device="cuda").to(torch.bfloat16)
    dy = torch.randn(16, HW, C, device="cuda").to(torch.bfloat16)
    g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    for _ in range(2):
        y, st = ops.groupnorm_fwd(x, g, b, 32, 1e-5, True)
        ops.groupnorm_bwd(x, dy, g, b, st, 32, 1e-5, True)
torch.cuda.synchronize()
