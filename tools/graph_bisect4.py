import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200"), os.path.join(ROOT, "tests")]
import torch
from oracle import e4t_oracle as O
import test_e2e_gpu as T
unet, sd = T._build_unet(O.TINY_UNET, 1)
lat = torch.randn(2, 4, 16, 16, device="cuda"); t = torch.tensor([3, 500], device="cuda"); ehs = torch.randn(2, 77, 64, device="cuda")
def fn():
    e = ehs.clone().requires_grad_(True)
    unet(lat, t, e).sample.sum().backward()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): fn()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        try:
            fn()
        except BaseException as e:
            print("INNER EXCEPTION:"); traceback.print_exc()
            raise
except BaseException as e:
    print("OUTER:", type(e).__name__, str(e)[:300])
