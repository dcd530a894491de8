import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch
from e4t_b200 import ops, functional as FN
from e4t.weightoffsets import WeightOffsets

def try_capture(name, fn, warm=2):
    try:
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warm): fn()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(f"[ok]   {name}")
    except Exception as e:
        print(f"[FAIL] {name}: {type(e).__name__}: {str(e).splitlines()[0][:120]}")
        try: torch.cuda.synchronize()
        except Exception: pass

bf = lambda *s: torch.randn(*s, device="cuda").to(torch.bfloat16)
x = bf(2, 256, 64).requires_grad_(True); w = bf(64, 64); bias = torch.zeros(64, device="cuda")
try_capture("raw ops in bwd thread? plain torch bwd", lambda: (x.float() * 2).sum().backward())
try_capture("LinearFn fwd+bwd", lambda: FN.LinearFn.apply(x, w, bias, None).float().sum().backward())
g_ = torch.ones(64, device="cuda"); b_ = torch.zeros(64, device="cuda")
try_capture("LayerNormFn", lambda: FN.LayerNormFn.apply(x, g_, b_, 1e-5).float().sum().backward())
try_capture("GroupNormFn", lambda: FN.GroupNormFn.apply(x, g_, b_, 32, 1e-5, True).float().sum().backward())
h = bf(2, 256, 128).requires_grad_(True)
try_capture("GEGLUFn", lambda: FN.GEGLUFn.apply(h).float().sum().backward())
qkv = bf(2, 256, 192).requires_grad_(True)
try_capture("AttentionFn self", lambda: FN.AttentionFn.apply(qkv, None, 4, 0.25).float().sum().backward())
xi = bf(2, 16, 16, 64).requires_grad_(True); w9 = bf(9, 64, 64)
try_capture("Conv3x3Fn", lambda: FN.Conv3x3Fn.apply(xi, w9, w9, bias, None, None).float().sum().backward())
try_capture("ResampleFn", lambda: FN.ResampleFn.apply(xi, 0).float().sum().backward())
wo = WeightOffsets(64, 64).cuda(); W = torch.randn(64, 64, device="cuda")
def wo_fb():
    weff, car = FN.WOEffectiveFn.apply(1, W, *wo.kernel_params())
    FN.WOLinearFn.apply(x, weff, car).float().sum().backward()
try_capture("WOEffective + WOLinear", wo_fb)
wc = torch.randn(4, 64, 3, 3, device="cuda"); bc = torch.zeros(4, device="cuda")
try_capture("ConvOutFn", lambda: FN.ConvOutFn.apply(xi, wc, bc).sum().backward())
try_capture("MeanPoolCatFn", lambda: FN.MeanPoolCatFn.apply(xi, xi).sum().backward())
