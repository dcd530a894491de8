import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200"), os.path.join(ROOT, "tests")]
import torch
from oracle import e4t_oracle as O
import test_e2e_gpu as T
from tools.graph_bisect2 import try_capture  # noqa  (re-runs its prints; ignore)
print("-----")
unet, sd = T._build_unet(O.TINY_UNET, 1)
bf = lambda *s: torch.randn(*s, device="cuda").to(torch.bfloat16)
temb = torch.randn(2, 256, device="cuda")
ehs = bf(2, 77, 64)
x64 = bf(2, 16, 16, 64)
def run(mod_fn):
    def f():
        xi = x64.clone().requires_grad_(True)
        out = mod_fn(xi)
        out.float().sum().backward()
    return f
r = unet.down_blocks[0].resnets[0]
try_capture("ResnetBlock2D 64->64", run(lambda xi: r(xi, temb)))
a = unet.down_blocks[0].attentions[0]
try_capture("Transformer2DModel", run(lambda xi: a(xi, encoder_hidden_states=ehs).sample))
blk = a.transformer_blocks[0]
tok = bf(2, 256, 64)
def tb():
    t = tok.clone().requires_grad_(True)
    blk(t, encoder_hidden_states=ehs).float().sum().backward()
try_capture("BasicTransformerBlock", tb)
def a1():
    t = tok.clone().requires_grad_(True)
    blk.attn1(t, residual=t).float().sum().backward()
try_capture("attn1 (self)", a1)
def a2():
    t = tok.clone().requires_grad_(True)
    blk.attn2(t, encoder_hidden_states=ehs, residual=t).float().sum().backward()
try_capture("attn2 (cross)", a2)
def a2g():
    t = tok.clone().requires_grad_(True); e = ehs.clone().requires_grad_(True)
    blk.attn2(t, encoder_hidden_states=e, residual=t).float().sum().backward()
try_capture("attn2 (cross, ehs grad)", a2g)
def ff():
    t = tok.clone().requires_grad_(True)
    blk.ff(t, residual=t).float().sum().backward()
try_capture("feed-forward", ff)
d = unet.down_blocks[0]
try_capture("CrossAttnDownBlock2D", run(lambda xi: d(xi, temb, encoder_hidden_states=ehs)[0]))
ds = d.downsamplers[0]
try_capture("Downsample2D", run(lambda xi: ds(xi)))
