"""Which part of the step breaks CUDA-graph capture?  Captures pieces separately (tiny models)."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch
from oracle import e4t_oracle as O
from e4t_b200 import ops, functional as FN
from e4t_b200.engine import PretrainStep
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_e2e_gpu as T

unet, enc, text, sds, cfgs, _ = T._build_step()
step = PretrainStep(unet, enc, text, O.PLACEHOLDER_ID, class_token_id=320, lr=1e-3, weight_dtype=torch.float32)
batch = {k: v.cuda() for k, v in O.synth_batch(2, 42, 16, 64).items()}
batch["placeholder_idxs"] = torch.tensor(step.placeholder_idxs(batch["input_ids"]), device="cuda")

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
        print(f"[FAIL] {name}: {type(e).__name__}: {str(e).splitlines()[0][:160]}")
        torch.cuda.synchronize()

x = torch.randn(4096, 64, device="cuda").to(torch.bfloat16); w = torch.randn(64, 64, device="cuda").to(torch.bfloat16)
try_capture("gemm kernel", lambda: ops.gemm(x, w))
gam = torch.ones(64, device="cuda"); bet = torch.zeros(64, device="cuda")
try_capture("groupnorm (memset + 2 kernels)", lambda: ops.groupnorm_fwd(x.view(4, 1024, 64), gam, bet, 32, 1e-5, True))
q = torch.randn(2, 256, 128, device="cuda").to(torch.bfloat16)
try_capture("attention fwd", lambda: ops.attn_fwd(q, q, q, 4))
lat, t, ehs = batch["latents"], batch["timesteps"], torch.randn(2, 77, 64, device="cuda")
def unet_fwd():
    with torch.no_grad(): return unet(lat, t, ehs).sample
try_capture("unet forward (no grad)", unet_fwd)
def unet_fb():
    e = ehs.clone().requires_grad_(True)
    unet(lat, t, e).sample.sum().backward()
try_capture("unet forward+backward", unet_fb)
def enc_f():
    with torch.no_grad():
        maps = unet(lat, t, ehs, return_encoder_outputs=True)["down_block_samples"]
        return enc(batch["pixel_values"], maps)
try_capture("unet enc-half + e4t encoder (no grad)", enc_f)
try_capture("text encoder", lambda: text(input_ids=batch["input_ids"])[0])
try_capture("forward_loss + backward", lambda: step.forward_loss(batch)["loss"].backward())
try_capture("optimizer", lambda: (step.opt.step(1.0), step.opt.zero_grad()))
try_capture("full step", lambda: step._eager_step(batch))
