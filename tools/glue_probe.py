"""Which Python call sites launch the torch (aten) glue kernels of one step?  torch.profiler with stacks over one eager
pre-training step; prints aten ops by device time with their innermost repo frames.   python tools/glue_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from e4t_b200.engine import PretrainStep  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    unet, enc, text = bench.build_models(dev)
    step = PretrainStep(unet, enc, text, placeholder_token_id=49408, class_token_id=320, lr=1.6e-5)
    bs = [bench.to_device(bench.host_batch(16, 42 + i), dev) for i in range(2)]
    for i in range(3):
        step(bs[i % 2])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True) as prof:
        step(bs[0])
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages(group_by_stack_n=12):
        t = getattr(e, "self_device_time_total", None)
        if t is None:
            t = getattr(e, "self_cuda_time_total", 0)
        if not e.key.startswith("aten::") or t <= 0:
            continue
        frames = [f for f in e.stack if "/e4t" in f or "bench.py" in f or "engine.py" in f]
        rows.append((t, e.count, e.key, frames[:3]))
    rows.sort(key=lambda r: -r[0])
    tot = sum(r[0] for r in rows)
    print(f"aten ops with device time: {tot / 1e3:.2f} ms over {sum(r[1] for r in rows)} calls")
    for t, n, k, fr in rows[:45]:
        print(f"{t / 1e3:7.3f} ms {n:4d} x {k:28s} " + " <- ".join(f.split('/')[-1][:60] for f in fr))


if __name__ == "__main__":
    main()
