"""How much of a step is GPU-idle gaps between kernels?  torch.profiler (CUPTI) over one step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch, bench
from torch.profiler import profile, ProfilerActivity
from e4t_b200.engine import PretrainStep
dev = torch.device("cuda", 0)
unet, enc, text = bench.build_models(dev)
step = PretrainStep(unet, enc, text, 49408, 320)
bs = [bench.to_device(bench.host_batch(16, 42 + i), dev) for i in range(2)]
for i in range(3):
    step(bs[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(bs[0]); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
span = ks[-1][1] - ks[0][0]
busy = 0; cur_s, cur_e = ks[0][0], ks[0][1]
for s, e, _ in ks[1:]:
    if s > cur_e:
        busy += cur_e - cur_s; cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print(f"kernels {len(ks)}  span {span/1e3:.2f} ms  busy(union) {busy/1e3:.2f} ms  idle {100*(1-busy/span):.1f}%  sum {sum(e-s for s,e,_ in ks)/1e3:.2f} ms")
import collections
agg = collections.Counter()
for s, e, n in ks: agg[n[:60]] += e - s
for n, t in agg.most_common(14): print(f"{t/1e3:8.2f} ms  {n}")
