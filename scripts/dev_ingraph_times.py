"""Per-kernel device time INSIDE the CUDA-graph replay of the bs=64 forward (torch.profiler / CUPTI timestamps):
development insight only -- numbers taken under a profiler are never reported by bench.py."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from tokenhmr_b200 import synth
from tokenhmr_b200.config import release_config
from tokenhmr_b200.engine import TokenHMREngine
cfg = release_config()
model = TokenHMREngine(cfg, synth.make_state_dict(cfg), synth.make_smpl(cfg), use_cuda_graph=True)
img = synth.make_images(64, cfg).cuda()
for _ in range(8): model.forward({"img": img}, alias_outputs=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): model.forward({"img": img}, alias_outputs=True)
e1.record(); torch.cuda.synchronize()
print(f"unprofiled graph step: {e0.elapsed_time(e1)/20:.3f} ms")
N = 5
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N): model.forward({"img": img}, alias_outputs=True)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name and "Memset" not in e.name]
evs.sort(key=lambda e: e.time_range.start)
print("cuda kernel events:", len(evs))
if not evs: sys.exit(0)
# split into replays by the im2col kernel
agg = collections.OrderedDict(); gaps = 0.0; busy = 0.0
t_first, t_last = evs[0].time_range.start, evs[-1].time_range.end
prev_end = None
for e in evs:
    d = e.time_range.end - e.time_range.start
    name = e.name.split("(")[0].replace("thmr::", "").replace("void ", "")[:60]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += d
    busy += d
    if prev_end is not None and e.time_range.start > prev_end: gaps += e.time_range.start - prev_end
    prev_end = max(prev_end or 0, e.time_range.end)
span = t_last - t_first
print(f"span per replay {span/N/1e3:.3f} ms, kernel-busy {busy/N/1e3:.3f} ms, gaps {gaps/N/1e3:.3f} ms")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} n/replay {n/N:6.1f}  us/launch {us/n:8.2f}  ms/replay {us/N/1e3:7.3f}  share {100*us/busy:5.1f}%")
