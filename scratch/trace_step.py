"""Kernel timeline of ONE graph-replayed train step (torch.profiler / CUPTI): start, duration, stream, name."""
import sys, json, torch
sys.path.insert(0, '.')
import bench
from elimrec_b200.data import Config
from elimrec_b200.model import EliMRec
from elimrec_b200.sampler import PairwiseSamplerV2
dev = torch.device('cuda:0')
ds, name = bench.build_dataset('tiktok')
conf = Config(**{"data.input.dataset": name, "topks": [20], "device": dev, "alpha": 0.5, "batch_size": 2048})
torch.manual_seed(2022)
model = EliMRec(conf, ds).to(dev)
model.make_optimizer()
smp = PairwiseSamplerV2(ds, batch_size=2048, mode="device", device=dev)
u, p, n = smp.sample_epoch_device(2048 * 8)
b = [(u[i*2048:(i+1)*2048], p[i*2048:(i+1)*2048], n[i*2048:(i+1)*2048]) for i in range(8)]
run = model.make_graphed_step()
for x in b[:4]:
    run(*x)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for x in b[4:7]:
        run(*x)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# keep the last replay: find kernels after the last prep_multi_kernel
idx = [i for i, e in enumerate(ev) if 'prep_multi' in e.name]
ev = ev[idx[-1]:] if idx else ev
t0 = ev[0].time_range.start
out = []
for e in ev:
    out.append(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f} s{getattr(e, 'stream', -1) if hasattr(e,'stream') else -1} {e.name[:90]}")
open('gpurun_out/trace_step.txt', 'w').write("\n".join(out) + "\n")
print("\n".join(out[:120]))
print("total span us", ev[-1].time_range.end - t0)
