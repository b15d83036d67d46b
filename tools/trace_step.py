#!/usr/bin/env python
"""Kernel timeline of ONE CUDA-graph-replayed training step (torch.profiler / CUPTI): start, duration, kernel.
Usage (GPU box): python tools/trace_step.py [workload] [lazy] > profiles/<name>.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from elimrec_b200.data import Config  # noqa: E402
from elimrec_b200.model import EliMRec  # noqa: E402
from elimrec_b200.sampler import PairwiseSamplerV2  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "tiktok"
    lazy = len(sys.argv) > 2 and sys.argv[2] == "lazy"
    dev = torch.device("cuda:0")
    ds, name = bench.build_dataset(workload)
    conf = Config(**{"data.input.dataset": name, "topks": [20], "device": dev, "alpha": 0.5, "batch_size": 2048, "lazy_tables": lazy})
    torch.manual_seed(2022)
    model = EliMRec(conf, ds).to(dev)
    model.make_optimizer()
    smp = PairwiseSamplerV2(ds, batch_size=2048, mode="device", device=dev)
    u, p, n = smp.sample_epoch_device(2048 * 8)
    b = [(u[i * 2048:(i + 1) * 2048], p[i * 2048:(i + 1) * 2048], n[i * 2048:(i + 1) * 2048]) for i in range(8)]
    run = model.make_graphed_step()
    for x in b[:4]:
        run(*x)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for x in b[4:7]:
            run(*x)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    idx = [i for i, e in enumerate(ev) if "prep_multi" in e.name]
    ev = ev[idx[-1]:] if idx else ev
    t0 = ev[0].time_range.start
    print(f"# one graph-replayed train step, {workload}-shape, batch 2048; columns: start_us duration_us kernel")
    for e in ev:
        nm = e.name.replace("(anonymous namespace)::", "").split("(")[0][:70]
        print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  {nm}")
    print(f"# span {ev[-1].time_range.end - t0:.1f} us")


if __name__ == "__main__":
    main()
