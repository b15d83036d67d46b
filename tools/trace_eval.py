#!/usr/bin/env python
"""Kernel timeline of ONE evaluate() right after a training step (table completion + ranking), torch.profiler / CUPTI.
Usage (GPU box): python tools/trace_eval.py [workload] > profiles/<name>.txt"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from elimrec_b200.data import Config  # noqa: E402
from elimrec_b200.model import EliMRec  # noqa: E402
from elimrec_b200.sampler import PairwiseSamplerV2  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "tiktok"
    dev = torch.device("cuda:0")
    ds, name = bench.build_dataset(workload)
    conf = Config(**{"data.input.dataset": name, "topks": [20], "device": dev, "alpha": 0.5, "batch_size": 2048})
    torch.manual_seed(2022)
    model = EliMRec(conf, ds).to(dev)
    model.make_optimizer()
    smp = PairwiseSamplerV2(ds, batch_size=2048, mode="device", device=dev)
    run = model.make_graphed_step(device_sampler=smp)
    for _ in range(4):
        run()
    model.eval()
    ev = model.test_evaluator.evaluator
    ev.evaluate(model)
    run()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    # host-side phases of the same call, without the profiler
    def phase(fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t)
    run()
    p1 = phase(lambda: model.all_users)
    p2 = phase(lambda: model.rank_tables())
    p3 = phase(lambda: model.rank_tc_tables())
    p4 = phase(lambda: ev.evaluate(model))
    print(f"# phases (ms, host wall incl. sync): complete tables {p1:.2f} | normalise heads {p2:.2f} | fp16 hi/lo split {p3:.2f} | rank + metrics {p4:.2f}")
    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        ev.evaluate(model)
        torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    evs = [e for e in prof.profiler.kineto_results.events() if e.device_type() == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.start_ns())
    t0 = evs[0].start_ns()
    print(f"# evaluate() right after a training step, {workload}-shape ({wall * 1e3:.2f} ms wall under the profiler); "
          "columns: start_us duration_us kernel")
    for e in evs:
        nm = e.name().replace("(anonymous namespace)::", "").replace("<unnamed>::", "").split("(")[0][:80]
        print(f"{(e.start_ns() - t0) / 1e3:10.1f} {e.duration_ns() / 1e3:9.1f}  {nm}")
    print(f"# span {(evs[-1].start_ns() + evs[-1].duration_ns() - t0) / 1e3:.1f} us")


if __name__ == "__main__":
    main()
