#!/usr/bin/env python
"""Kernel timeline of ONE CUDA-graph-replayed training step (torch.profiler / CUPTI): start, duration, kernel.
Usage (GPU box): python tools/trace_step.py [workload] [linear|rowsparse|reference] > profiles/<name>.txt
Columns: start_us duration_us stream kernel (the stream column shows which launches run side by side)."""
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
    sched = sys.argv[2] if len(sys.argv) > 2 else "linear"
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:       # torchrun: the column-sharded step, rank 0 prints its own timeline
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    ds, name = bench.build_dataset(workload)
    conf = Config(**{"data.input.dataset": name, "topks": [20], "device": dev, "alpha": 0.5, "batch_size": 2048, "lazy_tables": sched != "reference",
                     "linear_schedule": sched == "linear"})
    torch.manual_seed(2022)
    if world > 1:
        from elimrec_b200.colshard import ColShardedEliMRec
        model = ColShardedEliMRec(conf, ds).to(dev)
    else:
        model = EliMRec(conf, ds).to(dev)
    model.make_optimizer()
    smp = PairwiseSamplerV2(ds, batch_size=2048, mode="device", device=dev, seed=2022 + rank)
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
    ev = [e for e in prof.profiler.kineto_results.events() if e.device_type() == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.start_ns())
    idx = [i for i, e in enumerate(ev) if "adam_multi" in e.name()]      # a step = everything between two Adam launches
    ev = ev[idx[-2] + 1:idx[-1] + 1]
    t0 = ev[0].start_ns()
    streams = {}
    if rank != 0:
        ev = []
    print(f"# one graph-replayed train step ({sched} schedule" + (f", column-sharded over {world} GPUs, rank 0" if world > 1 else "") +
          f"), {workload}-shape, batch 2048; columns: start_us duration_us stream kernel") if rank == 0 else None
    for e in ev:
        nm = e.name().replace("(anonymous namespace)::", "").replace("<unnamed>::", "").split("(")[0][:70]
        sid = streams.setdefault(e.device_resource_id(), len(streams))
        print(f"{(e.start_ns() - t0) / 1e3:9.1f} {e.duration_ns() / 1e3:8.1f}  s{sid:<2d} {nm}")
    if rank == 0:
        print(f"# span {(ev[-1].start_ns() + ev[-1].duration_ns() - t0) / 1e3:.1f} us")
    if world > 1:
        import gc
        run = None
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        model.comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
