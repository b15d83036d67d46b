#!/usr/bin/env python
"""GPU BOX: the step's weight-gradient launch alone (tiktok-shape problem list, batch 2048): exact FFMA vs 3xTF32 tensor cores,
CUDA events around graphs of 20 back-to-back launches.   python tools/wgrad_bench.py [splits ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elimrec_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B = 2048
    R = 3 * B
    g = torch.Generator().manual_seed(0)
    ig = torch.randn(R, 256, generator=g).to(dev)
    Oin = torch.randn(R, 256, generator=g).to(dev)
    dO = torch.randn(R, 256, generator=g).to(dev)
    Kp = [132, 132, 772]
    Zg = torch.randn(R, sum(Kp), generator=g).to(dev)
    outs = [torch.empty(64, 256, device=dev), torch.empty(64, 256, device=dev)] + [torch.empty(64, 64, device=dev) for _ in range(3)] + \
           [torch.empty(64, k, device=dev) for k in Kp]
    bias = [torch.empty(64, device=dev) for _ in range(5)]
    pr = [(ig, 0, Oin, 0, 256, 0, B, outs[0], bias[0], True), (ig, 0, Oin, 0, 256, B, R, outs[1], bias[1], True)]
    for j in range(3):
        pr.append((ig, 64 * (j + 1), Oin, 64 * (j + 1), 64, 0, R, outs[2 + j], bias[2 + j], True))
    ko = 0
    for j, k in enumerate(Kp):
        pr.append((dO, 64 * (j + 1), Zg, ko, k, 0, R, outs[5 + j], None, False))
        ko += k
    gs = torch.tensor([0.5], device=dev)
    flops = 2 * sum((q[6] - q[5]) * 64 * q[4] for q in pr)
    for x3, splits in [(False, 24)] + [(True, int(a)) for a in (sys.argv[1:] or ["4", "8", "16"])]:
        ws = torch.empty(ops.wgrad_multi_ws_floats(pr, splits), device=dev)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(3):
                ops.wgrad_multi(pr, splits, ws, gs, x3=x3)
            st.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                for _ in range(20):
                    ops.wgrad_multi(pr, splits, ws, gs, x3=x3)
            gr.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(5):
                gr.replay()
            e1.record(st)
            st.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 100
        print(f"{'x3 tcgen05' if x3 else 'exact FFMA'} splits={splits}: {us:.1f} us per launch + reduction ({flops / us / 1e6:.1f} TFLOP/s algorithmic)")


if __name__ == "__main__":
    main()
