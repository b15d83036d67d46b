#!/usr/bin/env python
"""bench.py - BPR train triples/sec (+ full-rank eval users/sec) on synthetic Tiktok-shape data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload tiktok] [--impl ours|reference]

One JSON line on stdout (rank 0).  Contract: see the task statement; in short
  value        whole-job triples/s of K train steps, each = sample the batch (device Philox sampler, inside the step's CUDA
               graph) + fwd + bwd + Adam; dataset and parameters resident in HBM
  e2e          same metric through the reference-facing API exactly as main.py:92-102 drives it: whole epochs of
               `for users, pos, neg in PairwiseSamplerV2: loss = step(...); loss.item()` with the HOST compat sampler
               (bit-exact libc stream + numpy shuffle) INSIDE the timed region - overlapped on a prefetch thread
               (`e2e.value`) and serial (`e2e.serial_sampler_value`) - per step H2D of the batch from pinned memory and
               a D2H read of the loss
  eval         users/s of one full `evaluate()` (TIE, K=20, Precision/Recall/NDCG), device + e2e
  roofline     the propagation SpMM (wide launch) against the measured HBM copy bandwidth
  cpu_baseline the oracle port (torch-CPU restatement of the reference) timed on this host, rank 0
`--impl reference` times that same CPU port as the reference arm (the reference is Python + Cython
and cannot travel to the GPU box; the port is pinned to it by tests/golden).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 2048
TOPK = 20
METRIC = "bpr_train_triples_per_sec"
UNIT = "triples/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="tiktok")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--no-10x", action="store_true", help="skip the tiktok-10x scaling section (scale_10x)")
    ap.add_argument("--no-e2e", action="store_true", help="exploratory runs only: skip the host-sampler arm (e2e = null)")
    ap.add_argument("--eval-users", type=int, default=0, help="0 = all test users")
    ap.add_argument("--cuda-graph", type=int, default=1)
    ap.add_argument("--lazy-tables", type=int, default=1,
                    help="1: build all_users/all_items/all_s_embs on demand (at evaluation) instead of every step")
    ap.add_argument("--linear", type=int, default=1,
                    help="1 (default): linear schedule - modality graphs by linearity from one 64-wide propagation + constant "
                         "tables; 0: the row-sparse slab schedule of round 1")
    ap.add_argument("--set", action="append", default=[], metavar="KEY=VALUE",
                    help="exploratory runs only: extra model config entries (python literals), e.g. --set wgrad_groups=\"late\"")
    ap.add_argument("--two-hop-masks", type=int, default=0, help="linear schedule: layer L-1 / first backward hop under the two-hop row masks (1) or dense (0)")
    ap.add_argument("--fused-layer-grad", type=int, default=0,
                    help="1: layer-mean gradient added in the backward SpMM epilogues; 0 (default, measured faster): a scatter "
                         "kernel after each SpMM, hidden under the other stream's SpMM")
    ap.add_argument("--parallel", default="colshard", choices=["colshard", "dp", "rowshard"],
                    help="N>1: colshard (default) = column-sharded linear schedule: every rank owns 64/N embedding columns, draws "
                         "its own batch, propagation needs no communication (weak scaling); dp = data-parallel replicas with "
                         "averaged gradients (weak scaling); rowshard = the row-sharded all-gather design (strong scaling: every "
                         "rank works on the SAME batch)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.p, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_dataset(workload):
    from elimrec_b200 import synth
    from elimrec_b200.data import Dataset
    t0 = time.time()
    inter, feats = synth.make_shape(workload)
    name = "kwai" if workload == "kwai" else workload
    ds = Dataset(None, interactions=inter, features=feats, name=name)
    log(f"[bench] dataset {workload}: U={ds.num_users} I={ds.num_items} E_train={ds.train_matrix.nnz} ({time.time() - t0:.1f}s)")
    return ds, name


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_port_run(ds, name, steps, warmup, budget_s, eval_users=256):
    import torch
    from oracle.ref_model import OracleEliMRec
    from oracle import ref_eval
    torch.manual_seed(2022)
    U, I = ds.num_users, ds.num_items
    mods = "v" if name == "kwai" else "vat"
    dims = {m: getattr(ds, f"{m}_feat").shape[1] for m in mods}
    G = 1 + len(mods)
    xav = lambda *s: torch.nn.init.xavier_uniform_(torch.empty(*s))
    params = {"embedding_user.weight": xav(U, 64), "embedding_item.weight": xav(I, 64)}
    for m in mods:
        params[f"{m}_dense.weight"], params[f"{m}_dense.bias"] = xav(64, dims[m]), torch.zeros(64)
    for s in ("user", "item"):
        params[f"embedding_{s}_after_GCN.weight"], params[f"embedding_{s}_after_GCN.bias"] = xav(64, 64 * G), torch.zeros(64)
    for m in mods:
        params[f"s_dense_{m}.weight"], params[f"s_dense_{m}.bias"] = xav(64, 64), torch.zeros(64)
    orc = OracleEliMRec(params, {m: getattr(ds, f"{m}_feat") for m in mods}, ds.train_matrix, U, I, kwai=(name == "kwai"),
                        alpha=0.5)
    opt = torch.optim.Adam(list(orc.p.values()), lr=1e-3, weight_decay=1e-4)
    rng = np.random.default_rng(0)
    tm = ds.train_matrix
    times = []
    t_start = time.time()
    done = 0
    for s in range(warmup + steps):
        u = rng.integers(0, U, BATCH)
        p = tm.indices[tm.indptr[u]]
        n = rng.integers(0, I, BATCH)
        t0 = time.time()
        loss = orc.bpr_loss(u, p, n)
        opt.zero_grad()
        loss.backward()
        opt.step()
        float(loss)
        dt = time.time() - t0
        if s >= warmup:
            times.append(dt)
        done += 1
        if time.time() - t_start + dt > budget_s and len(times) >= 1:
            break
    step_s = float(np.mean(times))
    # eval sample
    users = list(ds.get_user_test_dict().keys())[:eval_users]
    truth = ds.get_user_test_dict()
    t0 = time.time()
    ref_eval.evaluate(lambda us: orc.predict(us, "TIE").numpy(), ds.get_user_train_dict(), {x: truth[x] for x in users},
                      top_k=[TOPK], batch_size=128)
    ev_s = time.time() - t0
    return dict(step_s=step_s, timed_steps=len(times), eval_users_per_s=len(users) / ev_s, eval_users=len(users),
                threads=torch.get_num_threads())


def load_peaks():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return pk["hbm_gbs"], pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1384.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1384.0, "fallback (B200_PROFILING.md)"


def steady_pair_us(model):
    """Average duration of the dense 64-wide propagation launch (elimrec_spmm64_pair over both CSR halves of the model's own
    graph and slabs), CUDA events around a CUDA graph of 10 back-to-back launches replayed 5 times: the kernel as it runs
    inside the graphed step (operands L2-resident, no host launch gaps), which per-launch events in eager mode overstate."""
    import torch
    from elimrec_b200 import ops
    ws, g, U = model._ws, model.graph, model.num_users
    Eu, Ei = model.embedding_user.weight.detach(), model.embedding_item.weight.detach()
    out = ws["P"][1]
    fn = lambda: ops.spmm64_pair(g.ui, g.iu, Ei, Eu, out[:U], out[U:])
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(10):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / 50


def rooflines_of(model, agg, workload, B, prefix=""):
    """Per-kernel rooflines from the serialised CUDA-event profile `agg` {tag: [ms, ...]} of a few training steps.
    Returns (headline roofline of the dominant propagation kernel, list for every family that has a stated bound)."""
    hbm, tf_bf16, which = load_peaks()
    g = model.graph
    U, I = model.num_users, model.num_items
    M = len(model.mods)
    Fw = 64 if model.linear else 64 * (1 + M)
    tot = sum(sum(v) for v in agg.values())
    out, head = [], None
    avg_s = lambda tag: 1e-3 * sum(agg[tag]) / len(agg[tag])

    def add(kernel, tag, bound, units, peak, unit, **extra):
        if tag not in agg:
            return None
        t = avg_s(tag)
        ach = units / t / (1e9 if unit == "GB/s" else 1e12)
        r = {"kernel": prefix + kernel, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
             "avg_launch_us": 1e6 * t, "launches_timed": len(agg[tag]), "share_of_step": sum(agg[tag]) / tot,
             ("bytes_per_launch" if unit == "GB/s" else "flops_per_launch"): units, **extra}
        out.append(r)
        return r

    # propagation SpMM, algorithmic bytes per launch: nnz*(4+4) + (rows_out+1)*4 + (rows_in + rows_out)*F*4.
    # Timed on launches that are exactly ONE kernel: the half without split rows ("w" tag).  Where both halves have split
    # rows, one event pair brackets the whole-row + split-row launches of a half (they run side by side).
    alg_of = lambda h: float(h.nnz * 8 + (h.n_rows + 1) * 4 + (h.n_cols + h.n_rows) * Fw * 4)
    half = g.ui if g.ui.n_heavy_seg == 0 else (g.iu if g.iu.n_heavy_seg == 0 else None)
    traffic = None
    try:      # L2->SM / DRAM bytes per launch from the committed ncu --set full capture of this kernel (tiktok shape)
        if workload == "tiktok":
            traffic = json.load(open(os.path.join(ROOT, "profiles", "spmm_traffic.json"))).get(f"dram_bytes_per_launch_w{Fw}")
    except Exception:
        pass
    fam = sum(sum(v) for k, v in agg.items() if k.startswith("spmm"))
    common = dict(traffic=traffic, peak_source=which, spmm_family_share_of_step=fam / tot)
    if model.linear and "spmm64_pair" in agg:
        # linear schedule: ONE launch per propagation layer covers both CSR halves (elimrec_spmm64_pair); timed on the
        # unmasked launches (forward layer 1).  Algorithmic bytes: both halves' indices + values, the [N x 64] slab in and out.
        alg = alg_of(g.ui) + alg_of(g.iu)
        head = add("spmm64_pair_kernel (one 64-wide propagation layer, both CSR halves in one launch, unmasked)", "spmm64_pair", "hbm",
                   alg, hbm, "GB/s", gathered_bytes_per_launch=float(g.nnz * 256), **common)
        if head is not None:
            head["in_step_event_us"] = head["avg_launch_us"]      # per-launch events of the eager profile pass (launch gaps inside)
            us = steady_pair_us(model)
            head.update(avg_launch_us=us, achieved=alg / (us * 1e-6) / 1e9, frac=alg / (us * 1e-6) / 1e9 / hbm, launches_timed=50,
                        timing="CUDA events around a graph of 10 back-to-back launches on the model's own graph / slabs, 5 replays")
            head["l2_gather_gbps"] = g.nnz * 256 / (us * 1e-6) / 1e9
        if "spmm64_pair+adam" in agg:      # last backward hop with the optimizer in its epilogue: + 7 Adam streams, - the slab store
            add("spmm64_pair_kernel + fused Adam epilogue (last backward hop: propagation + optimizer on both embedding tables)",
                "spmm64_pair+adam", "hbm", alg - (U + I) * 64 * 4 + 7.0 * (U + I) * 64 * 4 + (U + I) * 64 * 4, hbm, "GB/s",
                peak_source=which, note="algorithmic bytes = indices + gathered slab + 7 Adam streams + the pre-update snapshot")
    elif f"spmm{Fw}w" in agg and half is not None:
        head = add(f"spmm_seg_kernel<{Fw}> (whole rows, {'user' if half is g.ui else 'item'}-row half, unmasked)", f"spmm{Fw}w",
                   "hbm", alg_of(half), hbm, "GB/s", **common)
        other = g.iu if half is g.ui else g.ui
        add(f"spmm_seg_kernel<{Fw}> (whole-row + split-row launch, {'item' if half is g.ui else 'user'}-row half, unmasked)",
            f"spmm{Fw}", "hbm", alg_of(other), hbm, "GB/s", peak_source=which)
    elif f"spmm{Fw}" in agg:
        head = add(f"spmm_seg_kernel<{Fw}> (whole-row + split-row launch of one half, mean over the two halves)", f"spmm{Fw}",
                   "hbm", 0.5 * (alg_of(g.ui) + alg_of(g.iu)), hbm, "GB/s", **common)
    if head is not None:
        slab_mb = (U + I) * Fw * 4 / 1e6
        head["note"] = (f"operand slab {slab_mb:.0f} MB " + ("fits the 126 MB L2: the gather is bound by L2->SM bandwidth (~12.4 TB/s "
                        "cap, B300_MICROARCH.md), DRAM carries only the compulsory bytes" if slab_mb < 100 else "exceeds the 126 MB L2: "
                        "the gather misses to HBM at row granularity") + "; achieved = algorithmic bytes / launch time (DESIGN.md 3.1)")
    # Adam: 7 streams x 4 B per parameter element (read p, g, m, v; write p, m, v)
    n_par = sum(p.numel() for n_, p in model._params().items()
                if not (model.linear and "spmm64_pair+adam" in agg and n_.startswith(("embedding_user.w", "embedding_item.w"))))
    add("adam_multi_kernel (" + ("small weights; the tables ride in the last hop" if "spmm64_pair+adam" in agg else "all parameters") +
        ", one launch)", "elimrec_adam_apply_multi", "hbm", 28.0 * n_par, hbm, "GB/s", peak_source=which)
    if model.linear:
        Kt = model._lin_Ktot
        add("gather_rows_kernel (Zbar rows at the 3B instance rows)", "elimrec_gather_rows", "hbm", 2.0 * 3 * B * Kt * 4, hbm, "GB/s",
            peak_source=which)
        # TF32 runs at half the bf16 rate
        add("linear_tf32_fwd_kernel (modality blocks on the instance rows, tcgen05 kind::tf32)", "lin_modal_tc", "tensor",
            2.0 * 3 * B * 64 * Kt, tf_bf16 / 2, "TFLOP/s", peak_source=which + ", TF32 = bf16 / 2",
            note="3B x Ktot x 64 per step: launch-latency sized, not a throughput kernel")
        if "lin_modal_x3" in agg:      # one launch per modality; the widest one
            kmax = max(model._lin_Kp)
            t_big = max(agg["lin_modal_x3"])
            fl = 3 * 2.0 * 3 * B * 64 * kmax
            out.append({"kernel": prefix + f"linear_x3_fwd_kernel (widest modality block on the instance rows, 3xTF32, K={kmax})",
                        "bound": "tensor", "achieved": fl / (1e-3 * t_big) / 1e12, "peak": tf_bf16 / 2, "unit": "TFLOP/s",
                        "frac": fl / (1e-3 * t_big) / 1e12 / (tf_bf16 / 2), "avg_launch_us": 1e3 * t_big, "flops_per_launch": fl,
                        "share_of_step": sum(agg["lin_modal_x3"]) / tot, "peak_source": which + ", TF32 = bf16 / 2",
                        "note": "3B rows per step: launch-latency sized, not a throughput kernel"})
    else:
        feat_b = float(sum(model._feat[m].shape[1] for m in model.mods) * I * 4 + I * 64 * 4 * M)
        flops = 2.0 * I * 64 * sum(model._feat[m].shape[1] for m in model.mods)
        r = add("linear_tf32_fwd_kernel (all modal projections, one persistent launch): feature stream", "proj_fwd_tc", "hbm",
                feat_b, hbm, "GB/s", peak_source=which)
        if r is not None:
            r["tensor_tflops"] = flops / (r["avg_launch_us"] * 1e-6) / 1e12
            r["tensor_frac_of_tf32_peak"] = r["tensor_tflops"] / (tf_bf16 / 2)
            r["note"] = "32 flop per fp32 feature byte: HBM-bound on the feature stream, the tensor pipe only has to keep up"
        if "proj_wgrad_tc" in agg:      # one launch pair per modality; report the largest (text / visual) one
            dmax = max(model._feat[m].shape[1] for m in model.mods)
            t_big = max(agg["proj_wgrad_tc"])
            out.append({"kernel": prefix + f"linear_tf32_wgrad_kernel + reduce (widest modality, D={dmax})", "bound": "hbm",
                        "achieved": (I * dmax * 4 + I * 64 * 4) / (1e-3 * t_big) / 1e9, "peak": hbm, "unit": "GB/s",
                        "frac": (I * dmax * 4 + I * 64 * 4) / (1e-3 * t_big) / 1e9 / hbm, "avg_launch_us": 1e3 * t_big,
                        "share_of_step": sum(agg["proj_wgrad_tc"]) / tot, "peak_source": which,
                        "tensor_tflops": 2.0 * I * 64 * dmax / (1e-3 * t_big) / 1e12})
    return head, out


def measure_10x(args, rank, world, dev, log):
    """BASELINE.json configs[4]: the 10x-scaled Tiktok-shape graph.  At N > 1 both multi-GPU designs, device-timed, max over
    ranks: `colshard` (column-sharded linear schedule, every rank its own batch: weak scaling) and `rowshard` (the design the
    north star names: users / items row-sharded, an NCCL all-gather of the propagated shards per GCN layer, every rank the SAME
    batch: strong scaling) with its time split into compute and collectives.  At N = 1 the single-GPU step they scale from."""
    import torch
    import torch.distributed as dist
    from elimrec_b200.data import Config
    from elimrec_b200.model import EliMRec
    from elimrec_b200.sampler import PairwiseSamplerV2
    ds, name = build_dataset("tiktok10x")
    out = {"workload": f"tiktok10x-shape EliMRec train step, batch {BATCH}/GPU, U={ds.num_users} I={ds.num_items} "
                       f"E_train={ds.train_matrix.nnz}", "n_gpus": world}

    def timed(step, n_warm, n_steps):
        for _ in range(n_warm):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_steps):
            step()
        b.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / n_steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    conf = lambda: Config(**{"data.input.dataset": name, "topks": [TOPK], "device": dev, "alpha": 0.5, "batch_size": BATCH})
    torch.manual_seed(2022)
    if world == 1:
        model = EliMRec(conf(), ds).to(dev)
        model.make_optimizer()
        smp = PairwiseSamplerV2(ds, batch_size=BATCH, mode="device", device=dev, seed=2022)
        run = model.make_graphed_step(device_sampler=smp)
        ms = timed(run, 3, 10)
        out["single_gpu"] = {"ms_per_step": ms, "value": BATCH / (ms / 1e3), "unit": UNIT, "schedule": "linear, CUDA graph, device sampler"}
        return out
    from elimrec_b200.colshard import ColShardedEliMRec
    from elimrec_b200.sharded import ShardedEliMRec
    model = ColShardedEliMRec(conf(), ds).to(dev)
    model.make_optimizer()
    smp = PairwiseSamplerV2(ds, batch_size=BATCH, mode="device", device=dev, seed=2022 + rank)
    run = model.make_graphed_step(device_sampler=smp)
    ms = timed(run, 3, 10)
    out["colshard"] = {"ms_per_step": ms, "value": world * BATCH / (ms / 1e3), "unit": UNIT, "scaling": "weak",
                       "triples_per_optimizer_step": world * BATCH,
                       "collectives_per_step": "1 all-gather of the triples, 2 all-to-alls of the instance rows "
                                               f"({world * 3 * BATCH * 2 * (64 // world) * 4 / 1e6:.1f} MB per rank each), 1 all-reduce of the "
                                               "small weight gradients; no collective in the propagation"}
    log(f"[bench] rank {rank}: 10x colshard {ms:.3f} ms/step")
    run = None
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    model.comm.close()
    del model
    gc.collect()
    torch.cuda.empty_cache()
    model = ShardedEliMRec(conf(), ds).to(dev)
    model.make_optimizer()
    smp = PairwiseSamplerV2(ds, batch_size=BATCH, mode="device", device=dev, seed=2022)      # the same batch on every rank
    bu, bp, bn = (torch.empty(BATCH, dtype=torch.int64, device=dev) for _ in range(3))

    def step():
        smp.sample_batch_device(model._adam.step_dev, bu, bp, bn)
        model.train_step(bu, bp, bn)
    ms_full = timed(step, 2, 5)
    # the same step with every collective turned into a no-op: what is left is compute (results are garbage, timing is not)
    model._ag = lambda slab, blk: None
    model._ar = lambda t: None
    try:
        ms_comp = timed(step, 1, 5)
    finally:
        del model._ag, model._ar
    N = ds.num_users + ds.num_items
    out["rowshard"] = {"ms_per_step": ms_full, "value": BATCH / (ms_full / 1e3), "unit": UNIT, "scaling": "strong",
                       "ms_without_collectives": ms_comp, "collective_ms": ms_full - ms_comp,
                       "collectives_per_step": f"10 all-gathers of propagated slabs (5 x [N x 256] = {N * 256 * 4 / 1e6:.0f} MB and 5 x [N x 64] = "
                                               f"{N * 64 * 4 / 1e6:.0f} MB each, fp32), 1 all-reduce of the instance rows, 1 of the "
                                               "projection weight gradients",
                       "bound_by": "the per-layer all-gathers (NVLink), not the sharded SpMM"}
    log(f"[bench] rank {rank}: 10x rowshard {ms_full:.3f} ms/step ({ms_comp:.3f} without collectives)")
    del model
    gc.collect()
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: route everything else a library may print there (e.g. NCCL's version
    # banner) to stderr by pointing fd 1 at stderr and keeping a private handle on the real stdout
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(line):
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        import torch
        try:      # torchrun exports OMP_NUM_THREADS=1: the reference arm uses every core this process may run on
            torch.set_num_threads(len(os.sched_getaffinity(0)))
        except Exception:
            pass
        ds, name = build_dataset(args.workload)
        r = cpu_port_run(ds, name, args.steps, args.warmup, budget_s=200.0)
        v = BATCH / r["step_s"]
        sample = (f"{r['timed_steps']} of {args.steps} requested full train steps (fwd+bwd+Adam, batch {BATCH}) of the "
                  f"{args.workload}-shape graph, {args.warmup} warm-up; eval on the first {r['eval_users']} test users")
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": r["timed_steps"],
                "warmup": args.warmup, "ms_per_step": 1e3 * r["step_s"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{args.workload}-shape EliMRec train step (sample + fwd + bwd + Adam), batch {BATCH}/GPU, "
                                       f"layer_num 3, recdim 64, U={ds.num_users} I={ds.num_items} E_train={ds.train_matrix.nnz}"},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": sample},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "eval": {"value": r["eval_users_per_s"], "unit": "users/s", "n_users": r["eval_users"]},
                "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    from elimrec_b200 import _lib
    from elimrec_b200.data import Config
    from elimrec_b200.model import EliMRec
    from elimrec_b200.sampler import PairwiseSamplerV2

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ds, name = build_dataset(args.workload)
    conf = Config(**{"data.input.dataset": name, "topks": [TOPK], "device": dev, "alpha": 0.5, "batch_size": BATCH,
                     "lazy_tables": bool(args.lazy_tables), "fused_layer_grad": bool(args.fused_layer_grad),
                     "linear_schedule": bool(args.linear), "two_hop_masks": bool(args.two_hop_masks)})
    for kv in args.set:
        import ast
        k, v = kv.split("=", 1)
        conf[k] = ast.literal_eval(v)
    torch.manual_seed(2022)
    rowshard = world > 1 and args.parallel == "rowshard"
    colshard = world > 1 and args.parallel == "colshard" and bool(args.linear) and bool(args.lazy_tables)
    if rowshard:
        from elimrec_b200.sharded import ShardedEliMRec
        model = ShardedEliMRec(conf, ds).to(dev)
    elif colshard:
        from elimrec_b200.colshard import ColShardedEliMRec
        model = ColShardedEliMRec(conf, ds).to(dev)
    else:
        model = EliMRec(conf, ds).to(dev)
    model.make_optimizer()
    if world > 1 and not (rowshard or colshard):
        model.enable_data_parallel()
    log(f"[bench] rank {rank}/{world}: model ready")
    # data-parallel: every rank draws its own triples; row-sharded: all ranks work on the same batch
    sampler_dev = PairwiseSamplerV2(ds, batch_size=BATCH, mode="device", device=dev, seed=2022 + (0 if rowshard else rank))
    sampler_host = PairwiseSamplerV2(ds, batch_size=BATCH, mode="compat")

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device arm: every step = draw its batch on the device (Philox, inside the graph) + fwd + bwd + Adam ---------------
    use_graph = bool(args.cuda_graph) and not rowshard
    runner = model.make_graphed_step(device_sampler=sampler_dev) if use_graph else None
    bu, bp, bn = (torch.empty(BATCH, dtype=torch.int64, device=dev) for _ in range(3))

    def step_sampled():
        if runner is not None:
            return runner()
        sampler_dev.sample_batch_device(model._adam.step_dev, bu, bp, bn)
        return model.train_step(bu, bp, bn)

    for _ in range(args.warmup):
        step_sampled()
    sync_all()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    _lib.CALLS["launches"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step_sampled()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    log(f"[bench] rank {rank}: device arm {ms / args.steps:.3f} ms/step")
    launches = _lib.CALLS["launches"] if runner is None else runner.launches_per_step * args.steps
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms)
    units = 1 if rowshard else world          # batches processed per step by the whole job
    value = units * BATCH * args.steps / (ms / 1e3)
    final_loss = float(loss)

    # ---- sampler throughput on its own --------------------------------------------------------------------
    samp = None
    if rank == 0:
        t0 = time.perf_counter()
        hu_, hp_, hn_ = sampler_host.sample_epoch_host()
        t_draw = time.perf_counter() - t0
        t0 = time.perf_counter()
        perm = np.random.permutation(hu_.size)
        _ = hu_[perm], hp_[perm], hn_[perm]
        t_shuf = time.perf_counter() - t0
        sampler_dev.sample_epoch_device()
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        sampler_dev.sample_epoch_device()
        b_.record()
        torch.cuda.synchronize()
        samp = {"compat_triples_per_s": hu_.size / (t_draw + t_shuf), "compat_draw_s": t_draw, "compat_shuffle_s": t_shuf,
                "device_triples_per_s": hu_.size / (a_.elapsed_time(b_) / 1e3), "epoch_triples": int(hu_.size),
                "note": "compat = bit-exact replay of the reference's libc rand() stream (sequential by construction) + one numpy "
                        "permutation; device = Philox4x32-10, one thread per triple"}
        del hu_, hp_, hn_, perm

    # ---- e2e arm: main.py:92-102 as written - host sampler, per-batch H2D, loss.item() every step ---------------------------
    runner_h = model.make_graphed_step(host_loss=True) if use_graph else None      # loss lands in pinned memory inside the graph
    steps_per_epoch = len(sampler_host)

    def run_epochs(sampler, n_epochs):
        n = 0
        for _ in range(n_epochs):
            for hu, hp, hn in sampler:
                if hu.numel() != BATCH:           # the short last batch of an epoch (drop_last=False): eager step
                    float(model.train_step(hu, hp, hn))
                else:
                    float(runner_h(hu, hp, hn) if runner_h is not None else model.train_step(hu, hp, hn))   # .item(): main.py:102
                n += int(hu.numel())
        return n

    e2e_epochs = max(1, -(-args.steps // steps_per_epoch))
    e2e_value = e2e_serial = None
    e2e_steps = 0
    if not args.no_e2e:
        np.random.seed(2022 + (0 if rowshard else rank))
        sampler_pf = PairwiseSamplerV2(ds, batch_size=BATCH, mode="compat", prefetch=True, pin=True)
        sampler_pf.rng.seed(1 + (0 if rowshard else rank))        # replicas draw different triples
        run_epochs(sampler_pf, 1)                     # warm-up epoch; leaves the next epoch being sampled on the thread
        sync_all()
        t0 = time.perf_counter()
        n_e2e = run_epochs(sampler_pf, e2e_epochs)
        sync_all()
        e2e_s = time.perf_counter() - t0
        if sampler_pf._next is not None:
            sampler_pf._next[0].join()
        sampler_se = PairwiseSamplerV2(ds, batch_size=BATCH, mode="compat", prefetch=False, pin=True)
        sampler_se.rng.seed(101 + (0 if rowshard else rank))
        sync_all()
        t0 = time.perf_counter()
        n_ser = run_epochs(sampler_se, e2e_epochs)
        sync_all()
        ser_s = time.perf_counter() - t0
        # the same loop with the loss of step i read while step i+1 runs (make_graphed_step(host_loss="deferred")): same copies,
        # same per-step result on the host, one step late - what the per-step launch + sync latency of main.py:102 costs
        def_s = None
        if use_graph and not (rowshard or colshard) and world == 1:
            runner_d = model.make_graphed_step(host_loss="deferred")
            sampler_pf2 = PairwiseSamplerV2(ds, batch_size=BATCH, mode="compat", prefetch=True, pin=True)
            sampler_pf2.rng.seed(7)

            def run_deferred(n_epochs):
                n = 0
                for _ in range(n_epochs):
                    for hu, hp, hn in sampler_pf2:
                        if hu.numel() != BATCH:
                            runner_d.flush()
                            float(model.train_step(hu, hp, hn))
                        else:
                            prev = runner_d(hu, hp, hn)
                            if prev is not None:
                                float(prev)
                        n += int(hu.numel())
                last = runner_d.flush()
                if last is not None:
                    float(last)
                return n
            run_deferred(1)
            sync_all()
            t0 = time.perf_counter()
            n_def = run_deferred(e2e_epochs)
            sync_all()
            def_s = time.perf_counter() - t0
            if sampler_pf2._next is not None:
                sampler_pf2._next[0].join()
            log(f"[bench] rank {rank}: e2e with the loss read one step late: {1e3 * def_s / (e2e_epochs * steps_per_epoch):.3f} ms/step")
        te = torch.tensor([e2e_s, ser_s], device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_value = units * n_e2e / float(te[0])
        e2e_serial = units * n_ser / float(te[1])
        e2e_steps = e2e_epochs * steps_per_epoch
        log(f"[bench] rank {rank}: e2e arm done ({e2e_steps} steps: {1e3 * float(te[0]) / e2e_steps:.3f} ms/step overlapped, "
            f"{1e3 * float(te[1]) / e2e_steps:.3f} serial)")

    # fixed batches for the per-kernel profile and the reference-schedule comparison
    u, p, n = sampler_dev.sample_epoch_device((args.warmup + args.steps) * BATCH)
    n_steps = args.warmup + args.steps
    batches = [(u[i * BATCH:(i + 1) * BATCH], p[i * BATCH:(i + 1) * BATCH], n[i * BATCH:(i + 1) * BATCH]) for i in range(n_steps)]

    # ---- per-kernel CUDA-event profile + roofline of the wide SpMM -------------------------------
    kernels = {}
    roofline, rooflines = None, []
    if rank == 0 and not (rowshard or colshard):
        # rank-local, serialised launches, NO collective (the other ranks do not take part in this pass)
        dp_saved, model._dp = getattr(model, "_dp", False), False
        _lib.PROFILE["on"], _lib.PROFILE["events"] = True, []
        for b in batches[args.warmup:args.warmup + min(5, args.steps)]:
            model.train_step(*b)
        torch.cuda.synchronize()
        _lib.PROFILE["on"] = False
        model._dp = dp_saved
        agg = {}
        for tag, a, b_ in _lib.PROFILE["events"]:
            agg.setdefault(tag, []).append(a.elapsed_time(b_))
        nprof = min(5, args.steps)
        kernels = {k: {"ms_per_step": round(sum(v) / nprof, 4), "launches_per_step": len(v) // nprof,
                       "avg_us": round(1e3 * sum(v) / len(v), 2)} for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))}
        roofline, rooflines = rooflines_of(model, agg, args.workload, BATCH)
    clk = clocks.stop() if rank == 0 else None
    sync_all()
    if world > 1 and not (rowshard or colshard):   # the profile pass above stepped rank 0 only: put every replica back on the same weights
        for prm in model.parameters():
            dist.broadcast(prm.data, src=0)

    # ---- evaluation: full-ranking users/s ----------------------------------------------------------
    ev = None
    if not args.no_eval:
        model.eval()
        model.predict_type = "TIE"
        evalr = model.test_evaluator.evaluator
        users = list(ds.get_user_test_dict().keys())
        if args.eval_users:
            users = users[:args.eval_users]
        kw = dict(test_users=users) if args.eval_users else {}
        log(f"[bench] rank {rank}: evaluation warm-up")
        evalr.evaluate(model, **kw)  # warm-up (device CSR upload, smem opt-in)
        log(f"[bench] rank {rank}: evaluation warm-up done")

        def timed_eval():
            sync_all()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record()
            res_, _ = evalr.evaluate(model, **kw)   # returns host ndarray: the D2H of the result is inside
            b_.record()
            sync_all()
            tm = torch.tensor([a.elapsed_time(b_) / 1e3, time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            return float(tm[0]), float(tm[1]), res_

        # what main.py:111-127 does: evaluate right after training steps - the tables of the last forward are completed
        # (masked layers over every row, modality blocks, fusion Linear + heads), normalised and split inside the timed call
        import gc
        gc.collect()
        reps = []
        for _ in range(3):          # three (training step, evaluation) cycles; the median is reported
            step_sampled()
            reps.append(timed_eval())
        reps.sort(key=lambda x: x[0])
        t_dev, t_wall, res = reps[1]
        log(f"[bench] rank {rank}: evaluation after a step {1e3 * t_dev:.2f} ms (cycles: {[round(1e3 * x[0], 2) for x in reps]})")
        c_dev, c_wall, _ = timed_eval()          # again without a step in between: cached tables
        M_ = len(model.mods)
        flops = 2.0 * len(users) * ds.num_items * 64 * ((1 + M_) * 3 + 3)      # fp16 hi/lo: 3 MMA terms per dot product; + row-mean pass
        _, tfp, _ = load_peaks()
        ev = {"value": len(users) / t_dev, "unit": "users/s", "n_users": len(users), "e2e_value": len(users) / t_wall,
              "cached_tables_value": len(users) / c_dev, "ms": 1e3 * t_dev, "ms_cached_tables": 1e3 * c_dev,
              "ms_cycles": [1e3 * x[0] for x in reps],
              "topk": TOPK, "predict_type": "TIE", "result": [float(x) for x in res],
              "note": "value: evaluate() right after a training step, table completion included; cached_tables_value: a second "
                      "evaluate() of the same forward",
              "roofline": {"kernel": "rank_tc_kernel (tcgen05 kind::f16 hi/lo pairs; row-mean pass + TIE pass)", "bound": "tensor",
                           "achieved": flops / c_dev / 1e12, "peak": tfp, "unit": "TFLOP/s", "frac": flops / c_dev / 1e12 / tfp,
                           "flops_per_launch": flops,
                           "note": "MUFU-bound (1+M exp + 1 rcp per user-item pair), see DESIGN.md 3.6 and profiles/ for the "
                                   "tensor-pipe / XU-pipe percentages of the ncu capture"}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # reported at N=1 only (the other ranks would idle for it)
        r = cpu_port_run(ds, name, steps=2, warmup=1, budget_s=45.0)
        cpu = {"value": BATCH / r["step_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
               "sample": f"{r['timed_steps']} full train steps (batch {BATCH}) of the same {args.workload}-shape graph after 1 warm-up; "
                         f"eval {r['eval_users']} users", "eval_users_per_s": r["eval_users_per_s"]}

    # ---- the reference's schedule (every row of every layer + the full tables, every step) next to the default ----------
    dense = None
    if rank == 0 and world == 1 and args.lazy_tables and not rowshard:
        conf_d = Config(**{"data.input.dataset": name, "topks": [TOPK], "device": dev, "alpha": 0.5, "batch_size": BATCH,
                           "lazy_tables": False})
        torch.manual_seed(2022)
        md = EliMRec(conf_d, ds).to(dev)
        md.make_optimizer()
        rd = md.make_graphed_step() if use_graph else md.train_step
        for b in batches[:args.warmup]:
            rd(*b)
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for b in batches[args.warmup:]:
            ld = rd(*b)
        d1.record()
        torch.cuda.synchronize()
        dms = d0.elapsed_time(d1) / args.steps
        _lib.PROFILE["on"], _lib.PROFILE["events"] = True, []
        for b in batches[args.warmup:args.warmup + 3]:
            md.train_step(*b)
        torch.cuda.synchronize()
        _lib.PROFILE["on"] = False
        agg_d = {}
        for tag, a, b_ in _lib.PROFILE["events"]:
            agg_d.setdefault(tag, []).append(a.elapsed_time(b_))
        rl_d, rls_d = rooflines_of(md, agg_d, args.workload, BATCH, prefix="reference schedule: ")
        dense = {"ms_per_step": dms, "value": BATCH / (dms / 1e3), "unit": UNIT, "final_loss": float(ld), "roofline": rl_d,
                 "note": "lazy_tables=False: the reference's schedule - four graphs propagated (64 + 256 wide), every row of every "
                         "layer, projections over all items, fusion + heads over all U+I rows, every step"}
        rooflines += rls_d
        del md, rd

    tenx = None
    if not args.no_10x and args.workload == "tiktok" and bool(args.linear) and bool(args.lazy_tables) and not rowshard:
        try:
            tenx = measure_10x(args, rank, world, dev, log)
        except Exception as exc:      # the headline must survive a failure of the extra section
            tenx = {"error": f"{type(exc).__name__}: {exc}"}
            log(f"[bench] rank {rank}: scale_10x failed: {tenx['error']}")
    if rank == 0:
        N_ = ds.num_users + ds.num_items
        if model.linear:
            ws_mb = (N_ * 64 * 4 * (1 + 2 + 1 + model.n_layers + 2) + 2 * 3 * BATCH * model._lin_Ktot * 4) / 1e6
            sched = ("linear: the modality graphs by linearity from ONE 64-wide propagation + constant Zbar tables (built once); "
                     + ("layers L-1 / L" if args.two_hop_masks else "layer L") +
                     " and the fusion / head tables only at the rows the loss reads; identical loss / gradients / "
                     "parameters up to fp32 reassociation; full tables completed at evaluation")
        else:
            ws_mb = 900.0
            sched = ("row-sparse: loss-dead rows of the last two layers and of the fusion/head tables are not computed in the step"
                     if args.lazy_tables else "reference: every row of every layer and table, every step")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": ("strong" if rowshard else "weak"), "vs_baseline": None,
                "dtype": "f32 (fp32 propagation / loss / Adam; 3xTF32 tensor-core GEMMs = fp32 accuracy class)", "data": "synthetic",
                "config": {"workload": f"{args.workload}-shape EliMRec train step (sample + fwd + bwd + Adam), batch {BATCH}/GPU, "
                                       f"layer_num 3, recdim 64, U={ds.num_users} I={ds.num_items} E_train={ds.train_matrix.nnz}",
                           "parallelism": ((f"rowshard{world} (all-gather per GCN layer)" if rowshard else
                                            (f"colshard{world} (each rank: 64/{world} embedding columns of every row, its own batch; "
                                             "propagation without communication; 2 all-to-alls of the instance rows + 1 small "
                                             "all-reduce per step)" if colshard else f"dp{world}")) if world > 1 else "single"),
                           "l2_policy": f"no flush: the per-step working set ({ws_mb:.0f} MB: embedding tables + Adam moments + "
                                        "propagation / gradient slabs + gathered constant rows) exceeds the 126 MB L2",
                           "cuda_graph": bool(runner is not None), "lazy_tables": bool(args.lazy_tables),
                           "schedule": sched,
                           "triples_per_optimizer_step": BATCH * units,
                           "sampler": "value: device Philox sampler inside every step's graph; e2e: host compat sampler (bit-exact "
                                      "libc stream) + numpy shuffle, whole epochs, inside the timed region"},
                "e2e": None if args.no_e2e else {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 3 * BATCH * 8, "d2h_bytes_per_step": 4,
                        "steps": e2e_steps, "epochs": e2e_epochs, "serial_sampler_value": e2e_serial,
                        "deferred_loss_value": (units * n_def / def_s) if def_s else None,
                        "note": "main.py:92-102 loop over whole epochs; value: next epoch sampled on a prefetch thread while this one "
                                "trains; serial_sampler_value: each epoch sampled + shuffled before its first step (as the "
                                "reference does), both inside the timed region"},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "rooflines": rooflines, "cpu_baseline": cpu,
                "eval": ev, "sampler": samp, "kernels": kernels, "final_loss": final_loss, "dense_schedule": dense, "scale_10x": tenx}
        emit(line)
    if world > 1:
        # orderly teardown: graphs that captured collectives first, then the library's communicator, then torch's
        runner = runner_h = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        log(f"[bench] rank {rank}: teardown")
        if getattr(model, "comm", None) is not None:
            model.comm.close()
        dist.destroy_process_group()
        log(f"[bench] rank {rank}: done")
    return 0


if __name__ == "__main__":
    sys.exit(main())
