#!/usr/bin/env python
"""bench.py - BPR train triples/sec (+ full-rank eval users/sec) on synthetic Tiktok-shape data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload tiktok] [--impl ours|reference]

One JSON line on stdout (rank 0).  Contract: see the task statement; in short
  value        whole-job triples/s of K train steps (fwd + bwd + Adam), inputs resident in HBM
  e2e          same metric through the public API with HOST triples: per step H2D of the batch from
               pinned memory and a D2H read of the loss are inside the timed region
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
    ap.add_argument("--eval-users", type=int, default=0, help="0 = all test users")
    ap.add_argument("--cuda-graph", type=int, default=1)
    ap.add_argument("--lazy-tables", type=int, default=1,
                    help="1: build all_users/all_items/all_s_embs on demand (at evaluation) instead of every step")
    ap.add_argument("--fused-layer-grad", type=int, default=0,
                    help="1: layer-mean gradient added in the backward SpMM epilogues; 0 (default, measured faster): a scatter "
                         "kernel after each SpMM, hidden under the other stream's SpMM")
    ap.add_argument("--parallel", default="dp", choices=["dp", "rowshard"],
                    help="N>1: data-parallel replicas (weak scaling, default) or the row-sharded all-gather design "
                         "(strong scaling: every rank works on the SAME batch)")
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
                     "lazy_tables": bool(args.lazy_tables), "fused_layer_grad": bool(args.fused_layer_grad)})
    torch.manual_seed(2022)
    rowshard = world > 1 and args.parallel == "rowshard"
    if rowshard:
        from elimrec_b200.sharded import ShardedEliMRec
        model = ShardedEliMRec(conf, ds).to(dev)
    else:
        model = EliMRec(conf, ds).to(dev)
    model.make_optimizer()
    if world > 1 and not rowshard:
        model.enable_data_parallel()
    log(f"[bench] rank {rank}/{world}: model ready")
    # data-parallel: every rank draws its own triples; row-sharded: all ranks work on the same batch
    sampler_dev = PairwiseSamplerV2(ds, batch_size=BATCH, mode="device", device=dev, seed=2022 + (0 if rowshard else rank))
    sampler_host = PairwiseSamplerV2(ds, batch_size=BATCH, mode="compat")
    n_steps = args.warmup + args.steps

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: triples already in HBM ---------------------------------------------
    u, p, n = sampler_dev.sample_epoch_device(n_steps * BATCH)
    batches = [(u[i * BATCH:(i + 1) * BATCH], p[i * BATCH:(i + 1) * BATCH], n[i * BATCH:(i + 1) * BATCH]) for i in range(n_steps)]
    use_graph = bool(args.cuda_graph) and not rowshard
    runner = model.make_graphed_step() if use_graph else None

    def step(b):
        return runner(*b) if runner is not None else model.train_step(*b)

    for b in batches[:args.warmup]:
        step(b)
    sync_all()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    _lib.CALLS["launches"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for b in batches[args.warmup:]:
        loss = step(b)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    log(f"[bench] rank {rank}: device-resident arm {ms / args.steps:.3f} ms/step")
    launches = _lib.CALLS["launches"] if runner is None else runner.launches_per_step * args.steps
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms)
    units = 1 if rowshard else world          # batches processed per step by the whole job
    value = units * BATCH * args.steps / (ms / 1e3)
    final_loss = float(loss)

    # ---- e2e arm: host triples, H2D + loss D2H every step ----------------------------------------
    hu, hp, hn = sampler_host.sample_epoch_host()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a[:n_steps * BATCH])).pin_memory()
    hu, hp, hn = pin(hu), pin(hp), pin(hn)
    du, dp_, dn = (torch.empty(BATCH, dtype=torch.int64, device=dev) for _ in range(3))

    def e2e_step(i):
        sl = slice(i * BATCH, (i + 1) * BATCH)
        du.copy_(hu[sl], non_blocking=True)
        dp_.copy_(hp[sl], non_blocking=True)
        dn.copy_(hn[sl], non_blocking=True)
        return float(step((du, dp_, dn)))  # .item(): the D2H read main.py:102 does

    for i in range(args.warmup):
        e2e_step(i)
    sync_all()
    t0 = time.perf_counter()
    for i in range(args.warmup, n_steps):
        e2e_step(i)
    sync_all()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = units * BATCH * args.steps / float(te)
    log(f"[bench] rank {rank}: e2e arm done")

    # ---- per-kernel CUDA-event profile + roofline of the wide SpMM -------------------------------
    kernels = {}
    roofline = None
    if rank == 0 and not rowshard:
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
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, which = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
        Fw = 64 * (1 + len(model.mods))
        g = model.graph
        # dominant kernel: spmm_seg_kernel<Fw> (whole rows).  Timed on the launches that are exactly ONE such kernel - the
        # user-row half has no split rows, so its event pair brackets a single launch on the launching stream.
        alg_of = lambda h: float(h.nnz * 8 + (h.n_rows + 1) * 4 + (h.n_cols + h.n_rows) * Fw * 4)
        half = g.ui if g.ui.n_heavy_seg == 0 else (g.iu if g.iu.n_heavy_seg == 0 else None)
        tag = f"spmm{Fw}w"
        if tag in agg and half is not None:
            alg, what = alg_of(half), f"spmm_seg_kernel<{Fw}> (whole rows, {'user' if half is g.ui else 'item'}-row half)"
        elif f"spmm{Fw}" in agg:
            # both halves have split rows (kwai / movielens shapes): the unmasked wide SpMM runs once per half and step, each
            # as a whole-row launch + a split-row launch side by side; one event pair brackets the pair of launches
            tag, alg = f"spmm{Fw}", 0.5 * (alg_of(g.ui) + alg_of(g.iu))
            what = f"spmm_seg_kernel<{Fw}> (whole-row + split-row launch of one half, mean over the two halves)"
        else:
            tag = None
        if tag is not None:
            # algorithmic bytes per launch: nnz*(4+4) + (rows_out+1)*4 + (rows_in + rows_out)*F*4
            avg_s = 1e-3 * sum(agg[tag]) / len(agg[tag])
            ach = alg / avg_s / 1e9
            traffic = None
            try:   # DRAM bytes per wide launch from the committed ncu --set full capture (Tiktok shape only)
                if args.workload == "tiktok":
                    traffic = json.load(open(os.path.join(ROOT, "profiles", "spmm_traffic.json")))["avg_wide_launch_bytes"]
            except Exception:
                pass
            fam = sum(sum(v) for k, v in agg.items() if k.startswith("spmm"))
            tot = sum(sum(v) for v in agg.values())
            roofline = {"kernel": what, "bound": "hbm", "achieved": ach, "peak": peak,
                        "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": which, "bytes_per_launch": alg,
                        "avg_launch_us": 1e6 * avg_s, "launches_timed": len(agg[tag]),
                        "share_of_step": sum(sum(v) for k, v in agg.items() if k.startswith(f"spmm{Fw}")) / tot,
                        "spmm_family_share_of_step": fam / tot,
                        "note": "operand slab is L2-resident: the kernel is bound by L2->SM gather bandwidth (measured ~10-11 TB/s "
                                "of a ~12.4 TB/s L2 slice cap), not by HBM; see DESIGN.md 3.1"}
    clk = clocks.stop() if rank == 0 else None
    sync_all()
    if world > 1:   # the profile pass above stepped rank 0 only: put every replica back on the same weights
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
        evalr.evaluate(model, **kw)  # warm-up (device CSR upload, smem opt-in)
        sync_all()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        res, buf = evalr.evaluate(model, **kw)   # returns host ndarray: the D2H of the result is inside
        b_.record()
        sync_all()
        wall = time.perf_counter() - t0
        tm = torch.tensor([a.elapsed_time(b_) / 1e3, wall], device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ev = {"value": len(users) / float(tm[0]), "unit": "users/s", "n_users": len(users), "e2e_value": len(users) / float(tm[1]),
              "topk": TOPK, "predict_type": "TIE", "result": [float(x) for x in res]}

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
        dense = {"ms_per_step": dms, "value": BATCH / (dms / 1e3), "unit": UNIT, "final_loss": float(ld),
                 "note": "lazy_tables=False: all rows of the last layer, fusion + heads over all U+I rows each step"}
        del md, rd

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": ("strong" if rowshard else "weak"), "vs_baseline": None,
                "dtype": "f32 (tf32 tensor-core projections, 3xTF32 fusion/heads)", "data": "synthetic",
                "config": {"workload": f"{args.workload}-shape EliMRec train step (sample + fwd + bwd + Adam), batch {BATCH}/GPU, "
                                       f"layer_num 3, recdim 64, U={ds.num_users} I={ds.num_items} E_train={ds.train_matrix.nnz}",
                           "parallelism": (f"rowshard{world} (all-gather per GCN layer)" if rowshard else f"dp{world}") if world > 1 else "single",
                           "l2_policy": "per-step working set (features + propagation slabs, >0.9 GB) exceeds the 126 MB L2; no flush",
                           "cuda_graph": bool(runner is not None), "lazy_tables": bool(args.lazy_tables),
                           "row_sparse_step": ("loss-dead rows of the last two layers and of the fusion/head tables are not computed "
                                               "in the step (identical loss / gradients / parameters); full tables are completed "
                                               "at evaluation" if args.lazy_tables else "off"), "sampler": "device Philox (value) / compat libc stream (e2e)"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 3 * BATCH * 8, "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "eval": ev,
                "kernels": kernels, "final_loss": final_loss, "dense_schedule": dense}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
