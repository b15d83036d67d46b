#!/usr/bin/env python
"""Micro-benchmark of the 64-wide propagation layer (both CSR halves): spmm_seg_kernel<64> (one warp per row, two
launches + split rows) against elimrec_spmm64_pair (8-lane groups, one launch) at its tuning points, dense and masked.
Checked against torch.sparse (cuSPARSE) on the same device.  Usage (GPU box): python tools/spmm64_bench.py [workload]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from elimrec_b200 import ops  # noqa: E402
from elimrec_b200.graph import BipartiteGraph  # noqa: E402


def timeit(fn, reps=40):
    """us per call, measured on a CUDA graph of 10 back-to-back calls (no host launch latency, streams as in the step)"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps // 10):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / (10 * (reps // 10))


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "tiktok"
    dev = torch.device("cuda:0")
    ds, _ = bench.build_dataset(workload)
    g = BipartiteGraph(ds.train_matrix, dev)
    U, I = ds.num_users, ds.num_items
    torch.manual_seed(0)
    X = torch.randn(U + I, 64, device=dev)
    Y0, Y1 = torch.zeros(U + I, 64, device=dev), torch.zeros(U + I, 64, device=dev)
    A_ui = torch.sparse_csr_tensor(g.ui.indptr, g.ui.col.long(), g.ui.val, size=(U, I))
    A_iu = torch.sparse_csr_tensor(g.iu.indptr, g.iu.col.long(), g.iu.val, size=(I, U))
    ref = torch.cat([torch.sparse.mm(A_ui, X[U:]), torch.sparse.mm(A_iu, X[:U])])
    alg = g.nnz * 8 + (U + I + 2) * 4 + 2 * (U + I) * 64 * 4
    out = {"workload": workload, "U": U, "I": I, "nnz": g.nnz, "split_items": [g.ui.n_split64, g.iu.n_split64],
           "items": [g.ui.n_item64, g.iu.n_item64],
           "algorithmic_bytes_per_layer": alg, "gathered_bytes_per_layer": g.nnz * 256}

    def old():
        s = ops.fork_side(3)
        with torch.cuda.stream(s):
            ops.spmm(g.iu, X[:U], Y0[U:], 64)
        ops.spmm(g.ui, X[U:], Y0[:U], 64)
        ops.join_side(s)

    t = timeit(old)
    err = float((Y0 - ref).abs().max() / ref.abs().max())
    out["spmm_seg_kernel<64> x2 (+ split rows)"] = {"us": round(t, 2), "alg_GBps": round(alg / t / 1e3, 1), "err": err}
    for v in range(5):
        fn = lambda: ops.spmm64_pair(g.ui, g.iu, X[U:], X[:U], Y1[:U], Y1[U:], variant=v)
        Y1.zero_()
        t = timeit(fn)
        err = float((Y1 - ref).abs().max() / ref.abs().max())
        out[f"spmm64_pair v{v}"] = {"us": round(t, 2), "alg_GBps": round(alg / t / 1e3, 1), "gather_TBps": round(g.nnz * 256 / t / 1e6, 2),
                                    "err": err}
    # masks: rows at ~50 % (the layer L-1 of a step), columns at ~5 % (first backward hop)
    rm = (torch.rand(U + I, device=dev) < 0.5).to(torch.uint8)
    cm = (torch.rand(U + I, device=dev) < 0.05).to(torch.uint8)
    for name, kw_old, kw_new in (("row mask 50%", dict(row=rm), dict(row_mask_u=rm[:U], row_mask_i=rm[U:])),
                                 ("col mask 5%", dict(col=cm), dict(col_mask_u=cm[U:], col_mask_i=cm[:U]))):
        def old_m():
            s = ops.fork_side(3)
            with torch.cuda.stream(s):
                ops.spmm(g.iu, X[:U], Y0[U:], 64, row_mask=kw_old["row"][U:] if "row" in kw_old else None,
                         col_mask=kw_old["col"][:U] if "col" in kw_old else None)
            ops.spmm(g.ui, X[U:], Y0[:U], 64, row_mask=kw_old["row"][:U] if "row" in kw_old else None,
                     col_mask=kw_old["col"][U:] if "col" in kw_old else None)
            ops.join_side(s)
        Xm = X if "col" not in kw_old else X * cm.unsqueeze(1)
        refm = torch.cat([torch.sparse.mm(A_ui, Xm[U:]), torch.sparse.mm(A_iu, Xm[:U])])
        sel = rm.bool() if "row" in kw_old else torch.ones(U + I, dtype=torch.bool, device=dev)
        Y0.zero_(); Y1.zero_()
        t0 = timeit(old_m)
        fn = lambda: ops.spmm64_pair(g.ui, g.iu, X[U:], X[:U], Y1[:U], Y1[U:], **kw_new)
        t1 = timeit(fn)
        e0 = float((Y0 - refm)[sel].abs().max() / refm.abs().max())
        e1 = float((Y1 - refm)[sel].abs().max() / refm.abs().max())
        out[name] = {"spmm_seg_kernel us": round(t0, 2), "spmm64_pair us": round(t1, 2), "err": [e0, e1],
                     "untouched_rows_ok": bool((Y1[~sel] == 0).all())}
    # the last backward hop with the fused Adam epilogue, at its tuning points (variant 0 = the shipped one)
    P = torch.randn(U + I, 64, device=dev)
    M, V = torch.zeros_like(P), torch.zeros_like(P)
    E0 = torch.empty_like(P)
    consts = torch.tensor([1e-3, 1.0], dtype=torch.float64, device=dev)
    adam_bytes = alg - 2 * (U + I) * 64 * 4 + 8 * (U + I) * 64 * 4      # source slab + p, m, v read + p, m, v, snapshot written
    for v in range(3):      # 0 = shipped (<UNR 4, 4 CTAs/SM>, rows fetched after the gathers), 1 = prefetched rows, 2 = <UNR 2, 4>
        fn = lambda: ops.spmm64_pair(g.ui, g.iu, X[U:], X[:U], Y1[:U], Y1[U:], variant=v,
                                     adam_u=(P[:U], M[:U], V[:U], E0[:U]), adam_i=(P[U:], M[U:], V[U:], E0[U:]),
                                     adam_consts=(consts, 0.9, 0.999, 1e-8, 0.0))
        t = timeit(fn)
        out[f"fused Adam hop v{v}"] = {"us": round(t, 2), "alg_GBps": round(adam_bytes / t / 1e3, 1)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
