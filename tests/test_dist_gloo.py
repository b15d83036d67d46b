"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: user sharding of the evaluator and the
data-parallel gradient bucket."""
import os
import queue as _queue
import sys
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from elimrec_b200.dist import GradBucket, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 128, 36515):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert all(0 <= lo <= hi <= n for lo, hi in parts)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) gradient bucket: mean over ranks, strided inputs, views keep the parameter shapes
        # head = packed (strided table gradient), tail = an existing flat buffer the backward writes in place
        tail = torch.zeros(64 * 16 + 64)
        tail_views = {"v_dense.weight": tail[:1024].view(64, 16), "v_dense.bias": tail[1024:]}
        b = GradBucket({"embedding_user.weight": (50, 64)}, "cpu", tail_flat=tail, tail_views=tail_views)
        g = torch.Generator().manual_seed(rank)
        wide = torch.randn(50, 256, generator=g)
        tail_views["v_dense.weight"].copy_(torch.randn(64, 16, generator=g))
        tail_views["v_dense.bias"].fill_(float(rank + 1))
        grads = {"embedding_user.weight": wide[:, 64:128], **tail_views}
        b.pack(grads)
        assert b.names == ["embedding_user.weight", "v_dense.weight", "v_dense.bias"]
        w = b.all_reduce_mean_part(0, async_op=True)
        assert w is None        # gloo: synchronous
        b.all_reduce_mean_part(1)
        out = b.views
        exp_bias = sum(range(1, world + 1)) / world
        ok1 = bool(torch.allclose(out["v_dense.bias"], torch.full((64,), exp_bias)))
        mine = grads["embedding_user.weight"].clone()
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        ok2 = bool(torch.allclose(out["embedding_user.weight"], torch.stack(gathered).mean(0), atol=1e-7))
        # (2) evaluator sharding: per-shard metric SUMS all-reduced == sums over all users
        from oracle import ref_eval
        rng = np.random.default_rng(0)
        n_users, n_items, K = 37, 200, 20
        scores = rng.standard_normal((n_users, n_items)).astype(np.float32)
        truth = [np.sort(rng.choice(n_items, size=rng.integers(1, 9), replace=False)).tolist() for _ in range(n_users)]
        rows_all = ref_eval.metric_rows(ref_eval.topk_lowest_index(scores, K), truth, [1, 2, 4], K)
        lo, hi = shard_range(n_users, rank, world)
        rows = ref_eval.metric_rows(ref_eval.topk_lowest_index(scores[lo:hi], K), truth[lo:hi], [1, 2, 4], K)
        sums = torch.from_numpy(rows.astype(np.float64).sum(0))
        dist.all_reduce(sums)
        ok3 = bool(np.allclose(sums.numpy() / n_users, rows_all.astype(np.float64).mean(0), rtol=1e-12))
        q.put((rank, ok1, ok2, ok3))
    finally:
        dist.destroy_process_group()


def test_bucket_and_eval_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(all(r[1:]) for r in res), res


# ---- the multi-GPU training / evaluation modes end to end on 2 gloo ranks, kernels stood in for by tests/sim_ops.py ----------
def _install_sim():
    import sim_ops
    import elimrec_b200.evaluator as ev
    import elimrec_b200.linear as ln
    import elimrec_b200.model as md
    import elimrec_b200.optim as op
    import elimrec_b200.colshard as csm
    import elimrec_b200.sharded as sh

    class _Ev:
        def __init__(self, *a, **k):
            pass

        def record(self, *a):
            pass

    class _St:
        def wait_event(self, *a):
            pass
    for mod in (md, ev, op, sh, ln, csm):
        mod.ops = sim_ops
    md._require_cuda = lambda dev: None
    torch.cuda.Event = _Ev
    torch.cuda.current_stream = lambda *a: _St()


def _model_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from conftest import load_golden
        from helpers import golden_dataset, golden_params
        from elimrec_b200.data import Config
        from elimrec_b200.model import EliMRec
        from elimrec_b200.sharded import ShardedEliMRec
        _install_sim()
        g = load_golden("generic")
        g["_name"] = "generic"
        ds = golden_dataset(g)
        cfg = lambda **kw: Config(**{"data.input.dataset": "synthg", "topks": [20], "device": torch.device("cpu"), "alpha": 0.5,
                                     "test_batch_size": 16, "rank_backend": "fp32", "proj_precision": "fp32", **kw})
        load = lambda m: m.load_state_dict({k: torch.as_tensor(v) for k, v in golden_params(g).items()}, strict=False)
        batch = lambda i: tuple(torch.as_tensor(g[f"batch{i}_{k}"]) for k in ("users", "pos", "neg"))
        rel = lambda a, b: float((a.double() - torch.as_tensor(b).double()).abs().max() / torch.as_tensor(b).double().abs().max())
        # (1) data-parallel replicas, a different batch per rank == one replica stepping on the MEAN gradient
        model = EliMRec(cfg(), ds)
        load(model)
        model.make_optimizer(lr=1e-3, weight_decay=1e-4)
        model.enable_data_parallel()
        model.train_step(*batch(rank))
        ref = EliMRec(cfg(), ds)
        load(ref)
        opt = ref.make_optimizer(lr=1e-3, weight_decay=1e-4)
        for i in range(world):
            (ref.bpr_loss(*batch(i)) / world).backward()
        opt.step()
        worst_dp = max(rel(a, b) for a, b in zip(model.state_dict().values(), ref.state_dict().values()))
        # (2) user-sharded evaluation: all-reduced metric sums == the reference's numbers
        model = EliMRec(cfg(), ds)
        load(model)
        model.bpr_loss(*batch(0))
        model.eval()
        ok_eval = bool(np.abs(model.evaluate()[0] - g["evaluate_TIE"]).max() < 5e-5)
        # (3) row-sharded propagation (all-gather per layer): golden losses / parameters / metrics
        sm = ShardedEliMRec(cfg(), ds)
        load(sm)
        sm.make_optimizer(lr=1e-3, weight_decay=1e-4)
        l0 = float(sm._forward(*sm._triples(*batch(0))))
        sm.eval()
        ok_seval = bool(np.abs(sm.evaluate()[0] - g["evaluate_TIE"]).max() < 5e-5)
        losses = [float(sm.train_step(*batch(i))) for i in range(3)]
        ok_loss = bool(np.allclose(losses, g["losses"], rtol=2e-5)) and abs(l0 - float(g["loss0"])) < 1e-5
        worst_sh = max(rel(v, g["sd3/" + k]) for k, v in sm.state_dict().items())
        q.put((rank, worst_dp, ok_eval, ok_seval, ok_loss, worst_sh))
    finally:
        dist.destroy_process_group()


def test_data_parallel_and_row_sharded_models_world2():
    """N>1 host logic without a GPU: data-parallel replicas (gradient bucket, split all-reduce), user-sharded evaluation and
    the row-sharded model (per-layer all-gathers, instance-row all-reduce, sharded Adam, state_dict gather) on 2 gloo ranks."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29100 + (os.getpid() % 300)
    procs = [ctx.Process(target=_model_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, t0 = [], time.time()
    while len(res) < len(procs):        # fail fast if a worker dies
        try:
            res.append(q.get(timeout=2))
        except _queue.Empty:
            if any(not p.is_alive() and p.exitcode not in (0, None) for p in procs) or time.time() - t0 > 400:
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                raise AssertionError(f"worker failed (exit codes {[p.exitcode for p in procs]})")
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, worst_dp, ok_eval, ok_seval, ok_loss, worst_sh in res:
        assert worst_dp < 1e-4 and ok_eval and ok_seval and ok_loss and worst_sh < 1e-4, res


def _colshard_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from conftest import load_golden
        from helpers import golden_dataset, golden_params
        from elimrec_b200.colshard import ColShardedEliMRec
        from elimrec_b200.data import Config
        from elimrec_b200.model import EliMRec
        _install_sim()
        out = []
        for gname, dsname in (("generic", "synthg"), ("kwai", "kwai")):
            g = load_golden(gname)
            g["_name"] = gname
            ds = golden_dataset(g)
            cfg = lambda **kw: Config(**{"data.input.dataset": dsname, "topks": [20], "device": torch.device("cpu"), "alpha": 0.5,
                                         "test_batch_size": 16, "rank_backend": "fp32", "proj_precision": "fp32", **kw})
            load = lambda m: m.load_state_dict({k: torch.as_tensor(v) for k, v in golden_params(g).items()}, strict=False)
            batch = lambda i: tuple(torch.as_tensor(g[f"batch{i}_{k}"]) for k in ("users", "pos", "neg"))
            rel = lambda a, b: float((a.double() - torch.as_tensor(b).double()).abs().max() / torch.as_tensor(b).double().abs().max())
            # (1) the same batch on every rank: mean loss / mean gradient == the single-process step -> golden trajectory
            m = ColShardedEliMRec(cfg(), ds)
            load(m)
            m.make_optimizer(lr=1e-3, weight_decay=1e-4)
            losses = [float(m.train_step(*batch(i))) for i in range(3)]
            ok_loss = bool(np.allclose(losses, g["losses"], rtol=2e-5))
            m.eval()
            ok_eval = True
            worst = max(rel(v, g["sd3/" + k]) for k, v in m.state_dict().items())      # gathers the column shards
            # (2) a different batch per rank == one process stepping on the MEAN of the per-batch gradients; evaluation after it
            m = ColShardedEliMRec(cfg(), ds)
            load(m)
            m.make_optimizer(lr=1e-3, weight_decay=1e-4)
            lm = float(m.train_step(*batch(rank)))
            ref = EliMRec(cfg(), ds)
            load(ref)
            opt = ref.make_optimizer(lr=1e-3, weight_decay=1e-4)
            lr_ = 0.0
            for i in range(world):
                li = ref.bpr_loss(*batch(i)) / world
                li.backward()
                lr_ += float(li)
            opt.step()
            ok_mean = abs(lm - lr_) < 2e-5 * abs(lr_)
            worst2 = max(rel(a, b) for a, b in zip(m.state_dict().values(), ref.state_dict().values()))
            # tables of that forward (pre-update weights) through the all-gather of the column shards: user-sharded evaluation
            m.eval()
            ref0 = EliMRec(cfg(), ds)
            load(ref0)
            ref0.bpr_loss(*batch(rank))
            ok_tab = rel(m.all_users, ref0.all_users.detach().numpy()) < 2e-5 and rel(m.all_items, ref0.all_items.detach().numpy()) < 2e-5
            ev = m.evaluate()[0]
            out.append((gname, ok_loss, ok_eval, worst, ok_mean, worst2, ok_tab, bool(np.isfinite(ev).all())))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_column_sharded_model_world2():
    """column-sharded multi-GPU mode on 2 gloo ranks (32 columns each): all-gather of the triples, the two all-to-alls of the
    instance rows, the all-reduce of the small gradients, column-local fused Adam, state_dict / table gathers."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + (os.getpid() % 300)
    procs = [ctx.Process(target=_colshard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, t0 = [], time.time()
    while len(res) < len(procs):
        try:
            res.append(q.get(timeout=2))
        except _queue.Empty:
            if any(not p.is_alive() and p.exitcode not in (0, None) for p in procs) or time.time() - t0 > 400:
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                raise AssertionError(f"worker failed (exit codes {[p.exitcode for p in procs]})")
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in res:
        for gname, ok_loss, ok_eval, worst, ok_mean, worst2, ok_tab, fin in out:
            assert ok_loss and ok_eval and ok_mean and ok_tab and fin and worst < 1e-4 and worst2 < 1e-4, (rank, out)
