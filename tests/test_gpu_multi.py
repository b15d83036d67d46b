"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box): data-parallel replicas over NCCL + user-sharded evaluation."""
import os
import queue as _queue
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _collect(procs, q, timeout=240):
    """results of all workers, failing FAST if one of them dies"""
    res, t0 = [], time.time()
    while len(res) < len(procs):
        try:
            res.append(q.get(timeout=2))
        except _queue.Empty:
            dead = [p for p in procs if not p.is_alive() and p.exitcode not in (0, None)]
            if dead or time.time() - t0 > timeout:
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                raise AssertionError(f"worker failed (exit codes {[p.exitcode for p in procs]})")
    for p in procs:
        p.join(timeout=60)
    return res


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from conftest import load_golden
    from gpu_util import build_model, rel_err
    from helpers import golden_dataset, golden_params
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        g = load_golden("generic")
        g["_name"] = "generic"
        ds = golden_dataset(g)
        batch = lambda i: (g[f"batch{i}_users"], g[f"batch{i}_pos"], g[f"batch{i}_neg"])
        # (1) identical batches on every rank: averaged gradient == single-GPU gradient -> golden trajectory
        model = build_model(ds, golden_params(g), alpha=0.5, test_batch_size=16, device=dev)
        model.make_optimizer(lr=1e-3, weight_decay=1e-4)
        model.enable_data_parallel()
        losses = [float(model.train_step(*batch(i))) for i in range(3)]
        ok1 = bool(np.allclose(losses, g["losses"], rtol=2e-5))
        worst = max(rel_err(v, g["sd3/" + k]) for k, v in model.state_dict().items())
        # (2) different batches per rank == one replica stepping on the MEAN of the two gradients
        model = build_model(ds, golden_params(g), alpha=0.5, device=dev)
        model.make_optimizer(lr=1e-3, weight_decay=1e-4)
        model.enable_data_parallel()
        model.train_step(*batch(rank))
        ref = build_model(ds, golden_params(g), alpha=0.5, device=dev)
        opt = ref.make_optimizer(lr=1e-3, weight_decay=1e-4)
        for i in range(world):
            (ref.bpr_loss(*[torch.tensor(x) for x in batch(i)]) / world).backward()
        opt.step()
        worst2 = max(rel_err(a, b) for a, b in zip(model.state_dict().values(), ref.state_dict().values()))
        # (3) user-sharded evaluation == golden
        model = build_model(ds, golden_params(g), alpha=0.5, test_batch_size=16, device=dev)
        model.bpr_loss(*[torch.tensor(x) for x in batch(0)])
        model.eval()
        res, _ = model.evaluate()
        ok3 = bool(np.abs(res - g["evaluate_TIE"]).max() < 5e-5)
        q.put((rank, ok1, worst, worst2, ok3))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_and_sharded_eval_nccl():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = _collect(procs, q)
    for rank, ok1, worst, worst2, ok3 in res:
        assert ok1 and ok3, res
        assert worst < 1e-4 and worst2 < 1e-4, res


def _sharded_worker(rank, world, port, q):
    import torch.distributed as dist
    from conftest import load_golden
    from gpu_util import rel_err
    from helpers import golden_dataset, golden_params
    from elimrec_b200.data import Config
    from elimrec_b200.sharded import ShardedEliMRec
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        out = []
        for gname, dsname in (("generic", "synthg"), ("kwai", "kwai")):
            g = load_golden(gname)
            g["_name"] = gname
            ds = golden_dataset(g)
            conf = Config(**{"data.input.dataset": dsname, "topks": [20], "device": dev, "alpha": 0.5, "test_batch_size": 16,
                             "proj_precision": "fp32"})
            model = ShardedEliMRec(conf, ds).to(dev)
            model.load_state_dict({k: torch.as_tensor(v) for k, v in golden_params(g).items()}, strict=False)
            model.make_optimizer(lr=1e-3, weight_decay=1e-4)
            batch = lambda i: (g[f"batch{i}_users"], g[f"batch{i}_pos"], g[f"batch{i}_neg"])
            # evaluation after the first forward (tables cached from initial weights), users sharded over ranks
            l0 = float(model._forward(*model._triples(*batch(0))))
            model.eval()
            res, _ = model.evaluate()
            ok_eval = bool(np.abs(res - g["evaluate_TIE"]).max() < 5e-5)
            losses = [float(model.train_step(*batch(i))) for i in range(3)]
            ok_loss = bool(np.allclose(losses, g["losses"], rtol=2e-5)) and abs(l0 - float(g["loss0"])) < 1e-5
            sd = model.state_dict()          # all-gathers the owned rows
            worst = max(rel_err(v, g["sd3/" + k]) for k, v in sd.items())
            out.append((gname, ok_eval, ok_loss, worst))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_row_sharded_allgather_training_nccl():
    """North-star piece 5: users/items partitioned over 2 ranks, all-gather per GCN layer; must reproduce the
    single-GPU (= reference) losses, parameters and metrics."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = _collect(procs, q)
    for rank, out in res:
        for gname, ok_eval, ok_loss, worst in out:
            assert ok_eval and ok_loss and worst < 1e-4, (rank, gname, ok_eval, ok_loss, worst)


def _colshard_worker(rank, world, port, q):
    import torch.distributed as dist
    from conftest import load_golden
    from gpu_util import rel_err
    from helpers import golden_dataset, golden_params
    from elimrec_b200.colshard import ColShardedEliMRec
    from elimrec_b200.data import Config
    from elimrec_b200.model import EliMRec
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    comms = []
    try:
        out = []
        for gname, dsname in (("generic", "synthg"), ("kwai", "kwai")):
            g = load_golden(gname)
            g["_name"] = gname
            ds = golden_dataset(g)
            cfg = lambda: Config(**{"data.input.dataset": dsname, "topks": [20], "device": dev, "alpha": 0.5, "test_batch_size": 16})
            load = lambda m: m.load_state_dict({k: torch.as_tensor(v) for k, v in golden_params(g).items()}, strict=False)
            batch = lambda i: (g[f"batch{i}_users"], g[f"batch{i}_pos"], g[f"batch{i}_neg"])
            # (1) the same batch on both ranks == the single-GPU (= reference) trajectory; through the CUDA-graph runner,
            #     i.e. with the NCCL collectives of the C-ABI comm family captured inside the graph
            m = ColShardedEliMRec(cfg(), ds).to(dev)
            load(m)
            m.make_optimizer(lr=1e-3, weight_decay=1e-4)
            run = m.make_graphed_step(len(batch(0)[0]))
            losses = [float(run(*batch(i))) for i in range(3)]
            ok_loss = bool(np.allclose(losses, g["losses"], rtol=2e-5))
            worst = max(rel_err(v, g["sd3/" + k]) for k, v in m.state_dict().items())
            comms.append(m.comm)
            # (2) a different batch per rank == one GPU stepping on the mean gradient; tables + sharded evaluation of that forward
            m = ColShardedEliMRec(cfg(), ds).to(dev)
            load(m)
            m.make_optimizer(lr=1e-3, weight_decay=1e-4)
            m.train_step(*batch(rank))
            ref = EliMRec(cfg(), ds).to(dev)
            load(ref)
            opt = ref.make_optimizer(lr=1e-3, weight_decay=1e-4)
            for i in range(world):
                (ref.bpr_loss(*[torch.tensor(x) for x in batch(i)]) / world).backward()
            opt.step()
            worst2 = max(rel_err(a, b) for a, b in zip(m.state_dict().values(), ref.state_dict().values()))
            m.eval()
            res, _ = m.evaluate()
            ref0 = EliMRec(cfg(), ds).to(dev)
            load(ref0)
            ref0.bpr_loss(*[torch.tensor(x) for x in batch(rank)])
            ok_tab = rel_err(m.all_items, ref0.all_items) < 2e-5 and rel_err(m.all_users, ref0.all_users) < 2e-5
            out.append((gname, ok_loss, worst, worst2, bool(ok_tab), bool(np.isfinite(res).all())))
            comms.append(m.comm)
        q.put((rank, out))
    finally:
        # orderly teardown: the graphs that captured collectives, then the library's communicators, then torch's
        run = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        for c in comms:
            c.close()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_column_sharded_training_nccl():
    """column-sharded mode on 2 GPUs (32 embedding columns each), NCCL through the C-ABI comm family, step captured as one
    CUDA graph: golden trajectory, mean-of-batches step, table completion through the all-gather."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 150)
    procs = [ctx.Process(target=_colshard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = _collect(procs, q)
    for rank, out in res:
        for gname, ok_loss, worst, worst2, ok_tab, fin in out:
            assert ok_loss and ok_tab and fin and worst < 1e-4 and worst2 < 1e-4, (rank, out)
