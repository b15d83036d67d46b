"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box): data-parallel replicas over NCCL + user-sharded evaluation."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


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
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok1, worst, worst2, ok3 in res:
        assert ok1 and ok3, res
        assert worst < 1e-4 and worst2 < 1e-4, res
