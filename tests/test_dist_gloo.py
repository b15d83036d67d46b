"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: user sharding of the evaluator and the
data-parallel gradient bucket."""
import os

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
