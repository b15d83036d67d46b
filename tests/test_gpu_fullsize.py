"""-m gpu: BASELINE.json's FULL sizes through size-independent properties (the oracle would take minutes there).

Tiktok-shape (36,656 x 76,085, ~614K train edges, 128/128/768-d), Kwai-shape (v-only, 2048-d, F=128) and
Movielens-shape (2048/128/100-d: a feature width that is not a multiple of 32)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import build_model, rel_err


@pytest.fixture(scope="module")
def tiktok():
    from elimrec_b200 import synth
    from elimrec_b200.data import Dataset
    inter, feats = synth.make_shape("tiktok")
    ds = Dataset(None, interactions=inter, features=feats, name="tiktokshape")
    torch.manual_seed(2022)
    return ds, build_model(ds, None, dataset_name="tiktokshape", alpha=0.5, proj_precision="tf32")


def _batch(ds, B, seed):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, ds.num_users, B)
    tm = ds.train_matrix
    p = tm.indices[tm.indptr[u]]
    n = rng.integers(0, ds.num_items, B)
    return u, p, n


def test_spmm_fullsize_vs_cusparse_and_adjoint(tiktok):
    """Y = A X against torch.sparse.mm (cuSPARSE) on the same device, and <A_ui x, y> == <x, A_iu y> (A symmetric)."""
    from elimrec_b200 import ops
    ds, model = tiktok
    g, dev = model.graph, model.device_
    for half, other in ((g.ui, g.iu), (g.iu, g.ui)):
        for width in (64, 256):
            X = torch.randn(half.n_cols, width, device=dev)
            Y = torch.empty(half.n_rows, width, device=dev)
            ops.spmm(half, X, Y, width)
            crow = half.indptr
            A = torch.sparse_csr_tensor(crow, half.col.long(), half.val, size=(half.n_rows, half.n_cols))
            ref = torch.sparse.mm(A, X)
            assert rel_err(Y, ref) < 1e-5
            Yv = torch.randn(half.n_rows, width, device=dev)
            Z = torch.empty(half.n_cols, width, device=dev)
            ops.spmm(other, Yv, Z, width)
            lhs, rhs = float((Y.double() * Yv.double()).sum()), float((X.double() * Z.double()).sum())
            scale = float((Y.double() * Yv.double()).abs().sum())   # the two sums cancel heavily: compare to sum |terms|
            assert abs(lhs - rhs) < 1e-6 * scale
    # linearity
    X1, X2 = torch.randn(g.ui.n_cols, 256, device=dev), torch.randn(g.ui.n_cols, 256, device=dev)
    Y1, Y2, Y3 = (torch.empty(g.ui.n_rows, 256, device=dev) for _ in range(3))
    ops.spmm(g.ui, X1, Y1, 256); ops.spmm(g.ui, X2, Y2, 256); ops.spmm(g.ui, 2 * X1 - 3 * X2, Y3, 256)
    assert rel_err(Y3, 2 * Y1 - 3 * Y2) < 1e-5


def test_training_fullsize_descends_and_is_reproducible(tiktok):
    ds, model = tiktok
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model.make_optimizer(lr=1e-3, weight_decay=1e-4)
    b = [_batch(ds, 2048, s) for s in range(6)]
    l1 = [float(model.train_step(*x)) for x in b]
    assert all(np.isfinite(l1)) and abs(l1[0] - 4 * np.log(2) * 0 - l1[0]) == 0
    # same weights, same batches -> identical losses bit for bit (no float atomics on the reduction paths that matter
    # at this batch size would change the first step; later steps may differ in the last bits through the scatter-adds)
    model.load_state_dict(sd)
    model.make_optimizer(lr=1e-3, weight_decay=1e-4)
    l2 = [float(model.train_step(*x)) for x in b]
    assert l1[0] == l2[0] and np.allclose(l1, l2, rtol=1e-5)
    # the CUDA-graph runner replays the same step
    model.load_state_dict(sd)
    model.make_optimizer(lr=1e-3, weight_decay=1e-4)
    run = model.make_graphed_step(2048)          # consumes one warm-up step on a dummy batch of zeros
    model.load_state_dict(sd)
    model._adam.step_dev.zero_(); [t.zero_() for st in model._adam.state.values() for t in st]
    l3 = [float(run(*x)) for x in b]
    assert np.allclose(l1, l3, rtol=1e-5)
    # a few hundred steps on repeated batches must drive the loss down
    for _ in range(60):
        for x in b[:2]:
            last = float(run(*x))
    assert last < 0.9 * l1[0]


def test_gradient_directional_derivative_fullsize(tiktok):
    """<grad, d> against a central finite difference of the loss along a random direction d (fp32-exact path)."""
    ds, _ = tiktok
    torch.manual_seed(1)
    model = build_model(ds, None, dataset_name="tiktokshape", alpha=0.5, proj_precision="fp32", fuse_precision="fp32")
    u, p, n = _batch(ds, 2048, 9)
    loss = model.bpr_loss(u, p, n)
    loss.backward()
    params = [q for q in model.parameters() if q.grad is not None]
    torch.manual_seed(2)
    dirs = [torch.randn_like(q) * q.abs().mean() for q in params]
    analytic = sum(float((q.grad.double() * d.double()).sum()) for q, d in zip(params, dirs))
    eps = 0.05
    with torch.no_grad():
        for q, d in zip(params, dirs):
            q.add_(eps * d)
        lp = float(model.bpr_loss(u, p, n))
        for q, d in zip(params, dirs):
            q.sub_(2 * eps * d)
        lm = float(model.bpr_loss(u, p, n))
    fd = (lp - lm) / (2 * eps)
    assert abs(fd - analytic) < 2e-2 * abs(analytic) + 1e-6, (fd, analytic)


def test_eval_fullsize_properties(tiktok):
    ds, model = tiktok
    model.train_step(*_batch(ds, 2048, 3)) if model._adam else model.bpr_loss(*_batch(ds, 2048, 3))
    model.eval()
    model.predict_type = "TIE"
    ev = model.test_evaluator.evaluator
    res1, buf1, rows = ev.evaluate(model, return_rows=True)
    idx, val = ev.last_topk
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    users = list(ds.get_user_test_dict().keys())
    assert idx.shape == (len(users), 20) and (idx >= 0).all() and (idx < ds.num_items).all()
    assert (np.diff(val, axis=1) <= 0).all()                      # sorted by score, descending
    assert all(len(set(r)) == 20 for r in idx[:2000])              # no duplicates
    tm = ds.train_matrix
    sel = np.arange(0, len(users), 37)
    assert not np.asarray(tm[np.repeat(np.array(users)[sel], 20), idx[sel].ravel()]).any()   # train items never ranked
    res2, buf2 = ev.evaluate(model)
    assert buf1 == buf2                                            # idempotent / deterministic
    # scores of the ranked items agree with predict() (the un-fused path) and dominate every unranked, unmasked item
    some = [users[i] for i in (0, 5, 77)]
    sc = model.predict(some).numpy()
    for r, uix in zip((0, 5, 77), range(3)):
        np.testing.assert_allclose(sc[uix, idx[r]], val[r], rtol=1e-5)
        masked = sc[uix].copy()
        masked[tm.indices[tm.indptr[users[r]]:tm.indptr[users[r] + 1]]] = -np.inf
        assert masked.max() <= val[r, 0] * (1 + 1e-6) and np.sort(masked)[-20] >= val[r, -1] * (1 - 1e-6)
    # metric means equal the mean of the per-user rows; recall/precision consistency: P@K * K = hits = R@K * |truth|
    r = rows.cpu().numpy()
    np.testing.assert_allclose(r.astype(np.float64).mean(0).reshape(3, 20)[:, 19], res1, rtol=1e-6)   # device mean is fp64
    td = ds.get_user_test_dict()
    tl = np.array([len(td[x]) for x in users])
    np.testing.assert_allclose(r[:, 19] * 20, r[:, 39] * tl, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shape", ["kwai", "movielens"])
def test_other_shapes_step_and_eval(shape):
    """Kwai-shape: v-only model (F=128, D=2048).  Movielens-shape: 2048/128/100-d features, eval-heavy (55K users)."""
    from elimrec_b200 import synth
    from elimrec_b200.data import Dataset
    inter, feats = synth.make_shape(shape)
    ds = Dataset(None, interactions=inter, features=feats, name=shape)
    torch.manual_seed(0)
    model = build_model(ds, None, dataset_name=shape, alpha=0.5, proj_precision="tf32")
    model.make_optimizer()
    ls = [float(model.train_step(*_batch(ds, 2048, s))) for s in range(3)]
    assert np.isfinite(ls).all()
    # tensor-core projections vs the exact path on the same weights: TF32 class
    exact = build_model(ds, None, dataset_name=shape, alpha=0.5, proj_precision="fp32", fuse_precision="fp32")
    exact.load_state_dict(model.state_dict())
    b = _batch(ds, 2048, 11)
    la, lb = float(model.bpr_loss(*b)), float(exact.bpr_loss(*b))
    assert abs(la - lb) < 1e-3 * abs(lb)
    assert rel_err(model.all_items, exact.all_items) < 1e-3
    model.eval()
    res, buf = model.test()
    assert np.isfinite(res).all() and 0 <= res[1] <= 1
