"""-m gpu: every C-ABI kernel family against an independent fp64 / oracle computation."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

pytestmark = pytest.mark.gpu

from gpu_util import FP32_TOL, rel_err


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _rand_graph(U, I, nnz, heavy_rows, seed):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, U, size=nnz)
    i = rng.integers(0, I, size=nnz)
    for r, d in heavy_rows:  # rows much longer than the segment length, incl. > 32 segments
        u = np.concatenate([u, np.full(d, r)])
        i = np.concatenate([i, rng.choice(I, size=d, replace=False)])
    m = sp.csr_matrix((np.ones(u.size), (u, i)), shape=(U, I))
    m.sum_duplicates()
    m.data[:] = 1
    m[U - 1, :] = 0  # an empty row
    m.eliminate_zeros()
    return m


@pytest.mark.parametrize("width", [64, 128, 256])
def test_spmm_matches_fp64(dev, width):
    from elimrec_b200 import ops
    from elimrec_b200.graph import BipartiteGraph
    U, I = 700, 500
    m = _rand_graph(U, I, 6000, [(3, 480), (10, 130), (11, 65)], seed=width)
    g = BipartiteGraph(m, dev, seg_len=8 if width == 64 else 64)
    for half, n_in in ((g.ui, I), (g.iu, U)):
        X = torch.randn(n_in, width, device=dev)
        Y = torch.full((half.n_rows, width), float("nan"), device=dev)
        ops.spmm(half, X, Y, width)
        A = sp.csr_matrix((half.vals_host.astype(np.float64), half.indices_host, half.indptr_host), shape=(half.n_rows, n_in))
        ref = A @ X.double().cpu().numpy()
        assert rel_err(Y, ref) < FP32_TOL
        Y2 = torch.empty_like(Y)
        ops.spmm(half, X, Y2, width)
        assert torch.equal(Y, Y2), "split-row reduction must be deterministic"
        assert int(half.counter.abs().sum()) == 0


@pytest.mark.parametrize("width", [64, 128, 256])
def test_spmm_masked_bit_identical(dev, width):
    """Row-sparse last layer: masked rows / columns give the bits of the dense launch (whole and split rows); width 64 with
    a column mask regroups the two half-warp partial sums, i.e. agrees to rounding."""
    from elimrec_b200 import ops
    from elimrec_b200.graph import BipartiteGraph
    U, I = 700, 500
    g = BipartiteGraph(_rand_graph(U, I, 6000, [(3, 480), (10, 130), (11, 65)], seed=width), dev, seg_len=8 if width == 64 else 64)
    gen = torch.Generator().manual_seed(width)
    for half, n_in in ((g.ui, I), (g.iu, U)):
        X = torch.randn(n_in, width, generator=gen).to(dev)
        Y = torch.empty(half.n_rows, width, device=dev)
        ops.spmm(half, X, Y, width)
        # (a) row mask: marked rows equal the dense result, unmarked rows are untouched
        rm = (torch.rand(half.n_rows, generator=gen) < 0.1).to(torch.uint8)
        rm[[3, 10]] = 1
        rm[11] = 0
        rmd = rm.to(dev)
        for density in (5, 50):   # 256 / 32 segments per CTA in the sparse whole-row kernel
            Yr = torch.full_like(Y, 7.0)
            ops.spmm(half, X, Yr, width, row_mask=rmd, density=density)
            assert torch.equal(Yr[rmd.bool()], Y[rmd.bool()])
            assert float((Yr[~rmd.bool()] - 7.0).abs().max()) == 0
        # (b) column mask: equals the dense launch on X with the unmarked rows zeroed; NaNs in dropped rows are never read
        cm = (torch.rand(n_in, generator=gen) < 0.15).to(torch.uint8).to(dev)
        Xz = X * cm.float()[:, None]
        ops.spmm(half, Xz, Y, width)
        Xn = torch.where(cm.bool()[:, None], X, torch.full_like(X, float("nan")))
        Yc = torch.full_like(Y, 7.0)
        ops.spmm(half, Xn, Yc, width, col_mask=cm)
        if width == 64:   # two edges per warp-iteration: compaction changes which half-warp sums which edge
            assert rel_err(Yc, Y) < 1e-6 and not bool(torch.isnan(Yc).any())
        else:
            assert torch.equal(Yc, Y)
        assert int(half.counter.abs().sum()) == 0
    # instance rows + mask in one launch
    B = 37
    users = torch.randint(0, U, (B,), generator=gen).to(dev); pos = torch.randint(0, I, (B,), generator=gen).to(dev)
    neg = torch.randint(0, I, (B,), generator=gen).to(dev)
    rows = torch.empty(3 * B, dtype=torch.int32, device=dev); mask = torch.full((U + I,), 9, dtype=torch.uint8, device=dev)
    ops.inst_rows(users, pos, neg, U, rows, mask)
    m2 = torch.zeros(U + I, dtype=torch.uint8, device=dev)
    ops.mark_neighbors(g.ui, mask[:U], m2[U:])
    A = sp.csr_matrix((g.ui.vals_host, g.ui.indices_host, g.ui.indptr_host), shape=(U, I))
    ref2 = np.zeros(I, dtype=np.uint8); ref2[A[mask[:U].bool().cpu().numpy()].indices] = 1
    assert np.array_equal(m2[U:].cpu().numpy(), ref2) and int(m2[:U].sum()) == 0
    ref_rows = torch.cat([users, pos + U, neg + U]).int()
    assert torch.equal(rows, ref_rows)
    ref_mask = torch.zeros(U + I, dtype=torch.uint8, device=dev); ref_mask[ref_rows.long()] = 1
    assert torch.equal(mask, ref_mask)
    dst = torch.ones(I, 64, device=dev)
    ops.zero_rows(rows, U, U + I, U, dst, 64)
    ref = torch.ones(I, 64, device=dev); ref[torch.cat([pos, neg])] = 0
    assert torch.equal(dst, ref)


@pytest.mark.parametrize("width,G", [(256, 4), (64, 4), (128, 2), (64, 2)])
def test_spmm_mean_epilogue(dev, width, G):
    from elimrec_b200 import ops
    from elimrec_b200.graph import BipartiteGraph
    U, I, Fw = 300, 400, 64 * G
    g = BipartiteGraph(_rand_graph(U, I, 3000, [(5, 200)], seed=1), dev)
    half = g.ui
    X = torch.randn(I, width, device=dev)
    prev = [(torch.randn(U, 64, device=dev), 64), (torch.randn(U, Fw, device=dev), Fw), (torch.randn(U, 64, device=dev), 64)]
    out = torch.empty(U + 3, Fw, device=dev)[3:]  # offset view: row stride only
    ops.spmm(half, X, None, width, ops.mean_epilogue(prev, out, Fw, 0.25))
    A = sp.csr_matrix((half.vals_host.astype(np.float64), half.indices_host, half.indptr_host), shape=(U, I))
    y = torch.from_numpy(A @ X.double().cpu().numpy())
    if width == 64:
        y = y.repeat(1, G)
    acc = torch.zeros(U, Fw, dtype=torch.float64)
    for t, w in prev:
        acc += t.double().cpu().repeat(1, G) if w == 64 else t.double().cpu()
    ref = (acc + y) * 0.25
    assert rel_err(out, ref) < FP32_TOL


def test_row_ops(dev):
    from elimrec_b200 import ops
    src = torch.randn(50, 256, device=dev)
    rows = torch.tensor([3, 7, 3, 49, 20, 3], dtype=torch.int32, device=dev)
    dst = torch.empty(6, 256, device=dev)
    ops.gather_rows(rows, src, dst, 256)
    assert torch.equal(dst, src[rows.long()])
    acc = torch.zeros(30, 256, device=dev)
    ops.scatter_add_rows(rows, 0, 30, 0, dst, 256, acc, 256, 0.5)
    ref = torch.zeros(30, 256, dtype=torch.float64)
    for r, n in enumerate(rows.tolist()):
        if n < 30:
            ref[n] += 0.5 * dst[r].double().cpu()
    assert rel_err(acc, ref) < 1e-6
    acc = torch.zeros(40, 64, device=dev)
    ops.scatter_add_rows(rows, 10, 50, 10, dst, 256, acc, 64, 1.0)  # folded (sum of the 4 blocks), row window
    ref = torch.zeros(40, 64, dtype=torch.float64)
    for r, n in enumerate(rows.tolist()):
        if 10 <= n < 50:
            ref[n - 10] += dst[r].double().cpu().view(4, 64).sum(0)
    assert rel_err(acc, ref) < 1e-6
    a = torch.randn(9, 64, device=dev)
    slab = torch.zeros(9, 256, device=dev)
    ops.copy_2d(a, slab[:, 64:], 9, 64)
    assert torch.equal(slab[:, 64:128], a) and float(slab[:, :64].abs().sum()) == 0


@pytest.mark.parametrize("M,N,K,split", [(300, 64, 100, 1), (1000, 64, 24, 1), (128, 256, 64, 1), (768, 64, 5000, 16),
                                         (100, 64, 2048, 7), (64, 64, 300, 3), (5, 3, 2, 1)])
def test_gemm_strided(dev, M, N, K, split):
    from elimrec_b200 import ops
    gen = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=gen).to(dev)
    B = torch.randn(K, N, generator=gen).to(dev)
    bias = torch.randn(N, generator=gen).to(dev)
    ref = A.double() @ B.double()
    ws = torch.empty(max(1, split * M * N), device=dev)
    # NN, row-major everything
    C = torch.empty(M, N, device=dev)
    ops.gemm(M, N, K, A, K, 1, B, N, 1, C, N, 1, bias=bias, split_k=split, ws=ws)
    assert rel_err(C, ref + bias.double()) < FP32_TOL
    # NT (Linear forward: B given as [N x K]) into a column block of a wider slab, accumulate
    Bt = B.t().contiguous()
    slab = torch.ones(M, N + 64, device=dev)
    ops.gemm(M, N, K, A, K, 1, Bt, 1, K, slab, N + 64, 1, accumulate=True, split_k=split, ws=ws, c_off=64)
    assert rel_err(slab[:, 64:], ref + 1.0) < FP32_TOL and float((slab[:, :64] - 1).abs().sum()) == 0
    # TN with transposed output and a device scale (weight-gradient form)
    At = A.t().contiguous()
    Ct = torch.empty(N, M, device=dev)
    scale = torch.tensor([0.5], device=dev)
    ops.gemm(M, N, K, At, 1, M, B, N, 1, Ct, 1, M, split_k=split, ws=ws, scale=scale)
    assert rel_err(Ct.t(), 0.5 * ref) < FP32_TOL


def test_colsum(dev):
    from elimrec_b200 import ops
    A = torch.randn(3000, 256, device=dev)
    out = torch.ones(64, device=dev)
    ws = torch.empty(ops.colsum_ws_floats(3000, 64), device=dev)
    ops.colsum(3000, 64, A, 256, out, ws, accumulate=True, a_off=128)
    assert rel_err(out, A[:, 128:192].double().sum(0) + 1) < FP32_TOL


def test_bpr_forward_backward(dev):
    from elimrec_b200 import ops
    from oracle.ref_model import OracleEliMRec
    U, I, B, nt = 50, 80, 333, 4
    tabs = [torch.randn(U + I, 64) * (0.1 + t) for t in range(nt)]
    tabs[1][5] = 0.0  # zero row: the eps clamp of F.normalize
    w = [1.0, 0.5, 0.0, 0.25]
    u = torch.randint(0, U, (B,)); p = torch.randint(0, I, (B,)); n = torch.randint(0, I, (B,))
    u[0] = 5
    leaf = [t.clone().double().requires_grad_(True) for t in tabs]
    loss = sum(wt * OracleEliMRec._bpr(t[u], t[U + p], t[U + n]) for wt, t in zip(w, leaf))
    loss.backward()
    dt = [t.to(dev) for t in tabs]
    lo = torch.empty(1, device=dev)
    rows = torch.empty(3 * B, dtype=torch.int32, device=dev)
    ig = torch.empty(3 * B, 64 * nt, device=dev)
    ops.bpr(dt, w, u.to(dev), p.to(dev), n.to(dev), U, lo, rows, ig, torch.empty(nt * B, device=dev))
    assert abs(float(lo) - float(loss)) < 1e-6 * abs(float(loss))
    assert torch.equal(rows.cpu().long(), torch.cat([u, U + p, U + n]))
    for t in range(nt):
        dense = torch.zeros(U + I, 64, dtype=torch.float64)
        dense.index_add_(0, rows.cpu().long(), ig[:, 64 * t:64 * (t + 1)].double().cpu())
        assert rel_err(dense, leaf[t].grad) < FP32_TOL, t


def test_adam_matches_torch(dev):
    from elimrec_b200 import ops
    p0 = torch.randn(1000, 64)
    ref = p0.clone().to(dev).requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, weight_decay=1e-4)
    p = p0.clone().to(dev)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    consts = torch.zeros(2, dtype=torch.float64, device=dev)
    for s in range(5):
        gwide = torch.randn(1000, 256, device=dev) * 0.01
        ref.grad = gwide[:, 64:128].contiguous()
        opt.step()
        ops.adam_tick(step, consts, 1e-3, 0.9, 0.999)
        ops.adam_apply(p, gwide[:, 64:128], 64, 256, m, v, consts, 0.9, 0.999, 1e-8, 1e-4)  # strided gradient view
    assert int(step) == 5
    assert rel_err(p, ref) < 1e-6
    assert rel_err(p - p0.to(dev), ref.detach() - p0.to(dev)) < 1e-4  # the UPDATE itself, not just the weights


def _rank_inputs(dev, U, I, n_mod, seed=0):
    g = torch.Generator().manual_seed(seed)
    fu, fi = torch.randn(U, 64, generator=g) * 0.3, torch.randn(I, 64, generator=g) * 0.3
    su = [torch.randn(U, 64, generator=g) for _ in range(n_mod)]
    si = [torch.randn(I, 64, generator=g) for _ in range(n_mod)]
    return fu, fi, su, si


def _ref_scores(fu, fi, su, si, users, mode):
    import torch.nn.functional as F
    ui = torch.sigmoid(fu[users] @ fi.t())
    if mode == 0:
        return torch.sigmoid(ui)
    def cm(x):
        for a, b in zip(su, si):
            x = x * torch.sigmoid(F.normalize(a[users], dim=1) @ F.normalize(b, dim=1).t())
        return x
    if mode == 1:
        return torch.sigmoid(cm(ui))
    return torch.sigmoid(cm(ui) - cm(ui.mean(-1, keepdim=True)))


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("n_mod", [1, 3])
def test_rank_scores_topk_metrics(dev, mode, n_mod):
    from elimrec_b200 import ops
    from oracle import ref_eval
    from gpu_util import topk_sets_match
    U, I, K = 150, 1000 + 37, 20
    fu, fi, su, si = _rank_inputs(dev, U, I, n_mod)
    users = torch.randperm(U)[:101]
    ref = _ref_scores(fu, fi, su, si, users, mode).numpy()
    sun = [torch.empty_like(a, device=dev) for a in su]
    sin_ = [torch.empty_like(a, device=dev) for a in si]
    for a, b in zip(su + si, sun + sin_):
        ops.row_normalize(a.to(dev), b)
    t = ops.rank_tables(U, I, mode, fu.to(dev), fi.to(dev), sun if mode else [], sin_ if mode else [])
    eu = users.to(dev).int()
    mean = torch.empty(eu.numel(), device=dev)
    ops.rank_rowmean(t, eu, mean)
    assert rel_err(mean, torch.sigmoid(fu[users] @ fi.t()).mean(-1)) < FP32_TOL
    sc = torch.empty(eu.numel(), I, device=dev)
    ops.rank_scores(t, eu, mean, sc)
    assert rel_err(sc, ref) < FP32_TOL
    # train mask: random sorted item lists per USER id
    rng = np.random.default_rng(1)
    train = {u: np.sort(rng.choice(I, size=rng.integers(0, 90), replace=False)) for u in range(U)}
    train[int(users[0])] = np.arange(0, 200)  # a user masking whole tiles
    ptr = np.zeros(U + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([train[u].size for u in range(U)])
    flat = np.concatenate([train[u] for u in range(U)]).astype(np.int32)
    idx = torch.empty(eu.numel(), K, dtype=torch.int32, device=dev)
    val = torch.empty(eu.numel(), K, device=dev)
    ops.rank_topk(t, eu, mean, torch.from_numpy(ptr).to(dev), torch.from_numpy(flat).to(dev), K, idx, val)
    masked = [train[int(u)] for u in users]
    exact, explained, bad = topk_sets_match(idx.cpu().numpy(), ref, masked, K, tol=2e-6)
    assert bad == 0 and exact >= 0.9 * len(masked), (exact, explained, bad)
    # metric curves from the GPU top-K are bit-equal to the oracle's (and to the reference's own C++ when built)
    truth = [np.sort(rng.choice(I, size=rng.integers(1, 30), replace=False)).tolist() for _ in masked]
    tp = np.zeros(len(truth) + 1, dtype=np.int64)
    tp[1:] = np.cumsum([len(x) for x in truth])
    tf = np.concatenate(truth).astype(np.int32)
    rows = torch.empty(len(truth), 5 * K, device=dev)
    sums = torch.zeros(5 * K, dtype=torch.float64, device=dev)
    ops.metric_rows(idx, torch.from_numpy(tp).to(dev), torch.from_numpy(tf).to(dev), [1, 2, 3, 4, 5], K, rows, sums)
    want = ref_eval.metric_rows(idx.cpu().numpy(), truth, [1, 2, 3, 4, 5], K)
    assert np.array_equal(rows.cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert np.allclose(sums.cpu().numpy(), want.astype(np.float64).sum(0), rtol=1e-12)
    if ref_eval.ref_cpp_available():
        s = sc.cpu().numpy().copy()
        for r, mi in enumerate(masked):
            s[r, mi] = -np.inf
        cpp = ref_eval.ref_cpp_metric_rows(s, truth, [1, 2, 3, 4, 5], K)
        # the reference C++ ranks the GPU's own scores: identical curves unless a tie group straddles rank K
        assert (np.abs(cpp - rows.cpu().numpy()).max(axis=1) < 1e-6).mean() > 0.97


def test_topk_matrix_ties_lowest_index(dev):
    from elimrec_b200 import ops
    from oracle import ref_eval
    rng = np.random.default_rng(3)
    s = rng.integers(0, 12, size=(40, 333)).astype(np.float32)  # massive ties
    s[0, :] = 1.0
    s[1, 5] = -np.inf
    idx = torch.empty(40, 20, dtype=torch.int32, device=dev)
    val = torch.empty(40, 20, device=dev)
    ops.topk_matrix(torch.from_numpy(s).to(dev), 20, idx, val)
    assert np.array_equal(idx.cpu().numpy(), ref_eval.topk_lowest_index(s, 20))


def test_device_sampler_bit_exact_and_valid(dev):
    from elimrec_b200 import synth
    from elimrec_b200.data import Dataset
    from elimrec_b200.sampler import PairwiseSamplerV2
    from oracle import philox_sampler
    inter, feats = synth.make_shape("tiny")
    ds = Dataset(None, interactions=inter, features=feats, name="tiny")
    s = PairwiseSamplerV2(ds, batch_size=64, mode="device", device=dev, seed=2022)
    s.epoch = 3
    u, p, n = (x.cpu().numpy().copy() for x in s.sample_epoch_device(5000))
    ru, rp, rn = philox_sampler.sample_triples(2022, 3, 5000, s.users, s.ptr, s.items, ds.num_items)
    assert np.array_equal(u, ru) and np.array_equal(p, rp) and np.array_equal(n, rn)
    tm = ds.train_matrix
    assert np.asarray(tm[u, p]).all() and not np.asarray(tm[u, n]).any()
    # distribution: users uniform over users with train items (chi-square, very loose)
    cnt = np.bincount(u, minlength=ds.num_users)[s.users]
    exp = 5000 / s.users.size
    assert ((cnt - exp) ** 2 / exp).sum() < 3 * s.users.size
    batches = list(s)
    assert sum(b[0].numel() for b in batches) == s.num_trainings and batches[0][0].is_cuda


@pytest.mark.parametrize("M,K,ldx_extra,ldy", [(1000, 128, 0, 64), (76085 // 8, 768, 0, 256), (333, 100, 0, 64), (4096, 64, 192, 64),
                                               (130, 2048, 0, 64), (128, 32, 0, 64)])
def test_linear_tf32_fwd_tcgen05(dev, M, K, ldx_extra, ldy):
    """tcgen05 TF32 projection GEMM vs fp64; tolerance = the north star's 1e-3 for TF32 GEMMs."""
    from elimrec_b200 import ops
    from gpu_util import TC_TOL
    g = torch.Generator().manual_seed(M + K)
    Xfull = torch.randn(M, K + ldx_extra, generator=g).to(dev)
    X = Xfull[:, ldx_extra:] if ldx_extra else Xfull           # column-block view of a wider slab
    W = (torch.randn(64, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(64, generator=g).to(dev)
    Y = torch.full((M + 1, ldy), 7.0, device=dev)
    ops.linear_tf32_fwd(X, W, b, Y[:M], col=ldy - 64)
    ref = X.double() @ W.double().t() + b.double()
    assert rel_err(Y[:M, ldy - 64:], ref) < TC_TOL
    assert float((Y[M] - 7).abs().max()) == 0 and (ldy == 64 or float((Y[:M, :ldy - 64] - 7).abs().max()) == 0)


@pytest.mark.parametrize("M,K,lddy", [(5000, 128, 256), (76085 // 4, 768, 256), (777, 100, 64), (300, 2048, 64), (33, 160, 128)])
def test_linear_tf32_wgrad_tcgen05(dev, M, K, lddy):
    """tcgen05 TF32 weight gradient (MN-major operands, TMEM-resident partials) vs fp64."""
    from elimrec_b200 import ops
    from gpu_util import TC_TOL
    g = torch.Generator().manual_seed(M * 7 + K)
    X = torch.randn(M, K, generator=g).to(dev)
    dYfull = torch.randn(M, lddy, generator=g).to(dev)
    col = lddy - 64
    dW = torch.full((64, K), 3.0, device=dev)
    ws = torch.empty(ops.linear_tf32_wgrad_ws_floats(M, K), device=dev)
    ops.linear_tf32_wgrad(dYfull, X, dW, ws, col=col)
    ref = dYfull[:, col:].double().t() @ X.double()
    assert rel_err(dW, ref) < TC_TOL
    dW2 = torch.empty_like(dW)
    ops.linear_tf32_wgrad(dYfull, X, dW2, ws, col=col)
    assert torch.equal(dW, dW2)


@pytest.mark.parametrize("B,nt", [(100, 4), (2048, 4), (333, 2)])
def test_inst_backward_fused(dev, B, nt):
    """Fused backward of the fusion Linear + heads on the 3B instance rows vs torch autograd (fp64)."""
    from elimrec_b200 import ops
    F = 64 * nt
    g = torch.Generator().manual_seed(B + nt)
    ig = torch.randn(3 * B, F, generator=g, dtype=torch.float64) * 0.01
    Oin = torch.randn(3 * B, F, generator=g, dtype=torch.float64)
    Wu = torch.randn(64, F, generator=g, dtype=torch.float64)
    Wi = torch.randn(64, F, generator=g, dtype=torch.float64)
    Ws = [torch.randn(64, 64, generator=g, dtype=torch.float64) for _ in range(nt - 1)]
    gs = torch.tensor([0.7], dtype=torch.float64)
    # reference: Y_f = O W^T + b, Y_m = O[:, blk] Ws^T + b; given dY, dO = dY W etc.
    dF, dS = ig[:, :64], [ig[:, 64 * (m + 1):64 * (m + 2)] for m in range(nt - 1)]
    dO = torch.cat([dF[:B] @ Wu, dF[B:] @ Wi])
    for m in range(nt - 1):
        dO[:, 64 * (m + 1):64 * (m + 2)] += dS[m] @ Ws[m]
    dO *= gs
    want = dict(dWu=gs * dF[:B].t() @ Oin[:B], dWi=gs * dF[B:].t() @ Oin[B:], dbu=gs * dF[:B].sum(0), dbi=gs * dF[B:].sum(0))
    f = lambda t: t.float().to(dev).contiguous()
    out = dict(dO=torch.empty(3 * B, F, device=dev), dWu=torch.empty(64, F, device=dev), dWi=torch.empty(64, F, device=dev),
               dbu=torch.empty(64, device=dev), dbi=torch.empty(64, device=dev))
    dWs = [torch.empty(64, 64, device=dev) for _ in range(nt - 1)]
    dbs = [torch.empty(64, device=dev) for _ in range(nt - 1)]
    ws = torch.empty(ops.inst_backward_ws_floats(B, nt, F), device=dev)
    ops.inst_backward(B, nt, F, f(ig), f(Oin), f(gs), f(Wu), f(Wi), [f(w) for w in Ws], out["dO"], out["dWu"], out["dWi"],
                      out["dbu"], out["dbi"], dWs, dbs, ws)
    assert rel_err(out["dO"], dO) < FP32_TOL
    for k, v in want.items():
        assert rel_err(out[k], v) < FP32_TOL, k
    for m in range(nt - 1):
        blk = slice(64 * (m + 1), 64 * (m + 2))
        assert rel_err(dWs[m], gs * dS[m].t() @ Oin[:, blk]) < FP32_TOL
        assert rel_err(dbs[m], gs * dS[m].sum(0)) < FP32_TOL


def test_adam_multi_matches_torch(dev):
    from elimrec_b200 import ops
    shapes = [(1000, 64), (64, 768), (64,), (3000, 64), (64, 64)]
    ps = [torch.randn(*s) for s in shapes]
    refs = [p.clone().to(dev).requires_grad_(True) for p in ps]
    opt = torch.optim.Adam(refs, lr=1e-3, weight_decay=1e-4)
    mine = [p.clone().to(dev) for p in ps]
    st = [(torch.zeros_like(p), torch.zeros_like(p)) for p in mine]
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    consts = torch.zeros(2, dtype=torch.float64, device=dev)
    for it in range(4):
        gwide = torch.randn(3000, 256, device=dev) * 0.01
        grads = [torch.randn(*s, device=dev) * 0.01 for s in shapes]
        grads[3] = gwide[:, 128:192]                      # strided view
        for r, g in zip(refs, grads):
            r.grad = g.contiguous()
        opt.step()
        ops.adam_tick(step, consts, 1e-3, 0.9, 0.999)
        ops.adam_apply_multi([(p, g, m, v) for p, g, (m, v) in zip(mine, grads, st)], consts, 0.9, 0.999, 1e-8, 1e-4)
    for p, r in zip(mine, refs):
        assert rel_err(p, r) < 1e-6


@pytest.mark.parametrize("M,K,ldx_extra", [(1000, 256, 0), (112741 // 4, 64, 192), (333, 100, 0), (130, 768, 0), (128, 32, 0)])
def test_linear_x3_fwd_fp32_class(dev, M, K, ldx_extra):
    """3xTF32 tcgen05 linear: must sit in the fp32 tolerance class (1e-5), not the TF32 one."""
    from elimrec_b200 import ops
    g = torch.Generator().manual_seed(M + K)
    Xfull = torch.randn(M, K + ldx_extra, generator=g).to(dev)
    X = Xfull[:, ldx_extra:] if ldx_extra else Xfull
    W = (torch.randn(64, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(64, generator=g).to(dev)
    Wh, Wl = torch.empty_like(W), torch.empty_like(W)
    ops.split_tf32(W, Wh, Wl)
    assert float((W - (Wh + Wl)).abs().max()) < 1e-6 * float(W.abs().max())
    Y = torch.full((M + 1, 64), 7.0, device=dev)
    ops.linear_x3_fwd(X, Wh, Wl, b, Y[:M])
    ref = X.double() @ W.double().t() + b.double()
    assert rel_err(Y[:M], ref) < FP32_TOL
    assert float((Y[M] - 7).abs().max()) == 0


@pytest.mark.parametrize("rows,n_heads", [(1000, 3), (36656 // 3, 3), (130, 1), (128, 2)])
def test_fuse_heads_x3_one_pass(dev, rows, n_heads):
    """Fused fusion-Linear + heads over the layer-mean slab (3xTF32, fp32 tolerance class)."""
    from elimrec_b200 import ops
    F = 64 * (1 + n_heads)
    g = torch.Generator().manual_seed(rows + n_heads)
    Oall = torch.randn(rows + 5, F, generator=g).to(dev)
    O = Oall[5:]                                     # row-offset view (like O[U:])
    Wf = (torch.randn(64, F, generator=g) / F ** 0.5).to(dev)
    bf = torch.randn(64, generator=g).to(dev)
    Ws = [(torch.randn(64, 64, generator=g) / 8).to(dev) for _ in range(n_heads)]
    bs = [torch.randn(64, generator=g).to(dev) for _ in range(n_heads)]
    mk = lambda w: (torch.empty_like(w), torch.empty_like(w))
    (Fh, Fl), sp = mk(Wf), [mk(w) for w in Ws]
    ops.prep_weights_tf32([(Wf, Fh, Fl)] + [(w, h, l) for w, (h, l) in zip(Ws, sp)])
    Fo = torch.full((rows + 1, 64), 7.0, device=dev)
    So = [torch.full((rows + 1, 64), 7.0, device=dev) for _ in range(n_heads)]
    ops.fuse_heads_x3(O, Fh, Fl, bf, [h for h, _ in sp], [l for _, l in sp], bs, Fo[:rows], [s_[:rows] for s_ in So])
    assert rel_err(Fo[:rows], O.double() @ Wf.double().t() + bf.double()) < FP32_TOL
    for m in range(n_heads):
        ref = O[:, 64 * (m + 1):64 * (m + 2)].double() @ Ws[m].double().t() + bs[m].double()
        assert rel_err(So[m][:rows], ref) < FP32_TOL, m
        assert float((So[m][rows] - 7).abs().max()) == 0
    assert float((Fo[rows] - 7).abs().max()) == 0


@pytest.mark.parametrize("U,I,n_heads", [(300, 500, 3), (36656 // 4, 76085 // 4, 3), (128, 1, 1), (50, 70, 2)])
def test_fuse_heads_x3_all_persistent(dev, U, I, n_heads):
    """Persistent whole-slab variant (TMEM double buffering): user rows with Wu, item rows with Wi, fp32 class."""
    from elimrec_b200 import ops
    F = 64 * (1 + n_heads)
    g = torch.Generator().manual_seed(U + I)
    O = torch.randn(U + I, F, generator=g).to(dev)
    mk = lambda *s: (torch.randn(*s, generator=g) / s[-1] ** 0.5).to(dev)
    Wu, Wi, bu, bi = mk(64, F), mk(64, F), mk(64), mk(64)
    Ws, bs = [mk(64, 64) for _ in range(n_heads)], [mk(64) for _ in range(n_heads)]
    pair = lambda w: (torch.empty_like(w), torch.empty_like(w))
    Wus, Wis, Wss = pair(Wu), pair(Wi), [pair(w) for w in Ws]
    ops.prep_weights_tf32([(Wu, *Wus), (Wi, *Wis)] + [(w, h, l) for w, (h, l) in zip(Ws, Wss)])
    Fo = torch.full((U + I + 1, 64), 7.0, device=dev)
    So = [torch.full((U + I + 1, 64), 7.0, device=dev) for _ in range(n_heads)]
    ops.fuse_heads_x3_all(U, I, O, Wus, bu, Wis, bi, [h for h, _ in Wss], [l for _, l in Wss], bs, Fo, So)
    Od = O.double()
    assert rel_err(Fo[:U], Od[:U] @ Wu.double().t() + bu.double()) < FP32_TOL
    assert rel_err(Fo[U:U + I], Od[U:] @ Wi.double().t() + bi.double()) < FP32_TOL
    for m in range(n_heads):
        ref = Od[:, 64 * (m + 1):64 * (m + 2)] @ Ws[m].double().t() + bs[m].double()
        assert rel_err(So[m][:U + I], ref) < FP32_TOL, m
        assert float((So[m][U + I] - 7).abs().max()) == 0
    assert float((Fo[U + I] - 7).abs().max()) == 0


@pytest.mark.parametrize("mode,n_mod", [(2, 3), (1, 3), (0, 1), (2, 1)])
def test_rank_tc_matches_exact_path(dev, mode, n_mod):
    """Tensor-core evaluator (fp16 hi/lo pairs, tcgen05) vs the exact fp32 FFMA rank kernel and the fp64 scores."""
    from elimrec_b200 import ops
    from gpu_util import topk_sets_match
    U, I, K = 700, 5000 + 13, 20
    fu, fi, su, si = _rank_inputs(dev, U, I, n_mod, seed=5)
    fu, fi = fu * 0.5, fi * 0.5
    users = torch.randperm(U)[:300 + 7]
    eu = users.to(dev).int()
    norm = lambda t: torch.nn.functional.normalize(t, dim=1)
    sun, sin_ = [norm(a).to(dev) for a in su], [norm(a).to(dev) for a in si]
    fud, fid = fu.to(dev), fi.to(dev)
    rng = np.random.default_rng(2)
    train = {u: np.sort(rng.choice(I, size=rng.integers(0, 60), replace=False)) for u in range(U)}
    train[int(users[1])] = np.arange(100, 400)
    ptr_ = np.zeros(U + 1, dtype=np.int64); ptr_[1:] = np.cumsum([train[u].size for u in range(U)])
    flat = np.concatenate([train[u] for u in range(U)]).astype(np.int32)
    tp, tf = torch.from_numpy(ptr_).to(dev), torch.from_numpy(flat).to(dev)
    # exact path
    t = ops.rank_tables(U, I, mode, fud, fid, sun if mode else [], sin_ if mode else [])
    mean_x = torch.empty(eu.numel(), device=dev)
    ops.rank_rowmean(t, eu, mean_x)
    idx_x = torch.empty(eu.numel(), K, dtype=torch.int32, device=dev); val_x = torch.empty(eu.numel(), K, device=dev)
    ops.rank_topk(t, eu, mean_x, tp, tf, K, idx_x, val_x)
    # tensor-core path
    def split(x, amax):
        sc = 2.0 ** np.floor(np.log2(4096.0 / amax))
        hi = torch.empty(x.shape, dtype=torch.float16, device=dev); lo = torch.empty_like(hi)
        ops.split_fp16(x.contiguous(), float(sc), hi, lo)
        return hi, lo, sc
    tabs = []
    for a, b in [(fud, fid)] + (list(zip(sun, sin_)) if mode else []):
        ah, al, sa = split(a, float(a.abs().max()))
        bh, bl, sb = split(b, float(b.abs().max()))
        tabs.append((ah, al, bh, bl, float(1.0 / (sa * sb))))
    tt = ops.rank_tc_tables(U, I, mode, tabs)
    mean_t = torch.empty(eu.numel(), device=dev)
    ops.rank_tc(tt, 0, eu, None, None, None, K, None, None, mean_t)
    assert rel_err(mean_t, mean_x) < 2e-6
    idx_t = torch.empty(eu.numel(), K, dtype=torch.int32, device=dev); val_t = torch.empty(eu.numel(), K, device=dev)
    ops.rank_tc(tt, 1, eu, mean_t, tp, tf, K, idx_t, val_t, None)
    assert rel_err(val_t, val_x) < 2e-6
    same = (idx_t == idx_x).all(dim=1).float().mean()
    ref = _ref_scores(fu, fi, su, si, users, mode).numpy()
    exact, explained, bad = topk_sets_match(idx_t.cpu().numpy(), ref, [train[int(u)] for u in users], K, tol=2e-6)
    assert bad == 0 and float(same) > 0.97, (float(same), exact, explained, bad)


# ---- round 2: kernels of the linear schedule and the graph-replayable batch sampler -----------------------------------------
@pytest.mark.parametrize("n_mod,L", [(3, 3), (1, 2), (0, 1), (3, 8)])
def test_lin_assemble_and_seed_match_restatement(dev, n_mod, L):
    """elimrec_lin_assemble / elimrec_lin_seed against the few-line torch statements in tests/sim_ops.py (instance rows with
    repeats, all-rows mode, accumulate on / off, strided tables)."""
    import sim_ops
    from elimrec_b200 import ops
    U, I, Fw = 50, 70, 64 * (1 + max(n_mod, 1))
    g = torch.Generator().manual_seed(L)
    tabs = [(torch.randn(U, 64, generator=g), torch.randn(I, 96, generator=g)[:, 16:80]) for _ in range(L + 1)]
    rows = torch.randint(0, U + I, (333,), generator=g, dtype=torch.int32)
    rows[:7] = rows[0]
    for acc in (False, True):
        for rr, n_rows in ((rows, None), (None, U + I)):
            base = torch.randn(rows.numel() if rr is not None else U + I, Fw, generator=g)
            want = base.clone()
            sim_ops.lin_assemble(rr, U, sim_ops.lin_layers(tabs), 0.25, n_mod, acc, want, n_rows=n_rows)
            got = base.to(dev)
            dt = [(a.to(dev), b.to(dev)) for a, b in tabs]
            dt = [(a, torch.cat([torch.zeros(I, 16, device=dev), b, torch.zeros(I, 16, device=dev)], 1)[:, 16:80]) for a, b in dt]
            ops.lin_assemble(rr.to(dev) if rr is not None else None, U, ops.lin_layers(dt), 0.25, n_mod, acc, got, n_rows=n_rows)
            assert rel_err(got, want) < 1e-6
    dO = torch.randn(rows.numel(), Fw, generator=g)
    for layer in range(L + 1):
        want = torch.randn(U + I, 64, generator=g)
        got = want.to(dev)
        sim_ops.lin_seed(rows, U, layer, dO, n_mod, 0.25, want)
        ops.lin_seed(rows.to(dev), U, layer, dO.to(dev), n_mod, 0.25, got)
        assert rel_err(got, want) < 2e-6          # atomics: order of the repeated rows


def test_pack_proj_weights_and_axpy(dev):
    from elimrec_b200 import ops
    g = torch.Generator().manual_seed(0)
    items, want = [], []
    for dm in (16, 100, 768):
        kp = ((dm + 1 + 3) // 4) * 4
        W, b = torch.randn(64, dm, generator=g), torch.randn(64, generator=g)
        dst = torch.full((64, kp), float("nan"), device=dev)
        items.append((W.to(dev), b.to(dev), dst))
        w = torch.zeros(64, kp)
        w[:, :dm], w[:, dm] = W, b
        want.append(w)
    ops.pack_proj_weights(items, False)
    for (_, _, dst), w in zip(items, want):
        assert torch.equal(dst.cpu(), w)
    ops.pack_proj_weights(items, True)
    for (_, _, dst), w in zip(items, want):
        d = dst.cpu()
        assert (d.view(torch.int32) & 0x1FFF).eq(0).all() and rel_err(d, w) < 2 ** -11      # TF32: 10 mantissa bits, to nearest
    X, Y = torch.randn(300, 132, generator=g), torch.randn(300, 260, generator=g)
    for acc in (True, False):
        y = Y.to(dev)
        ops.axpy_2d(X.to(dev)[:, 4:], y[:, 8:], 300, 120, 0.25, accumulate=acc)
        ref = Y.clone()
        ref[:, 8:128] = (ref[:, 8:128] if acc else 0) + 0.25 * X[:, 4:124]
        assert rel_err(y, ref) < 1e-7


def test_batch_sampler_is_a_slice_of_the_epoch_stream(dev):
    """elimrec_sample_batch_device(batch index on the device) == the matching slice of elimrec_sample_triples_device == the
    numpy restatement (oracle/philox_sampler.py), bit for bit."""
    from elimrec_b200 import ops
    from oracle import philox_sampler
    rng = np.random.default_rng(3)
    U, I, B = 200, 300, 96
    deg = rng.integers(1, 12, U)
    ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    items = np.concatenate([np.sort(rng.choice(I, d, replace=False)) for d in deg]).astype(np.int32)
    uid = np.arange(U, dtype=np.int32)
    d = lambda a: torch.from_numpy(a).to(dev)
    eu, ep, en = (torch.empty(5 * B, dtype=torch.int64, device=dev) for _ in range(3))
    ops.sample_triples_device(11, 2, 5 * B, d(uid), d(ptr), d(items), I, eu, ep, en)
    ctr = torch.zeros(1, dtype=torch.int64, device=dev)
    for step in (0, 3, 4):
        ctr.fill_(step)
        bu, bp, bn = (torch.empty(B, dtype=torch.int64, device=dev) for _ in range(3))
        ops.sample_batch_device(11, 2, ctr, B, d(uid), d(ptr), d(items), I, bu, bp, bn)
        sl = slice(step * B, (step + 1) * B)
        assert torch.equal(bu, eu[sl]) and torch.equal(bp, ep[sl]) and torch.equal(bn, en[sl])
        wu, wp, wn = philox_sampler.sample_triples(11, 2, B, uid, ptr, items, I, first=step * B)
        assert np.array_equal(bu.cpu().numpy(), wu) and np.array_equal(bp.cpu().numpy(), wp) and np.array_equal(bn.cpu().numpy(), wn)


def test_wgrad_multi_matches_fp64(dev):
    """elimrec_wgrad_multi: several skinny A^T B problems (row ranges, column blocks, a K that is not a multiple of 64 and a
    B whose rows are not 16-byte aligned for the tail tile) + bias column sums, against fp64; deterministic."""
    from elimrec_b200 import ops
    g = torch.Generator().manual_seed(1)
    R = 1000
    A = torch.randn(R, 256, generator=g).to(dev)
    Bm = torch.randn(R, 300, generator=g).to(dev)
    Z = torch.randn(R, 140, generator=g).to(dev)
    gs = torch.tensor([0.37], device=dev)
    outs = [torch.full((64, 256), float("nan"), device=dev), torch.full((64, 64), float("nan"), device=dev),
            torch.full((64, 132), float("nan"), device=dev), torch.full((64, 100), float("nan"), device=dev)]
    b0, b1 = torch.full((64,), float("nan"), device=dev), torch.full((64,), float("nan"), device=dev)
    pr = [(A, 0, Bm, 0, 256, 0, 400, outs[0], b0, True), (A, 64, Bm, 64, 64, 0, R, outs[1], b1, True),
          (A, 128, Z, 4, 132, 0, R, outs[2], None, False), (A, 192, Bm, 1, 100, 123, 777, outs[3], None, False)]
    for splits in (1, 7, 24):
        ws = torch.empty(ops.wgrad_multi_ws_floats(pr, splits), device=dev)
        ops.wgrad_multi(pr, splits, ws, gs)
        first = [o.clone() for o in outs]
        for (A_, ac, B_, bc, K, r0, r1, out, bias, by_g) in pr:
            sc = 0.37 if by_g else 1.0
            want = sc * (A_[r0:r1, ac:ac + 64].double().t() @ B_[r0:r1, bc:bc + K].double())
            assert rel_err(out[:, :K], want) < 2e-6
            if bias is not None:
                assert rel_err(bias, sc * A_[r0:r1, ac:ac + 64].double().sum(0)) < 2e-6
        ops.wgrad_multi(pr, splits, ws, gs)
        assert all(torch.equal(a, b) for a, b in zip(first, outs))


@pytest.mark.parametrize("B,n_heads", [(2048, 3), (77, 3), (500, 2), (33, 0)])
def test_inst_forward_matches_fp64(dev, B, n_heads):
    """elimrec_inst_forward: fusion Linear (users / items take different weights) + heads on the 3B instance rows, exact
    fp32, against fp64 and deterministic; row counts that are not multiples of the 16-row CTA tile."""
    from elimrec_b200 import ops
    g = torch.Generator().manual_seed(B)
    nt = 1 + n_heads
    Fw = 64 * nt
    O = torch.randn(3 * B, Fw, generator=g).to(dev)
    Wu, Wi = (torch.randn(64, Fw, generator=g) * 0.1).to(dev), (torch.randn(64, Fw, generator=g) * 0.1).to(dev)
    bu, bi = torch.randn(64, generator=g).to(dev), torch.randn(64, generator=g).to(dev)
    Ws = [(torch.randn(64, 64, generator=g) * 0.1).to(dev) for _ in range(n_heads)]
    bs = [torch.randn(64, generator=g).to(dev) for _ in range(n_heads)]
    Fo = torch.full((3 * B, 64), float("nan"), device=dev)
    So = [torch.full((3 * B, 64), float("nan"), device=dev) for _ in range(n_heads)]
    ops.inst_forward(B, nt, Fw, O, Wu, Wi, Ws, bu, bi, bs, Fo, So)
    Od = O.double()
    want = torch.cat([Od[:B] @ Wu.double().t() + bu.double(), Od[B:] @ Wi.double().t() + bi.double()])
    assert rel_err(Fo, want) < 1e-6
    for m in range(n_heads):
        assert rel_err(So[m], Od[:, 64 * (m + 1):64 * (m + 2)] @ Ws[m].double().t() + bs[m].double()) < 1e-6
    first = Fo.clone()
    ops.inst_forward(B, nt, Fw, O, Wu, Wi, Ws, bu, bi, bs, Fo, So)
    assert torch.equal(first, Fo)


def test_inst_dO_seed_equals_two_launches(dev):
    """elimrec_inst_dout_seed = elimrec_inst_backward_part(1) + elimrec_lin_seed2 in one launch: the same d O[inst] bit for bit,
    the same seeds (float atomics: duplicates of a node may add in another order); bpr in two parts = bpr in one."""
    from elimrec_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, nt, Fw, N = 300, 4, 256, 500
    ig = torch.randn(3 * B, Fw, generator=g).to(dev) * 1e-3
    Wu, Wi = (torch.randn(64, Fw, generator=g) * 0.1).to(dev), (torch.randn(64, Fw, generator=g) * 0.1).to(dev)
    Ws = [(torch.randn(64, 64, generator=g) * 0.1).to(dev) for _ in range(3)]
    rows = torch.randint(0, N, (3 * B,), generator=g).to(torch.int32).to(dev)       # plenty of duplicates
    gs = torch.tensor([0.7], device=dev)
    dO1, dO2 = torch.empty(3 * B, Fw, device=dev), torch.empty(3 * B, Fw, device=dev)
    GA1, GB1, GA2, GB2 = (torch.zeros(N, 64, device=dev) for _ in range(4))
    dummy = torch.empty(1, device=dev)
    wsz = torch.empty(ops.inst_backward_ws_floats(B, nt, Fw), device=dev)
    ops.inst_backward(B, nt, Fw, ig, dO1, gs, Wu, Wi, Ws, dO1, dummy, dummy, dummy, dummy, [dummy] * 3, [dummy] * 3, wsz, part=1)
    ops.lin_seed2(rows, dO1, 3, 0.25, GA1, GB1)
    ops.inst_dO_seed(B, nt, Fw, ig, gs, Wu, Wi, Ws, dO2, rows, 3, 0.25, GA2, GB2)
    assert torch.equal(dO1, dO2)
    assert rel_err(GA2, GA1) < 1e-6 and rel_err(GB2, GB1) < 1e-6 and float(GA1.abs().max()) > 0
    # reference of the seeds in fp64
    want = torch.zeros(N, 64, dtype=torch.float64, device=dev)
    want.index_add_(0, rows.long(), 0.25 * dO1.double().reshape(3 * B, 4, 64).sum(1))
    assert rel_err(GA2, want) < 1e-6
    # bpr: parts 1 + 2 = 3
    T = [torch.randn(N, 64, generator=g).to(dev) for _ in range(4)]
    u = torch.randint(0, 200, (B,), generator=g).to(dev)
    p_ = torch.randint(0, 300, (B,), generator=g).to(dev)
    n = torch.randint(0, 300, (B,), generator=g).to(dev)
    outs = []
    for parts in ((3,), (1, 2)):
        loss, ir = torch.zeros(1, device=dev), torch.zeros(3 * B, dtype=torch.int32, device=dev)
        igo, terms = torch.zeros(3 * B, Fw, device=dev), torch.zeros(4 * B, device=dev)
        for part in parts:
            ops.bpr(T, [1.0, 0.5, 0.5, 0.5], u, p_, n, 200, loss, ir, igo, terms, part=part)
        outs.append((loss, ir, igo))
    assert all(torch.equal(a, b) for a, b in zip(*outs)) and float(outs[0][0]) > 0


def test_wgrad_multi_x3_matches_fp64(dev):
    """elimrec_wgrad_multi_x3 (tcgen05 3xTF32, MN-major operands split hi / lo in shared memory, bias sums through an all-ones
    tile): the same problem family as the exact kernel - row ranges that are not multiples of the 32-row stage, column blocks,
    K = 64 / 100 / 132 / 256 / 772 tiles, more row ranges than rows - against fp64 in the fp32 class, against the exact kernel,
    deterministic; a misaligned B is refused."""
    from elimrec_b200 import ops
    from elimrec_b200._lib import ElimrecError
    g = torch.Generator().manual_seed(2)
    R = 1000
    A = torch.randn(R, 256, generator=g).to(dev)
    Bm = torch.randn(R, 300, generator=g).to(dev)
    Z = (torch.randn(R, 780, generator=g) * torch.logspace(-3, 2, 780)).to(dev)
    gs = torch.tensor([0.37], device=dev)
    outs = [torch.full((64, 256), float("nan"), device=dev), torch.full((64, 64), float("nan"), device=dev),
            torch.full((64, 132), float("nan"), device=dev), torch.full((64, 100), float("nan"), device=dev),
            torch.full((64, 772), float("nan"), device=dev)]
    b0, b1 = torch.full((64,), float("nan"), device=dev), torch.full((64,), float("nan"), device=dev)
    pr = [(A, 0, Bm, 0, 256, 0, 400, outs[0], b0, True), (A, 64, Bm, 64, 64, 0, R, outs[1], b1, True),
          (A, 128, Z, 4, 132, 0, R, outs[2], None, False), (A, 192, Bm, 8, 100, 123, 777, outs[3], None, False),
          (A, 64, Z, 8, 772, 5, 995, outs[4], None, False)]
    exact = [torch.empty_like(o) for o in outs]
    ws = torch.empty(ops.wgrad_multi_ws_floats(pr, 7), device=dev)
    ops.wgrad_multi([(*q[:7], e, None, q[9]) for q, e in zip(pr, exact)], 7, ws, gs)
    for splits in (1, 7, 11, 40):
        ws = torch.empty(ops.wgrad_multi_ws_floats(pr, splits), device=dev)
        ops.wgrad_multi(pr, splits, ws, gs, x3=True)
        first = [o.clone() for o in outs]
        for (A_, ac, B_, bc, K, r0, r1, out, bias, by_g), ex in zip(pr, exact):
            sc = 0.37 if by_g else 1.0
            a64, b64 = A_[r0:r1, ac:ac + 64].double(), B_[r0:r1, bc:bc + K].double()
            want = sc * (a64.t() @ b64)
            bound = sc * (a64.abs().t() @ b64.abs())                 # per element: error relative to sum |a||b|
            assert float(((out[:, :K].double() - want).abs() / bound).max()) < 4e-6
            assert rel_err(out[:, :K], want) < 5e-6 and rel_err(out[:, :K], ex[:, :K]) < 5e-6
            if bias is not None:
                assert rel_err(bias, sc * a64.sum(0)) < 2e-6
        ops.wgrad_multi(pr, splits, ws, gs, x3=True)
        assert all(torch.equal(a, b) for a, b in zip(first, outs))
    with pytest.raises(ElimrecError):
        ops.wgrad_multi([(A, 192, Bm, 1, 100, 123, 777, outs[3], None, False)], 4, ws, gs, x3=True)


@pytest.mark.parametrize("variant", [0, 1, 3])
def test_spmm64_pair_matches_fp64(dev, variant):
    """elimrec_spmm64_pair (both halves in one launch, 8-lane groups, split rows with last-arriver reduction): dense, row- and
    column-masked, with the additive epilogue; against fp64; deterministic; masked rows carry the bits of the dense launch."""
    from elimrec_b200 import ops
    from elimrec_b200.graph import BipartiteGraph
    U, I = 700, 500
    m = _rand_graph(U, I, 6000, [(3, 480), (10, 130), (11, 65), (12, 64)], seed=variant)
    g = BipartiteGraph(m, dev)
    assert g.ui.n_split64 > 0 and g.iu.n_item64 >= I
    X = torch.randn(U + I, 64, device=dev)
    A_ui = sp.csr_matrix((g.ui.vals_host.astype(np.float64), g.ui.indices_host, g.ui.indptr_host), shape=(U, I))
    A_iu = sp.csr_matrix((g.iu.vals_host.astype(np.float64), g.iu.indices_host, g.iu.indptr_host), shape=(I, U))
    Xd = X.double().cpu().numpy()
    ref = np.concatenate([A_ui @ Xd[U:], A_iu @ Xd[:U]])
    Y = torch.full((U + I, 64), float("nan"), device=dev)
    ops.spmm64_pair(g.ui, g.iu, X[U:], X[:U], Y[:U], Y[U:], variant=variant)
    assert rel_err(Y, ref) < FP32_TOL
    Y2 = torch.empty_like(Y)
    ops.spmm64_pair(g.ui, g.iu, X[U:], X[:U], Y2[:U], Y2[U:], variant=variant)
    assert torch.equal(Y, Y2) and int(g.ui.counter64.abs().sum()) == 0 and int(g.iu.counter64.abs().sum()) == 0
    # row mask: marked rows = dense bits, the others untouched
    rm = (torch.rand(U + I, device=dev) < 0.4).to(torch.uint8)
    rm[3] = 1
    Y3 = torch.full_like(Y, 7.0)
    ops.spmm64_pair(g.ui, g.iu, X[U:], X[:U], Y3[:U], Y3[U:], row_mask_u=rm[:U], row_mask_i=rm[U:], variant=variant)
    assert torch.equal(Y3[rm.bool()], Y[rm.bool()]) and (Y3[~rm.bool()] == 7.0).all()
    # column mask: dropped columns may hold garbage; + additive epilogue on marked rows
    cm = (torch.rand(U + I, device=dev) < 0.3).to(torch.uint8)
    Xg = torch.where(cm.bool().unsqueeze(1), X, torch.full_like(X, float("nan")))
    add = torch.randn(U + I, 64, device=dev)
    am = (torch.rand(U + I, device=dev) < 0.5).to(torch.uint8)
    Y4 = torch.empty_like(Y)
    ops.spmm64_pair(g.ui, g.iu, Xg[U:], Xg[:U], Y4[:U], Y4[U:], col_mask_u=cm[U:], col_mask_i=cm[:U], addend_u=add[:U],
                    addend_i=add[U:], add_mask_u=am[:U], add_mask_i=am[U:], variant=variant)
    Xz = (X * cm.unsqueeze(1)).double().cpu().numpy()
    ref4 = np.concatenate([A_ui @ Xz[U:], A_iu @ Xz[:U]]) + (add * am.unsqueeze(1)).double().cpu().numpy()
    assert rel_err(Y4, ref4) < FP32_TOL


@pytest.mark.parametrize("width", [32, 16, 8])
def test_spmm64_pair_narrow_widths(dev, width):
    """the column-sharded widths (64 / world columns per rank): dense, masked, additive epilogue, fused Adam; vs fp64 / torch"""
    from elimrec_b200 import ops
    from elimrec_b200.graph import BipartiteGraph
    U, I = 700, 500
    g = BipartiteGraph(_rand_graph(U, I, 6000, [(3, 480), (10, 130), (11, 65)], seed=width), dev)
    X = torch.randn(U + I, width, device=dev)
    A_ui = sp.csr_matrix((g.ui.vals_host.astype(np.float64), g.ui.indices_host, g.ui.indptr_host), shape=(U, I))
    A_iu = sp.csr_matrix((g.iu.vals_host.astype(np.float64), g.iu.indices_host, g.iu.indptr_host), shape=(I, U))
    Xd = X.double().cpu().numpy()
    ref = np.concatenate([A_ui @ Xd[U:], A_iu @ Xd[:U]])
    Y = torch.full((U + I, width), float("nan"), device=dev)
    ops.spmm64_pair(g.ui, g.iu, X[U:], X[:U], Y[:U], Y[U:], width=width)
    assert rel_err(Y, ref) < FP32_TOL
    rm = (torch.rand(U + I, device=dev) < 0.4).to(torch.uint8)
    cm = (torch.rand(U + I, device=dev) < 0.5).to(torch.uint8)
    add = torch.randn(U + I, width, device=dev)
    Y2 = torch.full_like(Y, 3.0)
    ops.spmm64_pair(g.ui, g.iu, X[U:], X[:U], Y2[:U], Y2[U:], row_mask_u=rm[:U], row_mask_i=rm[U:], col_mask_u=cm[U:],
                    col_mask_i=cm[:U], addend_u=add[:U], addend_i=add[U:], add_mask_u=rm[:U], add_mask_i=rm[U:], width=width)
    Xz = (X * cm.unsqueeze(1)).double().cpu().numpy()
    ref2 = np.concatenate([A_ui @ Xz[U:], A_iu @ Xz[:U]]) + add.double().cpu().numpy()
    sel = rm.bool().cpu().numpy()
    assert rel_err(Y2[rm.bool()], ref2[sel]) < FP32_TOL and (Y2[~rm.bool()] == 3.0).all()
    # fused Adam on the finished rows == torch.optim.Adam on the gradient
    p0 = torch.randn(U + I, width, device=dev)
    prm = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([prm], lr=1e-3, weight_decay=1e-4)
    prm.grad = Y.clone()
    opt.step()
    consts, step = torch.zeros(2, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.int64, device=dev)
    ops.adam_tick(step, consts, 1e-3, 0.9, 0.999)
    pu, pi = p0[:U].clone(), p0[U:].clone()
    mu, vu, mi, vi = (torch.zeros_like(t) for t in (pu, pu, pi, pi))
    old = torch.empty_like(p0)
    ops.spmm64_pair(g.ui, g.iu, X[U:], X[:U], None, None, adam_u=(pu, mu, vu, old[:U]), adam_i=(pi, mi, vi, old[U:]),
                    adam_consts=(consts, 0.9, 0.999, 1e-8, 1e-4), width=width)
    assert rel_err(torch.cat([pu, pi]), prm.detach()) < 2e-6 and torch.equal(old, p0)


@pytest.mark.parametrize("world", [2, 8])
def test_colshard_glue_kernels_match_restatement(dev, world):
    import sim_ops
    from elimrec_b200 import ops
    w, U, I, B, L, nm = 64 // world, 50, 70, 33, 3, 3
    g = torch.Generator().manual_seed(world)
    T = torch.stack([torch.stack([torch.randint(0, U, (B,), generator=g), torch.randint(0, I, (B,), generator=g),
                                  torch.randint(0, I, (B,), generator=g)]) for _ in range(world)]).reshape(-1)
    rows_w, m1w, m2w = torch.zeros(world * 3 * B, dtype=torch.int32), torch.ones(U + I, dtype=torch.uint8), torch.ones(U + I, dtype=torch.uint8)
    sim_ops.cs_inst_rows(world, B, T, U, rows_w, m1w, m2w)
    rows, m1, m2 = torch.zeros_like(rows_w, device=dev), torch.ones(U + I, dtype=torch.uint8, device=dev), torch.ones(U + I, dtype=torch.uint8, device=dev)
    ops.cs_inst_rows(world, B, T.to(dev), U, rows, m1, m2)
    assert torch.equal(rows.cpu(), rows_w) and torch.equal(m1.cpu(), m1w) and torch.equal(m2.cpu(), m2w)
    tabs = [(torch.randn(U, w, generator=g), torch.randn(I, w, generator=g)) for _ in range(L + 1)]
    want = torch.zeros(world * 3 * B, 2 * w)
    sim_ops.cs_pack(rows_w, U, sim_ops.lin_layers(tabs), 0.25, w, want)
    got = torch.zeros(world * 3 * B, 2 * w, device=dev)
    ops.cs_pack(rows, U, ops.lin_layers([(a.to(dev), b.to(dev)) for a, b in tabs]), 0.25, w, got)
    assert rel_err(got, want) < 1e-6
    recv = torch.randn(world, 3 * B, 2 * w, generator=g)
    O0 = torch.randn(3 * B, 64 * (1 + nm), generator=g)
    want = O0.clone()
    sim_ops.cs_unpack(world, 3 * B, w, recv, nm, want)
    got = O0.to(dev)
    ops.cs_unpack(world, 3 * B, w, recv.to(dev), nm, got)
    assert rel_err(got, want) < 1e-6
    dO = torch.randn(3 * B, 64 * (1 + nm), generator=g)
    want = torch.zeros(world, 3 * B, 2 * w)
    sim_ops.cs_seed_pack(3 * B, world, w, dO, nm, 0.125, want)
    got = torch.zeros(world, 3 * B, 2 * w, device=dev)
    ops.cs_seed_pack(3 * B, world, w, dO.to(dev), nm, 0.125, got)
    assert rel_err(got, want) < 1e-6
    GA0, GB0 = torch.randn(U + I, w, generator=g), torch.randn(U + I, w, generator=g)
    wa, wb = GA0.clone(), GB0.clone()
    sim_ops.cs_seed_scatter(rows_w, w, recv, wa, wb)
    ga, gb = GA0.to(dev), GB0.to(dev)
    ops.cs_seed_scatter(rows, w, recv.to(dev), ga, gb)
    assert rel_err(ga, wa) < 2e-6 and rel_err(gb, wb) < 2e-6
