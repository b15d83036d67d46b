"""-m gpu: the kernels added for the SURVEY.md 8(f) rows, each against an independent torch computation.

(The end-to-end parity of those rows against the reference's golden vectors is tests/test_schedule_sim.py[cuda].)"""
import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gpu_util import FP32_TOL, rel_err
from test_gpu_kernels import _rank_inputs


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.mark.parametrize("rows,G", [(64, 4), (64, 2), (1000, 4), (36656 // 8 + 3, 4), (1, 1)])
def test_block_helpers(dev, rows, G):
    """tie_blocks / fold_blocks / broadcast_cols / axpy_rows / layer_mean on strided views, ragged row counts."""
    from elimrec_b200 import ops
    Fw = 64 * G
    src = torch.randn(rows, 64, device=dev)
    slab = torch.full((rows, Fw + 64), float("nan"), device=dev)       # row stride larger than the width
    ops.tie_blocks(src, slab, G, 1.0 / G)
    assert torch.equal(slab[:, :Fw], (src * (1.0 / G)).repeat(1, G)) and torch.isnan(slab[:, Fw:]).all()
    ops.broadcast_cols(src, slab, rows, G)
    assert torch.equal(slab[:, :Fw], src.repeat(1, G))
    wide = torch.randn(rows, Fw + 64, device=dev)
    out = torch.full((rows, 128), float("nan"), device=dev)
    ops.fold_blocks(wide, out, G, 0.5)
    ref = wide[:, :Fw].double().reshape(rows, G, 64).sum(1) * 0.5
    assert rel_err(out[:, :64], ref) < 1e-6 and torch.isnan(out[:, 64:]).all()
    s = torch.rand(rows, device=dev)
    X, Y = torch.randn(rows, Fw + 64, device=dev), torch.randn(rows, Fw, device=dev)
    ref = Y.double() + s.double().unsqueeze(1) * X[:, :Fw].double()
    keep = X.clone()
    ops.axpy_rows(s, X, Y, Fw)
    assert rel_err(Y, ref) < 1e-6 and torch.equal(X, keep)
    layers = [torch.randn(rows, Fw + 64 * (k % 2), device=dev) for k in range(4)]
    O = torch.full((rows, Fw + 64), float("nan"), device=dev)
    ops.layer_mean(layers, O, Fw, 0.25)
    ref = sum(t[:, :Fw].double() for t in layers) * 0.25
    assert rel_err(O[:, :Fw], ref) < 1e-6 and torch.isnan(O[:, Fw:]).all()


def _ref_scores_fm(fu, fi, su, si, users, pm, fm):
    """general_cm_fusion + predict (models/EliMRec.py:96-113,155-212) in fp64."""
    fu, fi = fu.double(), fi.double()
    ui = torch.sigmoid(fu[users] @ fi.t())
    if pm == 0:
        return torch.sigmoid(ui)
    cos = [F.normalize(a.double()[users], dim=1) @ F.normalize(b.double(), dim=1).t() for a, b in zip(su, si)]

    def fuse(x):
        if fm == 1:
            z = torch.sigmoid(x)
            for c in cos:
                z = z * torch.sigmoid(c)
            return torch.log(z + 1e-12) - torch.log1p(z)
        for c in cos:
            x = x + c
        return torch.log(torch.sigmoid(x) + 1e-12)
    if pm == 1:
        return torch.sigmoid(fuse(ui))
    return torch.sigmoid(fuse(ui) - fuse(ui.mean(-1, keepdim=True)))


@pytest.mark.parametrize("fm", [1, 2])
@pytest.mark.parametrize("pm", [0, 1, 2])
@pytest.mark.parametrize("n_mod", [1, 3])
def test_rank_hm_sum_epilogues(dev, fm, pm, n_mod):
    from elimrec_b200 import ops
    U, I, K = 150, 1000 + 37, 20
    fu, fi, su, si = _rank_inputs(dev, U, I, n_mod, seed=fm)
    users = torch.randperm(U)[:101]
    ref = _ref_scores_fm(fu, fi, su, si, users, pm, fm).numpy()
    sun = [torch.empty_like(a, device=dev) for a in su]
    sin_ = [torch.empty_like(a, device=dev) for a in si]
    for a, b in zip(su + si, sun + sin_):
        ops.row_normalize(a.to(dev), b)
    t = ops.rank_tables(U, I, pm + 4 * fm, fu.to(dev), fi.to(dev), sun if pm else [], sin_ if pm else [])
    eu = users.to(dev).int()
    mean = torch.empty(eu.numel(), device=dev)
    ops.rank_rowmean(t, eu, mean)
    sc = torch.empty(eu.numel(), I, device=dev)
    ops.rank_scores(t, eu, mean, sc)
    assert rel_err(sc, ref) < FP32_TOL
    ptr = torch.zeros(U + 1, dtype=torch.int64, device=dev)
    idx = torch.empty(eu.numel(), K, dtype=torch.int32, device=dev)
    val = torch.empty(eu.numel(), K, device=dev)
    ops.rank_topk(t, eu, mean, ptr, torch.zeros(1, dtype=torch.int32, device=dev), K, idx, val)
    # the fused top-K returns the K largest of the reference scores, lowest index first among (near-)equals
    from gpu_util import topk_sets_match
    exact, explained, bad = topk_sets_match(idx.cpu().numpy(), ref.astype(np.float32), [[] for _ in range(eu.numel())], K, tol=2e-6)
    assert bad == 0 and exact >= 0.9 * eu.numel(), (exact, explained, bad)


def test_rank_mode_validation(dev):
    from elimrec_b200 import ops
    from elimrec_b200._lib import ElimrecError
    fu, fi = torch.zeros(4, 64, device=dev), torch.zeros(4, 64, device=dev)
    out = torch.empty(4, device=dev)
    with pytest.raises(ElimrecError):
        ops.rank_rowmean(ops.rank_tables(4, 4, 3, fu, fi, [], []), torch.arange(4, device=dev).int(), out)     # predict mode 3
    with pytest.raises(ElimrecError):
        ops.rank_rowmean(ops.rank_tables(4, 4, 2 + 4 * 3, fu, fi, [], []), torch.arange(4, device=dev).int(), out)


def test_word_graph_scatter_mean(dev):
    """scatter-mean of word embeddings (models/EliMRec.py:376-378) as a 128-wide CSR SpMM, and its transpose."""
    from elimrec_b200 import ops
    from elimrec_b200.graph import CsrHalf
    rng = np.random.default_rng(0)
    I, V = 3000, 11574
    cnt = rng.integers(1, 40, size=I)
    item = np.repeat(np.arange(I), cnt)
    word = rng.integers(0, V, size=item.size)
    word[:500] = 7                                     # one very frequent word: a split row in the transposed graph
    c = np.bincount(item, minlength=I).astype(np.float32)
    m = sp.csr_matrix(((1.0 / c[item]).astype(np.float32), (item, word)), shape=(I, V))
    m.sum_duplicates(); m.sort_indices()
    mt = m.T.tocsr(); mt.sort_indices()
    W = torch.randn(V, 128, device=dev)
    T = torch.empty(I, 128, device=dev)
    ops.spmm(CsrHalf(m.indptr, m.indices, m.data, V, dev), W, T, 128)
    idx = torch.from_numpy(item).to(dev)
    ref = torch.zeros(I, 128, dtype=torch.float64, device=dev).index_add_(0, idx, W.double()[torch.from_numpy(word).to(dev)])
    ref /= torch.from_numpy(c).to(dev).double().unsqueeze(1)
    assert rel_err(T, ref) < FP32_TOL
    dT = torch.randn(I, 128, device=dev)
    dW = torch.full((V, 128), float("nan"), device=dev)
    ops.spmm(CsrHalf(mt.indptr, mt.indices, mt.data, I, dev), dT, dW, 128)
    ref = torch.from_numpy(mt.astype(np.float64) @ dT.double().cpu().numpy())
    assert rel_err(dW, ref) < FP32_TOL


@pytest.mark.parametrize("width", [64, 256])
@pytest.mark.parametrize("masks", ["none", "col", "row+col"])
def test_spmm_additive_epilogue(dev, width, masks):
    """elimrec_spmm_masked addend: Y[row] += G[row] on the rows add_mask marks, for whole rows, split rows, every mask mix;
    the addend slab holds NaN outside the marked rows (it is only valid there)."""
    from elimrec_b200 import ops
    from elimrec_b200.graph import BipartiteGraph
    from test_gpu_kernels import _rand_graph
    U, I = 700, 500
    m = _rand_graph(U, I, 6000, [(3, 480), (10, 130), (11, 65)], seed=width)
    g = BipartiteGraph(m, dev, seg_len=8 if width == 64 else 64)
    gen = torch.Generator(device="cpu").manual_seed(1)
    for half, n_in in ((g.ui, I), (g.iu, U)):
        X = torch.randn(n_in, width, generator=gen).to(dev)
        add_mask = (torch.rand(half.n_rows, generator=gen) < 0.3).to(torch.uint8).to(dev)
        add_mask[3] = 1                                          # a split row gets the addend too
        G = torch.full((half.n_rows, width + 64), float("nan"), device=dev)
        G[add_mask.bool(), :width] = torch.randn(int(add_mask.sum()), width, generator=gen).to(dev)
        col_mask = (torch.rand(n_in, generator=gen) < 0.5).to(torch.uint8).to(dev) if "col" in masks else None
        row_mask = (torch.rand(half.n_rows, generator=gen) < 0.4).to(torch.uint8).to(dev) if "row" in masks else None
        kw = dict(row_mask=row_mask, col_mask=col_mask)
        Y0 = torch.full((half.n_rows, width), 7.0, device=dev)
        if col_mask is None and row_mask is None:
            ops.spmm(half, X, Y0, width)
        else:
            ops.spmm(half, X, Y0, width, **kw)
        Y1 = torch.full((half.n_rows, width), 7.0, device=dev)
        ops.spmm(half, X, Y1, width, addend=G, add_mask=add_mask, **kw)
        want = Y0.clone()
        sel = add_mask.bool() if row_mask is None else (add_mask.bool() & row_mask.bool())
        want[sel] += G[sel, :width]
        assert not torch.isnan(Y1).any()
        assert torch.equal(Y1, want)
