"""Host-side logic against the golden vectors: id remap, CSR halves (bit-exact), segments, compat sampler."""
import ctypes

import numpy as np
import pytest
import torch

from elimrec_b200 import graph as G
from elimrec_b200.sampler import CompatRng, PairwiseSamplerV2
from helpers import csr_from_golden, dict_from_csr, golden_dataset, golden_feats


def test_dataset_remap_matches_reference(golden):
    ds = golden_dataset(golden)
    assert ds.num_users == int(golden["num_users"]) and ds.num_items == int(golden["num_items"])
    for split in ("train", "valid", "test"):
        m = getattr(ds, f"{split}_matrix")
        assert np.array_equal(m.indptr, golden[f"{split}_indptr"]) and np.array_equal(m.indices, golden[f"{split}_indices"])
    assert np.array_equal(np.array(list(ds.itemids.keys())), golden["itemids_raw_in_order"])
    feats = golden_feats(golden)
    for m, f in feats.items():
        assert np.array_equal(getattr(ds, f"{m}_feat").numpy(), f)
    assert ds.get_user_train_dict() == dict_from_csr(csr_from_golden(golden, "train"))


def test_adjacency_halves_bit_exact(golden):
    U, I = int(golden["num_users"]), int(golden["num_items"])
    h = G.normalized_halves(csr_from_golden(golden, "train"))
    (pu, iu, vu), (pi, ii, vi) = h["ui"], h["iu"]
    assert h["ui_t"] is None and h["self_u"] is None
    ru = np.repeat(np.arange(U), np.diff(pu))
    ri = np.repeat(np.arange(I), np.diff(pi)) + U
    row = np.concatenate([ru, ri])
    col = np.concatenate([iu + U, ii])
    val = np.concatenate([vu, vi])
    assert np.array_equal(row, golden["adj_row"]) and np.array_equal(col, golden["adj_col"])
    assert np.array_equal(val.view(np.uint32), golden["adj_val"].view(np.uint32))


@pytest.mark.parametrize("adj", ["plain", "norm", "gcmc", "mean"])
def test_other_adj_types_bit_exact(adj):
    """models/EliMRec.py:329-352: every adj_type against the COO the reference itself built (next.npz), plus A_hat^T."""
    from conftest import load_golden
    base, nxt = load_golden("generic"), load_golden("next")
    U, I = int(base["num_users"]), int(base["num_items"])
    g = G.BipartiteGraph(csr_from_golden(base, "train"), "cpu", adj)
    row, col, val = g.as_coo()
    # scipy leaves the columns of a row unsorted after `adj + sp.eye` ('norm'): compare in canonical (row, col) order
    rr, rc, rv = nxt[f"adj_{adj}/row"], nxt[f"adj_{adj}/col"], nxt[f"adj_{adj}/val"]
    o = np.lexsort((rc, rr))
    assert np.array_equal(row, rr[o]) and np.array_equal(col, rc[o])
    assert np.array_equal(val.view(np.uint32), rv[o].view(np.uint32))
    assert g.self_loops == (adj in ("norm", "mean")) and g.symmetric == (adj == "plain")
    # the transposed blocks really are the transpose
    import scipy.sparse as sp
    A = sp.csr_matrix((val, (row, col)), shape=(U + I, U + I))
    At = A.T.tocsr()
    ui_t = sp.csr_matrix((g.ui_t.vals_host, g.ui_t.indices_host, g.ui_t.indptr_host), shape=(U, I))
    iu_t = sp.csr_matrix((g.iu_t.vals_host, g.iu_t.indices_host, g.iu_t.indptr_host), shape=(I, U))
    assert abs(At[:U, U:] - ui_t).max() == 0 and abs(At[U:, :U] - iu_t).max() == 0


def test_grouped_evaluator_groups():
    """evaluator/grouped_evaluator.py:61-75: (lo, hi] groups by number of training items, larger users dropped."""
    from elimrec_b200.evaluator import GroupedEvaluator, ProxyEvaluator
    train = {u: list(range(n)) for u, n in enumerate([1, 3, 3, 4, 7, 9, 30])}
    test = {u: [50] for u in train}
    ge = GroupedEvaluator(train, test, None, metric=["Recall"], group_view=[3, 8], top_k=[5])
    assert list(ge.grouped_user.keys()) == ["(0,3]:".ljust(12), "(3,8]:".ljust(12)]
    assert list(ge.grouped_user.values()) == [[0, 1, 2], [3, 4]]
    assert ge.metrics_info() == "metrics:\t" + "Recall@5".ljust(12)
    with pytest.raises(ValueError):
        GroupedEvaluator(train, {6: [50]}, None, metric=["Recall"], group_view=[3, 8], top_k=[5])
    with pytest.raises(TypeError):
        GroupedEvaluator(train, test, None, group_view=(3, 8))
    assert isinstance(ProxyEvaluator(None, train, test, None, metric=["Recall"], group_view=[3, 8], top_k=5).evaluator,
                      GroupedEvaluator)


@pytest.mark.parametrize("seg_len", [4, 64])
def test_segments_partition_edges(seg_len):
    rng = np.random.default_rng(0)
    deg = rng.integers(0, 40, size=200)
    deg[7] = 1000
    deg[50] = 0
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    seg, heavy, n_hseg = G.build_segments(indptr, seg_len)
    covered = np.zeros(indptr[-1], dtype=np.int32)
    for r, b, e, h in seg:
        assert indptr[r] <= b <= e <= indptr[r + 1] and e - b <= seg_len
        covered[b:e] += 1
    assert (covered == 1).all()
    assert set(seg[:, 0].tolist()) == set(range(200))           # empty rows still get a (zero) segment
    hs = seg[:n_hseg]
    assert n_hseg % 8 == 0 and (hs[:, 3] >= 0).all() and (seg[n_hseg:, 3] == -1).all()
    for h, (first_cta, n_cta) in enumerate(heavy):
        blk = seg[8 * first_cta:8 * (first_cta + n_cta)]
        assert (blk[:, 3] == h).all() and len(set(blk[:, 0])) == 1   # a CTA of 8 warps never mixes rows
    if heavy.shape[0] > 1:
        d = [indptr[seg[8 * f, 0] + 1] - indptr[seg[8 * f, 0]] for f, _ in heavy]
        assert d == sorted(d, reverse=True)
    # row-range restricted lists (multi-GPU shards)
    seg2, _, _ = G.build_segments(indptr, seg_len, 50, 120)
    assert set(seg2[:, 0].tolist()) == set(range(50, 120))


def test_compat_rng_is_glibc_rand():
    libc = ctypes.CDLL("libc.so.6")
    for seed in (1, 7, 2022, 0):
        libc.srand(seed)
        r = CompatRng(seed)
        assert [r.rand() for _ in range(1000)] == [libc.rand() for _ in range(1000)]


def test_compat_sampler_bit_exact(golden):
    ds = golden_dataset(golden)
    s = PairwiseSamplerV2(ds, neg_num=1, batch_size=128, shuffle=True)
    assert s.num_trainings == golden["epoch_users"].size and len(s) == int(golden["n_batches"])
    u, p, n = s.sample_epoch_host()
    assert np.array_equal(u, golden["epoch_users"]) and np.array_equal(p, golden["epoch_pos"])
    assert np.array_equal(n, golden["epoch_neg"])
    u, p, n = s.sample_epoch_host()
    assert np.array_equal(u, golden["epoch2_users"]) and np.array_equal(p, golden["epoch2_pos"])
    # iterator: numpy permutation from the global RNG, as DataIterator does
    s = PairwiseSamplerV2(ds, neg_num=1, batch_size=128, shuffle=True)
    np.random.seed(2022)
    bs = list(s)
    assert len(bs) == int(golden["n_batches"])
    for i in range(min(4, len(bs))):
        for j, k in enumerate(("users", "pos", "neg")):
            assert np.array_equal(bs[i][j], golden[f"batch{i}_{k}"])
    assert bs[-1][0].size == s.num_trainings - 128 * (len(bs) - 1)  # last batch short, not dropped


def test_compat_sampler_invariants_medium():
    from elimrec_b200 import synth
    from elimrec_b200.data import Dataset
    inter, feats = synth.make_shape("small")
    ds = Dataset(None, interactions=inter, features=feats, name="small")
    s = PairwiseSamplerV2(ds, batch_size=2048)
    u, p, n = s.sample_epoch_host()
    tm = ds.train_matrix
    assert np.asarray(tm[u, p]).all()            # positives are train items
    assert not np.asarray(tm[u, n]).any()        # negatives never are
    # cross-check against the libc-driven oracle restatement
    from oracle import ref_sampler
    ref_sampler.srand(1)
    ou, op, on = ref_sampler.sample_epoch(ds.get_user_train_dict(), ds.num_items)
    assert np.array_equal(u, ou) and np.array_equal(p, op) and np.array_equal(n, on)


def test_eval_csr_helper():
    from elimrec_b200.evaluator import _dict_to_csr
    ptr, flat = _dict_to_csr({3: [5, 1], 9: [2]}, [9, 3, 4])
    assert ptr.tolist() == [0, 1, 3, 3] and flat.tolist() == [2, 1, 5]


def test_dataset_from_reference_file_layout(golden, tmp_path):
    """Dataset(conf) reads the reference's on-disk layout (data/dataset.py:164-185,207-212): CSV splits + generic .npy
    features / kwai_feat_v.pt; same id remap, CSRs and feature rows as the reference's own Dataset (golden)."""
    from elimrec_b200 import synth
    from elimrec_b200.data import Config, Dataset
    kwai = golden["_name"] == "kwai"
    name = "kwai" if kwai else "synthg"
    inter = synth.Interactions(int(golden["num_users"]), int(golden["num_items"]), golden["raw_train"], golden["raw_valid"],
                               golden["raw_test"])
    feats = [golden.get(f"raw_feat_{m}") for m in "vat"]
    synth.write_reference_files(str(tmp_path), name, inter, feats)
    ds = Dataset(Config(**{"data.input.path": str(tmp_path), "data.input.dataset": name}))
    assert ds.num_users == int(golden["num_users"]) and ds.num_items == int(golden["num_items"])
    for split in ("train", "valid", "test"):
        m = getattr(ds, f"{split}_matrix")
        assert np.array_equal(m.indptr, golden[f"{split}_indptr"]) and np.array_equal(m.indices, golden[f"{split}_indices"])
    for m, f in golden_feats(golden).items():
        assert np.array_equal(getattr(ds, f"{m}_feat").numpy(), f)
    u, i = ds.get_train_interactions()
    assert len(u) == ds.train_matrix.nnz == len(i)


def test_tiktok_file_layout(tmp_path):
    """the literal 'tiktok' layout: tiktok_{visual,audio,textual}_feat.pt; word pairs remapped like dataset.py:166-173"""
    from conftest import load_golden
    from elimrec_b200 import synth
    from elimrec_b200.data import Config, Dataset
    tk = load_golden("tiktok")
    inter = synth.Interactions(int(tk["num_users"]), int(tk["num_items"]), tk["raw_train"], tk["raw_valid"], tk["raw_test"])
    synth.write_tiktok_files(str(tmp_path), inter, tk["raw_feat_v"], tk["raw_feat_a"], tk["raw_words"])
    ds = Dataset(Config(**{"data.input.path": str(tmp_path), "data.input.dataset": "tiktok"}))
    assert np.array_equal(ds.words_tensor.numpy(), tk["words_tensor"]) and not hasattr(ds, "t_feat")
    assert ds.v_feat.shape == (ds.num_items, tk["raw_feat_v"].shape[1])
