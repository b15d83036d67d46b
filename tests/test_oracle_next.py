"""Pins the oracle's SURVEY.md 8(f) rows (adj_type, mm_fusion_mode, s_fusion_mode, candidate negatives, MAP/MRR, the
literal tiktok branch) against outputs of the reference itself (tests/golden/next.npz, tiktok.npz)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ref_eval, ref_model
from helpers import csr_from_golden, dict_from_csr, golden_feats, golden_params


@pytest.fixture(scope="module")
def base():
    g = load_golden("generic")
    g["_name"] = "generic"
    return g


@pytest.fixture(scope="module")
def nxt():
    return load_golden("next")


def _oracle(base, nxt, pre, **kw):
    return ref_model.OracleEliMRec(golden_params(nxt, f"{pre}/sd0/"), golden_feats(base), csr_from_golden(base, "train"),
                                   int(base["num_users"]), int(base["num_items"]), alpha=0.5, **kw)


def _check_run(o, nxt, pre, steps=3):
    opt = torch.optim.Adam(list(o.p.values()), lr=1e-3, weight_decay=1e-4)
    losses = []
    for i in range(steps):
        loss = o.bpr_loss(nxt[f"{pre}/batch{i}_users"], nxt[f"{pre}/batch{i}_pos"], nxt[f"{pre}/batch{i}_neg"])
        if i == 0:
            gr = o.grads(loss)
            for k, v in nxt.items():
                if k.startswith(f"{pre}/grad0/"):
                    np.testing.assert_allclose(gr[k[len(pre) + 7:]].numpy(), v, rtol=2e-5, atol=1e-8, err_msg=k)
        opt.zero_grad()
        loss.backward(retain_graph=True)
        opt.step()
        losses.append(float(loss))
    np.testing.assert_allclose(losses, nxt[f"{pre}/losses"], rtol=1e-6)
    for k, v in nxt.items():
        if k.startswith(f"{pre}/sd3/"):
            np.testing.assert_allclose(o.p[k[len(pre) + 5:]].detach().numpy(), v, rtol=1e-5, atol=1e-7, err_msg=k)


@pytest.mark.parametrize("adj", ["plain", "norm", "gcmc", "mean"])
def test_adjacency_types(base, nxt, adj):
    r, c, v = ref_model.norm_adj_coo(csr_from_golden(base, "train"), int(base["num_users"]), int(base["num_items"]), adj)
    assert np.array_equal(r, nxt[f"adj_{adj}/row"]) and np.array_equal(c, nxt[f"adj_{adj}/col"])
    assert np.array_equal(v.view(np.uint32), nxt[f"adj_{adj}/val"].view(np.uint32))
    _check_run(_oracle(base, nxt, f"adj_{adj}", adj_type=adj), nxt, f"adj_{adj}")


def test_mm_fusion_mean(base, nxt):
    o = _oracle(base, nxt, "mm_mean", mm_fusion_mode="mean")
    assert o.p["embedding_user_after_GCN.weight"].shape == (64, 64)
    o.bpr_loss(nxt["mm_mean/batch0_users"], nxt["mm_mean/batch0_pos"], nxt["mm_mean/batch0_neg"])
    np.testing.assert_allclose(o.cache["users"].detach().numpy(), nxt["mm_mean/all_users"], rtol=1e-6, atol=1e-7)
    for pt in ("TIE", "TE"):
        np.testing.assert_allclose(o.predict(nxt["mm_mean/predict_users"], pt).numpy(), nxt[f"mm_mean/predict_{pt}"],
                                   rtol=1e-6, atol=1e-7)
    _check_run(_oracle(base, nxt, "mm_mean", mm_fusion_mode="mean"), nxt, "mm_mean")


@pytest.mark.parametrize("pre,fm,modality", [("s_hm", "hm", "vat"), ("s_sum", "sum", "vat"), ("s_hm_va", "hm", "va")])
def test_score_fusion_modes(base, nxt, pre, fm, modality):
    o = _oracle(base, nxt, pre, s_fusion_mode=fm, modality=modality)
    loss = o.bpr_loss(nxt[f"{pre}/batch0_users"], nxt[f"{pre}/batch0_pos"], nxt[f"{pre}/batch0_neg"])
    assert abs(float(loss) - float(nxt[f"{pre}/loss0"])) < 1e-6
    train = dict_from_csr(csr_from_golden(base, "train"))
    valid = dict_from_csr(csr_from_golden(base, "valid"))
    for pt in ("TIE", "TE", "normal"):
        if f"{pre}/predict_{pt}" not in nxt:
            continue
        np.testing.assert_allclose(o.predict(nxt[f"{pre}/predict_users"], pt).numpy(), nxt[f"{pre}/predict_{pt}"],
                                   rtol=2e-6, atol=1e-6)
        res, _ = ref_eval.evaluate(lambda us: o.predict(us, pt).numpy(), train, valid, top_k=[20], batch_size=16)
        np.testing.assert_allclose(res, nxt[f"{pre}/evaluate_{pt}"], atol=1e-6)


def test_candidate_negatives_and_all_metrics(base, nxt):
    o = _oracle(base, nxt, "cand")
    o.bpr_loss(nxt["cand/batch0_users"], nxt["cand/batch0_pos"], nxt["cand/batch0_neg"])
    train = dict_from_csr(csr_from_golden(base, "train"))
    test = dict_from_csr(csr_from_golden(base, "test"))
    assert list(test.keys()) == nxt["cand/users"].tolist()
    neg = {u: n.tolist() for u, n in zip(nxt["cand/users"].tolist(), nxt["cand/neg"])}
    allm = ("Precision", "Recall", "MAP", "NDCG", "MRR")
    pf = lambda us: o.predict(us, "TIE").numpy()
    res, _ = ref_eval.evaluate(pf, train, test, metrics=allm, top_k=[5, 20], batch_size=16, user_neg=neg)
    np.testing.assert_allclose(res, nxt["cand/result"], atol=1e-6)
    res, _ = ref_eval.evaluate(pf, train, test, metrics=allm, top_k=[5, 20], batch_size=16)
    np.testing.assert_allclose(res, nxt["allmetrics/result"], atol=1e-6)
    res, _ = ref_eval.evaluate(pf, train, test, metrics=("MAP", "MRR"), top_k=7, batch_size=16)
    np.testing.assert_allclose(res, nxt["topk_int/result"], atol=1e-6)


def test_arg_topk(nxt):
    """util/cython/include/arg_topk.h:15-45 (partial_sort_copy on indices, '>' comparator)."""
    sc = nxt["arg_topk/scores"]
    got = ref_eval.topk_lowest_index(sc, 10)
    ref = nxt["arg_topk/idx"]
    same = got == ref
    # the reference's order among EQUAL scores is unspecified; everywhere else the indices agree
    assert np.array_equal(np.take_along_axis(sc, got, 1), np.take_along_axis(sc, ref, 1))
    assert same.mean() > 0.95


# ---- the literal tiktok branch ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def tk():
    return load_golden("tiktok")


def tiktok_setup(tk):
    """Dataset arrays + full-size word_embedding (rows the fixture does not hold stay zero)."""
    from elimrec_b200 import synth
    from elimrec_b200.data import Dataset
    inter = synth.Interactions(int(tk["num_users"]), int(tk["num_items"]), tk["raw_train"], tk["raw_valid"], tk["raw_test"])
    ds = Dataset(None, interactions=inter, features=[tk["raw_feat_v"], tk["raw_feat_a"], None], name="tiktok",
                 words=tk["raw_words"])

    def params(prefix):
        p = golden_params(tk, prefix)
        full = np.zeros((11574, 128), dtype=np.float32)
        full[tk["word_rows"]] = p["word_embedding.weight"]
        p["word_embedding.weight"] = full
        return p
    return ds, params


def test_tiktok_branch(tk):
    ds, params = tiktok_setup(tk)
    assert np.array_equal(ds.words_tensor.numpy(), tk["words_tensor"])
    feats = {"v": ds.v_feat.numpy(), "a": ds.a_feat.numpy()}
    o = ref_model.OracleEliMRec(params("tiktok/sd0/"), feats, ds.train_matrix, ds.num_users, ds.num_items, alpha=0.5,
                                words=tk["words_tensor"])
    np.testing.assert_allclose(o.feat["t"].detach().numpy(), tk["t_feat"], rtol=1e-6, atol=1e-8)
    opt = torch.optim.Adam(list(o.p.values()), lr=1e-3, weight_decay=1e-4)
    rows = tk["word_rows"]
    losses = []
    for i in range(3):
        loss = o.bpr_loss(tk[f"tiktok/batch{i}_users"], tk[f"tiktok/batch{i}_pos"], tk[f"tiktok/batch{i}_neg"])
        if i == 0:
            gr = o.grads(loss)
            for k, v in tk.items():
                if k.startswith("tiktok/grad0/"):
                    got = gr[k[13:]].numpy()
                    got = got[rows] if k.endswith("word_embedding.weight") else got
                    np.testing.assert_allclose(got, v, rtol=2e-5, atol=1e-8, err_msg=k)
            np.testing.assert_allclose(o.predict(tk["tiktok/predict_users"], "TIE").numpy(), tk["tiktok/predict_TIE"],
                                       rtol=1e-6, atol=1e-7)
        opt.zero_grad()
        loss.backward(retain_graph=True)
        opt.step()
        losses.append(float(loss))
    np.testing.assert_allclose(losses, tk["tiktok/losses"], rtol=1e-6)
    for k, v in tk.items():
        if k.startswith("tiktok/sd3/"):
            got = o.p[k[11:]].detach().numpy()
            got = got[rows] if k.endswith("word_embedding.weight") else got
            np.testing.assert_allclose(got, v, rtol=1e-5, atol=1e-7, err_msg=k)
