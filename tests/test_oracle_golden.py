"""Pins the oracle (oracle/) against outputs of the reference itself (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from oracle import ref_eval, ref_model, ref_sampler
from helpers import csr_from_golden, dict_from_csr, golden_feats, golden_params


def _oracle(g, **kw):
    kwai = g["_name"] == "kwai"
    return ref_model.OracleEliMRec(golden_params(g), golden_feats(g), csr_from_golden(g, "train"),
                                   int(g["num_users"]), int(g["num_items"]), kwai=kwai, alpha=0.5, **kw)


def test_adjacency_bit_exact(golden):
    r, c, v = ref_model.norm_adj_coo(csr_from_golden(golden, "train"), int(golden["num_users"]), int(golden["num_items"]))
    assert np.array_equal(r, golden["adj_row"]) and np.array_equal(c, golden["adj_col"])
    assert np.array_equal(v.view(np.uint32), golden["adj_val"].view(np.uint32))


def test_sampler_triples_bit_exact(golden):
    train = dict_from_csr(csr_from_golden(golden, "train"))
    ref_sampler.srand(1)
    u, p, n = ref_sampler.sample_epoch(train, int(golden["num_items"]))
    assert np.array_equal(u, golden["epoch_users"]) and np.array_equal(p, golden["epoch_pos"])
    assert np.array_equal(n, golden["epoch_neg"])
    u, p, n = ref_sampler.sample_epoch(train, int(golden["num_items"]))  # stream continues
    assert np.array_equal(u, golden["epoch2_users"]) and np.array_equal(n, golden["epoch2_neg"])
    # batches: same numpy permutation
    ref_sampler.srand(1)
    np.random.seed(2022)
    u, p, n = ref_sampler.sample_epoch(train, int(golden["num_items"]))
    bs = list(ref_sampler.epoch_batches(u, p, n, 128))
    assert len(bs) == int(golden["n_batches"])
    for i in range(min(4, len(bs))):
        assert np.array_equal(bs[i][0], golden[f"batch{i}_users"])
        assert np.array_equal(bs[i][1], golden[f"batch{i}_pos"])
        assert np.array_equal(bs[i][2], golden[f"batch{i}_neg"])


def test_loss_grads_tables(golden):
    o = _oracle(golden)
    loss = o.bpr_loss(golden["batch0_users"], golden["batch0_pos"], golden["batch0_neg"])
    assert abs(float(loss) - float(golden["loss0"])) < 1e-6
    gr = o.grads(loss)
    for k, v in golden.items():
        if k.startswith("grad0/"):
            np.testing.assert_allclose(gr[k[6:]].numpy(), v, rtol=1e-5, atol=1e-8, err_msg=k)
    np.testing.assert_allclose(o.cache["users"].detach().numpy(), golden["all_users"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(o.cache["items"].detach().numpy(), golden["all_items"], rtol=1e-6, atol=1e-7)
    for m in o.mods:
        np.testing.assert_allclose(o.cache["s"][m][1].detach().numpy(), golden[f"s_item_{m}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(o.cache["light"][m].detach().numpy(), golden[f"light_{m}"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("pt", ["TIE", "TE", "normal"])
def test_predict_and_evaluate(golden, pt):
    o = _oracle(golden)
    o.bpr_loss(golden["batch0_users"], golden["batch0_pos"], golden["batch0_neg"])
    sc = o.predict(golden["predict_users"], pt).numpy()
    np.testing.assert_allclose(sc, golden[f"predict_{pt}"], rtol=1e-6, atol=1e-7)
    train = dict_from_csr(csr_from_golden(golden, "train"))
    for split, key in (("valid", "evaluate"), ("test", "test")):
        truth = dict_from_csr(csr_from_golden(golden, split))
        res, buf = ref_eval.evaluate(lambda us: o.predict(us, pt).numpy(), train, truth, top_k=[20], batch_size=16)
        np.testing.assert_allclose(res, golden[f"{key}_{pt}"], rtol=0, atol=1e-6)


def test_metric_rows_bit_exact(golden):
    truth = dict_from_csr(csr_from_golden(golden, "valid"))
    users = golden["predict_users"].tolist()
    sc = golden["masked_scores_TIE"]
    tk = ref_eval.topk_lowest_index(sc, 20)
    rows = ref_eval.metric_rows(tk, [truth[u] for u in users], [1, 2, 4], 20)
    assert np.array_equal(rows.view(np.uint32), golden["metric_rows_TIE"].view(np.uint32))
    if ref_eval.ref_cpp_available():  # the reference's own C++, compiled from /root/reference
        rows2 = ref_eval.ref_cpp_metric_rows(sc.copy(), [truth[u] for u in users], [1, 2, 4], 20)
        assert np.array_equal(rows2.view(np.uint32), golden["metric_rows_TIE"].view(np.uint32))


def test_three_adam_steps(golden):
    o = _oracle(golden)
    opt = torch.optim.Adam(list(o.p.values()), lr=1e-3, weight_decay=1e-4)
    losses = []
    for i in range(3):
        loss = o.bpr_loss(golden[f"batch{i}_users"], golden[f"batch{i}_pos"], golden[f"batch{i}_neg"])
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    np.testing.assert_allclose(losses, golden["losses"], rtol=1e-6)
    for k, v in golden.items():
        if k.startswith("sd3/"):
            np.testing.assert_allclose(o.p[k[4:]].detach().numpy(), v, rtol=1e-5, atol=1e-7, err_msg=k)
