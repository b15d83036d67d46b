"""-m gpu: oracle parity at BASELINE.json's FULL sizes with the bench's DEFAULT configuration.

For the Tiktok, Kwai and Movielens shapes (``elimrec_b200.synth.SHAPES``): the CUDA path exactly as ``bench.py`` configures
it (linear schedule, 3xTF32 tensor-core GEMMs, tensor-core evaluator), its TF32 and exact-FFMA variants and the round-1
row-sparse schedule, against ``oracle.ref_model.OracleEliMRec`` / ``oracle.ref_eval`` (the CPU restatement pinned to the reference by
tests/golden) on the same weights and the same 2048-triple batch:

  * loss and every gradient: the 1e-5 class for the default (3xTF32) and the exact path, 1e-3 norm-wise with
    ``proj_precision='tf32'`` (models/EliMRec.py:115-142 and its autograd);
  * top-20 index sets of >= 1024 users: exact / explained by a near-tie within the score tolerance / unexplained, gated at
    zero unexplained (uni_evaluator.py:104-203, evaluate.h:23-64);
  * Recall@20 / NDCG@20 (and Precision@20) identical to 4 decimals.

The oracle costs about 3-6 s of CPU per shape."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from gpu_util import build_model, rel_err, topk_sets_match

N_EVAL = 1024


@pytest.fixture(scope="module", params=["tiktok", "kwai", "movielens"])
def setup(request):
    from elimrec_b200 import synth
    from elimrec_b200.data import Dataset
    from oracle.ref_model import OracleEliMRec
    shape = request.param
    inter, feats = synth.make_shape(shape)
    name = "kwai" if shape == "kwai" else shape + "shape"
    ds = Dataset(None, interactions=inter, features=feats, name=name)
    torch.manual_seed(2022)
    model = build_model(ds, None, dataset_name=name, alpha=0.5, proj_precision="auto")      # bench.py's defaults
    assert model.linear and model.lazy_tables and model.fuse_precision == "x3" and model.proj_precision == "x3"
    params = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    mods = "v" if shape == "kwai" else "vat"
    orc = OracleEliMRec(params, {m: getattr(ds, f"{m}_feat") for m in mods}, ds.train_matrix, ds.num_users, ds.num_items,
                        kwai=(shape == "kwai"), alpha=0.5)
    rng = np.random.default_rng(7)
    u = rng.integers(0, ds.num_users, 2048)
    tm = ds.train_matrix
    p = tm.indices[tm.indptr[u]]
    n = rng.integers(0, ds.num_items, 2048)
    lo = orc.bpr_loss(u, p, n)
    og = {k: v.detach().clone() for k, v in orc.grads(lo).items()}
    return dict(shape=shape, ds=ds, name=name, model=model, params=params, orc=orc, batch=(u, p, n), loss=float(lo), grads=og)


def _check_step(model, s, tol, bias_tol):
    for q in model.parameters():
        q.grad = None
    loss = model.bpr_loss(*s["batch"])
    loss.backward()
    assert abs(float(loss) - s["loss"]) < tol * abs(s["loss"]), (float(loss), s["loss"])
    worst = {}
    for name, prm in model.named_parameters():
        if name in s["grads"]:
            worst[name] = rel_err(prm.grad, s["grads"][name])
            # bias gradients are column sums with heavy cancellation: the summation order shows a few 1e-5 even in exact fp32
            assert worst[name] < (bias_tol if name.endswith("bias") else tol) * (2 if tol > 1e-4 else 1), (name, worst[name])
    return worst


def test_default_config_loss_and_gradients(setup):
    """bench default (linear schedule, 3xTF32 GEMMs): the 1e-5 class (2e-5 norm-wise, as everywhere in tests/; bias gradients
    are column sums with heavy cancellation: 1e-4)"""
    worst = _check_step(setup["model"], setup, 2e-5, 1e-4)
    assert len(worst) >= 8
    print(f"[{setup['shape']}] worst gradient error {max(worst.values()):.2e} ({max(worst, key=worst.get)})")
    assert rel_err(setup["model"].all_users, setup["orc"].cache["users"]) < 2e-5
    assert rel_err(setup["model"].all_items, setup["orc"].cache["items"]) < 2e-5


@pytest.mark.parametrize("cfg", [dict(proj_precision="fp32"), dict(proj_precision="fp32", linear_schedule=False),
                                 dict(proj_precision="tf32"), dict(proj_precision="tf32", linear_schedule=False)],
                         ids=["fp32-linear", "fp32-rowsparse", "tf32-linear", "tf32-rowsparse"])
def test_other_configs_loss_and_gradients(setup, cfg):
    """exact-FFMA GEMMs: 1e-5 class; TF32 GEMMs: 1e-3; both schedules"""
    s = setup
    if s["shape"] == "movielens" and cfg == dict(proj_precision="fp32", linear_schedule=False):
        pytest.skip("FFMA projections over 2048-d features of every item, twice per step: covered at the other shapes")
    model = build_model(s["ds"], s["params"], dataset_name=s["name"], alpha=0.5, **cfg)
    assert model.linear == cfg.get("linear_schedule", True)
    tol, btol = (1e-3, 1e-3) if cfg["proj_precision"] == "tf32" else (2e-5, 1e-4)
    _check_step(model, s, tol, btol)
    assert rel_err(model.all_users, s["orc"].cache["users"]) < tol
    assert rel_err(model.all_items, s["orc"].cache["items"]) < tol


@pytest.mark.parametrize("pt", ["TIE", "TE"])
def test_default_config_topk_sets_and_metrics(setup, pt):
    from oracle import ref_eval
    s = setup
    model, orc, ds = s["model"], s["orc"], s["ds"]
    if model._all_users is None and not model._tables_pending:
        model.bpr_loss(*s["batch"])
    model.eval()
    model.predict_type = pt
    test = ds.get_user_test_dict()
    train = ds.get_user_train_dict()
    users = list(test.keys())[:N_EVAL]
    ev = model.test_evaluator.evaluator
    res, _ = ev.evaluate(model, test_users=users)
    idx = ev.last_topk[0].cpu().numpy()
    ref = np.concatenate([orc.predict(users[i:i + 128], pt).numpy() for i in range(0, len(users), 128)])
    # measured score error of this configuration on a few users: the near-tie tolerance of the triage is twice that, and must
    # itself stay inside the fp32 class
    got = model.predict(users[:32]).numpy()
    score_err = float(np.abs(got - ref[:32]).max())
    assert score_err < 1e-5 * float(np.abs(ref[:32]).max()), score_err
    tol = max(2e-6, 2 * score_err)
    exact, explained, bad = topk_sets_match(idx, ref, [train.get(x, []) for x in users], 20, tol=tol)
    print(f"[{s['shape']} {pt}] score error {score_err:.2e}; top-20 sets: {exact} exact, {explained} near-tie (<= {tol:.1e}), "
          f"{bad} unexplained of {len(users)}")
    assert bad == 0 and exact + explained == len(users), (pt, exact, explained, bad, score_err)
    want, _ = ref_eval.evaluate(lambda us: orc.predict(us, pt).numpy(), train, {x: test[x] for x in users}, top_k=[20],
                                batch_size=128)
    print(f"[{s['shape']} {pt}] P/R/N@20 = {['%.4f' % a for a in res]} oracle {['%.4f' % a for a in want]}")
    assert np.abs(res - want).max() < 5e-5, (pt, res, want)         # identical to 4 decimals
