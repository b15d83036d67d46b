"""Round-2 host-logic checks: the linear schedule's launch profile, gradients of heads the loss does not use, workspace
sharing across batch sizes, getEmbedding.  ``[sim]`` runs here on the torch-CPU schedule simulator, ``[cuda]`` on the B200."""
import numpy as np
import pytest
import torch

from helpers import golden_dataset, golden_params
from test_schedule_sim import TOL, _oracle_for, backend, batch, build, rel, sim  # noqa: F401  (fixtures)


def _name(golden):
    return "kwai" if golden["_name"] == "kwai" else "synthg"


def test_linear_step_propagates_64_wide_only(sim, golden):
    """linear schedule: one 64-wide propagation forward and backward, no pass over the item features, the modality blocks
    from the constant Zbar tables gathered at the instance rows (elimrec_b200/linear.py)."""
    model = build(golden_dataset(golden), golden_params(golden), _name(golden))
    assert model.linear
    model.make_optimizer()
    sim.LAUNCH_LOG.clear()
    model.train_step(*batch(golden, 0))
    log = list(sim.LAUNCH_LOG)
    L = model.n_layers
    assert not any(x.startswith(("spmm128", "spmm256", "proj_fwd", "proj_wgrad", "colsum")) for x in log), log
    assert sum(x.startswith("spmm64") for x in log) == 4 * L          # two halves per layer, forward and backward
    # both seed vectors of the backward chain in one launch; the chain adds them in the propagation launches' epilogues
    assert log.count("lin_assemble") == 1 and log.count("lin_seed2") == 1 and log.count("gather_rows") == 1
    assert "fuse_heads_x3_all" not in log
    # completing the tables for an evaluation: the masked layer(s) again (every row) + one pass over Zbar
    sim.LAUNCH_LOG.clear()
    model.all_users
    log = list(sim.LAUNCH_LOG)
    assert "fuse_heads_x3_all" in log and log.count("lin_assemble") == 1
    assert sum(x.startswith("spmm64") for x in log) == 2 * (min(2, L) if model._lin_need2 else 1)


@pytest.mark.parametrize("prec", ["x3", "tf32", "fp32", "auto"])
def test_linear_schedule_precision_modes(backend, golden, prec):
    """host paths of the three GEMM precisions of the linear schedule (packed / split weights, stacked hi-lo operands of the
    weight gradient).  The simulator computes all of them exactly; on the B200 'tf32' is gated at its own class."""
    tol = 1e-3 if (backend == "cuda" and prec == "tf32") else TOL
    model = build(golden_dataset(golden), golden_params(golden), _name(golden), proj_precision=prec)
    assert model.linear and model.proj_precision == ("x3" if prec == "auto" else prec)
    loss = model.bpr_loss(*batch(golden, 0))
    loss.backward()
    assert abs(float(loss) - float(golden["loss0"])) < tol * abs(float(golden["loss0"]))
    for name, p in model.named_parameters():
        if ("grad0/" + name) in golden:
            assert rel(p.grad, golden["grad0/" + name]) < (2 * tol if prec == "tf32" else tol), name
    assert rel(model.all_users, golden["all_users"]) < tol and rel(model.all_items, golden["all_items"]) < tol


def test_zbar_is_the_propagated_feature_mean(sim, golden):
    """Zbar_m = mean_k A_hat^k [0 ; X_m | 1] against a dense restatement"""
    model = build(golden_dataset(golden), golden_params(golden), _name(golden))
    U, I, L = model.num_users, model.num_items, model.n_layers
    r, c, v = model.graph.as_coo()
    A = torch.zeros(U + I, U + I, dtype=torch.float64)
    A[torch.as_tensor(r), torch.as_tensor(c)] = torch.as_tensor(v, dtype=torch.float64)
    for j, m in enumerate(model.mods):
        X = model._feat[m].double()
        x = torch.cat([torch.zeros(U, X.shape[1] + 1, dtype=torch.float64), torch.cat([X, torch.ones(I, 1, dtype=torch.float64)], 1)])
        acc, cur = x.clone(), x
        for _ in range(L):
            cur = A @ cur
            acc += cur
        want = acc / (L + 1)
        ko, kp = model._lin_koff[j], model._lin_Kp[j]
        got = model._zbar[:, ko:ko + kp]
        assert rel(got[:, :X.shape[1] + 1], want.numpy()) < 1e-6
        assert not got[:, X.shape[1] + 1:].any()          # alignment padding stays zero


@pytest.mark.parametrize("variant", [dict(predict_type="normal"), dict(modality="va"), dict(modality="t", linear_schedule=False)],
                         ids=lambda v: ",".join(f"{k}={x}" for k, x in v.items()))
def test_unused_heads_get_no_gradient(backend, golden, variant):
    """heads whose loss term has weight 0: the reference never calls them, .grad stays None and Adam leaves them alone
    (models/EliMRec.py:125-140)."""
    if golden["_name"] == "kwai" and "modality" in variant:
        pytest.skip("kwai forces modality='v'")
    model = build(golden_dataset(golden), golden_params(golden), _name(golden), **variant)
    dead = model._dead_params()
    mods = model.mods
    want_dead = {f"s_dense_{m}.{p}" for m in mods for p in ("weight", "bias")
                 if variant.get("predict_type") == "normal" or m not in variant.get("modality", "vat")}
    assert dead == want_dead and dead
    loss = model.bpr_loss(*batch(golden, 0))
    loss.backward()
    for n, p in model.named_parameters():
        if n in dead:
            assert p.grad is None, n
        elif n in model._param_names:
            assert p.grad is not None, n
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    model.make_optimizer(lr=1e-3, weight_decay=1e-4)
    model.train_step(*batch(golden, 1))
    for n, p in model.named_parameters():
        if n in dead:
            assert torch.equal(p.detach(), before[n]), n        # no weight-decay drift either
            assert n not in model._adam.state
        elif n in model._param_names:
            assert not torch.equal(p.detach(), before[n]), n


@pytest.mark.parametrize("linear", [True, False])
def test_short_last_batch_shares_the_workspace(sim, golden, linear):
    """PairwiseSamplerV2 has drop_last=False: the last batch of an epoch is short.  Only the 3B-row buffers are allocated
    per batch size; slabs / tables / gradient buffers are shared, and both sizes give the oracle's numbers."""
    model = build(golden_dataset(golden), golden_params(golden), _name(golden), linear_schedule=linear)
    o = _oracle_for(golden)
    u, p, n = batch(golden, 0)
    seen = {}
    for size in (u.numel(), 5, u.numel()):
        loss = model.bpr_loss(u[:size], p[:size], n[:size])
        loss.backward()
        lo = o.bpr_loss(u[:size], p[:size], n[:size])
        og = o.grads(lo)
        assert abs(float(loss) - float(lo)) < TOL * abs(float(lo))
        for nm, prm in model.named_parameters():
            if nm in og:
                assert rel(prm.grad, og[nm].numpy()) < (5e-5 if nm.endswith("bias") else TOL), (size, nm)
            prm.grad = None
        ws = model._ws
        assert ws["B"] == size
        seen.setdefault(size, ws)
        assert seen[size] is ws                                  # cached per batch size
    a, b = seen[u.numel()], seen[5]
    shared = ("O", "F_all", "g_flat", "mask") + (("H", "E0") if linear else ("dW", "X0_i"))
    for k in shared:
        ta, tb = (a[k][0], b[k][0]) if isinstance(a[k], list) else (a[k], b[k])
        assert ta.data_ptr() == tb.data_ptr(), k
    assert a["O_inst"].data_ptr() != b["O_inst"].data_ptr() and a["cache"] is b["cache"]


def test_get_embedding(backend, golden):
    """models/EliMRec.py:274-289: (all_users[u], all_items[p], all_items[n], E_u[u], E_i[p], E_i[n])"""
    model = build(golden_dataset(golden), golden_params(golden), _name(golden))
    o = _oracle_for(golden)
    u, p, n = batch(golden, 0)
    o.bpr_loss(u, p, n)
    out = model.getEmbedding(u, p, n)
    assert len(out) == 6
    wu, wi = o.cache["users"].detach(), o.cache["items"].detach()
    for got, want in zip(out[:3], (wu[u], wi[p], wi[n])):
        assert rel(got, want.numpy()) < TOL
    sd = golden_params(golden)
    for got, want in zip(out[3:], (sd["embedding_user.weight"][u], sd["embedding_item.weight"][p], sd["embedding_item.weight"][n])):
        assert np.array_equal(got.detach().cpu().numpy(), want)


def test_fused_adam_epilogue_matches_separate_pass(backend, golden):
    """linear schedule on one GPU: Adam on the two embedding tables runs in the epilogue of the last backward hop (the table
    gradients are never stored; the pre-update rows are recorded for tables completed later).  Same parameters, same
    moments, same evaluation as the separate optimizer pass, and both reproduce the reference's three steps."""
    from test_schedule_sim import check_steps
    name = _name(golden)
    fused = build(golden_dataset(golden), golden_params(golden), name)
    plain = build(golden_dataset(golden), golden_params(golden), name, fused_adam=False)
    for m in (fused, plain):
        m.make_optimizer(lr=1e-3, weight_decay=1e-4)
    for i in range(3):
        lf, lp = float(fused.train_step(*batch(golden, i))), float(plain.train_step(*batch(golden, i)))
        assert abs(lf - lp) <= 1e-6 * abs(lp)
    for (k, a), b in zip(fused.state_dict().items(), plain.state_dict().values()):
        assert rel(a, b.detach().cpu().numpy()) < 1e-6, k
    for n in ("embedding_user.weight", "embedding_item.weight"):
        for a, b in zip(fused._adam.state[n], plain._adam.state[n]):
            assert rel(a, b.detach().cpu().numpy()) < 1e-5, n
    # tables of the LAST forward (pre-update embeddings): the fused epilogue recorded them
    assert rel(fused.all_users, plain.all_users.detach().cpu().numpy()) < 1e-6
    assert rel(fused.all_items, plain.all_items.detach().cpu().numpy()) < 1e-6
    check_steps(build(golden_dataset(golden), golden_params(golden), name), golden, "")


def test_work_list_segment_length_follows_the_mean_degree(monkeypatch):
    """graph.py: a CSR half is cut into work items of 64 edges, 128 when its mean degree exceeds 32 (most rows would be split
    at 64); every edge is in exactly one item either way and ELIMREC_SEG64_LEN overrides the choice."""
    from elimrec_b200 import graph
    monkeypatch.delenv("ELIMREC_SEG64_LEN", raising=False)
    rng = np.random.default_rng(0)
    for mean, want in ((8, 64), (150, 128)):
        deg = rng.poisson(mean, size=400)
        deg[:3] = (1000, 0, 65)
        indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
        seg = graph.seg64_len_for(indptr)
        assert seg == want
        items, hrow, n_hit = graph.build_segments64(indptr, seg)
        covered = np.zeros(indptr[-1], dtype=np.int32)
        for r, b, e, h in items:
            assert 0 <= e - b <= seg and indptr[r] <= b and e <= indptr[r + 1]      # (an empty row is an item too: it writes zeros)
            covered[b:e] += 1
        assert (covered == 1).all()
        assert n_hit == int(hrow[:, 1].sum()) and (items[:n_hit, 3] >= 0).all() and (items[n_hit:, 3] < 0).all()
    monkeypatch.setenv("ELIMREC_SEG64_LEN", "256")
    assert graph.seg64_len_for(indptr) == 256


@pytest.mark.parametrize("knobs", [dict(wgrad_groups="early"), dict(wgrad_groups="late"), dict(wgrad_overlap=False),
                                   dict(fused_seed=False), dict(inst_fuse="ffma"), dict(wgrad_precision="fp32"),
                                   dict(fused_adam=False), dict(two_hop_masks=True)],
                         ids=lambda k: ",".join(f"{a}={b}" for a, b in k.items()))
def test_linear_schedule_knobs_leave_the_trajectory_alone(sim, golden, knobs):
    """Every scheduling knob of the linear step (how the weight gradients are launched, where the seeds and the loss
    reduction run, which kernel does fusion + heads on the instance rows, Adam fused or separate, two-hop masks) reorders or
    regroups launches only: three Adam steps reproduce the reference's losses and parameters."""
    from test_schedule_sim import check_steps
    model = build(golden_dataset(golden), golden_params(golden), _name(golden), proj_precision="x3", **knobs)
    assert model.linear
    check_steps(model, golden, "")
