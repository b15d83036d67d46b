import numpy as np
import torch


def rel_err(a, b):
    """norm-wise relative error max|a-b| / max|b| (the parity metric of BASELINE.json's tolerances)."""
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


FP32_TOL = 1e-5   # north star: fp32 paths within 1e-5 relative
TC_TOL = 1e-3     # TF32 / bf16 tensor-core GEMMs within 1e-3 relative


def build_model(ds, params=None, dataset_name="synthg", **cfg):
    from elimrec_b200.data import Config
    from elimrec_b200.model import EliMRec
    cfg.setdefault("proj_precision", "fp32")   # exact path unless a test asks for the tensor-core projections
    cfg.setdefault("device", torch.device("cuda:0"))
    conf = Config(**{"data.input.dataset": dataset_name, "topks": [20], **cfg})
    model = EliMRec(conf, ds).to(conf.device)
    if params is not None:
        sd = {k: torch.as_tensor(v) for k, v in params.items()}
        missing = model.load_state_dict(sd, strict=False)
        assert not missing.unexpected_keys, missing
    return model


def topk_sets_match(idx_gpu, scores_ref, masked, k, tol):
    """exact-set matches / tolerance-explained swaps / unexplained swaps (top-K ties)."""
    exact = explained = bad = 0
    for r in range(idx_gpu.shape[0]):
        s = scores_ref[r].copy()
        s[masked[r]] = -np.inf
        ref = np.argsort(-s, kind="stable")[:k]
        got = idx_gpu[r]
        if np.array_equal(ref, got):
            exact += 1
            continue
        kth = s[ref[-1]]
        diff = set(got.tolist()) ^ set(ref.tolist())
        # every item that differs, and every out-of-order pair, must sit within tol of its neighbour
        ok = all(abs(s[i] - kth) <= tol for i in diff) and np.all(np.abs(np.sort(-s[got]) + s[got]) <= tol)
        if ok:
            explained += 1
        else:
            bad += 1
    return exact, explained, bad
