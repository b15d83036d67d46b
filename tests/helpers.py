"""Shared helpers: rebuild datasets / models from the golden fixtures."""
import numpy as np
import scipy.sparse as sp


def csr_from_golden(g, split):
    U, I = int(g["num_users"]), int(g["num_items"])
    indptr, indices = g[f"{split}_indptr"], g[f"{split}_indices"]
    return sp.csr_matrix((np.ones(indices.size, dtype=np.float64), indices, indptr), shape=(U, I))


def dict_from_csr(m):
    out = {}
    for u in range(m.shape[0]):
        a, b = m.indptr[u], m.indptr[u + 1]
        if b > a:
            out[u] = m.indices[a:b].tolist()
    return out


def golden_params(g, prefix="sd0/"):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def golden_feats(g):
    order = g["itemids_raw_in_order"]
    return {m: g[f"raw_feat_{m}"][order] for m in "vat" if f"raw_feat_{m}" in g}


def golden_dataset(g):
    """An elimrec_b200.data.Dataset built from the raw splits stored in the fixture."""
    from elimrec_b200 import synth
    from elimrec_b200.data import Dataset
    inter = synth.Interactions(int(g["num_users"]), int(g["num_items"]), g["raw_train"], g["raw_valid"], g["raw_test"])
    feats = [g.get(f"raw_feat_{m}") for m in "vat"]
    name = "kwai" if g["_name"] == "kwai" else "synthg"
    return Dataset(None, interactions=inter, features=feats, name=name)
