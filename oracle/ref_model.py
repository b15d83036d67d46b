"""ORACLE (test infrastructure) - torch-CPU restatement of the reference model math.

Follows, op for op and in the same order (so fp32 rounding matches to ~1e-7):
  * adjacency  : ``models/EliMRec.py:309-354`` (``create_adj_mat``, 'pre'), COO->torch ``:80-84``
  * forward    : ``models/EliMRec.py:228-272`` (``compute`` / ``compute_graph``)
  * heads      : ``models/EliMRec.py:144-153`` (``gcn_cf``)
  * loss       : ``models/EliMRec.py:115-142`` (``bpr_loss``), ``:291-297`` (``original_bpr_loss``)
  * scoring    : ``models/EliMRec.py:96-113`` (``predict``), ``:155-188`` (``general_cm_fusion``, rubi)
  * optimiser  : ``main.py:49,99-101`` (torch.optim.Adam, coupled L2)
Pinned against the reference's own outputs in ``tests/golden/*.npz`` (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F

MODS = "vat"


def norm_adj_coo(train_csr: sp.csr_matrix, num_users: int, num_items: int):
    """(row, col, val) of D^-1/2 A D^-1/2 exactly as scipy produces it in the reference
    (EliMRec.py:310-347): fp32 values, row-major, columns sorted inside a row."""
    coo = train_csr.tocoo()
    u = coo.row.astype(np.int32)
    i = coo.col.astype(np.int32)
    n = num_users + num_items
    half = sp.csr_matrix((np.ones_like(u, dtype=np.float32), (u, i + num_users)), shape=(n, n))
    adj = half + half.T
    deg = np.array(adj.sum(1))
    with np.errstate(divide="ignore"):
        dinv = np.power(deg, -0.5).flatten()
    dinv[np.isinf(dinv)] = 0.0
    dm = sp.diags(dinv)
    out = dm.dot(adj).dot(dm).tocoo()
    return out.row.astype(np.int64), out.col.astype(np.int64), out.data.astype(np.float32)


class OracleEliMRec:
    """Functional restatement; parameters are plain tensors keyed like the reference state_dict."""

    def __init__(self, params: dict, feats: dict, train_csr, num_users, num_items, *, kwai=False,
                 alpha=0.5, n_layers=3, modality="vat", predict_type="TIE"):
        self.p = {k: torch.as_tensor(v).clone().float().requires_grad_(True) for k, v in params.items()}
        self.U, self.I = int(num_users), int(num_items)
        self.kwai = kwai
        self.mods = "v" if kwai else MODS
        self.alpha = alpha
        self.L = n_layers
        self.modality = "v" if kwai else modality
        self.predict_type = predict_type
        # features are L2-row-normalised once at init (EliMRec.py:366-381)
        self.feat = {m: F.normalize(torch.as_tensor(feats[m]).float(), dim=1) for m in self.mods}
        r, c, v = norm_adj_coo(train_csr, self.U, self.I)
        n = self.U + self.I
        # legacy-constructor equivalent: an UNcoalesced COO (EliMRec.py:81-83)
        self.adj = torch.sparse_coo_tensor(torch.from_numpy(np.stack([r, c])), torch.from_numpy(v), (n, n))
        self.cache = None

    # -- EliMRec.py:238-248
    def _graph(self, u_emb, i_emb):
        x = torch.cat([u_emb, i_emb])
        layers = [x]
        for _ in range(self.L):
            x = torch.sparse.mm(self.adj, x)
            layers.append(x)
        return torch.mean(torch.stack(layers, dim=1), dim=1)

    # -- EliMRec.py:228-272 + 144-153
    def forward(self):
        p = self.p
        eu, ei = p["embedding_user.weight"], p["embedding_item.weight"]
        light = {"i": self._graph(eu, ei)}
        for m in self.mods:
            proj = F.linear(self.feat[m], p[f"{m}_dense.weight"], p[f"{m}_dense.bias"])
            light[m] = self._graph(eu, proj)
        blocks = ["i"] + list(self.mods)
        cat_u = torch.cat([light[b][: self.U] for b in blocks], dim=1)
        cat_i = torch.cat([light[b][self.U:] for b in blocks], dim=1)
        users = F.linear(cat_u, p["embedding_user_after_GCN.weight"], p["embedding_user_after_GCN.bias"])
        items = F.linear(cat_i, p["embedding_item_after_GCN.weight"], p["embedding_item_after_GCN.bias"])
        s = {}
        for m in self.mods:
            sm = F.linear(light[m], p[f"s_dense_{m}.weight"], p[f"s_dense_{m}.bias"])
            s[m] = (sm[: self.U], sm[self.U:])
        self.cache = dict(users=users, items=items, s=s, light=light)
        return self.cache

    # -- EliMRec.py:291-297
    @staticmethod
    def _bpr(u, p, n):
        u, p, n = F.normalize(u, dim=1), F.normalize(p, dim=1), F.normalize(n, dim=1)
        return torch.mean(F.softplus(torch.sum(u * n, dim=1) - torch.sum(u * p, dim=1)))

    # -- EliMRec.py:115-142
    def bpr_loss(self, users, pos, neg):
        users, pos, neg = (torch.as_tensor(x).long() for x in (users, pos, neg))
        c = self.forward()
        fusion = self._bpr(c["users"][users], c["items"][pos], c["items"][neg])
        if self.predict_type == "normal":
            return fusion
        single = 0
        for m in self.modality:
            su, si = c["s"][m]
            single = single + self._bpr(su[users], si[pos], si[neg])
        return fusion + self.alpha * single

    def grads(self, loss):
        names = [k for k in self.p]
        gs = torch.autograd.grad(loss, [self.p[k] for k in names], allow_unused=True, retain_graph=True)
        return {k: g for k, g in zip(names, gs) if g is not None}

    # -- EliMRec.py:155-188 (rubi) and 96-113
    def _cm(self, logits, users):
        c = self.cache
        z = logits
        order = self.mods  # reference multiplies z_v, then z_a, then z_t
        for m in order:
            if m in self.modality:
                su, si = c["s"][m]
                zu = F.normalize(su[users], dim=1)
                zi = F.normalize(si, dim=1)
                z = z * torch.sigmoid(zu @ zi.t())
        return z

    @torch.no_grad()
    def predict(self, user_ids, predict_type=None):
        pt = predict_type or self.predict_type
        users = torch.as_tensor(np.asarray(user_ids)).long()
        c = self.cache
        ui = torch.sigmoid(c["users"][users] @ c["items"].t())
        if pt == "TE":
            return torch.sigmoid(self._cm(ui, users))
        if pt == "TIE":
            fixed = torch.mean(ui, -1, True)
            return torch.sigmoid(self._cm(ui, users) - self._cm(fixed, users))
        return torch.sigmoid(ui)


def adam_reference(params: dict, grads_fn, steps, lr=1e-3, weight_decay=1e-4):
    """main.py:49,99-101 restated with torch.optim.Adam itself (third-party arithmetic, torch wheel)."""
    opt = torch.optim.Adam(list(params.values()), lr=lr, weight_decay=weight_decay)
    for s in range(steps):
        opt.zero_grad()
        grads_fn(s).backward()
        opt.step()
    return params
