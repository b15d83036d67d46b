"""ORACLE (test infrastructure) - torch-CPU restatement of the reference model math.

Follows, op for op and in the same order (so fp32 rounding matches to ~1e-7):
  * adjacency  : ``models/EliMRec.py:309-354`` (``create_adj_mat``, all five ``adj_type``s), COO->torch ``:80-84``
  * forward    : ``models/EliMRec.py:228-272`` (``compute`` / ``compute_graph``)
  * heads      : ``models/EliMRec.py:144-153`` (``gcn_cf``)
  * loss       : ``models/EliMRec.py:115-142`` (``bpr_loss``), ``:291-297`` (``original_bpr_loss``)
  * scoring    : ``models/EliMRec.py:96-113`` (``predict``), ``:155-212`` (``general_cm_fusion``: rubi / hm / sum)
  * fusion     : ``models/EliMRec.py:221-226`` (``mm_fusion``: concat / mean)
  * tiktok text: ``models/EliMRec.py:371-378`` (``t_feat = scatter_mean(word_embedding(words[1]), words[0])``, once, with
                 autograd history; ``torch_scatter`` is an un-vendored dependency, restated as sum / count)
  * optimiser  : ``main.py:49,99-101`` (torch.optim.Adam, coupled L2)
Pinned against the reference's own outputs in ``tests/golden/*.npz`` (tests/test_oracle_golden.py,
tests/test_oracle_next.py).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F

MODS = "vat"


def norm_adj_coo(train_csr: sp.csr_matrix, num_users: int, num_items: int, adj_type: str = "pre"):
    """(row, col, val) of the propagation matrix exactly as scipy produces it in the reference
    (EliMRec.py:310-354): fp32 values, row-major, columns sorted inside a row."""
    coo = train_csr.tocoo()
    u = coo.row.astype(np.int32)
    i = coo.col.astype(np.int32)
    n = num_users + num_items
    half = sp.csr_matrix((np.ones_like(u, dtype=np.float32), (u, i + num_users)), shape=(n, n))
    adj = half + half.T

    def single(a):      # normalized_adj_single, EliMRec.py:319-327
        rowsum = np.array(a.sum(1))
        with np.errstate(divide="ignore"):
            d_inv = np.power(rowsum, -1).flatten()
        d_inv[np.isinf(d_inv)] = 0.0
        return sp.diags(d_inv).dot(a).tocoo()

    if adj_type == "plain":
        out = adj
    elif adj_type == "norm":
        out = single(adj + sp.eye(n))
    elif adj_type == "gcmc":
        out = single(adj)
    elif adj_type == "pre":
        deg = np.array(adj.sum(1))
        with np.errstate(divide="ignore"):
            dinv = np.power(deg, -0.5).flatten()
        dinv[np.isinf(dinv)] = 0.0
        dm = sp.diags(dinv)
        out = dm.dot(adj).dot(dm)
    else:               # 'mean' (the reference's else branch, EliMRec.py:349-352)
        out = single(adj) + sp.eye(n)
    out = out.tocoo()
    return out.row.astype(np.int64), out.col.astype(np.int64), out.data.astype(np.float32)


def scatter_mean(src: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """torch_scatter.scatter(src, index, dim=0, reduce='mean'): rows = max(index)+1, mean = sum / max(count, 1)."""
    n = int(index.max()) + 1
    out = torch.zeros(n, src.shape[1], dtype=src.dtype).index_add(0, index, src)
    cnt = torch.zeros(n, dtype=src.dtype).index_add(0, index, torch.ones_like(index, dtype=src.dtype))
    return out / cnt.clamp(min=1).unsqueeze(1)


class OracleEliMRec:
    """Functional restatement; parameters are plain tensors keyed like the reference state_dict."""

    def __init__(self, params: dict, feats: dict, train_csr, num_users, num_items, *, kwai=False,
                 alpha=0.5, n_layers=3, modality="vat", predict_type="TIE", adj_type="pre", mm_fusion_mode="concat",
                 s_fusion_mode="rubi", words=None):
        self.p = {k: torch.as_tensor(v).clone().float().requires_grad_(True) for k, v in params.items()}
        self.U, self.I = int(num_users), int(num_items)
        self.kwai = kwai
        self.mods = "v" if kwai else MODS
        self.alpha = alpha
        self.L = n_layers
        self.modality = "v" if kwai else modality
        self.predict_type = predict_type
        # features are L2-row-normalised once at init (EliMRec.py:366-381)
        self.feat = {m: F.normalize(torch.as_tensor(feats[m]).float(), dim=1) for m in self.mods if m in feats}
        self.mm_fusion_mode, self.fusion_mode = mm_fusion_mode, s_fusion_mode
        if words is not None:   # literal 'tiktok': text feature from word ids, NOT normalised, keeps its autograd history
            w = torch.as_tensor(np.asarray(words)).long()
            self.feat["t"] = scatter_mean(self.p["word_embedding.weight"][w[1]], w[0])
        r, c, v = norm_adj_coo(train_csr, self.U, self.I, adj_type)
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
        if self.mm_fusion_mode == "concat":
            fuse = lambda reps: torch.cat(reps, dim=1)
        else:           # 'mean', EliMRec.py:224-225
            fuse = lambda reps: torch.mean(torch.stack(reps), dim=0)
        cat_u = fuse([light[b][: self.U] for b in blocks])
        cat_i = fuse([light[b][self.U:] for b in blocks])
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

    # -- EliMRec.py:155-212 and 96-113
    def _cm(self, logits, users):
        c = self.cache
        cos = {}
        for m in self.mods:     # reference order: z_v, then z_a, then z_t
            su, si = c["s"][m]
            cos[m] = F.normalize(su[users], dim=1) @ F.normalize(si, dim=1).t()
        if self.fusion_mode == "rubi":
            z = logits
            for m in self.mods:
                if m in self.modality:
                    z = z * torch.sigmoid(cos[m])
            return z
        if self.fusion_mode == "hm":    # every modality, whatever `modality` says (EliMRec.py:190-199)
            z = torch.sigmoid(logits)
            for m in self.mods:
                z = z * torch.sigmoid(cos[m])
            return torch.log(z + 1e-12) - torch.log1p(z)
        z = logits                      # 'sum' (EliMRec.py:201-210)
        for m in self.mods:
            z = z + cos[m]
        return torch.log(torch.sigmoid(z) + 1e-12)

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
