"""Host-side input contract of the hot path: ``Config`` and ``Dataset``.

These mirror, attribute for attribute, what ``models/EliMRec.py`` and ``evaluator/`` read from the
reference's ``util/configurator.py:Configurator`` (dict / attribute / ``in`` access,
``configurator.py:117-149``) and ``data/dataset.py:Dataset`` (``num_users``, ``num_items``,
``train_matrix``, ``v_feat``/``a_feat``/``t_feat``, ``get_user_{train,valid,test}_dict``,
``get_train_interactions``; ``dataset.py:105-192, 319-354``).  The reference's own objects can be
passed to the model instead (duck typing) - that is the drop-in case driven by ``main.py``.

Vectorised: no per-row Python loops (the reference's ``csr_to_user_dict`` / dok iteration,
``util/tool.py:70-79`` and ``dataset.py:347-354``, are replaced by CSR slicing).
"""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sp
import torch

# Defaults = NeuRec.properties + conf/EliMRec.properties of the reference
DEFAULTS = {
    "recommender": "EliMRec", "layer_num": 3, "batch_size": 2048, "recdim": 64, "lr": 0.001,
    "weight_decay": 1e-4, "topks": [10], "num_epoch": 1000, "seed": 2022, "test_step": 10,
    "temp": 0.2, "adj_type": "pre", "logits": "cosin", "stop_cnt": 50, "save_flag": True,
    "data.input.path": "./dataset", "data.input.dataset": "tiktok", "data.column.format": "UI",
    "data.convert.separator": ",", "splitter": "given", "metric": ["Precision", "Recall", "NDCG"],
    "group_view": None, "rec.evaluate.neg": 0, "test_batch_size": 128, "num_thread": 8,
    "no_cuda": False, "suffix": "", "path": "./saved_models", "with_item_vat": True,
    "pretrain": False, "loss": "bpr_loss", "alpha": 0.5, "verbose": 1,
}


class Config:
    """Dict/attribute config with the access semantics of the reference ``Configurator``."""

    def __init__(self, **overrides):
        object.__setattr__(self, "_d", dict(DEFAULTS))
        self._d.update(overrides)

    def __getitem__(self, k):
        if not isinstance(k, str):
            raise TypeError("index must be a str")
        if k not in self._d:
            raise KeyError("There are not the parameter named '%s'" % k)
        return self._d[k]

    def __setitem__(self, k, v):
        self._d[k] = v

    def __getattr__(self, k):
        try:
            return self._d[k]
        except KeyError:
            raise KeyError("There are not the parameter named '%s'" % k)

    def __setattr__(self, k, v):
        self._d[k] = v

    def __contains__(self, k):
        return k in self._d


def _remap_first_appearance(col_all: np.ndarray):
    """ids -> 0..n-1 in order of first appearance (``dataset.py:219-232`` via ``Series.unique``)."""
    uniq, first = np.unique(col_all, return_index=True)
    order = np.argsort(first, kind="stable")
    raw_in_order = uniq[order]
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return raw_in_order, uniq, rank  # new_id = rank[searchsorted(uniq, raw)]


def csr_rows_to_dict(m: sp.csr_matrix) -> dict:
    """``util/tool.py:70-79`` (``csr_to_user_dict``) without the per-row scipy slicing."""
    indptr, indices = m.indptr, m.indices
    out = {}
    nz = np.nonzero(np.diff(indptr))[0]
    idx_list = indices.tolist()
    ptr = indptr.tolist()
    for u in nz.tolist():
        out[u] = idx_list[ptr[u]:ptr[u + 1]]
    return out


class Dataset:
    """Same public surface as the reference ``Dataset`` for the fields the hot path reads."""

    def __init__(self, conf=None, *, interactions=None, features=None, name=None, words=None):
        self.conf = conf
        self.userids = self.itemids = None
        if interactions is not None:
            self.dataset_name = name or (conf["data.input.dataset"] if conf is not None else "synthetic")
            self._from_arrays(interactions.train, interactions.valid, interactions.test, features, words)
        else:
            self.dataset_name = conf["data.input.dataset"]
            self._load_files(conf)

    # ---- construction -------------------------------------------------------------------------
    def _load_files(self, conf):
        prefix = os.path.join(conf["data.input.path"], self.dataset_name)
        sep = conf["data.convert.separator"]
        rd = lambda f: np.loadtxt(f, delimiter=sep, dtype=np.int64, ndmin=2)[:, :2]
        train, valid, test = rd(prefix + ".train"), rd(prefix + ".valid"), rd(prefix + ".test")
        path = conf["data.input.path"]
        feats = [None, None, None]
        words = None
        if "with_item_vat" not in conf or conf["with_item_vat"]:
            if self.dataset_name == "tiktok":
                # dataset.py:164-174: visual / audio tensors by raw item id, text as [2 x n] (raw item id, word id) pairs
                feats[0] = torch.load(os.path.join(path, "tiktok_visual_feat.pt")).numpy()
                feats[1] = torch.load(os.path.join(path, "tiktok_audio_feat.pt")).numpy()
                words = torch.load(os.path.join(path, "tiktok_textual_feat.pt")).detach().cpu().numpy()
            elif self.dataset_name == "kwai":
                feats[0] = torch.load(os.path.join(path, "kwai_feat_v.pt")).numpy()
            else:
                feats[0] = np.load(os.path.join(path, f"{self.dataset_name}_FeatureVideo_normal.npy"))
                feats[1] = np.load(os.path.join(path, f"{self.dataset_name}_FeatureAudio_avg_normal.npy"))
                feats[2] = np.load(os.path.join(path, f"{self.dataset_name}_FeatureText_stl_normal.npy"))
        self._from_arrays(train, valid, test, feats, words)

    def _from_arrays(self, train, valid, test, feats, words=None):
        # reference order of concatenation is [train, test, valid] (dataset.py:219)
        all_u = np.concatenate([train[:, 0], test[:, 0], valid[:, 0]])
        all_i = np.concatenate([train[:, 1], test[:, 1], valid[:, 1]])
        raw_u, uq_u, rk_u = _remap_first_appearance(all_u)
        raw_i, uq_i, rk_i = _remap_first_appearance(all_i)
        self.userids = dict(zip(raw_u.tolist(), range(raw_u.size)))
        self.itemids = dict(zip(raw_i.tolist(), range(raw_i.size)))
        self._raw_items_in_order = raw_i
        mu = lambda a: rk_u[np.searchsorted(uq_u, a)]
        mi = lambda a: rk_i[np.searchsorted(uq_i, a)]
        self.num_users = int(raw_u.size)
        self.num_items = int(raw_i.size)
        self.num_ratings = int(all_u.size)

        def csr(pairs):
            m = sp.csr_matrix((np.ones(pairs.shape[0], dtype=np.float64), (mu(pairs[:, 0]), mi(pairs[:, 1]))),
                              shape=(self.num_users, self.num_items))
            m.sum_duplicates()
            m.sort_indices()
            return m

        self.train_matrix, self.valid_matrix, self.test_matrix = csr(train), csr(valid), csr(test)
        self.trainDataSize = int(train.shape[0])
        self.negative_matrix = None
        names = ("v_feat", "a_feat", "t_feat")
        for nm, f in zip(names, feats if feats is not None else [None] * 3):
            if f is not None:
                setattr(self, nm, torch.from_numpy(np.ascontiguousarray(f[raw_i])))
        if words is not None:
            # literal 'tiktok' (dataset.py:166-173): keep the pairs whose raw item id is known, in file order, item id remapped
            w = np.asarray(words, dtype=np.int64)
            pos = np.minimum(np.searchsorted(uq_i, w[0]), uq_i.size - 1)
            known = uq_i[pos] == w[0]
            self.words_tensor = torch.from_numpy(np.stack([rk_i[pos[known]], w[1][known]]).astype(np.int64))

    # ---- accessors (same names as the reference) ------------------------------------------------
    def get_user_train_dict(self, by_time=False):
        return csr_rows_to_dict(self.train_matrix)

    def get_user_valid_dict(self):
        return csr_rows_to_dict(self.valid_matrix)

    def get_user_test_dict(self):
        return csr_rows_to_dict(self.test_matrix)

    def get_train_interactions(self):
        coo = self.train_matrix.tocoo()
        return coo.row.tolist(), coo.col.tolist()

    def to_csr_matrix(self):
        return self.train_matrix.copy()
