"""ORACLE (test infrastructure) - numpy restatement of the reference full-ranking evaluator.

Follows ``evaluator/backend/cpp/uni_evaluator.py:104-203`` (user order, batching, train-item
masking to -inf, mean over users, top_show selection, "%.8f" formatting) and
``evaluator/backend/cpp/include/evaluate.h:23-42`` + ``include/metric.h:17-43,66-83``
(per-user top-K and the Precision / Recall / NDCG curves with their float accumulators).

Tie rule: the reference's ``partial_sort_copy`` order among EQUAL scores is unspecified; the contract
(BASELINE.json) is lowest index first, which is what this restatement and the CUDA path implement.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

METRIC_ID = {"Precision": 1, "Recall": 2, "MAP": 3, "NDCG": 4, "MRR": 5}


def topk_lowest_index(scores: np.ndarray, k: int) -> np.ndarray:
    """indices of the k largest entries per row, ties -> lowest index first."""
    order = np.argsort(-scores, axis=1, kind="stable")  # stable => equal scores keep index order
    return order[:, :k].astype(np.int32)


def metric_rows(topk: np.ndarray, truth: list, metrics, k: int) -> np.ndarray:
    """metric.h:17-43,66-83 - one [len(metrics)*k] fp32 row per user."""
    n = topk.shape[0]
    out = np.zeros((n, len(metrics) * k), dtype=np.float32)
    for r in range(n):
        tset = set(truth[r])
        tl = len(tset)
        hit = np.array([int(x) in tset for x in topk[r]], dtype=bool)
        for mi, m in enumerate(metrics):
            row = out[r, mi * k:(mi + 1) * k]
            if m == 1:    # precision: 1.0*hits/(i+1) in double, stored as float
                row[:] = (np.cumsum(hit).astype(np.float64) / np.arange(1, k + 1)).astype(np.float32)
            elif m == 2:  # recall
                row[:] = (np.cumsum(hit).astype(np.float64) / tl).astype(np.float32)
            elif m == 4:  # ndcg: float accumulators, double addends
                dcg = np.float32(0)
                idcg = np.float32(0)
                for i in range(k):
                    if hit[i]:
                        dcg = np.float32(np.float64(dcg) + 1.0 / np.log2(i + 2.0))
                    if i < tl:
                        idcg = np.float32(np.float64(idcg) + 1.0 / np.log2(i + 2.0))
                    row[i] = np.float32(dcg) / np.float32(idcg)
            elif m == 3:  # ap
                hits = 0
                s = np.float32(0)
                for i in range(k):
                    if hit[i]:
                        hits += 1
                        s = np.float32(np.float64(s) + np.float64(np.float32(1.0 * hits / (i + 1))))
                    row[i] = 0.0 if hits == 0 else np.float32(s) / np.float32(hits)
            elif m == 5:  # mrr
                first = np.nonzero(hit)[0]
                if first.size:
                    row[first[0]:] = np.float32(1.0 / (first[0] + 1))
            else:
                raise ValueError(m)
    return out


def evaluate(predict_fn, user_train: dict, user_test: dict, metrics=("Precision", "Recall", "NDCG"),
             top_k=(20,), batch_size=128, return_rows=False, user_neg: dict | None = None, test_users=None):
    """uni_evaluator.py:104-203.  ``predict_fn(list_of_users) -> [B x I] fp32 ndarray``.

    ``user_neg`` (candidate-negatives mode, ``:132-140``): the reference hands ``predict`` a candidate list per user,
    ``EliMRec.predict`` ignores it (``models/EliMRec.py:96``) and returns all I scores, so what the evaluator ranks is
    the full UNMASKED row against the truth ``set(range(len(pos_test[u])))`` - restated literally."""
    mids = [METRIC_ID[m] for m in metrics]
    max_top = top_k if isinstance(top_k, int) else max(top_k)
    top_show = np.arange(max_top) + 1 if isinstance(top_k, int) else np.sort(top_k)
    users = list(user_test.keys()) if test_users is None else list(test_users)
    rows = []
    for b in range(0, len(users), batch_size):
        bu = users[b:b + batch_size]
        sc = np.array(predict_fn(bu), dtype=np.float32)
        if user_neg is not None:
            truth = [list(range(len(user_test[u]))) for u in bu]
        else:
            truth = [user_test[u] for u in bu]
            for r, u in enumerate(bu):
                sc[r, user_train.get(u, [])] = -np.inf
        tk = topk_lowest_index(sc, max_top)
        rows.append(metric_rows(tk, truth, mids, max_top))
    allrows = np.concatenate(rows, axis=0)
    final = np.mean(allrows, axis=0).reshape(len(mids), max_top)[:, top_show - 1].reshape(-1)
    buf = "\t".join([("%.8f" % x).ljust(12) for x in final])
    if return_rows:
        return final, buf, allrows
    return final, buf


# ---- optional second checker: the reference's own C++ compiled from /root/reference -----------------
_REF_SO = os.path.join(os.path.dirname(__file__), "_ref", "libref_eval.so")


def ref_cpp_available() -> bool:
    return os.path.exists(_REF_SO)


def ref_cpp_metric_rows(scores: np.ndarray, truth: list, metrics, k: int, threads: int = 8) -> np.ndarray:
    """Calls ``cpp_evaluate_matrix`` (evaluate.h:45-64) through oracle/ref_shim.cpp."""
    lib = ctypes.CDLL(_REF_SO)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n, rating_len = scores.shape
    ptr = np.zeros(n + 1, dtype=np.int32)
    ptr[1:] = np.cumsum([len(t) for t in truth])
    flat = np.concatenate([np.asarray(t, dtype=np.int32) for t in truth]) if ptr[-1] else np.zeros(0, np.int32)
    mids = np.asarray(metrics, dtype=np.int32)
    out = np.zeros((n, len(metrics) * k), dtype=np.float32)
    lib.ref_evaluate_matrix(scores.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n), ctypes.c_int(rating_len),
                            ptr.ctypes.data_as(ctypes.c_void_p), flat.ctypes.data_as(ctypes.c_void_p),
                            mids.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(metrics)), ctypes.c_int(k),
                            ctypes.c_int(threads), out.ctypes.data_as(ctypes.c_void_p))
    return out
