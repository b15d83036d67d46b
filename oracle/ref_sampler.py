"""ORACLE (test infrastructure) - ctypes wrapper of ``oracle/ref_sampler.c`` plus the batch iterator.

``sample_epoch``  : data/sampler.py:93-126 via libc rand() (see ref_sampler.c header).
``epoch_batches`` : data/sampler.py:336-344 + util/data_iterator.py:58-60,133-152 - ONE
                    ``np.random.permutation(E)`` from the global numpy RNG, consecutive chunks of
                    ``batch_size``, last batch short.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libref_sampler.so")


def _lib():
    if not os.path.exists(_SO):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, os.path.join(_DIR, "ref_sampler.c")], check=True)
    return ctypes.CDLL(_SO)


def srand(seed: int = 1):
    _lib().ref_srand(ctypes.c_uint(seed))


def sample_epoch(user_train: dict, num_items: int, num_samples: int | None = None):
    lib = _lib()
    users = np.fromiter(user_train.keys(), dtype=np.int32)
    lens = np.fromiter((len(v) for v in user_train.values()), dtype=np.int64)
    ptr = np.zeros(users.size + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(lens)
    items = np.concatenate([np.asarray(v, dtype=np.int32) for v in user_train.values()])
    n = int(ptr[-1]) if num_samples is None else int(num_samples)
    ou, op, on = (np.empty(n, dtype=np.int32) for _ in range(3))
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.ref_sample_epoch(ctypes.c_int(users.size), vp(users), vp(ptr), vp(items), ctypes.c_int(num_items),
                              ctypes.c_longlong(n), vp(ou), vp(op), vp(on))
    if rc != 0:
        raise ValueError("a user owns every item")
    return ou.astype(np.int64), op.astype(np.int64), on.astype(np.int64)


def epoch_batches(users, pos, neg, batch_size: int, shuffle: bool = True):
    n = len(users)
    perm = np.random.permutation(n) if shuffle else np.arange(n)
    for b in range(0, n, batch_size):
        idx = perm[b:b + batch_size]
        yield users[idx], pos[idx], neg[idx]
