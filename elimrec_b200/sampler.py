"""BPR triple samplers - drop-in for ``data/sampler.py:PairwiseSamplerV2`` (``:297-351``).

``PairwiseSamplerV2``  : same constructor / iteration protocol / batch contents as the reference.
    ``mode='compat'`` (default) replays the reference's libc ``rand()`` stream bit-exactly
    (``elimrec_sample_epoch_compat``; the generator state is owned by this object, seed 1 = what an
    unseeded glibc process gives) and shuffles with ONE ``np.random.permutation`` from the global
    numpy RNG exactly like ``DataIterator`` (``util/data_iterator.py:58-60``), so the same
    ``set_seed`` yields the same batches.  Vectorised: no per-user / per-sample Python loops.
    ``mode='device'`` draws the epoch on the GPU (Philox4x32-10, same distribution) and yields device
    tensors - the throughput mode; its numpy restatement is ``oracle/philox_sampler.py``.
    ``prefetch=True`` (compat mode, off by default): the NEXT epoch is sampled and shuffled on a host thread while the
    current one trains (the C sampler releases the GIL); ``pin=True`` yields pinned torch tensors so the per-batch
    host->device copies are asynchronous.  Both keep the stream: the draws happen in the same order, just earlier.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np
import torch

from . import _lib, ops


class CompatRng:
    """glibc rand() stream (TYPE_3) with caller-owned state; seed 1 == never-seeded libc."""

    def __init__(self, seed: int = 1):
        self.state = np.zeros(40, dtype=np.uint32)
        self.seed(seed)

    def seed(self, seed: int):
        _lib.lib().elimrec_compat_rng_seed(self.state.ctypes.data, seed)

    def rand(self) -> int:
        return int(_lib.lib().elimrec_compat_rng_next(self.state.ctypes.data))


def _train_csr(user_pos_dict: dict):
    users = np.fromiter(user_pos_dict.keys(), dtype=np.int32, count=len(user_pos_dict))
    lens = np.fromiter((len(v) for v in user_pos_dict.values()), dtype=np.int64, count=len(user_pos_dict))
    ptr = np.zeros(users.size + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    items = np.concatenate([np.asarray(v, dtype=np.int32) for v in user_pos_dict.values()]) if users.size else np.zeros(0, np.int32)
    return users, ptr, items


class PairwiseSamplerV2:
    def __init__(self, dataset, neg_num=1, batch_size=1024, shuffle=True, drop_last=False, mode="compat",
                 device=None, seed=2022, prefetch=False, pin=False):
        if neg_num <= 0:
            raise ValueError("'neg_num' must be a positive integer.")
        if neg_num != 1:
            raise NotImplementedError("neg_num > 1 is never used by main.py (main.py:59)")
        self.batch_size, self.drop_last, self.shuffle, self.neg_num = batch_size, drop_last, shuffle, neg_num
        self.item_num = int(dataset.num_items)
        user_pos_dict = dataset.get_user_train_dict()
        if not user_pos_dict:
            raise ValueError("'user_pos_dict' cannot be empty.")
        self.users, self.ptr, self.items = _train_csr(user_pos_dict)
        # the reference keeps the per-user lists as given (sorted CSR rows); membership test needs sorted
        self.num_trainings = int(self.ptr[-1])
        self.mode = mode
        self.rng = CompatRng(1)
        self.epoch = 0
        self.seed = seed
        self.device = device
        self._dev = None
        self.prefetch, self.pin = bool(prefetch), bool(pin)
        self._next = None           # (thread, result holder) of the epoch being sampled ahead

    def __len__(self):
        n = self.num_trainings
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    # -- one epoch of triples ------------------------------------------------------------------
    def sample_epoch_host(self):
        n = self.num_trainings
        ou, op, on = (np.empty(n, dtype=np.int64) for _ in range(3))
        vp = lambda a: a.ctypes.data
        rc = _lib.lib().elimrec_sample_epoch_compat(vp(self.rng.state), self.users.size, vp(self.users), vp(self.ptr),
                                                   vp(self.items), self.item_num, n, vp(ou), vp(op), vp(on))
        if rc != 0:
            raise _lib.ElimrecError(_lib.lib().elimrec_last_error().decode())
        return ou, op, on

    def ensure_device(self, dev=None):
        """upload the train CSR once (must happen outside CUDA-graph capture)"""
        if self._dev is None:
            dev = dev or self.device or torch.device("cuda", torch.cuda.current_device())
            self._dev = (torch.from_numpy(self.users).to(dev), torch.from_numpy(self.ptr).to(dev),
                         torch.from_numpy(self.items).to(dev))
        return self._dev

    def sample_epoch_device(self, n=None):
        dev = self.device or torch.device("cuda", torch.cuda.current_device())
        self.ensure_device(dev)
        n = self.num_trainings if n is None else n
        ou, op, on = (torch.empty(n, dtype=torch.int64, device=dev) for _ in range(3))
        ops.sample_triples_device(self.seed, self.epoch, n, *self._dev, self.item_num, ou, op, on)
        return ou, op, on

    def sample_batch_device(self, batch_index_dev, ou, op, on):
        """One batch (``ou.numel()`` triples) of the current epoch's device stream; its position in the stream is read from
        the device counter ``batch_index_dev`` (int64[1]) - capturable in a CUDA graph, a new batch at every replay."""
        self.ensure_device(ou.device)
        ops.sample_batch_device(self.seed, self.epoch, batch_index_dev, ou.numel(), *self._dev, self.item_num, ou, op, on)

    def _host_epoch(self):
        """sample + shuffle one epoch on the host (the two steps of data/sampler.py:336-344), optionally into pinned memory"""
        u, p, n = self.sample_epoch_host()
        total = u.size
        perm = np.random.permutation(total) if self.shuffle else None
        if self.pin:
            # pinned and PACKED: full batch b is [users | pos | neg] in 3 * batch_size consecutive words, so that a consumer can
            # move it to the device in one copy (EliMRec.make_graphed_step does); the short last batch follows, same layout
            bs = self.batch_size
            nb, tail = divmod(total, bs)
            buf = torch.empty(3 * total, dtype=torch.int64).pin_memory()
            full = buf[:3 * nb * bs].view(nb, 3, bs).numpy()
            rest = buf[3 * nb * bs:].view(3, tail).numpy()
            for j, src in enumerate((u, p, n)):
                src = src if perm is None else src[perm]
                full[:, j, :] = src[:nb * bs].reshape(nb, bs)
                rest[j, :] = src[nb * bs:]
            return ("packed", buf, None), total
        if perm is not None:
            u, p, n = u[perm], p[perm], n[perm]
        return (u, p, n), total

    def _start_prefetch(self):
        box = {}

        def work():
            try:
                box["epoch"] = self._host_epoch()
            except BaseException as exc:      # surfaced by the consumer
                box["error"] = exc
        t = threading.Thread(target=work, daemon=True)
        t.start()
        self._next = (t, box)

    def __iter__(self):
        bs = self.batch_size
        if self.mode == "compat":
            if self._next is not None:
                t, box = self._next
                t.join()
                self._next = None
                if "error" in box:
                    raise box["error"]
                (u, p, n), total = box["epoch"]
            else:
                (u, p, n), total = self._host_epoch()
            if self.prefetch:
                self._start_prefetch()
        else:
            u, p, n = self.sample_epoch_device()
            total = u.numel()  # i.i.d. draws: already in random order, no shuffle pass needed
        self.epoch += 1
        if isinstance(u, str):      # packed pinned epoch (see _host_epoch): views, no copies
            buf = p
            nb, tail = divmod(total, bs)
            for b in range(nb):
                blk = buf[3 * b * bs:3 * (b + 1) * bs]
                yield blk[:bs], blk[bs:2 * bs], blk[2 * bs:]
            if tail and not self.drop_last:
                blk = buf[3 * nb * bs:]
                yield blk[:tail], blk[tail:2 * tail], blk[2 * tail:]
            return
        for b in range(0, total, bs):
            if self.drop_last and b + bs > total:
                break
            yield u[b:b + bs], p[b:b + bs], n[b:b + bs]
