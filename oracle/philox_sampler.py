"""ORACLE (test infrastructure) - numpy restatement of the DEVICE sampler (Philox4x32-10).

The device sampler is new design (the reference samples on the host with libc rand(),
data/sampler.py:93-126); what it must preserve is the reference's distribution: user uniform over
users with train items, positive uniform over that user's items, negative uniform over the
non-train items (rejection).  This file restates elimrec_b200/csrc/sampler.cu bit for bit so the
kernel can be checked exactly; the distribution itself is checked statistically in tests.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32).copy() for x in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _u64(hi, lo):
    return (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)


def sample_triples(seed, epoch, n, user_ids, row_ptr, items, num_items, first=0):
    """samples first .. first+n-1 of the epoch's stream (elimrec_sample_batch_device: first = batch index * batch size)"""
    k = np.arange(first, first + n, dtype=np.uint64)
    c0, c1 = (k & MASK).astype(np.uint32), (k >> np.uint64(32)).astype(np.uint32)
    c3 = np.full(n, epoch & 0xFFFFFFFF, dtype=np.uint32)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    o = philox4x32_10(c0, c1, np.zeros(n, np.uint32), c3, k0, k1)
    s = (_u64(o[0], o[1]) % np.uint64(len(user_ids))).astype(np.int64)
    beg = row_ptr[s]
    deg = (row_ptr[s + 1] - beg).astype(np.uint64)
    pos = items[beg + (_u64(o[2], o[3]) % deg).astype(np.int64)]
    neg = np.full(n, -1, dtype=np.int64)
    todo = np.arange(n)
    j = 1
    while todo.size:
        o = philox4x32_10(c0[todo], c1[todo], np.full(todo.size, j, np.uint32), c3[todo], k0, k1)
        for h in range(2):
            a = (_u64(o[2 * h], o[2 * h + 1]) % np.uint64(num_items)).astype(np.int64)
            for t, idx in enumerate(todo):
                if neg[idx] < 0:
                    b, e = row_ptr[s[idx]], row_ptr[s[idx] + 1]
                    row = items[b:e]
                    p = np.searchsorted(row, a[t])
                    if not (p < row.size and row[p] == a[t]):
                        neg[idx] = a[t]
        todo = todo[neg[todo] < 0]
        j += 1
    return user_ids[s].astype(np.int64), pos.astype(np.int64), neg
