"""Normalised bipartite adjacency in the layout the SpMM kernel consumes.

Replaces ``create_adj_mat('pre')`` + the COO->torch conversion of the reference
(``models/EliMRec.py:309-354`` and ``:80-84``).  The reference builds one (U+I)^2 COO matrix; here
the two off-diagonal blocks are kept as separate CSR halves

    ui : user rows -> item columns   [U x I]      iu : item rows -> user columns   [I x U]

because every propagation layer only ever multiplies one block at a time (the graph is bipartite),
which is also what makes the 4-graph dedup possible.  Edge values for the default
``adj_type='pre'`` are ``fl32(fl32(d_r * 1) * d_c)`` with ``d = deg^-1/2`` computed in numpy fp32 exactly as
scipy does in the reference, so the arrays are bit-equal to the reference COO; the other ``adj_type``s
(plain / gcmc: still bipartite; norm / mean: plus a diagonal kept as a separate vector) follow the same rule
(tests/test_host_logic.py checks all five against the reference's own COO).

One-time, host-side (vectorised numpy); only the result lives on the device.
"""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sp
import torch

SEG_LEN = 96        # rows with at most this many edges are ONE warp work item
HEAVY_SEG_LEN = 96  # longer rows are split into a multiple of 8 segments of at most this many edges (8 per CTA, reduced in
                    # shared memory; rows needing several CTAs add a deterministic 2-stage reduce through scratch)
HEAVY_BALANCED = True   # a split row's edges are dealt EVENLY over its segments: no empty padding segments, and every row
                        # of up to 8 * HEAVY_SEG_LEN edges is exactly ONE CTA (no partials, no counter).
# Measured on B200 (ms / train step, tiktok | kwai | movielens shapes):
#   64, 16, fixed-length  0.682 | 1.137 | 0.815      (short segments: most mid-degree rows paid the 2-stage reduce)
#   64, 64, balanced      0.661 | 0.882 | 0.723
#   96, 96, balanced      0.661 | 0.875 | 0.717      128, 32, balanced  0.685 | 0.958 | 0.759
if os.environ.get("ELIMREC_SEG"):   # tuning sweeps: "seg_len,heavy_seg_len,balanced"
    _a, _b, _c = os.environ["ELIMREC_SEG"].split(",")
    SEG_LEN, HEAVY_SEG_LEN, HEAVY_BALANCED = int(_a), int(_b), bool(int(_c))


SEG64_LEN = 64      # work items of the 64-wide pair kernel (csrc/spmm64.cu): rows of at most this many edges are one item, longer
                    # rows are dealt evenly over ceil(deg / seg_len) items whose partial sums the last-arriving item adds up
SEG64_LEN_DENSE = 128       # ... for a CSR half whose MEAN degree exceeds SEG64_DENSE_DEGREE (most of its rows would be split at 64:
SEG64_DENSE_DEGREE = 32     # Movielens items 169, Kwai users 149): measured per dense launch 61.1 -> 56.1 us (Movielens shape),
                            # 41.8 -> 39.3 us (Kwai); the Tiktok shape (means 17 / 8) stays at 64 (36.9 us; 39.7 at 128, 55.6 at 256)


def seg64_len_for(indptr: np.ndarray, row_lo: int = 0, row_hi: int | None = None) -> int:
    """segment length of a CSR half's work list (ELIMREC_SEG64_LEN overrides, for experiments)"""
    if "ELIMREC_SEG64_LEN" in os.environ:
        return int(os.environ["ELIMREC_SEG64_LEN"])
    row_hi = indptr.size - 1 if row_hi is None else row_hi
    n = max(1, row_hi - row_lo)
    return SEG64_LEN_DENSE if (int(indptr[row_hi]) - int(indptr[row_lo])) / n > SEG64_DENSE_DEGREE else SEG64_LEN


def build_segments64(indptr: np.ndarray, seg_len: int = SEG64_LEN, row_lo: int = 0, row_hi: int | None = None):
    """Work list of elimrec_spmm64_pair for one CSR half: item = (row, edge_begin, edge_end, split_row_id or -1).
    Split rows first (longest first: their fixed-order reduction then overlaps the bulk of the launch), then the whole rows by
    DESCENDING degree - the four rows a warp works on in lock step have the same length, and the shortest rows form the tail.
    Returns (items [n x 4] int32, split_rows [h x 2] int32 = (first item, number of items), n_split_items)."""
    n_rows = indptr.size - 1
    row_hi = n_rows if row_hi is None else row_hi
    rows = np.arange(row_lo, row_hi, dtype=np.int64)
    beg = indptr[rows]
    deg = indptr[rows + 1] - beg
    heavy = deg > seg_len
    h_order = np.argsort(-deg[heavy], kind="stable")
    h_rows, h_beg, h_deg = rows[heavy][h_order], beg[heavy][h_order], deg[heavy][h_order]
    h_n = -(-h_deg // seg_len)
    h_len = -(-h_deg // np.maximum(h_n, 1))
    n_hitems = int(h_n.sum())
    parts = []
    hrow = np.zeros((h_rows.size, 2), dtype=np.int32)
    if h_rows.size:
        first = np.concatenate([[0], np.cumsum(h_n)[:-1]])
        hrow[:, 0], hrow[:, 1] = first, h_n
        k = np.arange(n_hitems) - np.repeat(first, h_n)
        ln = np.repeat(h_len, h_n)
        row_end = np.repeat(h_beg + h_deg, h_n)
        sb = np.minimum(np.repeat(h_beg, h_n) + k * ln, row_end)
        parts.append(np.stack([np.repeat(h_rows, h_n), sb, np.minimum(sb + ln, row_end), np.repeat(np.arange(h_rows.size), h_n)], axis=1))
    l_order = np.argsort(-deg[~heavy], kind="stable")
    l_rows, l_beg, l_deg = rows[~heavy][l_order], beg[~heavy][l_order], deg[~heavy][l_order]
    parts.append(np.stack([l_rows, l_beg, l_beg + l_deg, np.full(l_rows.size, -1)], axis=1))
    items = np.concatenate(parts, axis=0).astype(np.int32)
    return np.ascontiguousarray(items), hrow, n_hitems


class CsrHalf:
    """One CSR block + its segment work-list on the device."""

    def __init__(self, indptr: np.ndarray, indices: np.ndarray, vals: np.ndarray, n_cols: int, device,
                 seg_len: int = SEG_LEN, row_lo: int = 0, row_hi: int | None = None):
        self.n_rows = int(indptr.size - 1)
        self.n_cols = int(n_cols)
        self.nnz = int(indices.size)
        self.indptr_host = indptr.astype(np.int64)
        self.indices_host = indices.astype(np.int32)
        self.vals_host = vals.astype(np.float32)
        seg, heavy, n_hseg = build_segments(self.indptr_host, seg_len, row_lo, self.n_rows if row_hi is None else row_hi)
        self.n_seg = int(seg.shape[0])
        self.n_heavy_seg = int(n_hseg)
        self.seg = torch.from_numpy(seg).to(device)
        self.heavy = torch.from_numpy(heavy if heavy.size else np.zeros((1, 2), np.int32)).to(device)
        self.counter = torch.zeros(max(1, heavy.shape[0]), dtype=torch.int32, device=device)
        self.partial = torch.empty(max(1, n_hseg // 8) * 256, dtype=torch.float32, device=device)
        self.col = torch.from_numpy(self.indices_host).to(device)
        self.val = torch.from_numpy(self.vals_host).to(device)
        self.indptr = torch.from_numpy(self.indptr_host).to(device)
        # work list of the 64-wide pair kernel (its own split: 8-lane groups, no CTA padding)
        it, hrow, n_hit = build_segments64(self.indptr_host, seg64_len_for(self.indptr_host, row_lo, row_hi), row_lo,
                                           self.n_rows if row_hi is None else row_hi)
        self.n_item64, self.n_split64 = int(it.shape[0]), int(n_hit)
        self.item64 = torch.from_numpy(it if it.size else np.zeros((1, 4), np.int32)).to(device)
        self.hrow64 = torch.from_numpy(hrow if hrow.size else np.zeros((1, 2), np.int32)).to(device)
        self.counter64 = torch.zeros(max(1, hrow.shape[0]), dtype=torch.int32, device=device)
        self.partial64 = torch.empty(max(1, n_hit) * 64, dtype=torch.float32, device=device)

    def with_values(self, vals: np.ndarray) -> "CsrHalf":
        """Same structure, work-lists and scratch (never used concurrently with this one), different edge values."""
        import copy
        other = copy.copy(self)
        other.vals_host = vals.astype(np.float32)
        other.val = torch.from_numpy(other.vals_host).to(self.val.device)
        return other


def build_segments(indptr: np.ndarray, seg_len: int, row_lo: int = 0, row_hi: int | None = None,
                   heavy_seg_len: int | None = None):
    """seg[s] = (row, edge_begin, edge_end, heavy_id or -1); split rows first (longest first), each padded to a
    multiple of 8 segments so that a CTA of 8 warps never mixes rows; heavy[h] = (first CTA, number of CTAs)."""
    n_rows = indptr.size - 1
    row_hi = n_rows if row_hi is None else row_hi
    rows = np.arange(row_lo, row_hi, dtype=np.int64)
    beg = indptr[rows]
    deg = indptr[rows + 1] - beg
    hsl = min(seg_len, heavy_seg_len or HEAVY_SEG_LEN)
    heavy_mask = deg > seg_len
    nseg = np.where(heavy_mask, -(-deg // hsl), 1)
    # heavy rows, longest first so that their tails start early
    h_rows = rows[heavy_mask]
    order = np.argsort(-deg[heavy_mask], kind="stable")
    h_rows = h_rows[order]
    h_beg = beg[heavy_mask][order]
    h_deg = deg[heavy_mask][order]
    h_nseg = nseg[heavy_mask][order]
    h_nseg_pad = -(-h_nseg // 8) * 8            # one CTA (8 warps) = 8 segments of ONE row; pad with empty segments
    h_len = np.full(h_rows.size, hsl, dtype=np.int64)
    if HEAVY_BALANCED:                          # ... or none: equal shares of the row for all of its segments
        h_len = -(-h_deg // h_nseg_pad)
    n_hseg = int(h_nseg_pad.sum())
    segs = []
    heavy = np.zeros((h_rows.size, 2), dtype=np.int32)
    if h_rows.size:
        first = np.concatenate([[0], np.cumsum(h_nseg_pad)[:-1]])
        heavy[:, 0] = first // 8                 # first CTA (= partial slot) of the row
        heavy[:, 1] = h_nseg_pad // 8            # CTAs of the row
        hid = np.repeat(np.arange(h_rows.size), h_nseg_pad)
        k = np.arange(n_hseg) - np.repeat(first, h_nseg_pad)
        row_end = np.repeat(h_beg + h_deg, h_nseg_pad)
        ln = np.repeat(h_len, h_nseg_pad)
        sb = np.minimum(np.repeat(h_beg, h_nseg_pad) + k * ln, row_end)
        se = np.minimum(sb + ln, row_end)
        segs.append(np.stack([np.repeat(h_rows, h_nseg_pad), sb, se, hid], axis=1))
    l_rows = rows[~heavy_mask]
    l_beg = beg[~heavy_mask]
    segs.append(np.stack([l_rows, l_beg, l_beg + deg[~heavy_mask], np.full(l_rows.size, -1)], axis=1))
    seg = np.concatenate(segs, axis=0).astype(np.int32)
    return np.ascontiguousarray(seg), heavy, n_hseg


ADJ_TYPES = ("plain", "norm", "gcmc", "pre", "mean")


def normalized_halves(train_csr: sp.csr_matrix, adj_type: str = "pre"):
    """Both off-diagonal blocks of the propagation matrix of ``create_adj_mat`` (models/EliMRec.py:309-354), values
    bit-equal to the reference COO.  Returns a dict:

        ui, iu        (indptr, indices, vals) of the user-row / item-row block of A_hat
        ui_t, iu_t    vals of the same blocks of A_hat^T (the backward pass), or None when A_hat is symmetric
        self_u/self_i the diagonal of A_hat per user / item (adj_type 'norm' / 'mean'), or None

    'plain': A.  'pre': D^-1/2 A D^-1/2.  'gcmc': D^-1 A.  'norm': (D+I)^-1 (A+I).  'mean' (the reference's else
    branch, any other string): D^-1 A + I.  scipy's dtypes are followed: float32 throughout, except 'norm' whose
    ``sp.eye`` promotes the row sums to float64 before the final float32 cast (EliMRec.py:81-83)."""
    m = train_csr.tocsr().astype(np.float32)
    m.sum_duplicates()
    m.sort_indices()
    m.data[:] = 1.0
    mt = m.T.tocsr()
    mt.sort_indices()
    deg_u = np.asarray(m.sum(1), dtype=np.float32).ravel()
    deg_i = np.asarray(mt.sum(1), dtype=np.float32).ravel()
    row_u = np.repeat(np.arange(m.shape[0]), np.diff(m.indptr))
    row_i = np.repeat(np.arange(mt.shape[0]), np.diff(mt.indptr))
    one = np.float32(1.0)
    out = dict(ui_t=None, iu_t=None, self_u=None, self_i=None)

    def inv(deg, power, dtype=np.float32):
        with np.errstate(divide="ignore"):
            d = np.power(deg.astype(dtype), dtype(power) if power != -1 else -1)
        d[np.isinf(d)] = 0.0
        return d

    if adj_type == "pre":
        du, di = inv(deg_u, -0.5), inv(deg_i, -0.5)
        val_ui = ((du[row_u] * one) * di[m.indices]).astype(np.float32)
        val_iu = ((di[row_i] * one) * du[mt.indices]).astype(np.float32)
    elif adj_type == "plain":
        val_ui = np.ones(m.nnz, dtype=np.float32)
        val_iu = np.ones(mt.nnz, dtype=np.float32)
    else:
        if adj_type == "norm":      # rows of A + I, normalised in float64, stored as float32
            du = inv(deg_u.astype(np.float64) + 1.0, -1, np.float64).astype(np.float32)
            di = inv(deg_i.astype(np.float64) + 1.0, -1, np.float64).astype(np.float32)
            out["self_u"], out["self_i"] = du.copy(), di.copy()
        else:                       # 'gcmc', and 'mean' = gcmc + I
            du, di = inv(deg_u, -1), inv(deg_i, -1)
            if adj_type != "gcmc":
                out["self_u"] = np.ones(m.shape[0], dtype=np.float32)
                out["self_i"] = np.ones(mt.shape[0], dtype=np.float32)
        val_ui = (du[row_u] * one).astype(np.float32)       # A_hat[u, i] = d_u
        val_iu = (di[row_i] * one).astype(np.float32)       # A_hat[i, u] = d_i
        out["ui_t"] = di[m.indices].astype(np.float32)      # A_hat^T[u, i] = A_hat[i, u]
        out["iu_t"] = du[mt.indices].astype(np.float32)
    out["ui"] = (m.indptr, m.indices, val_ui)
    out["iu"] = (mt.indptr, mt.indices, val_iu)
    return out


class BipartiteGraph:
    def __init__(self, train_csr: sp.csr_matrix, device, adj_type: str = "pre", seg_len: int = SEG_LEN,
                 user_rows=None, item_rows=None):
        self.num_users, self.num_items = train_csr.shape
        self.adj_type = adj_type
        h = normalized_halves(train_csr, adj_type)
        (pu, iu_, vu), (pi, ii_, vi) = h["ui"], h["iu"]
        ur = user_rows or (0, self.num_users)
        ir = item_rows or (0, self.num_items)
        self.ui = CsrHalf(pu, iu_, vu, self.num_items, device, seg_len, *ur)
        self.iu = CsrHalf(pi, ii_, vi, self.num_users, device, seg_len, *ir)
        # blocks of A_hat^T for the backward pass: same structure (and work-lists), other values when A_hat is asymmetric
        self.ui_t = self.ui if h["ui_t"] is None else self.ui.with_values(h["ui_t"])
        self.iu_t = self.iu if h["iu_t"] is None else self.iu.with_values(h["iu_t"])
        self.symmetric = h["ui_t"] is None
        # diagonal of A_hat ('norm' / 'mean'): one vector over all N nodes, users first
        self.self_loops = h["self_u"] is not None
        self.self_host = np.concatenate([h["self_u"], h["self_i"]]).astype(np.float32) if self.self_loops else None
        self.self_all = torch.from_numpy(self.self_host).to(device) if self.self_loops else None
        self.nnz = self.ui.nnz + self.iu.nnz

    def as_coo(self):
        """(row, col, val) in the reference's (U+I)^2 indexing, row-major, columns sorted - for parity tests."""
        U = self.num_users
        ru = np.repeat(np.arange(U), np.diff(self.ui.indptr_host))
        ri = np.repeat(np.arange(self.num_items), np.diff(self.iu.indptr_host)) + U
        row = np.concatenate([ru, ri]).astype(np.int64)
        col = np.concatenate([self.ui.indices_host.astype(np.int64) + U, self.iu.indices_host.astype(np.int64)])
        val = np.concatenate([self.ui.vals_host, self.iu.vals_host])
        if self.self_loops:
            n = np.arange(U + self.num_items, dtype=np.int64)
            row, col, val = np.concatenate([row, n]), np.concatenate([col, n]), np.concatenate([val, self.self_host])
            order = np.lexsort((col, row))
            row, col, val = row[order], col[order], val[order]
        return row, col, val
