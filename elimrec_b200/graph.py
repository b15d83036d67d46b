"""Normalised bipartite adjacency in the layout the SpMM kernel consumes.

Replaces ``create_adj_mat('pre')`` + the COO->torch conversion of the reference
(``models/EliMRec.py:309-354`` and ``:80-84``).  The reference builds one (U+I)^2 COO matrix; here
the two off-diagonal blocks are kept as separate CSR halves

    ui : user rows -> item columns   [U x I]      iu : item rows -> user columns   [I x U]

because every propagation layer only ever multiplies one block at a time (the graph is bipartite),
which is also what makes the 4-graph dedup possible.  Edge values are
``fl32(fl32(d_r * 1) * d_c)`` with ``d = deg^-1/2`` computed in numpy fp32 exactly as scipy does in
the reference, so the arrays are bit-equal to the reference COO (tests/test_graph.py).

One-time, host-side (vectorised numpy); only the result lives on the device.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch

SEG_LEN = 64        # rows with at most this many edges are ONE warp work item
HEAVY_SEG_LEN = 16  # longer rows are split into segments of this many edges (8 per CTA, deterministic 2-stage reduce):
                    # short segments keep the per-warp dependent-load chain short and the launch at full occupancy


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
        sb = np.minimum(np.repeat(h_beg, h_nseg_pad) + k * hsl, row_end)
        se = np.minimum(sb + hsl, row_end)
        segs.append(np.stack([np.repeat(h_rows, h_nseg_pad), sb, se, hid], axis=1))
    l_rows = rows[~heavy_mask]
    l_beg = beg[~heavy_mask]
    segs.append(np.stack([l_rows, l_beg, l_beg + deg[~heavy_mask], np.full(l_rows.size, -1)], axis=1))
    seg = np.concatenate(segs, axis=0).astype(np.int32)
    return np.ascontiguousarray(seg), heavy, n_hseg


def normalized_halves(train_csr: sp.csr_matrix, adj_type: str = "pre"):
    """(indptr, indices, vals) of both blocks of D^-1/2 A D^-1/2; bit-equal to the reference COO."""
    if adj_type != "pre":
        raise NotImplementedError(f"adj_type={adj_type!r}: only 'pre' (the conf/EliMRec.properties default) is built; "
                                  "plain/norm/gcmc/mean are SURVEY.md row f2")
    m = train_csr.tocsr().astype(np.float32)
    m.sum_duplicates()
    m.sort_indices()
    m.data[:] = 1.0
    mt = m.T.tocsr()
    mt.sort_indices()
    deg_u = np.asarray(m.sum(1), dtype=np.float32).ravel()
    deg_i = np.asarray(mt.sum(1), dtype=np.float32).ravel()
    with np.errstate(divide="ignore"):
        du = np.power(deg_u, np.float32(-0.5)).astype(np.float32)
        di = np.power(deg_i, np.float32(-0.5)).astype(np.float32)
    du[np.isinf(du)] = 0.0
    di[np.isinf(di)] = 0.0
    one = np.float32(1.0)
    row_u = np.repeat(np.arange(m.shape[0]), np.diff(m.indptr))
    val_ui = ((du[row_u] * one) * di[m.indices]).astype(np.float32)
    row_i = np.repeat(np.arange(mt.shape[0]), np.diff(mt.indptr))
    val_iu = ((di[row_i] * one) * du[mt.indices]).astype(np.float32)
    return (m.indptr, m.indices, val_ui), (mt.indptr, mt.indices, val_iu)


class BipartiteGraph:
    def __init__(self, train_csr: sp.csr_matrix, device, adj_type: str = "pre", seg_len: int = SEG_LEN,
                 user_rows=None, item_rows=None):
        self.num_users, self.num_items = train_csr.shape
        (pu, iu_, vu), (pi, ii_, vi) = normalized_halves(train_csr, adj_type)
        ur = user_rows or (0, self.num_users)
        ir = item_rows or (0, self.num_items)
        self.ui = CsrHalf(pu, iu_, vu, self.num_items, device, seg_len, *ur)
        self.iu = CsrHalf(pi, ii_, vi, self.num_users, device, seg_len, *ir)
        self.nnz = self.ui.nnz + self.iu.nnz

    def as_coo(self):
        """(row, col, val) in the reference's (U+I)^2 indexing, row-major - for parity tests."""
        U = self.num_users
        ru = np.repeat(np.arange(U), np.diff(self.ui.indptr_host))
        ri = np.repeat(np.arange(self.num_items), np.diff(self.iu.indptr_host)) + U
        row = np.concatenate([ru, ri]).astype(np.int64)
        col = np.concatenate([self.ui.indices_host.astype(np.int64) + U, self.iu.indices_host.astype(np.int64)])
        val = np.concatenate([self.ui.vals_host, self.iu.vals_host])
        return row, col, val
