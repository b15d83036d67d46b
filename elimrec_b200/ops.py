"""Tensor-level wrappers of the C-ABI (one Python function per ``elimrec_*`` entry point).

Every wrapper takes CUDA tensors, passes raw device pointers + the current torch stream, and
raises on failure.  No wrapper allocates on the hot path; callers pass preallocated outputs.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import MeanEpilogue, RankTables, call, ptr, stream

F32 = torch.float32

# ---- two-stream fork/join: a layer's WIDE and NARROW SpMM are independent and run concurrently (the narrow one is
# latency-bound and hides under the bandwidth-bound wide one); under CUDA-graph capture these become parallel branches.
_SIDE = {}


def fork_side(slot=0, high_priority=False):
    """A side stream that starts after everything queued on the current one.  ``high_priority``: for branches of SMALL kernels
    that run beside a propagation launch - their CTAs are then placed as soon as SM slots free up instead of queueing behind
    the big grid (stream priority is kept by the nodes of a captured CUDA graph)."""
    cur = torch.cuda.current_stream()
    if _lib.PROFILE["on"]:      # per-kernel event timing needs the launches serialised
        return cur
    key = (cur.device.index, cur.cuda_stream, slot)
    side = _SIDE.get(key)
    if side is None:
        side = _SIDE[key] = torch.cuda.Stream(device=cur.device, priority=-1 if high_priority else 0)
    ev = torch.cuda.Event()
    ev.record(cur)
    side.wait_event(ev)
    return side


def join_side(side):
    if side == torch.cuda.current_stream():
        return
    ev = torch.cuda.Event()
    ev.record(side)
    torch.cuda.current_stream().wait_event(ev)



def spmm(half, X: torch.Tensor, Y, width: int, epi: MeanEpilogue | None = None, row_mask=None, col_mask=None, density=50,
         addend=None, add_mask=None):
    """Y[rows of half] = A_half @ X  (+ optional fused layer-mean epilogue).  The split (Zipf-head) rows run as their
    own launch on a side stream, concurrently with the whole rows.  ``row_mask`` / ``col_mask`` (uint8 per row / column
    of this half): the row-sparse last-layer variant, elimrec_spmm_masked; ``density`` = expected % of marked rows.
    ``addend`` / ``add_mask``: Y[row] += addend[row] on the rows add_mask marks (fused layer-mean gradient)."""
    masked = row_mask is not None or col_mask is not None or addend is not None

    def go(part, launches):
        args = [width, part, half.n_seg, half.n_heavy_seg, ptr(half.seg), ptr(half.heavy), ptr(half.counter),
                ptr(half.col), ptr(half.val), ptr(X, F32), X.stride(0), ptr(Y, F32, True), (Y.stride(0) if Y is not None else 0),
                ptr(half.partial), (C.byref(epi) if epi is not None else None)]
        if masked:
            call("elimrec_spmm_masked", *args, ptr(row_mask, torch.uint8, True), ptr(col_mask, torch.uint8, True), int(density),
                 ptr(addend, F32, True), (addend.stride(0) if addend is not None else 0), ptr(add_mask, torch.uint8, True),
                 stream(), launches=launches, tag=f"spmm{width}m")
        else:
            # "w": the half has no split rows -> ONE launch of the whole-row kernel, timed exactly by its event pair
            call("elimrec_spmm", *args, stream(), launches=launches, tag=f"spmm{width}" + ("w" if half.n_heavy_seg == 0 else ""))

    if half.n_heavy_seg == 0:
        go(0, 1)
        return
    prof = _lib.PROFILE["on"]
    if prof:      # one event pair around the whole op (both launches, concurrent as in the step)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.PROFILE["on"] = False
    side = fork_side(1)
    with torch.cuda.stream(side):
        go(1, 1)
    go(2, 1)
    join_side(side)
    if prof:
        e1.record()
        _lib.PROFILE["on"] = True
        _lib.PROFILE["events"].append((f"spmm{width}m" if masked else f"spmm{width}", e0, e1))


def spmm64_half(h, X, Y, row_mask=None, col_mask=None, addend=None, add_mask=None, adam=None):
    d = _lib.Spmm64Half()
    d.n_item, d.n_split_item, d.item, d.split_rows = h.n_item64, h.n_split64, ptr(h.item64), ptr(h.hrow64)
    d.counter, d.partial, d.col, d.val = ptr(h.counter64), ptr(h.partial64), ptr(h.col), ptr(h.val)
    d.X, d.ldx, d.Y, d.ldy = ptr(X, F32), X.stride(0), ptr(Y, F32, True), (Y.stride(0) if Y is not None else 0)
    d.row_mask, d.col_mask = ptr(row_mask, torch.uint8, True), ptr(col_mask, torch.uint8, True)
    d.addend, d.ld_add, d.add_mask = ptr(addend, F32, True), (addend.stride(0) if addend is not None else 0), ptr(add_mask, torch.uint8, True)
    if adam is not None:
        p, m, v, old = adam
        for t in (p, m, v) + ((old,) if old is not None else ()):
            if not t.is_contiguous():
                raise _lib.ElimrecError("fused Adam takes contiguous [rows x width] tensors")
        d.adam_param, d.adam_exp_avg, d.adam_exp_avg_sq, d.adam_old_out = ptr(p, F32), ptr(m, F32), ptr(v, F32), ptr(old, F32, True)
    return d


def spmm64_pair(half_u, half_i, X_for_u, X_for_i, Y_u, Y_i, row_mask_u=None, row_mask_i=None, col_mask_u=None, col_mask_i=None,
                addend_u=None, addend_i=None, add_mask_u=None, add_mask_i=None, adam_u=None, adam_i=None, adam_consts=None,
                variant=0, width=64):
    """Both halves of a 64-wide (``width``: 64 / 32 / 16 / 8 columns per rank when column-sharded) propagation layer in ONE launch (elimrec_spmm64_pair): Y_u = half_u @ X_for_u (user rows gather
    item rows), Y_i = half_i @ X_for_i.  ``col_mask_u``: mask over the COLUMNS of half_u (item rows), etc.  ``addend_*`` /
    ``add_mask_*``: Y[row] += addend[row] on the marked rows (the gradient entering this layer of the backward chain).
    ``adam_u`` / ``adam_i`` = (param, exp_avg, exp_avg_sq, old_out or None) with ``adam_consts`` = (consts_dev, beta1, beta2, eps,
    weight_decay): the finished rows are gradients of those tables and Adam is applied in the epilogue (Y_* may be None)."""
    da = spmm64_half(half_u, X_for_u, Y_u, row_mask_u, col_mask_u, addend_u, add_mask_u, adam_u)
    db = spmm64_half(half_i, X_for_i, Y_i, row_mask_i, col_mask_i, addend_i, add_mask_i, adam_i)
    ac = None
    if adam_consts is not None:
        ac = _lib.AdamConsts()
        ac.consts_dev, ac.beta1, ac.beta2, ac.eps, ac.weight_decay = (ptr(adam_consts[0], torch.float64), adam_consts[1], adam_consts[2],
                                                                     adam_consts[3], adam_consts[4])
    if X_for_u.shape[-1] < width or X_for_i.shape[-1] < width:
        raise _lib.ElimrecError("spmm64_pair: operand narrower than the propagated width")
    call("elimrec_spmm64_pair", C.byref(da), C.byref(db), (C.byref(ac) if ac is not None else None), int(width), int(variant), stream(),
         tag="spmm64_pair" + ("+adam" if ac is not None else ("_masked" if any(m is not None for m in (
             row_mask_u, row_mask_i, col_mask_u, col_mask_i)) else "")))


def mark_rows(rows, mask):
    call("elimrec_mark_rows", rows.numel(), ptr(rows, torch.int32), mask.numel(), ptr(mask, torch.uint8), stream(), launches=2)


def inst_rows(users, pos, neg, num_users, rows, mask=None, mask2=None):
    call("elimrec_inst_rows", users.numel(), ptr(users, torch.int64), ptr(pos, torch.int64), ptr(neg, torch.int64), num_users,
         ptr(rows, torch.int32), (mask.numel() if mask is not None else 0), ptr(mask, torch.uint8, True),
         ptr(mask2, torch.uint8, True), stream(), launches=1 + (mask is not None) + (mask2 is not None))


def mark_neighbors(half, row_mask, out_mask):
    """out_mask[col] = 1 for every edge of the rows of ``half`` with row_mask != 0."""
    call("elimrec_mark_neighbors", half.n_seg, ptr(half.seg), ptr(half.col), ptr(row_mask, torch.uint8),
         ptr(out_mask, torch.uint8), stream())


def zero_rows(rows, lo, hi, off, dst, width):
    call("elimrec_zero_rows", rows.numel(), ptr(rows, torch.int32), lo, hi, off, ptr(dst, F32), dst.stride(0), width, stream())


def mean_epilogue(prev, out: torch.Tensor, width: int, scale: float) -> MeanEpilogue:
    e = MeanEpilogue()
    e.n_prev = len(prev)
    for k, (t, w) in enumerate(prev):
        e.prev[k] = ptr(t, F32)
        e.prev_ld[k] = t.stride(0)
        e.prev_width[k] = w
    e.mean_out = ptr(out, F32)
    e.mean_ld = out.stride(0)
    e.mean_width = width
    e.mean_scale = scale
    e._keepalive = (prev, out)
    return e


def scatter_add_rows(rows, lo, hi, off, src, src_width, dst, width, scale):
    call("elimrec_scatter_add_rows", rows.numel(), ptr(rows, torch.int32), lo, hi, off, ptr(src, F32), src.stride(0),
         src_width, ptr(dst, F32), dst.stride(0), width, scale, stream())


def gather_rows(rows, src, dst, width):
    call("elimrec_gather_rows", rows.numel(), ptr(rows, torch.int32), ptr(src, F32), src.stride(0), ptr(dst, F32),
         dst.stride(0), width, stream())


def copy_2d(src, dst, n_rows, width):
    call("elimrec_copy_2d", n_rows, width, ptr(src, F32), src.stride(0), ptr(dst, F32), dst.stride(0), stream())


def broadcast_cols(src, dst, n_rows, n_rep):
    """dst[r, 64g + c] = src[r, c] for g < n_rep."""
    call("elimrec_broadcast_cols", n_rows, ptr(src, F32), src.stride(0), ptr(dst, F32), dst.stride(0), n_rep, stream())


def tie_blocks(src, dst, n_rep, scale):
    """dst[r, 64g + c] = scale * src[r, c]  (tied weight of mm_fusion_mode='mean')."""
    call("elimrec_tie_blocks", src.shape[0], ptr(src, F32), src.stride(0), ptr(dst, F32), dst.stride(0), n_rep, scale, stream())


def fold_blocks(src, dst, n_rep, scale=1.0):
    """dst[r, c] = scale * sum_g src[r, 64g + c]."""
    call("elimrec_fold_blocks", src.shape[0], ptr(src, F32), src.stride(0), n_rep, scale, ptr(dst, F32), dst.stride(0), stream())


def axpy_rows(row_scale, X, Y, width):
    """Y[r, :width] += row_scale[r] * X[r, :width]  (self-loop term of adj_type 'norm' / 'mean')."""
    call("elimrec_axpy_rows", X.shape[0], width, ptr(row_scale, F32), ptr(X, F32), X.stride(0), ptr(Y, F32), Y.stride(0),
         stream())


def layer_mean(layers, out, width, scale):
    n = len(layers)
    lp = (C.c_void_p * n)(*[ptr(t, F32) for t in layers])
    ld = (C.c_int64 * n)(*[t.stride(0) for t in layers])
    call("elimrec_layer_mean", out.shape[0], width, n, lp, ld, scale, ptr(out, F32), out.stride(0), stream())


def lin_layers(tables):
    """tables: list over k = 0..L of (user rows [U x 64], item rows [I x 64]) of p_k = A_hat^k [E_u ; E_i]"""
    lay = _lib.LinLayers()
    lay.n = len(tables)
    for k, (tu, ti) in enumerate(tables):
        lay.user[k], lay.user_ld[k] = ptr(tu, F32), tu.stride(0)
        lay.item[k], lay.item_ld[k] = ptr(ti, F32), ti.stride(0)
    lay._keepalive = tables
    return lay


def lin_assemble(rows, num_users, layers, scale, n_mod, accumulate, out, n_rows=None):
    """out[j, 0:64] = scale * sum_k p_k[rows[j]];  out[j, 64(1+m):64(2+m)] (+)= scale * (parity layers) - elimrec_lin_assemble.
    rows None: out row j = node j (n_rows of them)."""
    n = rows.numel() if rows is not None else n_rows
    call("elimrec_lin_assemble", n, ptr(rows, torch.int32, True), num_users, C.byref(layers), scale, n_mod, int(accumulate),
         ptr(out, F32), out.stride(0), stream())


def lin_seed(rows, num_users, layer, dO, n_mod, scale, dst):
    """dst[rows[j]] += scale * (dO[j, :64] + [parity] * sum_m dO[j, 64(1+m):64(2+m)])  (adjoint of lin_assemble, layer `layer`)"""
    call("elimrec_lin_seed", rows.numel(), ptr(rows, torch.int32), num_users, layer, ptr(dO, F32), dO.stride(0), n_mod, scale,
         ptr(dst, F32), dst.stride(0), stream())


def lin_seed2(rows, dO, n_mod, scale, GA, GB):
    """GA[rows[j]] += scale * (sum of all 64-column blocks of dO[j]);  GB[rows[j]] += scale * dO[j, :64]"""
    call("elimrec_lin_seed2", rows.numel(), ptr(rows, torch.int32), ptr(dO, F32), dO.stride(0), n_mod, scale, ptr(GA, F32),
         ptr(GB, F32), GA.stride(0), stream())


def pack_proj_weights(items, round_tf32):
    """items: list of (W [64 x Dm], b [64], dst [64 x Kp]) -> dst = [W | b | 0..], one launch"""
    arr = (_lib.PackProj * len(items))()
    for k, (W, b, dst) in enumerate(items):
        arr[k].W, arr[k].b, arr[k].dst, arr[k].Dm, arr[k].Kp = ptr(W, F32), ptr(b, F32), ptr(dst, F32), W.shape[1], dst.shape[1]
    call("elimrec_pack_proj_weights", len(items), arr, int(round_tf32), stream())


def split3_rows(src, dst, n_rows, width, pattern):
    """dst[3 n_rows x width] = [hi ; hi ; lo] (pattern 0) / [hi ; lo ; hi] (pattern 1) TF32 parts of src"""
    call("elimrec_split3_rows", n_rows, width, ptr(src, F32), src.stride(0), ptr(dst, F32), dst.stride(0), pattern, stream())


def axpy_2d(X, Y, n_rows, width, scale=1.0, accumulate=True):
    """Y[:, :width] = [Y +] scale * X[:, :width]"""
    call("elimrec_axpy_2d", n_rows, width, scale, ptr(X, F32), X.stride(0), ptr(Y, F32), Y.stride(0), int(accumulate), stream())


def gemm(M, N, K, A, a_sm, a_sk, B, b_sk, b_sn, Cm, c_sm, c_sn, bias=None, accumulate=False, split_k=1, ws=None,
         scale=None, a_off=0, b_off=0, c_off=0, tag=None):
    """C(m,n) = [C +] scale * sum_k A(m,k) B(k,n) [+ bias(n)]; *_off are element offsets into the tensors."""
    if split_k > 1:
        need = split_k * M * N
        if ws is None or ws.numel() < need:
            raise _lib.ElimrecError(f"gemm workspace too small: need {need} floats")
    call("elimrec_gemm", M, N, K, ptr(A, F32) + 4 * a_off, a_sm, a_sk, ptr(B, F32) + 4 * b_off, b_sk, b_sn,
         ptr(Cm, F32) + 4 * c_off, c_sm, c_sn, ptr(bias, F32, True), int(accumulate), split_k, ptr(ws, F32, True),
         ptr(scale, F32, True), stream(), launches=(2 if split_k > 1 else 1), tag=tag or "gemm")


def linear_tf32_fwd(X, W, b, Y, col=0, tag="proj_fwd_tc"):
    """Y[:, col:col+64] = X @ W^T + b on the tcgen05 tensor cores (TF32 inputs, fp32 accumulate).
    X may be a column-block view (row stride != K)."""
    M, K = X.shape
    call("elimrec_linear_tf32_fwd", M, K, ptr(X, F32), X.stride(0), ptr(W, F32), ptr(b, F32, True),
         ptr(Y, F32) + 4 * col, Y.stride(0), stream(), tag=tag)


def linear_tf32_fwd_multi(problems, tag="proj_fwd_tc"):
    """problems: list of (X, W, b, Y, col) - all in ONE persistent launch (grid <= SM count)."""
    arr = (_lib.LinearDesc * len(problems))()
    for k, (X, W, b, Y, col) in enumerate(problems):
        arr[k].M, arr[k].K = X.shape
        arr[k].X, arr[k].ldx, arr[k].W, arr[k].b = ptr(X, F32), X.stride(0), ptr(W, F32), ptr(b, F32, True)
        arr[k].Y, arr[k].ldy = ptr(Y, F32) + 4 * col, Y.stride(0)
    call("elimrec_linear_tf32_fwd_multi", len(problems), arr, stream(), tag=tag)


def prep_weights_tf32(items):
    """items: list of (src, hi, lo_or_None) - all rounded / split in one launch."""
    arr = (_lib.PrepTensor * len(items))()
    for k, (src, hi, lo) in enumerate(items):
        arr[k].src, arr[k].hi, arr[k].lo, arr[k].numel = ptr(src, F32), ptr(hi, F32), ptr(lo, F32, True), src.numel()
    call("elimrec_prep_weights_tf32", len(items), arr, stream())


def fuse_heads_x3(O_rows, Wf_hi, Wf_lo, bf, Ws_hi, Ws_lo, bs, F_out, S_out, tag="fuse_heads_x3"):
    """F_out = O_rows @ Wf^T + bf and S_out[m] = O_rows[:, 64(m+1):64(m+2)] @ Ws[m]^T + bs[m] in ONE pass over O_rows."""
    n = len(Ws_hi)
    arr = lambda ts: (C.c_void_p * 3)(*([ptr(t, F32) for t in ts] + [None] * (3 - n)))
    call("elimrec_fuse_heads_x3", O_rows.shape[0], n, ptr(O_rows, F32), O_rows.stride(0), ptr(Wf_hi, F32), ptr(Wf_lo, F32),
         ptr(bf, F32), arr(Ws_hi), arr(Ws_lo), arr(bs), ptr(F_out, F32), arr(S_out), stream(), tag=tag)


def fuse_heads_x3_all(U, I, O, Wu, bu, Wi, bi, Ws_hi, Ws_lo, bs, F_out, S_out, tag="fuse_heads_x3_all"):
    """Persistent one-pass fusion Linear + heads over the whole slab O [(U+I) x F]; Wu/Wi are (hi, lo) pairs."""
    n = len(Ws_hi)
    arr = lambda ts: (C.c_void_p * 3)(*([ptr(t, F32) for t in ts] + [None] * (3 - n)))
    call("elimrec_fuse_heads_x3_all", U, I, n, ptr(O, F32), O.stride(0), ptr(Wu[0], F32), ptr(Wu[1], F32), ptr(bu, F32),
         ptr(Wi[0], F32), ptr(Wi[1], F32), ptr(bi, F32), arr(Ws_hi), arr(Ws_lo), arr(bs), ptr(F_out, F32), arr(S_out), stream(),
         tag=tag)


def split_tf32(src, hi, lo):
    call("elimrec_split_tf32", src.numel(), ptr(src, F32), ptr(hi, F32), ptr(lo, F32), stream())


def linear_x3_fwd(X, W_hi, W_lo, b, Y, col=0, tag="linear_x3"):
    """Y[:, col:col+64] = X @ W^T + b with 3xTF32 on tcgen05 (fp32-class accuracy); X may be a strided column block."""
    M, K = X.shape
    call("elimrec_linear_x3_fwd", M, K, ptr(X, F32), X.stride(0), ptr(W_hi, F32), ptr(W_lo, F32), ptr(b, F32, True),
         ptr(Y, F32) + 4 * col, Y.stride(0), stream(), tag=tag)


def round_tf32(src, dst):
    call("elimrec_round_tf32", src.numel(), ptr(src, F32), ptr(dst, F32), stream())


def linear_tf32_wgrad(dY, X, dW, ws, col=0, tag="proj_wgrad_tc"):
    """dW[64 x K] = dY[:, col:col+64]^T @ X on the tcgen05 tensor cores; deterministic row-range reduction."""
    M, K = X.shape
    call("elimrec_linear_tf32_wgrad", M, K, ptr(dY, F32) + 4 * col, dY.stride(0), ptr(X, F32), X.stride(0), ptr(dW, F32),
         ptr(ws, F32), stream(), launches=2, tag=tag)


def linear_tf32_wgrad_ws_floats(M, K):
    return int(_lib.lib().elimrec_linear_tf32_wgrad_workspace_floats(M, K))


def colsum(M, N, A, ld, out, ws, accumulate=False, scale=None, a_off=0):
    call("elimrec_colsum", M, N, ptr(A, F32) + 4 * a_off, ld, ptr(out, F32), ptr(ws, F32), int(accumulate),
         ptr(scale, F32, True), stream())


def colsum_ws_floats(M, N):
    return int(_lib.lib().elimrec_colsum_workspace_floats(M, N))


def bpr(tables, weights, users, pos, neg, num_users, loss_out, inst_rows, inst_grad, ws, part=3):
    """part 1 = per-triple terms + instance gradients, part 2 = their fixed-order reduction to the loss scalar, 3 = both"""
    n = len(tables)
    tp = (C.c_void_p * n)(*[ptr(t, F32) for t in tables])
    wp = (C.c_float * n)(*weights)
    call("elimrec_bpr_forward_backward_part", part, users.numel(), n, tp, wp, ptr(users, torch.int64), ptr(pos, torch.int64),
         ptr(neg, torch.int64), num_users, ptr(loss_out, F32), ptr(inst_rows, torch.int32), ptr(inst_grad, F32),
         ptr(ws, F32), stream(), launches={1: 1, 2: 1, 3: 2}[part], tag="elimrec_bpr_forward_backward")


def inst_dO_seed(B, nt, F, inst_grad, gscale, Wu, Wi, Ws, dO_inst, rows, n_mod, scale, GA, GB):
    """d O[inst] with the linear schedule's backward seeds in its epilogue (elimrec_inst_dout_seed): inst_backward(part=1)
    followed by lin_seed2, in one launch"""
    wsp = (C.c_void_p * max(nt - 1, 1))(*[ptr(t, F32) for t in Ws])
    call("elimrec_inst_dout_seed", B, nt, F, ptr(inst_grad, F32), ptr(gscale, F32, True), ptr(Wu, F32), ptr(Wi, F32), wsp,
         ptr(dO_inst, F32), ptr(rows, torch.int32), n_mod, scale, ptr(GA, F32), ptr(GB, F32), GA.stride(0), stream(), launches=1,
         tag="inst_dO_seed")


def inst_backward_ws_floats(B, nt, F):
    return int(_lib.lib().elimrec_inst_backward_workspace_floats(B, nt, F))


def inst_backward(B, nt, F, inst_grad, O_inst, gscale, Wu, Wi, Ws, dO_inst, dWu, dWi, dbu, dbi, dWs, dbs, ws, part=3):
    """Backward of the fusion Linear + heads on the 3B instance rows (3 launches).  part 1 = d O[inst] only (1 launch),
    part 2 = weight / bias gradients only (2 launches)."""
    n = nt - 1
    wsp = (C.c_void_p * max(n, 1))(*[ptr(t, F32) for t in Ws])
    dwp = (C.c_void_p * max(n, 1))(*[ptr(t, F32) for t in dWs])
    dbp = (C.c_void_p * max(n, 1))(*[ptr(t, F32) for t in dbs])
    call("elimrec_inst_backward_part", part, B, nt, F, ptr(inst_grad, F32), ptr(O_inst, F32), ptr(gscale, F32, True),
         ptr(Wu, F32), ptr(Wi, F32), wsp, ptr(dO_inst, F32), ptr(dWu, F32), ptr(dWi, F32), ptr(dbu, F32), ptr(dbi, F32), dwp,
         dbp, ptr(ws, F32), stream(), launches={1: 1, 2: 2, 3: 3}[part], tag=f"inst_backward{part}")


def inst_forward(B, nt, F, O_inst, Wu, Wi, Ws, bu, bi, bs, F_out, S_out):
    """fusion Linear + heads on the 3B instance rows (users first), exact fp32, one launch (elimrec_inst_forward)"""
    n = nt - 1
    arr = lambda ts: (C.c_void_p * max(n, 1))(*[ptr(t, F32) for t in ts])
    call("elimrec_inst_forward", B, nt, F, ptr(O_inst, F32), ptr(Wu, F32), ptr(Wi, F32), arr(Ws), ptr(bu, F32), ptr(bi, F32),
         arr(bs), ptr(F_out, F32), arr(S_out), stream(), launches=1, tag="inst_forward")


def _wgrad_problems(problems):
    arr = (_lib.WgradProblem * len(problems))()
    for k, (A, a_col, B, b_col, K, r0, r1, out, bias, by_g) in enumerate(problems):
        arr[k].A, arr[k].lda = ptr(A, F32) + 4 * a_col, A.stride(0)
        arr[k].B, arr[k].ldb, arr[k].K = ptr(B, F32) + 4 * b_col, B.stride(0), K
        arr[k].row_begin, arr[k].row_end = r0, r1
        arr[k].out, arr[k].ldo, arr[k].bias_out, arr[k].scale_by_g = ptr(out, F32), out.stride(0), ptr(bias, F32, True), int(by_g)
    return arr


def wgrad_multi_ws_floats(problems, splits):
    return int(_lib.lib().elimrec_wgrad_multi_workspace_floats(len(problems), _wgrad_problems(problems), splits))


def wgrad_multi(problems, splits, ws, gscale=None, x3=False):
    """problems: list of (A, a_col, B, b_col, K, row_begin, row_end, out [64 x K], bias_out or None, scale_by_g):
    out = g * A[r0:r1, a_col:a_col+64]^T B[r0:r1, b_col:b_col+K], bias_out = g * column sums of that A block - all problems in
    one launch + one fixed-order reduction.  x3=False: exact fp32 FFMA (elimrec_wgrad_multi); x3=True: 3xTF32 on the tensor
    cores (elimrec_wgrad_multi_x3, fp32-class accuracy)."""
    arr = _wgrad_problems(problems)
    if x3:
        call("elimrec_wgrad_multi_x3", len(problems), arr, splits, ptr(ws, F32), ptr(gscale, F32, True), stream(), launches=2,
             tag="wgrad_multi_x3")
    else:
        call("elimrec_wgrad_multi", len(problems), arr, splits, ptr(ws, F32), ptr(gscale, F32, True), stream(), launches=2,
             tag="wgrad_multi")


def adam_apply_multi(items, consts_dev, b1, b2, eps, wd):
    """items: list of (param, grad_view, exp_avg, exp_avg_sq); one launch for all of them."""
    arr = (_lib.AdamTensor * len(items))()
    for k, (p, g, m, v) in enumerate(items):
        row_len = p.shape[-1]
        arr[k].param, arr[k].grad, arr[k].exp_avg, arr[k].exp_avg_sq = ptr(p, F32), ptr(g, F32), ptr(m, F32), ptr(v, F32)
        if g.dim() == 1 and g.numel() > 1 and g.stride(0) != 1:      # a column of a wider matrix (packed [W | b] gradient)
            row_len, gld = 1, g.stride(0)
        else:
            gld = g.stride(0) if g.dim() == 2 else row_len
        arr[k].numel, arr[k].row_len = p.numel(), row_len
        arr[k].grad_ld = gld
    call("elimrec_adam_apply_multi", len(items), arr, ptr(consts_dev, torch.float64), b1, b2, eps, wd, stream())


def adam_tick(step_dev, consts_dev, lr, b1, b2):
    call("elimrec_adam_tick", ptr(step_dev, torch.int64), ptr(consts_dev, torch.float64), lr, b1, b2, stream())


def adam_apply(p, g, row_len, g_ld, m, v, consts_dev, b1, b2, eps, wd, g_off=0):
    call("elimrec_adam_apply", p.numel(), ptr(p, F32), ptr(g, F32) + 4 * g_off, row_len, g_ld, ptr(m, F32), ptr(v, F32),
         ptr(consts_dev, torch.float64), b1, b2, eps, wd, stream())


def row_normalize(src, dst):
    call("elimrec_row_normalize", src.shape[0], ptr(src, F32), ptr(dst, F32), stream())


def rank_tables(num_users, num_items, mode, f_user, f_item, s_user, s_item) -> RankTables:
    t = RankTables()
    t.num_users, t.num_items, t.n_mod, t.mode = num_users, num_items, len(s_user), mode
    t.f_user, t.f_item = ptr(f_user, F32), ptr(f_item, F32)
    for m, (a, b) in enumerate(zip(s_user, s_item)):
        t.s_user[m], t.s_item[m] = ptr(a, F32), ptr(b, F32)
    t._keepalive = (f_user, f_item, list(s_user), list(s_item))  # the descriptor only holds raw pointers
    return t


def rank_rowmean(t, eval_users, out):
    call("elimrec_rank_rowmean", C.byref(t), eval_users.numel(), ptr(eval_users, torch.int32), ptr(out, F32), stream())


def rank_scores(t, eval_users, ui_mean, out):
    call("elimrec_rank_scores", C.byref(t), eval_users.numel(), ptr(eval_users, torch.int32), ptr(ui_mean, F32, True),
         ptr(out, F32), stream())


def rank_topk(t, eval_users, ui_mean, train_ptr, train_items, K, idx, val):
    call("elimrec_rank_topk", C.byref(t), eval_users.numel(), ptr(eval_users, torch.int32), ptr(ui_mean, F32, True),
         ptr(train_ptr, torch.int64), ptr(train_items, torch.int32), K, ptr(idx, torch.int32), ptr(val, F32), stream())


def split_fp16(src, scale, hi, lo):
    call("elimrec_split_fp16", src.numel(), ptr(src, F32), scale, ptr(hi, torch.float16), ptr(lo, torch.float16), stream())


def rank_tc_tables(num_users, num_items, mode, tables):
    """tables: list over t of (user_hi, user_lo, item_hi, item_lo, inv_scale); t = 0 fused, then the active heads."""
    t = _lib.RankTcTables()
    t.num_users, t.num_items, t.n_mod, t.mode = num_users, num_items, len(tables) - 1, mode
    for k, (uh, ul, ih, il, inv) in enumerate(tables):
        t.user_hi[k], t.user_lo[k] = ptr(uh, torch.float16), ptr(ul, torch.float16)
        t.item_hi[k], t.item_lo[k] = ptr(ih, torch.float16), ptr(il, torch.float16)
        t.inv_scale[k] = inv
    t._keepalive = tables
    return t


_RANK_TC_WS = {}


def rank_tc(t, what, eval_users, ui_mean, train_ptr, train_items, K, idx, val, mean_out):
    n = eval_users.numel()
    need = int(_lib.lib().elimrec_rank_tc_workspace_bytes(n))
    key = eval_users.device.index
    ws = _RANK_TC_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = _RANK_TC_WS[key] = torch.empty(need, dtype=torch.uint8, device=eval_users.device)
    call("elimrec_rank_tc", C.byref(t), what, n, ptr(eval_users, torch.int32), ptr(ui_mean, F32, True),
         ptr(train_ptr, torch.int64, True), ptr(train_items, torch.int32, True), K, ptr(idx, torch.int32, True),
         ptr(val, F32, True), ptr(mean_out, F32, True), ptr(ws), stream(), tag="rank_tc")


def topk_matrix(scores, K, idx, val):
    call("elimrec_topk_matrix", scores.shape[0], scores.shape[1], ptr(scores, F32), K, ptr(idx, torch.int32),
         ptr(val, F32), stream())


def mask_train(scores, users, train_ptr, train_items):
    """scores[r, train items of users[r]] = -inf"""
    call("elimrec_mask_train", scores.shape[0], scores.shape[1], ptr(users, torch.int32), ptr(train_ptr, torch.int64),
         ptr(train_items, torch.int32), ptr(scores, F32), stream())


def metric_rows(topk_idx, truth_ptr, truth_items, metric_ids, K, rows, sums):
    ids = np.ascontiguousarray(metric_ids, dtype=np.int32)
    inv = np.ascontiguousarray(1.0 / np.log2(np.arange(K, dtype=np.float64) + 2.0))
    call("elimrec_metric_rows", topk_idx.shape[0], K, ptr(topk_idx, torch.int32), ptr(truth_ptr, torch.int64),
         ptr(truth_items, torch.int32), ids.size, ids.ctypes.data, inv.ctypes.data, ptr(rows, F32),
         ptr(sums, torch.float64, True), stream())


def sample_triples_device(seed, epoch, n, user_ids, row_ptr, items, num_items, ou, op, on):
    call("elimrec_sample_triples_device", seed, epoch, n, user_ids.numel(), ptr(user_ids, torch.int32),
         ptr(row_ptr, torch.int64), ptr(items, torch.int32), num_items, ptr(ou, torch.int64), ptr(op, torch.int64),
         ptr(on, torch.int64), stream())


def sample_batch_device(seed, epoch, batch_index_dev, n, user_ids, row_ptr, items, num_items, ou, op, on):
    """one batch of the epoch's Philox stream, its position read from a device counter (graph-replayable)"""
    call("elimrec_sample_batch_device", seed, epoch, ptr(batch_index_dev, torch.int64), n, user_ids.numel(),
         ptr(user_ids, torch.int32), ptr(row_ptr, torch.int64), ptr(items, torch.int32), num_items, ptr(ou, torch.int64),
         ptr(op, torch.int64), ptr(on, torch.int64), stream())


# ---- column-sharded multi-GPU step (csrc/colshard.cu) --------------------------------------------------------------------
def cs_pack(rows, num_users, layers, scale, w, out, n_rows=None):
    n = rows.numel() if rows is not None else n_rows
    call("elimrec_cs_pack", n, ptr(rows, torch.int32, True), num_users, C.byref(layers), scale, w, ptr(out, F32), stream())


def cs_unpack(world, n, w, recv, n_mod, O):
    call("elimrec_cs_unpack", world, n, w, ptr(recv, F32), n_mod, ptr(O, F32), O.stride(0), stream())


def cs_seed_pack(n, world, w, dO, n_mod, scale, send):
    call("elimrec_cs_seed_pack", n, world, w, ptr(dO, F32), dO.stride(0), n_mod, scale, ptr(send, F32), stream())


def cs_seed_scatter(rows_all, w, recv, GA, GB):
    call("elimrec_cs_seed_scatter", rows_all.numel(), ptr(rows_all, torch.int32), w, ptr(recv, F32), ptr(GA, F32), ptr(GB, F32),
         GA.stride(0), stream())


def cs_inst_rows(world, B, triples, num_users, rows, mask=None, mask2=None):
    call("elimrec_cs_inst_rows", world, B, ptr(triples, torch.int64), num_users, ptr(rows, torch.int32),
         (mask.numel() if mask is not None else 0), ptr(mask, torch.uint8, True), ptr(mask2, torch.uint8, True), stream(),
         launches=1 + (mask is not None) + (mask2 is not None))
