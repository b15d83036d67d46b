"""TEST INFRASTRUCTURE - a torch-CPU stand-in for ``elimrec_b200.ops`` (the tensor-level wrappers of the C-ABI).

It states, in a few lines of torch each, WHAT every entry point computes.  ``tests/test_schedule_sim.py`` swaps it
in for the real wrappers so that the host-side schedule of ``elimrec_b200/model.py`` / ``evaluator.py`` / ``optim.py``
(which halves, which rows, which masks, which epilogues, in which order) can be checked against the oracle in the
GPU-less build container.  It is never imported by the package, by ``bench.py`` or by the ``-m gpu`` tests: the product
path has no CPU implementation, and nothing here says anything about the CUDA kernels themselves (those are compared
with the oracle on the B200, ``tests/test_gpu_*.py``).  Tensor-core paths are simulated in exact fp32.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from oracle import ref_eval

LAUNCH_LOG = []          # names of the simulated entry points, in call order


def _log(name):
    LAUNCH_LOG.append(name)


# ---- streams: the simulator is sequential ------------------------------------------------------------------------------
def fork_side(slot=0, high_priority=False):
    return None


def join_side(side):
    return None


class _Epi:
    def __init__(self, prev, out, width, scale):
        self.prev, self.out, self.width, self.scale = prev, out, width, scale


def mean_epilogue(prev, out, width, scale):
    return _Epi(list(prev), out, width, scale)


def _csr(half):
    n = half.n_rows
    crow = torch.from_numpy(half.indptr_host.astype(np.int64))
    return torch.sparse_csr_tensor(crow, half.col.long(), half.val, size=(n, half.n_cols))


def spmm(half, X, Y, width, epi=None, row_mask=None, col_mask=None, density=50, addend=None, add_mask=None):
    _log(f"spmm{width}" + ("m" if row_mask is not None or col_mask is not None or addend is not None else ""))
    Xs = X[:, :width]
    if col_mask is not None:     # edges to unmarked columns are dropped before the gather (their rows may hold garbage)
        Xs = torch.where(col_mask.bool().unsqueeze(1), Xs, torch.zeros_like(Xs))
    acc = torch.sparse.mm(_csr(half), Xs.contiguous())
    if addend is not None:       # Y[row] += addend[row] on the marked rows only (elsewhere the addend slab holds garbage)
        sel = torch.ones(half.n_rows, dtype=torch.bool) if add_mask is None else add_mask.bool()
        acc[sel] += addend[sel, :width]
    rows = torch.unique(half.seg[:, 0].long())       # the work-list's rows: all of them, or a rank's block when row-sharded
    if row_mask is not None:
        rows = rows[row_mask[rows].bool()]
    if Y is not None:
        Y[rows, :width] = acc[rows]
    if epi is not None:
        reps = epi.width // width
        s = None
        for t, w in epi.prev:    # ((x0 + x1) + ...) + x_L, narrow layers broadcast over the graph blocks
            v = t[:, :w]
            v = v.repeat(1, epi.width // w) if w != epi.width else v
            s = v.clone() if s is None else s + v
        s = s + (acc.repeat(1, reps) if reps > 1 else acc)
        epi.out[rows, :epi.width] = (s * epi.scale)[rows]


def spmm64_pair(half_u, half_i, X_for_u, X_for_i, Y_u, Y_i, row_mask_u=None, row_mask_i=None, col_mask_u=None, col_mask_i=None,
                addend_u=None, addend_i=None, add_mask_u=None, add_mask_i=None, adam_u=None, adam_i=None, adam_consts=None,
                variant=0, width=64):
    for half, X, Y, rm, cm, ad, am, adam in ((half_u, X_for_u, Y_u, row_mask_u, col_mask_u, addend_u, add_mask_u, adam_u),
                                             (half_i, X_for_i, Y_i, row_mask_i, col_mask_i, addend_i, add_mask_i, adam_i)):
        if adam is None:
            spmm(half, X, Y, width, row_mask=rm, col_mask=cm, addend=ad, add_mask=am)
            continue
        grad = torch.zeros(half.n_rows, width)
        spmm(half, X, grad, width, row_mask=rm, col_mask=cm, addend=ad, add_mask=am)
        p, m, v, old = adam
        if old is not None:
            old.copy_(p)
        consts, b1, b2, eps, wd = adam_consts
        adam_apply_multi([(p, grad, m, v)], consts, b1, b2, eps, wd)


def inst_rows(users, pos, neg, num_users, rows, mask=None, mask2=None):
    _log("inst_rows")
    r = torch.cat([users, num_users + pos, num_users + neg]).to(torch.int32)
    rows.copy_(r)
    for m in (mask, mask2):
        if m is not None:
            m.zero_()
            m[r.long()] = 1


def mark_rows(rows, mask):
    mask.zero_()
    mask[rows.long()] = 1


def mark_neighbors(half, row_mask, out_mask):
    _log("mark_neighbors")
    ptr = half.indptr_host
    for r in torch.nonzero(row_mask).flatten().tolist():
        out_mask[half.col[ptr[r]:ptr[r + 1]].long()] = 1


def zero_rows(rows, lo, hi, off, dst, width):
    _log("zero_rows")
    r = rows.long()
    r = r[(r >= lo) & (r < hi)] - off
    dst[r, :width] = 0


def scatter_add_rows(rows, lo, hi, off, src, src_width, dst, width, scale):
    _log("scatter_add_rows")
    fold = src_width // width
    r = rows.long()
    sel = (r >= lo) & (r < hi)
    v = src[sel, :src_width].reshape(-1, fold, width).sum(1) * scale
    dst[:, :width].index_add_(0, r[sel] - off, v)


def gather_rows(rows, src, dst, width):
    _log("gather_rows")
    dst[:, :width] = src[rows.long(), :width]


def copy_2d(src, dst, n_rows, width):
    _log("copy_2d")
    dst[:n_rows, :width] = src[:n_rows, :width]


def broadcast_cols(src, dst, n_rows, n_rep):
    _log("broadcast_cols")
    dst[:n_rows, :64 * n_rep] = src[:n_rows, :64].repeat(1, n_rep)


def tie_blocks(src, dst, n_rep, scale):
    _log("tie_blocks")
    dst[:, :64 * n_rep] = (src[:, :64] * scale).repeat(1, n_rep)


def fold_blocks(src, dst, n_rep, scale=1.0):
    _log("fold_blocks")
    dst[:, :64] = src[:, :64 * n_rep].reshape(src.shape[0], n_rep, 64).sum(1) * scale


def axpy_rows(row_scale, X, Y, width):
    _log("axpy_rows")
    Y[:, :width] += row_scale.unsqueeze(1) * X[:, :width]


def layer_mean(layers, out, width, scale):
    _log("layer_mean")
    s = layers[0][:, :width].clone()
    for t in layers[1:]:
        s = s + t[:, :width]
    out[:, :width] = s * scale


def lin_layers(tables):
    return list(tables)


def lin_assemble(rows, num_users, layers, scale, n_mod, accumulate, out, n_rows=None):
    _log("lin_assemble")
    node = rows.long() if rows is not None else torch.arange(n_rows)
    is_user = node < num_users
    sa = sp = None
    for k, (tu, ti) in enumerate(layers):
        v = torch.where(is_user.unsqueeze(1), tu[torch.where(is_user, node, 0), :64], ti[torch.where(is_user, 0, node - num_users), :64])
        sa = v.clone() if sa is None else sa + v
        par = torch.where(is_user, k % 2 == 0, k % 2 == 1).unsqueeze(1)
        sp = torch.where(par, v, torch.zeros_like(v)) if sp is None else sp + torch.where(par, v, torch.zeros_like(v))
    out[:, :64] = sa * scale
    for m in range(n_mod):
        blk = slice(64 * (m + 1), 64 * (m + 2))
        out[:, blk] = (out[:, blk] if accumulate else 0) + sp * scale


def lin_seed(rows, num_users, layer, dO, n_mod, scale, dst):
    _log("lin_seed")
    node = rows.long()
    par = torch.where(node < num_users, layer % 2 == 0, layer % 2 == 1).unsqueeze(1)
    v = dO[:, :64].clone()
    if n_mod:
        v = v + torch.where(par, dO[:, 64:64 * (1 + n_mod)].reshape(-1, n_mod, 64).sum(1), torch.zeros_like(v))
    dst[:, :64].index_add_(0, node, v * scale)


def lin_seed2(rows, dO, n_mod, scale, GA, GB):
    _log("lin_seed2")
    node = rows.long()
    a = dO[:, :64 * (1 + n_mod)].reshape(-1, 1 + n_mod, 64).sum(1)
    GA[:, :64].index_add_(0, node, a * scale)
    GB[:, :64].index_add_(0, node, dO[:, :64] * scale)


def pack_proj_weights(items, round_tf32):
    _log("pack_proj")
    for W, b, dst in items:
        dm = W.shape[1]
        dst.zero_()
        dst[:, :dm] = W
        dst[:, dm] = b


def split3_rows(src, dst, n_rows, width, pattern):
    """(exact-fp32 simulation: hi = x, lo = 0 - the three stacked products then sum to x*y)"""
    _log("split3")
    z = torch.zeros(n_rows, width)
    parts = [src[:n_rows, :width], z, src[:n_rows, :width]] if pattern else [src[:n_rows, :width], src[:n_rows, :width], z]
    for b, part in enumerate(parts):
        dst[b * n_rows:(b + 1) * n_rows, :width] = part


def linear_x3_fwd(X, W_hi, W_lo, b, Y, col=0, tag="linear_x3"):
    _log(tag)
    Y[:, col:col + 64] = X @ (W_hi + W_lo).t() + (b if b is not None else 0)


def axpy_2d(X, Y, n_rows, width, scale=1.0, accumulate=True):
    Y[:n_rows, :width] = (Y[:n_rows, :width] if accumulate else 0) + scale * X[:n_rows, :width]


def _view(t, off, shape, strides):
    return torch.as_strided(t, shape, strides, t.storage_offset() + off)


def gemm(M, N, K, A, a_sm, a_sk, B, b_sk, b_sn, Cm, c_sm, c_sn, bias=None, accumulate=False, split_k=1, ws=None,
         scale=None, a_off=0, b_off=0, c_off=0, tag=None):
    _log(tag or "gemm")
    a = _view(A, a_off, (M, K), (a_sm, a_sk))
    b = _view(B, b_off, (K, N), (b_sk, b_sn))
    c = _view(Cm, c_off, (M, N), (c_sm, c_sn))
    r = a @ b
    if scale is not None:
        r = r * scale
    if bias is not None:
        r = r + bias
    c.copy_(c + r if accumulate else r)


def linear_tf32_fwd_multi(problems, tag="proj_fwd_tc"):
    _log(tag)
    for X, W, b, Y, col in problems:
        Y[:, col:col + 64] = X @ W.t() + (b if b is not None else 0)


def linear_tf32_fwd(X, W, b, Y, col=0, tag="proj_fwd_tc"):
    linear_tf32_fwd_multi([(X, W, b, Y, col)], tag)


def linear_tf32_wgrad(dY, X, dW, ws, col=0, tag="proj_wgrad_tc"):
    _log(tag)
    dW.copy_(dY[:, col:col + 64].t() @ X)


def linear_tf32_wgrad_ws_floats(M, K):
    return 1


def colsum_ws_floats(M, N):
    return 1


def inst_backward_ws_floats(B, nt, Fw):
    return 1


def round_tf32(src, dst):
    dst.copy_(src)


def prep_weights_tf32(items):
    _log("prep_weights")
    for src, hi, lo in items:
        hi.copy_(src)
        if lo is not None:
            lo.zero_()


def fuse_heads_x3(O_rows, Wf_hi, Wf_lo, bf, Ws_hi, Ws_lo, bs, F_out, S_out, tag="fuse_heads_x3"):
    _log(tag)
    F_out.copy_(O_rows @ (Wf_hi + Wf_lo).t() + bf)
    for m in range(len(Ws_hi)):
        S_out[m].copy_(O_rows[:, 64 * (m + 1):64 * (m + 2)] @ (Ws_hi[m] + Ws_lo[m]).t() + bs[m])


def inst_forward(B, nt, F, O_inst, Wu, Wi, Ws, bu, bi, bs, F_out, S_out):
    _log("inst_forward")
    F_out[:B] = O_inst[:B] @ Wu.t() + bu
    F_out[B:3 * B] = O_inst[B:3 * B] @ Wi.t() + bi
    for m in range(nt - 1):
        S_out[m][:3 * B] = O_inst[:3 * B, 64 * (m + 1):64 * (m + 2)] @ Ws[m].t() + bs[m]


def fuse_heads_x3_all(U, I, O, Wu, bu, Wi, bi, Ws_hi, Ws_lo, bs, F_out, S_out, tag="fuse_heads_x3_all"):
    fuse_heads_x3(O[:U], Wu[0], Wu[1], bu, Ws_hi, Ws_lo, bs, F_out[:U], [s[:U] for s in S_out], tag)
    fuse_heads_x3(O[U:U + I], Wi[0], Wi[1], bi, Ws_hi, Ws_lo, bs, F_out[U:U + I], [s[U:U + I] for s in S_out], tag)


def colsum(M, N, A, ld, out, ws, accumulate=False, scale=None, a_off=0):
    _log("colsum")
    s = _view(A, a_off, (M, N), (ld, 1)).sum(0)
    if scale is not None:
        s = s * scale
    out[:N] = out[:N] + s if accumulate else s


def inst_dO_seed(B, nt, F, inst_grad, gscale, Wu, Wi, Ws, dO_inst, rows, n_mod, scale, GA, GB):
    inst_backward(B, nt, F, inst_grad, None, gscale, Wu, Wi, Ws, dO_inst, None, None, None, None, None, None, None, part=1)
    lin_seed2(rows, dO_inst, n_mod, scale, GA, GB)


def bpr(tables, weights, users, pos, neg, num_users, loss_out, inst_rows_out, inst_grad, terms, part=3):
    if part == 2:       # (the simulator computes the loss with the terms: nothing left to reduce)
        return
    """1 + M normalised-cosine BPR losses (EliMRec.py:291-297) and their gradient w.r.t. the gathered rows."""
    _log("bpr")
    B = users.numel()
    total = 0
    for t, (tab, w) in enumerate(zip(tables, weights)):
        rows = [tab[users], tab[num_users + pos], tab[num_users + neg]]
        with torch.enable_grad():       # (called from inside an autograd.Function forward, where grad mode is off)
            rows = [r.detach().clone().requires_grad_(True) for r in rows]
            u, p, n = (F.normalize(r, dim=1) for r in rows)
            loss = torch.mean(F.softplus(torch.sum(u * n, 1) - torch.sum(u * p, 1)))
            gs = torch.autograd.grad(loss, rows)
        for k in range(3):
            inst_grad[k * B:(k + 1) * B, 64 * t:64 * (t + 1)] = w * gs[k]
        total = total + w * loss.detach()
    loss_out[0] = total
    inst_rows_out.copy_(torch.cat([users, num_users + pos, num_users + neg]).to(torch.int32))


def inst_backward(B, nt, Fw, inst_grad, O_inst, gscale, Wu, Wi, Ws, dO_inst, dWu, dWi, dbu, dbi, dWs, dbs, ws, part=3):
    _log(f"inst_backward{part}")
    g = 1.0 if gscale is None else float(gscale)
    ig = inst_grad * g
    if part & 1:
        dO_inst[:B] = ig[:B, :64] @ Wu
        dO_inst[B:] = ig[B:, :64] @ Wi
        for m in range(nt - 1):
            blk = slice(64 * (m + 1), 64 * (m + 2))
            dO_inst[:, blk] += ig[:, blk] @ Ws[m]
    if part & 2:
        dWu.copy_(ig[:B, :64].t() @ O_inst[:B])
        dWi.copy_(ig[B:, :64].t() @ O_inst[B:])
        dbu.copy_(ig[:B, :64].sum(0))
        dbi.copy_(ig[B:, :64].sum(0))
        for m in range(nt - 1):
            blk = slice(64 * (m + 1), 64 * (m + 2))
            dWs[m].copy_(ig[:, blk].t() @ O_inst[:, blk])
            dbs[m].copy_(ig[:, blk].sum(0))


def wgrad_multi_ws_floats(problems, splits):
    return 1


def wgrad_multi(problems, splits, ws, gscale=None, x3=False):
    _log("wgrad_multi")
    for A, a_col, B, b_col, K, r0, r1, out, bias, by_g in problems:
        g = float(gscale) if (by_g and gscale is not None) else 1.0
        a = A[r0:r1, a_col:a_col + 64]
        out[:, :K] = g * (a.t() @ B[r0:r1, b_col:b_col + K])
        if bias is not None:
            bias.copy_(g * a.sum(0))


def adam_tick(step_dev, consts_dev, lr, b1, b2):
    step_dev += 1
    t = int(step_dev)
    consts_dev[0] = lr / (1 - b1 ** t)
    consts_dev[1] = (1 - b2 ** t) ** 0.5


def adam_apply_multi(items, consts_dev, b1, b2, eps, wd):
    _log("adam")
    step_size, bc2_sqrt = float(consts_dev[0]), float(consts_dev[1])
    for p, g, m, v in items:
        g = g.reshape(p.shape) + wd * p
        m.lerp_(g, 1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        p.addcdiv_(m, v.sqrt() / bc2_sqrt + eps, value=-step_size)


def row_normalize(src, dst):
    dst.copy_(F.normalize(src, dim=1))


# ---- rank ---------------------------------------------------------------------------------------------------------------
class _RankTables:
    pass


def rank_tables(num_users, num_items, mode, f_user, f_item, s_user, s_item):
    t = _RankTables()
    t.mode, t.f_user, t.f_item, t.s_user, t.s_item = mode, f_user, f_item, list(s_user), list(s_item)
    return t


def _ui(t, users):
    return torch.sigmoid(t.f_user[users.long()] @ t.f_item.t())


def rank_rowmean(t, eval_users, out):
    _log("rank_rowmean")
    out.copy_(_ui(t, eval_users).mean(1))


def _scores(t, users, mean):
    pm, fm = t.mode & 3, t.mode >> 2
    ui = _ui(t, users)
    if pm == 0:
        return torch.sigmoid(ui)
    cos = [su[users.long()] @ si.t() for su, si in zip(t.s_user, t.s_item)]

    def fuse(x):
        if fm == 0:
            for c in cos:
                x = x * torch.sigmoid(c)
            return x
        if fm == 1:
            z = torch.sigmoid(x)
            for c in cos:
                z = z * torch.sigmoid(c)
            return torch.log(z + 1e-12) - torch.log1p(z)
        for c in cos:
            x = x + c
        return torch.log(torch.sigmoid(x) + 1e-12)

    if pm == 1:
        return torch.sigmoid(fuse(ui))
    return torch.sigmoid(fuse(ui) - fuse(mean.unsqueeze(1).expand_as(ui)))


def rank_scores(t, eval_users, ui_mean, out):
    _log("rank_scores")
    out.copy_(_scores(t, eval_users, ui_mean))


def rank_topk(t, eval_users, ui_mean, train_ptr, train_items, K, idx, val):
    _log("rank_topk")
    sc = _scores(t, eval_users, ui_mean).numpy().copy()
    tp, ti = train_ptr.numpy(), train_items.numpy()
    for r, u in enumerate(eval_users.tolist()):
        sc[r, ti[tp[u]:tp[u + 1]]] = -np.inf
    top = ref_eval.topk_lowest_index(sc, K)
    idx.copy_(torch.from_numpy(top))
    val.copy_(torch.from_numpy(np.take_along_axis(sc, top.astype(np.int64), 1)))


def mask_train(scores, users, train_ptr, train_items):
    _log("mask_train")
    tp, ti = train_ptr.numpy(), train_items.numpy()
    for r, u in enumerate(users.tolist()):
        scores[r, torch.from_numpy(ti[tp[u]:tp[u + 1]].astype(np.int64))] = -np.inf


def topk_matrix(scores, K, idx, val):
    _log("topk_matrix")
    sc = scores.numpy()
    top = ref_eval.topk_lowest_index(sc, K)
    idx.copy_(torch.from_numpy(top))
    val.copy_(torch.from_numpy(np.take_along_axis(sc, top.astype(np.int64), 1)))


def metric_rows(topk_idx, truth_ptr, truth_items, metric_ids, K, rows, sums):
    _log("metric_rows")
    tp, ti = truth_ptr.numpy(), truth_items.numpy()
    truth = [ti[tp[r]:tp[r + 1]].tolist() for r in range(topk_idx.shape[0])]
    out = ref_eval.metric_rows(topk_idx.numpy(), truth, list(metric_ids), K)
    rows[:out.shape[0]] = torch.from_numpy(out)
    if sums is not None:
        sums += torch.from_numpy(out.astype(np.float64).sum(0))


# ---- column-sharded multi-GPU step ---------------------------------------------------------------------------------------
def cs_pack(rows, num_users, layers, scale, w, out, n_rows=None):
    _log("cs_pack")
    node = rows.long() if rows is not None else torch.arange(n_rows)
    is_user = node < num_users
    sa = sp = None
    for k, (tu, ti) in enumerate(layers):
        v = torch.where(is_user.unsqueeze(1), tu[torch.where(is_user, node, 0), :w], ti[torch.where(is_user, 0, node - num_users), :w])
        sa = v.clone() if sa is None else sa + v
        par = torch.where(is_user, k % 2 == 0, k % 2 == 1).unsqueeze(1)
        pv = torch.where(par, v, torch.zeros_like(v))
        sp = pv if sp is None else sp + pv
    out.view(-1, 2 * w)[:, :w] = sa * scale
    out.view(-1, 2 * w)[:, w:] = sp * scale


def cs_unpack(world, n, w, recv, n_mod, O):
    _log("cs_unpack")
    r = recv.view(world, n, 2 * w)
    O[:n, :64] = r[:, :, :w].permute(1, 0, 2).reshape(n, 64)
    par = r[:, :, w:].permute(1, 0, 2).reshape(n, 64)
    for m in range(n_mod):
        O[:n, 64 * (m + 1):64 * (m + 2)] += par


def cs_seed_pack(n, world, w, dO, n_mod, scale, send):
    _log("cs_seed_pack")
    b = dO[:n, :64]
    a = dO[:n, :64 * (1 + n_mod)].reshape(n, 1 + n_mod, 64).sum(1)
    s = send.view(world, n, 2 * w)
    s[:, :, :w] = (scale * a).reshape(n, world, w).permute(1, 0, 2)
    s[:, :, w:] = (scale * b).reshape(n, world, w).permute(1, 0, 2)


def cs_seed_scatter(rows_all, w, recv, GA, GB):
    _log("cs_seed_scatter")
    r = recv.view(-1, 2 * w)
    GA[:, :w].index_add_(0, rows_all.long(), r[:, :w])
    GB[:, :w].index_add_(0, rows_all.long(), r[:, w:])


def cs_inst_rows(world, B, triples, num_users, rows, mask=None, mask2=None):
    _log("cs_inst_rows")
    t = triples.view(world, 3, B).clone()
    t[:, 1:] += num_users
    rows.copy_(t.reshape(-1).to(torch.int32))
    for m in (mask, mask2):
        if m is not None:
            m.zero_()
            m[rows.long()] = 1
