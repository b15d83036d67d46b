"""Row-sharded EliMRec: users and items partitioned over the ranks, NCCL all-gather of the propagated
embedding shards for every GCN layer (BASELINE.json north star, piece 5).

The reference has no multi-GPU path at all, so there is no reference behaviour to match
beyond the single-device numbers: a G-rank run must produce the same loss and the same parameters as
one GPU on the same triples (tests/test_gpu_multi.py).

Partition.  Rank r owns the contiguous user block [r*Ub, (r+1)*Ub) and item block [r*Ib, (r+1)*Ib)
(Ub = ceil(U/G), Ib = ceil(I/G); slabs are padded to G*Ub / G*Ib rows so that shards are equal-sized and
``all_gather_into_tensor`` works IN PLACE on the slab).  It owns those rows of A_ui / A_iu (segment lists
restricted to the block), of both embedding tables and their Adam state, and the matching rows of the
item features.

Step.
  forward   owned rows of layer 0 (E_u block; [E_i | P_v | P_a | P_t] block) -> all-gather;
            for k = 1..L-1: SpMM over the owned rows (wide + narrow) -> all-gather of both slabs;
            layer L: SpMM with the fused layer-mean epilogue -> O[owned rows]; fusion Linear + heads on them.
  loss      every rank draws the SAME triples.  Each contributes its owned rows of O[instance rows] (others
            zero); ONE all-reduce of the [3B x F] buffer gives everybody all instance rows, from which the
            loss, the instance gradients and the fusion / head weight gradients are computed redundantly
            and identically (no further exchange for them).
  backward  the seed gradient slab is built in full on every rank from the instance gradients; each hop
            computes the owned rows and all-gathers them, except the last hop; projection weight
            gradients over the owned item rows are all-reduced (small).
  Adam      owned rows of the tables, all small tensors (identical on every rank).

Collectives per step at L = 3: 6 forward + 4 backward all-gathers (half of them 64 wide), one 6 MB
all-reduce, one small all-reduce.  At these sizes the all-gathers dominate (1.15 GB per
wide layer on the 10x graph vs a 0.1 ms local SpMM); data-parallel replicas (``enable_data_parallel``) are the
throughput mode, this is the capacity mode for graphs that do not fit one GPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from .graph import BipartiteGraph
from .model import D, EliMRec, ElimrecError


class ShardedEliMRec(EliMRec):
    def _lin_init(self):
        self.linear = False      # the wide/narrow slab schedule, row-sharded

    def _init_weight(self):
        if not (dist.is_available() and dist.is_initialized()):
            raise ElimrecError("ShardedEliMRec needs an initialised torch.distributed process group")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        super()._init_weight()
        if self._generic or not self.graph.symmetric or self.mm_fusion_mode != "concat" or self.tiktok:
            raise NotImplementedError("ShardedEliMRec covers the default configuration (symmetric adj_type, concat fusion, "
                                      "feature-file text modality); the SURVEY.md 8f variants run on EliMRec")
        U, I, G = self.num_users, self.num_items, self.world
        self.Ub, self.Ib = -(-U // G), -(-I // G)
        self.u0, self.u1 = min(U, self.rank * self.Ub), min(U, (self.rank + 1) * self.Ub)
        self.i0, self.i1 = min(I, self.rank * self.Ib), min(I, (self.rank + 1) * self.Ib)
        # segment lists restricted to the owned rows (the CSR arrays themselves are replicated: 10 MB)
        self.graph = BipartiteGraph(self.dataset.train_matrix, self.device_, self.config["adj_type"],
                                    user_rows=(self.u0, self.u1), item_rows=(self.i0, self.i1))
        self._initial_sync_done = False

    def _initial_sync(self):
        """identical starting weights on every rank (done lazily: parameters move to the GPU after construction)"""
        if not self._initial_sync_done:
            for p in self.parameters():
                dist.broadcast(p.data, src=0)
            self._initial_sync_done = True

    # ---- helpers ---------------------------------------------------------------------------------
    def _ag(self, slab, blk):
        """in-place all-gather of equal row blocks of a padded slab"""
        r = self.rank
        dist.all_gather_into_tensor(slab, slab[r * blk:(r + 1) * blk])

    def _ar(self, t):
        dist.all_reduce(t)

    def _workspace(self, B):
        ws = self._ws
        if ws is not None and ws["B"] == B:
            return ws
        dev, L, G = self.device_, self.n_layers, self.world
        Up, Ip = G * self.Ub, G * self.Ib
        Gm = 1 + len(self.mods)
        Fw = D * Gm
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        ws = dict(B=B, G=Gm, F=Fw, nt=Gm, cache={})
        ws["Eu"], ws["X0_i"] = z(Up, D), z(Ip, Fw)
        rows = lambda side: Up if side == "u" else Ip
        ws["XW"], ws["XN"] = {}, {}
        for k in range(1, L):
            side = "u" if k % 2 == 1 else "i"
            ws["XW"][k] = z(rows(side), Fw)
            ws["XN"][k] = z(rows("i" if side == "u" else "u"), D)
        ws["O_u"], ws["O_i"] = z(Up, Fw), z(Ip, Fw)       # rows outside the owned block stay zero for ever
        ws["F_u"], ws["F_i"] = z(Up, D), z(Ip, D)
        ws["S_u"], ws["S_i"] = [z(Up, D) for _ in self.mods], [z(Ip, D) for _ in self.mods]
        ws["loss"], ws["terms"] = z(1), z(Gm * B)
        ws["inst_rows"] = torch.empty(3 * B, dtype=torch.int32, device=dev)
        ws["inst_dummy"] = torch.empty(3 * B, dtype=torch.int32, device=dev)
        ws["inst_grad"], ws["O_inst"], ws["dO_inst"] = z(3 * B, D * Gm), z(3 * B, Fw), z(3 * B, Fw)
        ws["F_c"], ws["S_c"] = z(3 * B, D), [z(3 * B, D) for _ in self.mods]
        ar = torch.arange(B, dtype=torch.int64, device=dev)
        ws["c_users"], ws["c_pos"], ws["c_neg"] = ar.clone(), ar.clone(), ar + B
        R = max(Up, Ip)
        ws["dW"], ws["dN"] = [z(R, Fw), z(R, Fw)], [z(R, D), z(R, D)]
        ws["g"] = {n: torch.zeros_like(p, device=dev) for n, p in self._params().items()
                   if not n.startswith("embedding_user.w") and not n.startswith("embedding_item.w")}
        ws["g_proj_bias"] = z(D * len(self.mods))
        for j, m in enumerate(self.mods):
            ws["g"][f"{m}_dense.bias"] = ws["g_proj_bias"][D * j:D * (j + 1)]
        dmax = max(self._feat[m].shape[1] for m in self.mods)
        ws["split_proj"] = max(1, min(256, (self.Ib + 1023) // 1024))
        ws["gemm_ws"] = z(ws["split_proj"] * dmax * D)
        ws["W_tf32"] = {m: z(D, self._feat[m].shape[1]) for m in self.mods}
        ws["inst_ws"] = z(ops.inst_backward_ws_floats(B, Gm, Fw))
        ws["W_split"] = {"u": (z(D, Fw), z(D, Fw)), "i": (z(D, Fw), z(D, Fw))}
        for m in self.mods:
            ws["W_split"][m] = (z(D, D), z(D, D))
        ws["wgrad_ws_m"] = {m: z(max(1, ops.linear_tf32_wgrad_ws_floats(self.Ib, self._feat[m].shape[1]))) for m in self.mods}
        ws["gemm_ws_m"] = {m: z(ws["split_proj"] * self._feat[m].shape[1] * D) for m in self.mods}
        ws["colsum_ws"] = z(ops.colsum_ws_floats(self.Ib, D * len(self.mods)))
        # flat buffer for the projection-weight gradients (summed over ranks)
        names = [f"{m}_dense.weight" for m in self.mods]
        ws["proj_flat"] = z(sum(ws["g"][n].numel() for n in names) + D * len(self.mods))
        self._ws = ws
        return ws

    def _fuse_heads(self, P, ws, O_rows, Fout, Sout, who):
        """fusion Linear + heads on a block of rows of the layer-mean slab (``who`` = 'u' or 'i')"""
        Wf, bf = P[f"embedding_{'user' if who == 'u' else 'item'}_after_GCN.weight"].detach(), \
            P[f"embedding_{'user' if who == 'u' else 'item'}_after_GCN.bias"].detach()
        Fw, n = ws["F"], O_rows.shape[0]
        if n == 0:
            return
        if self.fuse_precision == "x3":
            sp = ws["W_split"]
            ops.fuse_heads_x3(O_rows, sp[who][0], sp[who][1], bf, [sp[m][0] for m in self.mods], [sp[m][1] for m in self.mods],
                              [P[f"s_dense_{m}.bias"].detach() for m in self.mods], Fout, Sout)
        else:
            ops.gemm(n, D, Fw, O_rows, Fw, 1, Wf, 1, Fw, Fout, D, 1, bias=bf, tag="fuse_fwd")
            for j, m in enumerate(self.mods):
                Wsm, bsm = P[f"s_dense_{m}.weight"].detach(), P[f"s_dense_{m}.bias"].detach()
                ops.gemm(n, D, D, O_rows[:, D * (j + 1):], Fw, 1, Wsm, 1, D, Sout[j], D, 1, bias=bsm, tag="head_fwd")

    # ---- forward -----------------------------------------------------------------------------------
    def _forward(self, users, pos, neg):
        P = self._params()
        U, I, L = self.num_users, self.num_items, self.n_layers
        B = int(users.numel())
        self._initial_sync()
        ws = self._workspace(B)
        Fw = ws["F"]
        g = self.graph
        u0, u1, i0, i1, Ub, Ib = self.u0, self.u1, self.i0, self.i1, self.Ub, self.Ib
        self._prep_weights(P, ws)
        Eu, X0_i = ws["Eu"], ws["X0_i"]
        # layer 0: owned rows, then all-gather
        if u1 > u0:
            ops.copy_2d(P["embedding_user.weight"].detach()[u0:u1], Eu[u0:u1], u1 - u0, D)
        if i1 > i0:
            ops.copy_2d(P["embedding_item.weight"].detach()[i0:i1], X0_i[i0:i1], i1 - i0, D)
            self._proj_forward(P, ws, X0_i, i0, i1)
        self._ag(Eu, Ub)
        self._ag(X0_i, Ib)
        prev_u, prev_i = [(Eu, D)], [(X0_i, Fw)]
        wide_in, narrow_in = X0_i, Eu
        inv = 1.0 / (L + 1)
        for k in range(1, L + 1):
            users_wide = (k % 2 == 1)
            half_w, half_n = (g.ui, g.iu) if users_wide else (g.iu, g.ui)
            blk_w, blk_n = (Ub, Ib) if users_wide else (Ib, Ub)
            if k < L:
                Yw, Yn = ws["XW"][k], ws["XN"][k]
                ops.spmm(half_w, wide_in, Yw, Fw)
                ops.spmm(half_n, narrow_in, Yn, D)
                self._ag(Yw, blk_w)
                self._ag(Yn, blk_n)
                if users_wide:
                    prev_u.append((Yw, Fw)); prev_i.append((Yn, D))
                else:
                    prev_i.append((Yw, Fw)); prev_u.append((Yn, D))
                wide_in, narrow_in = Yw, Yn
            else:
                out_w, out_n = (ws["O_u"], ws["O_i"]) if users_wide else (ws["O_i"], ws["O_u"])
                pw, pn = (prev_u, prev_i) if users_wide else (prev_i, prev_u)
                ops.spmm(half_w, wide_in, None, Fw, ops.mean_epilogue(pw, out_w, Fw, inv))
                ops.spmm(half_n, narrow_in, None, D, ops.mean_epilogue(pn, out_n, Fw, inv))
        # fusion + heads on the owned rows
        self._fuse_heads(P, ws, ws["O_u"][u0:u1], ws["F_u"][u0:u1], [s_[u0:u1] for s_ in ws["S_u"]], "u")
        self._fuse_heads(P, ws, ws["O_i"][i0:i1], ws["F_i"][i0:i1], [s_[i0:i1] for s_ in ws["S_i"]], "i")
        self._tables_version = getattr(self, "_tables_version", 0) + 1
        self._tables_gathered = False
        self.all_users, self.all_items = ws["F_u"][:U], ws["F_i"][:I]      # complete only after _gather_tables()
        # instance rows: every rank contributes the rows it owns (all others are zero), one all-reduce
        rows, Oin = ws["inst_rows"], ws["O_inst"]
        rows[:B] = users.int()
        rows[B:2 * B] = (pos + U).int()
        rows[2 * B:] = (neg + U).int()
        ops.gather_rows(rows[:B], ws["O_u"], Oin[:B], Fw)
        item_idx = (rows[B:] - U).contiguous()
        ops.gather_rows(item_idx, ws["O_i"], Oin[B:], Fw)
        self._ar(Oin)
        self._fuse_heads(P, ws, Oin[:B], ws["F_c"][:B], [s_[:B] for s_ in ws["S_c"]], "u")
        self._fuse_heads(P, ws, Oin[B:], ws["F_c"][B:], [s_[B:] for s_ in ws["S_c"]], "i")
        if self.kwai:
            self.modality = "v"
        alpha = float(self.config.alpha)
        weights = [1.0] + ([0.0] * len(self.mods) if self.predict_type == "normal"
                           else [alpha * self.modality.count(m) for m in self.mods])
        # compact tables of 3B rows: triple k uses rows k, B+k, 2B+k
        ops.bpr([ws["F_c"]] + ws["S_c"], weights, ws["c_users"][:B], ws["c_pos"][:B], ws["c_neg"][:B], B, ws["loss"],
                ws["inst_dummy"], ws["inst_grad"], ws["terms"])
        return ws["loss"][0]

    # ---- backward ----------------------------------------------------------------------------------
    def _backward(self, gscale=None):
        P = self._params()
        ws = self._ws
        U, I, L, B = self.num_users, self.num_items, self.n_layers, ws["B"]
        Fw, nt, gr = ws["F"], ws["nt"], ws["g"]
        g = self.graph
        u0, u1, i0, i1, Ub, Ib = self.u0, self.u1, self.i0, self.i1, self.Ub, self.Ib
        rows, ig, Oin, dOin = ws["inst_rows"], ws["inst_grad"], ws["O_inst"], ws["dO_inst"]
        if gscale is not None:
            gscale = gscale.reshape(1)
        ops.inst_backward(B, nt, Fw, ig, Oin, gscale, P["embedding_user_after_GCN.weight"].detach(),
                          P["embedding_item_after_GCN.weight"].detach(), [P[f"s_dense_{m}.weight"].detach() for m in self.mods],
                          dOin, gr["embedding_user_after_GCN.weight"], gr["embedding_item_after_GCN.weight"],
                          gr["embedding_user_after_GCN.bias"], gr["embedding_item_after_GCN.bias"],
                          [gr[f"s_dense_{m}.weight"] for m in self.mods], [gr[f"s_dense_{m}.bias"] for m in self.mods],
                          ws["inst_ws"])
        inv = 1.0 / (L + 1)
        full = {"u": (0, U, 0), "i": (U, U + I, U)}
        own = {"u": (u0, u1, 0), "i": (U + i0, U + i1, U)}
        padded = {"u": self.world * Ub, "i": self.world * Ib}
        blk = {"u": Ub, "i": Ib}

        def add_G(dst, side, wide, rng):
            a, b, off = rng[side]
            if b > a:
                ops.scatter_add_rows(rows, a, b, off, dOin, Fw, dst, Fw if wide else D, inv)

        s_w = "u" if L % 2 == 1 else "i"
        s_n = "i" if s_w == "u" else "u"
        dWc, dNc = ws["dW"][0][:padded[s_w]], ws["dN"][0][:padded[s_n]]
        dWc.zero_(); dNc.zero_()
        add_G(dWc, s_w, True, full)          # the seed is known in full everywhere: no exchange
        add_G(dNc, s_n, False, full)
        flip = 1
        for k in range(L, 0, -1):
            s = "u" if k % 2 == 1 else "i"
            o = "i" if s == "u" else "u"
            half_o, half_s = (g.iu, g.ui) if s == "u" else (g.ui, g.iu)
            nW, nN = ws["dW"][flip][:padded[o]], ws["dN"][flip][:padded[s]]
            ops.spmm(half_o, dWc, nW, Fw)
            ops.spmm(half_s, dNc, nN, D)
            add_G(nW, o, True, own)
            add_G(nN, s, False, own)
            if k > 1:
                self._ag(nW, blk[o])
                self._ag(nN, blk[s])
            dWc, dNc, flip = nW, nN, flip ^ 1
        # owned rows of d x_0: [dE_i | dP_m] (items), dE_u (users); projection weights summed over ranks
        if i1 > i0:
            self._proj_wgrad(ws, dWc, i0, i1)
        else:
            for m in self.mods:
                gr[f"{m}_dense.weight"].zero_()
            ws["g_proj_bias"].zero_()
        flat, o_ = ws["proj_flat"], 0
        for m in self.mods:
            n_ = gr[f"{m}_dense.weight"].numel()
            flat[o_:o_ + n_] = gr[f"{m}_dense.weight"].reshape(-1)
            o_ += n_
        flat[o_:] = ws["g_proj_bias"]
        self._ar(flat)
        o_ = 0
        for m in self.mods:
            n_ = gr[f"{m}_dense.weight"].numel()
            gr[f"{m}_dense.weight"].copy_(flat[o_:o_ + n_].view_as(gr[f"{m}_dense.weight"]))
            o_ += n_
        ws["g_proj_bias"].copy_(flat[o_:])
        self._owned_grads = {"embedding_user.weight": dNc, "embedding_item.weight": dWc}
        grads = dict(gr)
        return grads

    # ---- training step: Adam on the owned rows -----------------------------------------------------
    def train_step(self, users, pos_items, neg_items):
        if self._adam is None:
            self.make_optimizer()
        users, pos, neg = self._triples(users, pos_items, neg_items)
        with torch.no_grad():
            loss = self._forward(users, pos, neg)
            grads = self._backward(None)
            ad = self._adam
            P = self._params()
            ops.adam_tick(ad.step_dev, ad.consts, ad.lr, ad.betas[0], ad.betas[1])
            items = []
            for name, gview in grads.items():
                p = P[name]
                m, v = ad._st(name, p)
                items.append((p.data, gview, m, v))
            for name, (r0, r1) in (("embedding_user.weight", (self.u0, self.u1)), ("embedding_item.weight", (self.i0, self.i1))):
                if r1 > r0:
                    p = P[name]
                    m, v = ad._st(name, p)
                    gslab = self._owned_grads[name]
                    gview = gslab[r0:r1] if name.startswith("embedding_user") else gslab[r0:r1, :D]
                    items.append((p.data[r0:r1], gview, m[r0:r1], v[r0:r1]))
            ops.adam_apply_multi(items, ad.consts, ad.betas[0], ad.betas[1], ad.eps, ad.wd)
            self._params_synced = False
        return loss

    def bpr_loss(self, users, pos_items, neg_items):
        raise NotImplementedError("ShardedEliMRec trains through train_step() (Adam state is sharded with the rows)")

    # ---- parameter / table exchange for checkpoints and evaluation ---------------------------------------
    def sync_parameters(self):
        """all-gather the owned rows of both embedding tables so that every rank holds the full, current tables"""
        if getattr(self, "_params_synced", True):
            return
        P = self._params()
        for name, n, blk, (r0, r1) in (("embedding_user.weight", self.num_users, self.Ub, (self.u0, self.u1)),
                                       ("embedding_item.weight", self.num_items, self.Ib, (self.i0, self.i1))):
            pad = torch.zeros(self.world * blk, D, dtype=torch.float32, device=self.device_)
            pad[r0:r1] = P[name].data[r0:r1]
            self._ag(pad, blk)
            P[name].data.copy_(pad[:n])
        self._params_synced = True

    def state_dict(self, *a, **k):
        self.sync_parameters()
        return super().state_dict(*a, **k)

    def _tables(self):
        ws = self._ws
        if not self._tables_gathered:      # every rank ranks its own users against ALL items
            for t, blk in [(ws["F_u"], self.Ub), (ws["F_i"], self.Ib)] + [(s_, self.Ub) for s_ in ws["S_u"]] + \
                          [(s_, self.Ib) for s_ in ws["S_i"]]:
                self._ag(t, blk)
            self._tables_gathered = True
        U, I = self.num_users, self.num_items
        return (ws["F_u"][:U], ws["F_i"][:I], [s_[:U] for s_ in ws["S_u"]], [s_[:I] for s_ in ws["S_i"]])

    def enable_data_parallel(self):
        raise ElimrecError("ShardedEliMRec is the row-sharded mode; data-parallel replicas use EliMRec")

    def make_graphed_step(self, batch_size=None):
        raise ElimrecError("the row-sharded step contains NCCL collectives and is not captured as one CUDA graph")
