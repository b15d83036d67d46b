"""The LINEAR schedule of the training step (default where it applies; ``linear_schedule=False`` switches it off).

What the reference computes every step (``models/EliMRec.py:228-258``): four graphs ``x^g_0 = [E_u ; Z_g]`` with
``Z_id = E_i`` and ``Z_m = X_m W_m^T + b_m``, three propagation layers each, ``light_out_g = mean_k A_hat^k x^g_0``.
Propagation is linear and the item features ``X_m`` are constants of the dataset, so for a modality graph

    light_out_m = mean_k A_hat^k [E_u ; 0]  +  (mean_k A_hat^k [0 ; X_m | 1]) [W_m | b_m]^T
                = (parity part of p)        +            Zbar_m               W'_m^T

where ``p_k = A_hat^k [E_u ; E_i]`` is the id graph's own 64-wide propagation: ``A_hat`` is bipartite, so
``A_hat^k [E_u ; 0]`` lives on the user rows for even k and on the item rows for odd k - exactly the rows of ``p_k`` that
do not depend on ``E_i``.  ``Zbar_m`` ([N x (D_m + 1)], the bias rides along as a column of ones) is built once
(``_lin_build_zbar``).  A step then is

    forward   ONE 64-wide propagation (layers L-1 / L only at the rows the loss reaches), Zbar rows gathered at the 3B
              instance rows, a [3B x D_m] x [D_m x 64] GEMM per modality, ``lin_assemble`` -> O[inst rows]
    backward  dW'_m = dO_m[inst]^T Zbar_m[inst] (one small GEMM: weight AND bias gradient), one 64-wide backward chain
              seeded by ``lin_seed`` -> dE_u, dE_i

instead of one 64-wide plus three 256-wide propagations and two passes over the feature matrices (311 MB each at the
Tiktok shape).  Same loss and gradients up to fp32 reassociation (gated at the fp32 tolerance against the reference's
golden vectors, tests/test_linear_schedule.py).  The full tables for evaluation are completed on demand from the same
pieces (``_lin_materialize``).  Applies to the bipartite adjacency types ('pre', 'plain', 'gcmc') with ``lazy_tables``;
the self-loop types and the literal-tiktok dead word gradient keep the slab schedules of model.py.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import ElimrecError

D = 64


def _cfg(config, key, default):
    return config[key] if key in config else default


class LinearSchedule:
    """Mixin of ``EliMRec`` (model.py): everything specific to the linear schedule."""
    _lin_prefork = True      # _lin_forward consumes a stream forked at the start of a captured step (see there)

    # ---- construction ---------------------------------------------------------------------------------------------------
    def _lin_init(self):
        want = bool(_cfg(self.config, "linear_schedule", True))
        self.linear = bool(want and self.lazy_tables and not self._generic and not (self.tiktok and self.word_grad))
        # two_hop_masks: layer L-1 / the first backward hop only at the rows the instance rows reach (need2).  Off (default): those
        # layers run dense - no mark_neighbors launches, no dependency of layer L-1 on the batch.  Measured on B200, ms/step with
        # masks | dense: tiktok 0.410 | 0.399, kwai 0.481 | 0.446, movielens 0.549 | 0.550 (need2 covers most rows anyway: the
        # positives are popular items, and their neighbours are most users)
        self._lin_need2 = bool(_cfg(self.config, "two_hop_masks", False))
        self._lin_w = getattr(self, "_lin_w", D)        # columns of the propagated slabs held here (64; 64 / world when column-sharded)
        if self.proj_precision == "auto":
            self.proj_precision = "x3" if self.linear else "tf32"
        if not self.linear:
            if self.proj_precision == "x3":
                raise ElimrecError("proj_precision='x3' needs the linear schedule; the slab schedules take 'tf32' or 'fp32'")
            return
        dims = [self._feat[m].shape[1] for m in self.mods]
        self._lin_Kp = [((d + 1 + 3) // 4) * 4 for d in dims]          # [X_m | 1 | 0-pad] : multiple of 4 floats (TMA)
        self._lin_koff = [sum(self._lin_Kp[:j]) for j in range(len(dims))]
        self._lin_Ktot = sum(self._lin_Kp)
        self._lin_build_zbar()

    @torch.no_grad()
    def _lin_build_zbar(self):
        """Zbar [N x Ktot] (users first), block m = mean_k A_hat^k [0 ; X_m | 1 | 0] - constant.  Built with the propagation
        kernel itself, 256 / 128 / 64 columns at a time; in TF32 mode it is rounded (to nearest) once, like the features."""
        U, I, L, dev = self.num_users, self.num_items, self.n_layers, self.device_
        g, inv = self.graph, 1.0 / (self.n_layers + 1)
        Z = torch.zeros(U + I, self._lin_Ktot, dtype=torch.float32, device=dev)
        tmp = {}
        for j, m in enumerate(self.mods):
            X = self._feat[m]
            Dm, Kp, koff = X.shape[1], self._lin_Kp[j], self._lin_koff[j]
            Kc = -(-Kp // 64) * 64
            T0 = torch.zeros(I, Kc, dtype=torch.float32, device=dev)
            T0[:, :Dm].copy_(X)
            T0[:, Dm] = 1.0
            c0 = 0
            while c0 < Kc:
                w = 256 if Kc - c0 >= 256 else (128 if Kc - c0 >= 128 else 64)
                wa = min(w, Kp - c0)                 # columns of this chunk that exist in Zbar (the rest is padding)
                if w not in tmp:
                    tmp[w] = (torch.empty(U, w, dtype=torch.float32, device=dev), torch.empty(I, w, dtype=torch.float32, device=dev))
                tu, ti = tmp[w]
                zc = koff + c0
                cur = T0[:, c0:c0 + w]               # layer 0 lives on the item rows
                ops.axpy_2d(cur, Z[U:, zc:], I, wa, inv, accumulate=False)
                on_items = True
                for k in range(1, L + 1):
                    if on_items:
                        ops.spmm(g.ui, cur, tu, w)
                        ops.axpy_2d(tu, Z[:U, zc:], U, wa, inv, accumulate=(k > 1))
                        cur = tu
                    else:
                        ops.spmm(g.iu, cur, ti, w)
                        ops.axpy_2d(ti, Z[U:, zc:], I, wa, inv, accumulate=True)
                        cur = ti
                    on_items = not on_items
                c0 += w
            del T0
        if self.proj_precision == "tf32":
            ops.round_tf32(Z, Z)
        self._zbar = Z

    # ---- workspace ------------------------------------------------------------------------------------------------------
    def _lin_workspace_shared(self, ws):
        dev, U, I, L = self.device_, self.num_users, self.num_items, self.n_layers
        N = U + I
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        w = self._lin_w
        ws["P"] = [None] + [e(N, w) for _ in range(L)]      # p_1 .. p_L (users first); p_0 are the embedding tables themselves
        ws["H"] = [e(N, w), e(N, w)]                        # backward chain, ping-pong
        # the two seed vectors of the chain (lin_seed2), valid on the instance rows of the current step only
        ws["GA"], ws["GB"] = torch.zeros(N, w, dtype=torch.float32, device=dev), torch.zeros(N, w, dtype=torch.float32, device=dev)
        ws["E0"] = e(N, w)                                  # [E_u ; E_i] as THIS forward saw them (Adam overwrites the tables)
        ws["Wp"] = {m: e(D, kp) for m, kp in zip(self.mods, self._lin_Kp)}     # [W_m | b_m | 0] of this forward
        if self.proj_precision == "x3":
            ws["Wp_split"] = {m: (e(D, kp), e(D, kp)) for m, kp in zip(self.mods, self._lin_Kp)}   # its TF32 hi / lo parts
        # gradients of the small parameters, one flat buffer (what the data-parallel all-reduce sends in place): first the
        # packed projection gradients d[W_m | b_m] (weight and bias gradients are strided views of them), then the rest
        small = {n: p for n, p in self._params().items()
                 if not n.startswith("embedding_user.w") and not n.startswith("embedding_item.w")}
        n_pack = sum(D * kp for kp in self._lin_Kp)
        rest = [n for n in small if "_dense." not in n or n.startswith("s_dense")]
        ws["g_flat"] = torch.zeros(n_pack + sum(small[n].numel() for n in rest), dtype=torch.float32, device=dev)
        ws["g"], ws["dWp"], o = {}, {}, 0
        for m, kp in zip(self.mods, self._lin_Kp):
            blk = ws["g_flat"][o:o + D * kp].view(D, kp)
            dm = self._feat[m].shape[1]
            ws["dWp"][m] = blk
            ws["g"][f"{m}_dense.weight"], ws["g"][f"{m}_dense.bias"] = blk[:, :dm], blk[:, dm]
            o += D * kp
        for n in rest:
            ws["g"][n] = ws["g_flat"][o:o + small[n].numel()].view(small[n].shape)
            o += small[n].numel()
        assert o == ws["g_flat"].numel()
        ws["split_proj"], ws["dmax"] = 64, max(self._lin_Kp)       # sizes the split-K scratch of the exact-fp32 GEMMs

    def _lin_workspace_batch(self, ws, B):
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device_)
        ws["Zg"] = e(3 * B, self._lin_Ktot)                 # Zbar gathered at the instance rows
        ws["B"] = B
        ws["wg_splits"] = max(1, min(32, (3 * B + 255) // 256))
        if self.mm_fusion_mode == "mean":
            gWu, gWi = ws["g_eff"]["u"], ws["g_eff"]["i"]
        else:
            gWu, gWi = ws["g"]["embedding_user_after_GCN.weight"], ws["g"]["embedding_item_after_GCN.weight"]
        probs = self._lin_wgrad_problems(ws, gWu, gWi)
        n0 = 2 + len(self.mods)          # problems [:n0] contract the instance gradients, [n0:] contract dO[inst]
        groups = (probs, probs[:n0], probs[n0:])
        if self._lin_wgrad_x3():
            # tensor-core form: one CTA per SM (four 48 KB stages each), so the row ranges are sized to fill the SMs once
            fill = lambda pr: max(1, min(64, 148 // max(1, sum((p[4] + 127) // 128 for p in pr)), (3 * B + 31) // 32))
            ws["wg_splits"] = fill(probs)
            ws["wg_splits_g"] = (fill(groups[1]), fill(groups[2]))
        else:
            ws["wg_splits_g"] = (ws["wg_splits"], ws["wg_splits"])
        need = [ops.wgrad_multi_ws_floats(pr, sp) for pr, sp in zip(groups, (ws["wg_splits"],) + ws["wg_splits_g"]) if pr]
        ws["wg_ws"] = e(max(1, *need))

    def _lin_wgrad_x3(self):
        """weight gradients on the tensor cores (3xTF32) in the default accuracy class; exact FFMA with proj_precision='fp32'
        (and in 'tf32' mode, whose weight gradients were always exact) or wgrad_precision='fp32'"""
        return self.proj_precision == "x3" and str(_cfg(self.config, "wgrad_precision", "x3")) == "x3"

    def _lin_tc(self):
        return self.proj_precision == "tf32"

    def _lin_wgrad_problems(self, ws, gWu, gWi):
        """(A, a_col, B, b_col, K, row_begin, row_end, out, bias_out, scale_by_g) for elimrec_wgrad_multi"""
        B, Fw, gr = ws["B"], ws["F"], ws["g"]
        ig, Oin, dOin, Zg = ws["inst_grad"], ws["O_inst"], ws["dO_inst"], ws["Zg"]
        pr = [(ig, 0, Oin, 0, Fw, 0, B, gWu, gr["embedding_user_after_GCN.bias"], True),
              (ig, 0, Oin, 0, Fw, B, 3 * B, gWi, gr["embedding_item_after_GCN.bias"], True)]
        for j, m in enumerate(self.mods):
            c = D * (j + 1)
            pr.append((ig, c, Oin, c, D, 0, 3 * B, gr[f"s_dense_{m}.weight"], gr[f"s_dense_{m}.bias"], True))
        for j, (m, kp, ko) in enumerate(zip(self.mods, self._lin_Kp, self._lin_koff)):      # dO[inst] already carries g
            pr.append((dOin, D * (j + 1), Zg, ko, kp, 0, 3 * B, ws["dWp"][m], None, False))
        return pr

    def _lin_pack_weights(self, P, ws):
        """[W_m | b_m | 0] of this forward (TF32-rounded in 'tf32' mode; hi / lo split by _prep_weights in 'x3' mode)"""
        ops.pack_proj_weights([(P[f"{m}_dense.weight"].detach(), P[f"{m}_dense.bias"].detach(), ws["Wp"][m]) for m in self.mods],
                              self._lin_tc())
        extra = [(ws["Wp"][m], *ws["Wp_split"][m]) for m in self.mods] if self.proj_precision == "x3" else []
        self._prep_weights(P, ws, proj=False, extra=extra)

    def _lin_tables(self, ws, p0_u, p0_i):
        U = self.num_users
        return [(p0_u, p0_i)] + [(p[:U], p[U:]) for p in ws["P"][1:]]

    def _lin_modal_gemm(self, ws, Zrows, out, n_rows):
        """out[:, 64(1+j) : 64(2+j)] = Zrows[:, block j] @ W'_j^T for every modality (bias included: ones column)"""
        Fw = ws["F"]
        if self._lin_tc():
            ops.linear_tf32_fwd_multi([(Zrows[:, ko:ko + kp], ws["Wp"][m], None, out, D * (j + 1))
                                       for j, (m, kp, ko) in enumerate(zip(self.mods, self._lin_Kp, self._lin_koff))],
                                      tag="lin_modal_tc")
        elif self.proj_precision == "x3":
            sts = []
            for j, (m, kp, ko) in enumerate(zip(self.mods, self._lin_Kp, self._lin_koff)):     # disjoint output columns: side by side
                hi, lo = ws["Wp_split"][m]
                go = lambda: ops.linear_x3_fwd(Zrows[:, ko:ko + kp], hi, lo, None, out, col=D * (j + 1), tag="lin_modal_x3")
                if j == 0:
                    go()
                else:
                    sts.append(ops.fork_side(10 + j, high_priority=True))
                    with torch.cuda.stream(sts[-1]):
                        go()
            for st in sts:
                ops.join_side(st)
        else:
            ld = Zrows.stride(0)
            for j, (m, kp, ko) in enumerate(zip(self.mods, self._lin_Kp, self._lin_koff)):
                ops.gemm(n_rows, D, kp, Zrows, ld, 1, ws["Wp"][m], 1, kp, out, out.stride(0), 1, a_off=ko, c_off=D * (j + 1),
                         tag="lin_modal")

    # ---- forward --------------------------------------------------------------------------------------------------------
    def _lin_forward(self, users, pos, neg):
        P = self._params()
        U, I, L = self.num_users, self.num_items, self.n_layers
        B = int(users.numel())
        ws = self._workspace(B)
        g = self.graph
        Eu = P["embedding_user.weight"].detach()
        Ei = P["embedding_item.weight"].detach()
        mask, need2 = ws["mask"], ws["need2"]
        rows = ws["inst_rows"]
        # Small kernels run on HIGH-PRIORITY side streams: beside a propagation launch that fills every SM they would otherwise
        # queue behind its whole grid.
        #   side: instance rows + masks, then Zbar gathered at them and the modality GEMMs (nothing here needs the propagation)
        #   aux : what tables completed later must use (weights, E_u / E_i of THIS forward), zeroed seed rows of the backward
        side = ops.fork_side(0, high_priority=True)
        with torch.cuda.stream(side):
            ops.inst_rows(users, pos, neg, U, rows, mask, need2)
            ev_rows = torch.cuda.Event()
            ev_rows.record(side)
            if L >= 2 and self._lin_need2:    # layer L-1 is read at the graph neighbours of the instance rows
                ops.mark_neighbors(g.ui, mask[:U], need2[U:])
                ops.mark_neighbors(g.iu, mask[U:], need2[:U])
            ev_masks = torch.cuda.Event()
            ev_masks.record(side)
            ops.gather_rows(rows, self._zbar, ws["Zg"], self._lin_Ktot)
        aux = ops.fork_side(7, high_priority=True)
        with torch.cuda.stream(aux):
            self._lin_pack_weights(P, ws)
            if getattr(self, "_tick_early", False):      # train_step: the optimizer's step counter / bias corrections
                self._adam.tick()
            self._snapshot(P, ws)
            if not getattr(self, "_fuse_adam_now", False):      # (the fused Adam epilogue writes the pre-update rows itself)
                ops.copy_2d(Eu, ws["E0"][:U], U, D)
                ops.copy_2d(Ei, ws["E0"][U:], I, D)
            torch.cuda.current_stream().wait_event(ev_rows)
            ops.zero_rows(rows, 0, U + I, 0, ws["GA"], D)
            ops.zero_rows(rows, 0, U + I, 0, ws["GB"], D)
            ws["seed_zeroed"] = True
        with torch.cuda.stream(side):      # the modality GEMMs: packed weights (aux) x gathered Zbar rows (side)
            ops.join_side(aux)
            self._lin_modal_gemm(ws, ws["Zg"], ws["O_inst"], 3 * B)
        # the propagation: p_k = A_hat p_{k-1}, both halves 64 wide in one launch
        in_u, in_i = Eu, Ei
        # make_graphed_step forks a stream at the very start of the captured step (before the batch is drawn): the unmasked
        # layers depend on the tables only, so they start at t = 0 instead of behind the sampler and the first small nodes
        prop, self._prop_stream = getattr(self, "_prop_stream", None), None
        cur = prop if prop is not None else torch.cuda.current_stream()
        with torch.cuda.stream(prop):      # (None: the current stream)
            for k in range(1, L + 1):
                out = ws["P"][k]
                rm = mask if k == L else (need2 if (k == L - 1 and self._lin_need2) else None)
                if rm is not None:
                    cur.wait_event(ev_rows if k == L else ev_masks)
                ops.spmm64_pair(g.ui, g.iu, in_i, in_u, out[:U], out[U:], row_mask_u=rm[:U] if rm is not None else None,
                                row_mask_i=rm[U:] if rm is not None else None)
                in_u, in_i = out[:U], out[U:]
        if prop is not None:
            ops.join_side(prop)
        ops.join_side(side)
        lay = ops.lin_layers(self._lin_tables(ws, Eu, Ei))
        ops.lin_assemble(rows, U, lay, 1.0 / (L + 1), len(self.mods), True, ws["O_inst"])
        ops.join_side(aux)
        self._tables_version = getattr(self, "_tables_version", 0) + 1
        return self._loss(P, ws, users, pos, neg, gathered=True)

    # ---- backward -------------------------------------------------------------------------------------------------------
    def _lin_backward(self, gscale=None, split=False, fuse_adam=False):
        """``fuse_adam`` (train_step on one GPU): the last hop of the chain applies Adam to both embedding tables in its
        epilogue (their gradients are never stored) and records the pre-update rows for tables completed later."""
        P = self._params()
        ws = self._ws
        U, I, L = self.num_users, self.num_items, self.n_layers
        B, G, Fw, nt = ws["B"], ws["G"], ws["F"], ws["nt"]
        g = self.graph
        ig, gr = ws["inst_grad"], ws["g"]
        Oin, dOin, rows = ws["O_inst"], ws["dO_inst"], ws["inst_rows"]
        if gscale is not None:
            gscale = gscale.reshape(1)
        Wu, Wi = self._fusion_weights(P, ws)
        tied = self.mm_fusion_mode == "mean"
        gWu = ws["g_eff"]["u"] if tied else gr["embedding_user_after_GCN.weight"]
        gWi = ws["g_eff"]["i"] if tied else gr["embedding_item_after_GCN.weight"]
        ib = lambda part: ops.inst_backward(
            B, nt, Fw, ig, Oin, gscale, Wu, Wi, [P[f"s_dense_{m}.weight"].detach() for m in self.mods], dOin,
            gWu, gWi, gr["embedding_user_after_GCN.bias"], gr["embedding_item_after_GCN.bias"],
            [gr[f"s_dense_{m}.weight"] for m in self.mods], [gr[f"s_dense_{m}.bias"] for m in self.mods], ws["inst_ws"], part=part)

        def weights(group=None):
            """every gradient that is not an embedding table's (elimrec_wgrad_multi: one launch + its reduction per call) - none
            of it waits for the propagation backward.  group 0: fusion Linears and heads (they contract the instance
            gradients and can start before d O[inst] exists); group 1: d[W_m | b_m] = dO_m[inst]^T Zbar_m[inst] (weight AND
            bias: the ones column of Zbar); None: both in one launch."""
            probs = self._lin_wgrad_problems(ws, gWu, gWi)
            n0 = 2 + len(self.mods)
            sel = probs if group is None else (probs[:n0] if group == 0 else probs[n0:])
            splits = ws["wg_splits"] if group is None else ws["wg_splits_g"][group]
            if sel:
                ops.wgrad_multi(sel, splits, ws["wg_ws"], gscale, x3=self._lin_wgrad_x3())
            if tied and group != 1:
                ops.fold_blocks(gWu, gr["embedding_user_after_GCN.weight"], G, 1.0 / G)
                ops.fold_blocks(gWi, gr["embedding_item_after_GCN.weight"], G, 1.0 / G)

        # wgrad_overlap (default): the weight gradients run on a side stream beside the backward chain (the chain itself on a
        # high-priority stream).  Measured on tiktok-shape with the 54-register tensor-core kernel (ms/step): one launch after
        # d O[inst] 0.375 (default, wgrad_groups="one"); two launches, the instance-gradient problems already beside the
        # d O[inst] kernel 0.384 ("early": that kernel is on the critical path and slows down by what the hop gains); two
        # launches after d O[inst] 0.384 ("late"); in sequence before the chain 0.386 (wgrad_overlap=False).
        overlap = bool(_cfg(self.config, "wgrad_overlap", True)) and not split
        early = str(_cfg(self.config, "wgrad_groups", "one"))      # "one" | "early" | "late"
        sw = None
        if overlap and early == "early":
            sw = ops.fork_side(5)
            with torch.cuda.stream(sw):
                weights(0)
        # d O[inst] and the backward seeds (two vectors per instance row, see below) in one launch when the shapes allow
        fused_seed = bool(_cfg(self.config, "fused_seed", True)) and Fw == 256
        if fused_seed:
            if not ws.pop("seed_zeroed", False):
                ops.zero_rows(rows, 0, U + I, 0, ws["GA"], D)
                ops.zero_rows(rows, 0, U + I, 0, ws["GB"], D)
            ops.inst_dO_seed(B, nt, Fw, ig, gscale, Wu, Wi, [P[f"s_dense_{m}.weight"].detach() for m in self.mods], dOin, rows,
                             len(self.mods), 1.0 / (L + 1), ws["GA"], ws["GB"])
        else:
            ib(1)
        pending = None
        if overlap:
            sw = ops.fork_side(5)      # the same side stream, now also behind d O[inst]
            with torch.cuda.stream(sw):
                if early == "one":
                    weights(None)
                else:
                    if early != "early":
                        weights(0)
                    weights(1)
            pending = []
        elif not split:
            weights(None)      # in sequence, before the chain (measured with the exact FFMA kernel: 0.417 vs 0.399 ms/step)
            pending = []
        # the 64-wide backward chain: h_L = g_L, h_{k-1} = A_hat^T h_k + g_{k-1}.  g_k lives on the instance rows and takes two
        # values per row: GA (all blocks of dO) where layer k carries the modality graphs' E_u part (users: k even, items:
        # k odd), GB (id block) elsewhere.  h_L is read straight from the G slabs through the column mask; every later g_k is the
        # additive epilogue of the propagation launch.
        inv, nm = 1.0 / (L + 1), len(self.mods)
        GA, GB = ws["GA"], ws["GB"]
        if not fused_seed:
            if not ws.pop("seed_zeroed", False):      # normally done by the forward, off the critical path
                ops.zero_rows(rows, 0, U + I, 0, GA, D)
                ops.zero_rows(rows, 0, U + I, 0, GB, D)
            ops.lin_seed2(rows, dOin, nm, inv, GA, GB)
        mask, need2 = ws["mask"], ws["need2"]
        g_u = lambda k: (GA if k % 2 == 0 else GB)[:U]
        g_i = lambda k: (GA if k % 2 == 1 else GB)[U:]
        h_u, h_i, flip = g_u(L), g_i(L), 0
        def hops():
            nonlocal h_u, h_i, flip
            for k in range(L, 0, -1):
                nxt = ws["H"][flip]
                # h_L is valid on the instance rows only, h_{L-1} on need2 only: the first two hops drop every other column
                # (without the two-hop masks the first hop writes every row - exact zeros where nothing arrives - and the
                # second one runs dense)
                cm = mask if k == L else (need2 if (k == L - 1 and self._lin_need2) else None)
                rm = need2 if (k == L and L >= 2 and self._lin_need2) else None
                kw = {}
                if k == 1 and fuse_adam:
                    ad = self._adam
                    Eu, Ei = P["embedding_user.weight"], P["embedding_item.weight"]
                    kw = dict(adam_u=(Eu.data, *ad._st("embedding_user.weight", Eu), ws["E0"][:U]),
                              adam_i=(Ei.data, *ad._st("embedding_item.weight", Ei), ws["E0"][U:]),
                              adam_consts=(ad.consts, ad.betas[0], ad.betas[1], ad.eps, ad.wd))
                ops.spmm64_pair(g.ui_t, g.iu_t, h_i, h_u, nxt[:U], nxt[U:],
                                row_mask_u=rm[:U] if rm is not None else None, row_mask_i=rm[U:] if rm is not None else None,
                                col_mask_u=cm[U:] if cm is not None else None, col_mask_i=cm[:U] if cm is not None else None,
                                addend_u=g_u(k - 1), addend_i=g_i(k - 1), add_mask_u=mask[:U], add_mask_i=mask[U:], **kw)
                h_u, h_i, flip = nxt[:U], nxt[U:], flip ^ 1

        if overlap:
            chain = ops.fork_side(8, high_priority=True)
            with torch.cuda.stream(chain):
                hops()
            ops.join_side(chain)
            ops.join_side(sw)
        else:
            hops()
        grads = {} if fuse_adam else {"embedding_user.weight": h_u, "embedding_item.weight": h_i}
        ws["bw_pending"] = (weights, pending)
        if split:
            return grads
        grads.update(self._lin_backward_weights())
        return grads

    def _lin_backward_weights(self):
        ws = self._ws
        weights, pending = ws.pop("bw_pending")
        if pending is None:
            weights(None)
        dead = self._dead_params()
        return {n: gv for n, gv in ws["g"].items() if n not in dead} if dead else ws["g"]

    # ---- full tables for evaluation, on demand ------------------------------------------------------------------------------
    def _lin_materialize(self):
        """O over ALL rows from the pieces of the last forward: the masked layers completed (every row, same inputs), the
        modality blocks as Zbar W'^T with that forward's packed weights, then the fusion Linear + heads."""
        ws = self._ws
        U, I, L = self.num_users, self.num_items, self.n_layers
        N, g = U + I, self.graph
        E0 = ws["E0"]
        for k in range(max(1, L - 1 if self._lin_need2 else L), L + 1):      # the layers the step computed under a row mask
            src = E0 if k == 1 else ws["P"][k - 1]
            out = ws["P"][k]
            ops.spmm64_pair(g.ui, g.iu, src[U:], src[:U], out[:U], out[U:])
        self._lin_modal_gemm(ws, self._zbar, ws["O"], N)
        lay = ops.lin_layers(self._lin_tables(ws, E0[:U], E0[U:]))
        ops.lin_assemble(None, U, lay, 1.0 / (L + 1), len(self.mods), True, ws["O"], n_rows=N)
        self._dense_tables(None, ws, from_snapshot=True)
