"""Column-sharded ("feature-sharded") EliMRec - the multi-GPU mode built on the linear schedule.

BASELINE.json's north star names a row-sharded design (users / items partitioned, an NCCL all-gather of the propagated
shards per GCN layer: ``elimrec_b200/sharded.py``).  SURVEY.md section 7 asked to evaluate the alternative as well, and with
the linear schedule it wins outright: propagation acts on every column independently, so when rank r of G owns columns
``[r w, (r+1) w)``, ``w = 64 / G``, of both embedding tables and of every propagation / gradient slab (CSR replicated),

  * a propagation layer needs NO communication (the row-sharded design all-gathers the whole slab six times per step);
  * every rank draws its own batch, and what crosses NVLink is only what the losses read and write at the instance rows:
    one all-to-all of ``[G x 3B x 2w]`` floats forward (each rank receives all 64 columns of ITS 3B rows), one backward (each
    rank receives its columns of everybody's seed gradients), plus one all-reduce of the small weight gradients and one
    all-gather of the triples (3 x 2048 int64 per rank) - about 3 MB per rank and step at G = 8;
  * the propagation - the part of a step that does not depend on the batch at all - is done ONCE for the G batches instead
    of once per replica: a step of G x 2048 triples costs each rank 1/G of the columns of one propagation.

One optimizer step is taken on the mean loss of the G batches (exactly what G data-parallel replicas with averaged gradients
compute: ``tests/test_dist_gloo.py`` checks it against the single-process model).  Modality GEMMs, fusion Linear, heads, BPR
and their gradients run on the rank's own batch exactly as on one GPU (``linear.py``); Adam on the tables is column-local
(fused into the last backward hop), Adam on the small weights is replicated after the all-reduce.
Collectives: ``comm.py`` (NCCL behind the C-ABI on CUDA, capturable in the step's CUDA graph; gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops
from .comm import Comm
from .model import D, EliMRec, ElimrecError


class ColShardedEliMRec(EliMRec):
    _lin_prefork = False      # (its own _lin_forward)

    def _init_weight(self):
        if not (dist.is_available() and dist.is_initialized()):
            raise ElimrecError("ColShardedEliMRec needs an initialised torch.distributed process group")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if D % self.world or self.world > 8:
            raise ElimrecError(f"world size {self.world}: the 64 embedding columns are dealt in equal blocks of 32 / 16 / 8")
        self._lin_w = D // self.world
        super()._init_weight()
        if not self.linear:
            raise ElimrecError("the column-sharded mode is built on the linear schedule (bipartite adj_type, lazy_tables)")
        self._cs = None           # column shards of the two tables + their Adam moments
        self.comm = None

    # ---- shards ---------------------------------------------------------------------------------------------------------
    def _shards(self):
        """identical starting weights on every rank, then this rank's columns of the two embedding tables (done lazily:
        parameters move to the GPU / are loaded after construction)"""
        if self._cs is None:
            if self.comm is None:
                self.comm = Comm(self.device_)
            for p in self.parameters():
                dist.broadcast(p.data, src=0)
            c0, w = self.rank * self._lin_w, self._lin_w
            z = lambda t: torch.zeros_like(t)
            eu = self.embedding_user.weight.data[:, c0:c0 + w].contiguous()
            ei = self.embedding_item.weight.data[:, c0:c0 + w].contiguous()
            self._cs = dict(eu=eu, ei=ei, mu=z(eu), vu=z(eu), mi=z(ei), vi=z(ei), synced=True)
        return self._cs

    def _extra_state(self):
        cs = self._shards()
        return [cs[k] for k in ("eu", "ei", "mu", "vu", "mi", "vi")]

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self._cs = None
        return out

    def sync_parameters(self):
        """all-gather the column shards so that every rank holds the full, current embedding tables"""
        cs = self._cs
        if cs is None or cs["synced"]:
            return
        for shard, prm in ((cs["eu"], self.embedding_user.weight), (cs["ei"], self.embedding_item.weight)):
            buf = torch.empty(self.world, *shard.shape, dtype=shard.dtype, device=shard.device)
            self.comm.all_gather(buf, shard)
            prm.data.copy_(buf.permute(1, 0, 2).reshape(prm.shape))
        cs["synced"] = True

    def state_dict(self, *a, **k):
        self.sync_parameters()
        return super().state_dict(*a, **k)

    # ---- workspace ------------------------------------------------------------------------------------------------------
    def _lin_workspace_batch(self, ws, B):
        super()._lin_workspace_batch(ws, B)
        G, w, dev = self.world, self._lin_w, self.device_
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        ws["T_own"] = torch.zeros(3 * B, dtype=torch.int64, device=dev)            # users | pos | neg of this rank's batch
        ws["T_all"] = torch.zeros(G * 3 * B, dtype=torch.int64, device=dev)        # [G][3][B]
        ws["rows_all"] = torch.zeros(G * 3 * B, dtype=torch.int32, device=dev)     # node ids, [G][3B]
        ws["cs_send"], ws["cs_recv"] = e(G, 3 * B, 2 * w), e(G, 3 * B, 2 * w)

    _WS_PER_BATCH = EliMRec._WS_PER_BATCH + ("T_own", "T_all", "rows_all", "cs_send", "cs_recv")

    def _lin_workspace_shared(self, ws):
        super()._lin_workspace_shared(ws)
        # + one slot behind the small gradients: the rank's loss rides along in their all-reduce
        flat = torch.zeros(ws["g_flat"].numel() + 1, dtype=torch.float32, device=self.device_)
        off = {n: (v.data_ptr() - ws["g_flat"].data_ptr()) // 4 for n, v in ws["g"].items()}
        for n, v in list(ws["g"].items()):
            ws["g"][n] = torch.as_strided(flat, v.shape, v.stride(), off[n])
        for m, v in list(ws["dWp"].items()):
            ws["dWp"][m] = torch.as_strided(flat, v.shape, v.stride(), (v.data_ptr() - ws["g_flat"].data_ptr()) // 4)
        ws["g_flat"] = flat
        ws["cs_all"] = None      # [N x 2w] / [G x N x 2w] buffers of the table completion, allocated on first evaluation

    # ---- forward --------------------------------------------------------------------------------------------------------
    def _lin_forward(self, users, pos, neg):
        P = self._params()
        U, I, L, G, w = self.num_users, self.num_items, self.n_layers, self.world, self._lin_w
        B = int(users.numel())
        cs = self._shards()
        ws = self._workspace(B)
        g = self.graph
        Eu, Ei = cs["eu"], cs["ei"]
        mask, need2 = ws["mask"], ws["need2"]
        T_own, T_all, rows_all = ws["T_own"], ws["T_all"], ws["rows_all"]
        rows = rows_all[self.rank * 3 * B:(self.rank + 1) * 3 * B]            # this rank's instance rows
        ws["inst_rows"] = rows
        # Layer 1 does not depend on the batch: it starts at once on the current stream.  Everything about the instance rows
        # runs beside it on a high-priority stream: every rank needs every batch's rows (it propagates its columns for all).
        side = ops.fork_side(0, high_priority=True)
        with torch.cuda.stream(side):
            if users.data_ptr() != T_own.data_ptr():
                T_own[:B].copy_(users); T_own[B:2 * B].copy_(pos); T_own[2 * B:].copy_(neg)
            self.comm.all_gather(T_all, T_own)
            ops.cs_inst_rows(G, B, T_all, U, rows_all, mask, need2)
            ev_rows = torch.cuda.Event()
            ev_rows.record(side)
            if L >= 2 and self._lin_need2:
                ops.mark_neighbors(g.ui, mask[:U], need2[U:])
                ops.mark_neighbors(g.iu, mask[U:], need2[:U])
            ev_masks = torch.cuda.Event()
            ev_masks.record(side)
            ops.gather_rows(rows, self._zbar, ws["Zg"], self._lin_Ktot)
        aux = ops.fork_side(7, high_priority=True)
        with torch.cuda.stream(aux):
            self._lin_pack_weights(P, ws)
            if getattr(self, "_tick_early", False):
                self._adam.tick()
            self._snapshot(P, ws)
            torch.cuda.current_stream().wait_event(ev_rows)
            ops.zero_rows(rows_all, 0, U + I, 0, ws["GA"], w)
            ops.zero_rows(rows_all, 0, U + I, 0, ws["GB"], w)
            ws["seed_zeroed"] = True
        with torch.cuda.stream(side):
            ops.join_side(aux)
            self._lin_modal_gemm(ws, ws["Zg"], ws["O_inst"], 3 * B)
        # the propagation of this rank's w columns: no communication
        in_u, in_i = Eu, Ei
        cur = torch.cuda.current_stream()
        for k in range(1, L + 1):
            out = ws["P"][k]
            rm = mask if k == L else (need2 if (k == L - 1 and self._lin_need2) else None)
            if rm is not None:
                cur.wait_event(ev_rows if k == L else ev_masks)
            ops.spmm64_pair(g.ui, g.iu, in_i, in_u, out[:U], out[U:], row_mask_u=rm[:U] if rm is not None else None,
                            row_mask_i=rm[U:] if rm is not None else None, width=w)
            in_u, in_i = out[:U], out[U:]
        cur.wait_event(ev_rows)
        # exchange: my columns of everybody's instance rows  ->  all 64 columns of my instance rows
        lay = ops.lin_layers(self._lin_tables(ws, Eu, Ei))
        ops.cs_pack(rows_all, U, lay, 1.0 / (L + 1), w, ws["cs_send"])
        self.comm.all_to_all(ws["cs_recv"], ws["cs_send"])
        ops.join_side(side)
        ops.cs_unpack(G, 3 * B, w, ws["cs_recv"], len(self.mods), ws["O_inst"])
        ops.join_side(aux)
        self._tables_version = getattr(self, "_tables_version", 0) + 1
        cs["synced"] = False
        return self._loss(P, ws, users, pos, neg, gathered=True)

    # ---- backward -------------------------------------------------------------------------------------------------------
    def _lin_backward(self, gscale=None, split=False, fuse_adam=False):
        """always applies Adam to this rank's columns of the two tables in the last hop; returns the (averaged) small gradients"""
        if gscale is not None or split:
            raise ElimrecError("ColShardedEliMRec trains through train_step() (the table gradients never leave their rank)")
        P = self._params()
        ws, cs = self._ws, self._cs
        U, I, L, G, w = self.num_users, self.num_items, self.n_layers, self.world, self._lin_w
        B, Fw, nt = ws["B"], ws["F"], ws["nt"]
        g = self.graph
        gr = ws["g"]
        dOin, rows_all = ws["dO_inst"], ws["rows_all"]
        Wu, Wi = self._fusion_weights(P, ws)
        tied = self.mm_fusion_mode == "mean"
        gWu = ws["g_eff"]["u"] if tied else gr["embedding_user_after_GCN.weight"]
        gWi = ws["g_eff"]["i"] if tied else gr["embedding_item_after_GCN.weight"]
        ops.inst_backward(B, nt, Fw, ws["inst_grad"], ws["O_inst"], None, Wu, Wi, [P[f"s_dense_{m}.weight"].detach() for m in self.mods],
                          dOin, gWu, gWi, gr["embedding_user_after_GCN.bias"], gr["embedding_item_after_GCN.bias"],
                          [gr[f"s_dense_{m}.weight"] for m in self.mods], [gr[f"s_dense_{m}.bias"] for m in self.mods], ws["inst_ws"],
                          part=1)
        # exchange: the seeds of my rows, cut into column slices  ->  my columns of everybody's seeds (loss = mean over ranks)
        inv, nm = 1.0 / (L + 1), len(self.mods)
        ops.cs_seed_pack(3 * B, G, w, dOin, nm, inv / G, ws["cs_send"])
        self.comm.all_to_all(ws["cs_recv"], ws["cs_send"])
        GA, GB = ws["GA"], ws["GB"]
        if not ws.pop("seed_zeroed", False):
            ops.zero_rows(rows_all, 0, U + I, 0, GA, w)
            ops.zero_rows(rows_all, 0, U + I, 0, GB, w)
        ops.cs_seed_scatter(rows_all, w, ws["cs_recv"], GA, GB)
        mask, need2 = ws["mask"], ws["need2"]
        g_u = lambda k: (GA if k % 2 == 0 else GB)[:U]
        g_i = lambda k: (GA if k % 2 == 1 else GB)[U:]
        h_u, h_i, flip = g_u(L), g_i(L), 0
        ad = self._adam
        chain = ops.fork_side(8, high_priority=True)
        with torch.cuda.stream(chain):
            for k in range(L, 0, -1):
                nxt = ws["H"][flip]
                cm = mask if k == L else (need2 if (k == L - 1 and self._lin_need2) else None)
                rm = need2 if (k == L and L >= 2 and self._lin_need2) else None
                kw = {}
                if k == 1:
                    kw = dict(adam_u=(cs["eu"], cs["mu"], cs["vu"], ws["E0"][:U]), adam_i=(cs["ei"], cs["mi"], cs["vi"], ws["E0"][U:]),
                              adam_consts=(ad.consts, ad.betas[0], ad.betas[1], ad.eps, ad.wd))
                ops.spmm64_pair(g.ui_t, g.iu_t, h_i, h_u, nxt[:U], nxt[U:],
                                row_mask_u=rm[:U] if rm is not None else None, row_mask_i=rm[U:] if rm is not None else None,
                                col_mask_u=cm[U:] if cm is not None else None, col_mask_i=cm[:U] if cm is not None else None,
                                addend_u=g_u(k - 1), addend_i=g_i(k - 1), add_mask_u=mask[:U], add_mask_i=mask[U:], width=w, **kw)
                h_u, h_i, flip = nxt[:U], nxt[U:], flip ^ 1
        # small weights: gradients of my batch, averaged over the ranks (the loss value rides along)
        ops.wgrad_multi(self._lin_wgrad_problems(ws, gWu, gWi), ws["wg_splits"], ws["wg_ws"], None, x3=self._lin_wgrad_x3())
        if tied:
            ops.fold_blocks(gWu, gr["embedding_user_after_GCN.weight"], ws["G"], 1.0 / ws["G"])
            ops.fold_blocks(gWi, gr["embedding_item_after_GCN.weight"], ws["G"], 1.0 / ws["G"])
        ws["g_flat"][-1:].copy_(ws["loss"])
        self.comm.all_reduce(ws["g_flat"], average=True)
        ops.join_side(chain)
        dead = self._dead_params()
        return {n: gv for n, gv in gr.items() if n not in dead}

    def train_step(self, users, pos_items, neg_items):
        """one optimizer step on the mean loss of the `world` batches; returns that mean loss (device scalar)"""
        if self._adam is None:
            self.make_optimizer()
        users, pos, neg = self._triples(users, pos_items, neg_items)
        with torch.no_grad():
            self._tick_early = True
            try:
                self._forward(users, pos, neg)
            finally:
                self._tick_early = False
            grads = self._backward(None)
            self._adam.apply(grads, tick=False)
        return self._ws["g_flat"][-1]

    def bpr_loss(self, users, pos_items, neg_items):
        raise NotImplementedError("ColShardedEliMRec trains through train_step(): Adam state is sharded with the columns")

    def enable_data_parallel(self):
        raise ElimrecError("ColShardedEliMRec is itself the multi-GPU mode")

    # ---- full tables for evaluation ------------------------------------------------------------------------------------------
    def _lin_materialize(self):
        ws = self._ws
        U, I, L, G, w = self.num_users, self.num_items, self.n_layers, self.world, self._lin_w
        N, g = U + I, self.graph
        E0 = ws["E0"]
        for k in range(max(1, L - 1 if self._lin_need2 else L), L + 1):
            src = E0 if k == 1 else ws["P"][k - 1]
            out = ws["P"][k]
            ops.spmm64_pair(g.ui, g.iu, src[U:], src[:U], out[:U], out[U:], width=w)
        if ws.get("cs_all") is None:
            ws["cs_all"] = (torch.empty(N, 2 * w, dtype=torch.float32, device=self.device_),
                            torch.empty(G, N, 2 * w, dtype=torch.float32, device=self.device_))
        mine, everyone = ws["cs_all"]
        lay = ops.lin_layers(self._lin_tables(ws, E0[:U], E0[U:]))
        ops.cs_pack(None, U, lay, 1.0 / (L + 1), w, mine, n_rows=N)
        self.comm.all_gather(everyone, mine)
        self._lin_modal_gemm(ws, self._zbar, ws["O"], N)
        ops.cs_unpack(G, N, w, everyone, len(self.mods), ws["O"])
        self._dense_tables(None, ws, from_snapshot=True)
