"""``BasicModel`` / ``EliMRec`` - drop-in for the reference's ``models/BasicModel.py`` and
``models/EliMRec.py``, running entirely on hand-written sm_100a kernels (``libelimrec_b200.so``).

Same constructor, parameter names/shapes/init order (so ``state_dict()`` interchanges and
``torch.manual_seed`` gives the same initial weights), same public methods:

    EliMRec(config, dataset)                     models/EliMRec.py:32-36
    .bpr_loss(users, pos, neg) -> scalar Tensor  :115-142   (supports .backward(retain_graph=True))
    .predict(user_ids, candidate_items=None)     :96-113    ([B x I] fp32 CPU tensor)
    .evaluate() / .test() / .getFileName()       models/BasicModel.py:34-50
    .predict_type = 'TIE' | 'TE' | 'normal'      main.py:121,140
    .all_users / .all_items / .all_s_embs        tables cached by the last training forward

What is different underneath:
  * the 4 modality graphs are propagated as ONE wide SpMM per layer plus one 64-wide SpMM for the
    rows that are identical across graphs (bipartite dedup, 0.625x the edge work), with the layer
    mean fused into the last layer's epilogue;
  * the BPR losses and their backward are one fused kernel that emits a ROW-SPARSE gradient
    (3B instance rows); the dense [N x 64] table gradients of the reference never exist;
  * ``train_step`` = forward + backward + fused Adam without going through autograd at all
    (and is CUDA-graph capturable); ``bpr_loss`` wraps the same kernels in one autograd.Function
    for the unmodified ``main.py`` loop.

There is no CPU path: constructing the model without a CUDA device raises.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from ._lib import CALLS, ElimrecError
from .graph import BipartiteGraph
from .linear import LinearSchedule

D = 64
PREDICT_MODE = {"normal": 0, "TE": 1, "TIE": 2}
FUSION_MODE = {"rubi": 0, "hm": 1, "sum": 2}      # s_fusion_mode; elimrec_rank_tables_t.mode = predict + 4 * fusion
N_WORDS, WORD_DIM = 11574, 128                    # literal 'tiktok' text branch (EliMRec.py:374-375)


def _require_cuda(dev):
    if dev.type != "cuda":
        raise ElimrecError("elimrec_b200 has no CPU path: config.device must be a CUDA device")


def _cfg(config, key, default):
    return config[key] if key in config else default


class BasicModel(nn.Module):
    """models/BasicModel.py:9-50 (constructor, evaluate/test, getFileName)."""

    def __init__(self, dataset, config):
        super().__init__()
        from .evaluator import ProxyEvaluator
        self.config = config
        self.dataset = dataset
        train = dataset.get_user_train_dict()
        kw = dict(metric=config["metric"], group_view=config["group_view"], top_k=config["topks"],
                  batch_size=config["test_batch_size"], num_thread=config["num_thread"])
        self.valid_evaluator = ProxyEvaluator(dataset, train, dataset.get_user_valid_dict(), None, **kw)
        self.test_evaluator = ProxyEvaluator(dataset, train, dataset.get_user_test_dict(), None, **kw)

    def getFileName(self):
        suffix = self.config["suffix"]
        if not os.path.exists(self.config.path):
            os.mkdir(self.config.path)
        file = f"{self.config.recommender}-{self.config['data.input.dataset']}-{self.config['loss']}-{suffix}.pth.tar"
        return os.path.join(self.config.path, file)

    def predict(self, user_ids, candidate_items=None):
        raise NotImplementedError

    def evaluate(self):
        return self.valid_evaluator.evaluate(self)

    def test(self):
        return self.test_evaluator.evaluate(self)


class _StepFunction(torch.autograd.Function):
    """One autograd node for the whole training forward; backward runs the fused backward chain."""

    @staticmethod
    def forward(ctx, model, users, pos, neg, *params):
        ctx.model = model
        ctx.names = model._param_names
        return model._forward(users, pos, neg).clone()

    @staticmethod
    def backward(ctx, grad_out):
        model = ctx.model
        grads = model._backward(grad_out.contiguous().float())
        # autograd may keep what we return as .grad -> hand out copies, never workspace views
        return (None, None, None, None) + tuple(
            (grads[n].clone() if grads.get(n) is not None else None) for n in ctx.names)


class EliMRec(LinearSchedule, BasicModel):
    def __init__(self, config, dataset):
        super().__init__(dataset, config)
        self._init_weight()

    # ------------------------------------------------------------------------------------------
    # construction: same order of parameter creation / initialisation as EliMRec.py:38-93,356-407
    # ------------------------------------------------------------------------------------------
    def _init_weight(self):
        cfg, ds = self.config, self.dataset
        self.num_users, self.num_items = int(ds.num_users), int(ds.num_items)
        self.latent_dim = cfg["recdim"]
        if self.latent_dim != D:
            raise ElimrecError(f"recdim={self.latent_dim}: kernels are built for recdim=64 (conf/EliMRec.properties:6)")
        self.n_layers = cfg["layer_num"]
        if not 1 <= self.n_layers <= 8:
            raise ElimrecError("layer_num must be in [1, 8]")
        self.predict_type = _cfg(cfg, "predict_type", "TIE")
        self.mm_fusion_mode = _cfg(cfg, "mm_fusion_mode", "concat")
        self.fusion_mode = _cfg(cfg, "s_fusion_mode", "rubi")
        if self.mm_fusion_mode not in ("concat", "mean"):       # EliMRec.py:221-226
            raise ElimrecError(f"mm_fusion_mode={self.mm_fusion_mode!r}: expected 'concat' or 'mean'")
        if self.fusion_mode not in FUSION_MODE:                 # EliMRec.py:171-210
            raise ElimrecError(f"s_fusion_mode={self.fusion_mode!r}: expected 'rubi', 'hm' or 'sum'")
        self.modality = _cfg(cfg, "modality", "vat")
        # modal projections (v/a/t_dense): 'tf32' = tcgen05 tensor cores, TF32 operands (1e-3 parity class); 'fp32' = exact FFMA
        # path (1e-5 class); 'x3' = 3xTF32 on the tensor cores (1e-5 class; linear schedule only, where the GEMM runs on the 3B
        # instance rows and costs next to nothing).  Default 'auto': 'x3' with the linear schedule, else 'tf32'.
        self.proj_precision = _cfg(cfg, "proj_precision", "auto")
        if self.proj_precision not in ("auto", "tf32", "fp32", "x3"):
            raise ElimrecError("proj_precision must be 'auto', 'tf32', 'x3' or 'fp32'")
        # fusion Linear + single-modal heads: 'x3' = 3xTF32 on the tensor cores (fp32-class accuracy), 'fp32' = FFMA
        self.fuse_precision = _cfg(cfg, "fuse_precision", "x3")
        if self.fuse_precision not in ("x3", "fp32"):
            raise ElimrecError("fuse_precision must be 'x3' or 'fp32'")
        # the same two layers on the 3B instance rows of a row-sparse step: 'x3' = the tensor-core kernel on 128-row tiles, users
        # and items side by side (default: 21 us at Tiktok shape), 'ffma' = one exact-fp32 launch over 3B/16 CTAs (29 us)
        self.inst_fuse = _cfg(cfg, "inst_fuse", "x3")
        if self.inst_fuse not in ("ffma", "x3"):
            raise ElimrecError("inst_fuse must be 'ffma' or 'x3'")
        # lazy_tables (default on): a training step computes exactly what its loss and gradients read.  The BPR loss
        # indexes the fused / single-modal tables with the <= 3B sampled rows (EliMRec.py:120-128), so the step produces
        #   * the last propagation layer and its layer mean at those rows only (row-masked SpMM),
        #   * layer L-1's wide side at the rows those read (their graph neighbours),
        #   * fusion Linear + heads on the 3B gathered rows,
        # and the backward skips the columns where d x_L / d x_{L-1} are structurally zero.  Loss, gradients and updated
        # parameters are the reference's (same sums, same order).  The full all_users / all_items / all_s_embs tables -
        # which the reference materialises every step (EliMRec.py:261-272,144-153) and only predict() reads - are completed
        # on first access from the layer inputs and weights THAT forward saw: same values, once per evaluation instead of
        # once per step.  lazy_tables=False runs the reference's schedule (every row, every step).
        self.lazy_tables = bool(_cfg(cfg, "lazy_tables", True))
        # row-sparse step only: the layer-mean gradient G (non-zero on the <= 3B instance rows) is kept in two slabs and added by
        # the backward SpMMs' epilogue instead of by a scatter kernel after every SpMM (2 scatters per step instead of 8).
        # Measured on B200 (Tiktok shape): 0.6845 vs 0.6800 ms/step - the scatters already hide under the other stream's
        # SpMM, while the epilogue costs the dense launches registers - so it is off by default.
        self.fused_layer_grad = bool(_cfg(cfg, "fused_layer_grad", False))
        self.kwai = cfg["data.input.dataset"] == "kwai"
        # literal 'tiktok' (EliMRec.py:371-378): the text feature is the mean word embedding of the item's words, computed
        # ONCE from the initial word_embedding; the parameter keeps receiving gradients / Adam updates that never reach
        # the outputs.  word_grad=False drops that dead gradient (the parameter then only stays at its initial value).
        self.tiktok = cfg["data.input.dataset"] == "tiktok" and hasattr(ds, "words_tensor")
        self.word_grad = bool(_cfg(cfg, "word_grad", True))
        self.mods = "v" if self.kwai else "vat"
        # main.py:36 sets `config.device` as a plain ATTRIBUTE of the Configurator (not a key: `"device" in config` is False)
        try:
            dev = cfg.device
        except (KeyError, AttributeError):
            dev = None
        dev = torch.device(dev) if dev is not None else torch.device("cuda", torch.cuda.current_device())
        _require_cuda(dev)
        self.device_ = dev

        self.embedding_user = nn.Embedding(self.num_users, D)
        self.embedding_item = nn.Embedding(self.num_items, D)
        nn.init.xavier_uniform_(self.embedding_user.weight)
        nn.init.xavier_uniform_(self.embedding_item.weight)
        # item features, L2-row-normalised once (EliMRec.py:366-381); constant afterwards
        self._feat = {m: F.normalize(getattr(ds, f"{m}_feat").to(dev).float(), dim=1).contiguous() for m in self.mods
                      if not (self.tiktok and m == "t")}
        if self.tiktok:
            self.words_tensor = ds.words_tensor.to(dev)
            self.word_embedding = nn.Embedding(N_WORDS, WORD_DIM)
            nn.init.xavier_normal_(self.word_embedding.weight)
            self._build_word_graph()
            self._feat["t"] = torch.empty(self.num_items, WORD_DIM, dtype=torch.float32, device=dev)
            self.rebuild_text_feature()
        for m in self.mods:
            setattr(self, f"{m}_feat", self._feat[m])
        for m in self.mods:
            setattr(self, f"{m}_dense", nn.Linear(self._feat[m].shape[1], D))
        self.item_feat_dim = D * (1 + len(self.mods)) if self.mm_fusion_mode == "concat" else D
        for m in self.mods:
            nn.init.xavier_uniform_(getattr(self, f"{m}_dense").weight)
        self.embedding_user_after_GCN = nn.Linear(self.item_feat_dim, D)
        nn.init.xavier_uniform_(self.embedding_user_after_GCN.weight)
        self.embedding_item_after_GCN = nn.Linear(self.item_feat_dim, D)
        nn.init.xavier_uniform_(self.embedding_item_after_GCN.weight)
        self._all_users = self._all_items = self._all_s_embs = None
        self._tables_pending = False
        self.graph = BipartiteGraph(ds.train_matrix, dev, cfg["adj_type"])
        # adj_type 'norm' / 'mean' put a diagonal into A_hat: user rows then depend on user rows, the bipartite dedup and
        # the row-sparse step no longer apply -> generic schedule (every layer F wide on both sides, every row)
        self._generic = self.graph.self_loops
        if self._generic:
            self.lazy_tables = False
        self.f = nn.Sigmoid()
        self.s_dense_v = nn.Linear(D, D)
        self.s_dense_a = nn.Linear(D, D)
        self.s_dense_t = nn.Linear(D, D)
        nn.init.xavier_uniform_(self.s_dense_v.weight)
        nn.init.xavier_uniform_(self.s_dense_a.weight)
        nn.init.xavier_uniform_(self.s_dense_t.weight)
        self._ws = None
        self._adam = None
        # linear schedule (linear.py): the modality graphs by linearity from one 64-wide propagation + constant tables
        self._lin_init()

    # ---- literal 'tiktok' text branch ---------------------------------------------------------------
    def _build_word_graph(self):
        """t_feat = M @ word_embedding with M[i, w] = (#times word w is listed for item i) / (#words of item i): the
        scatter-mean of EliMRec.py:377-378 as a CSR SpMM, and M^T for the gradient that flows back into word_embedding."""
        import scipy.sparse as sp
        from .graph import CsrHalf
        w = self.words_tensor.cpu().numpy()
        I = self.num_items
        cnt = np.bincount(w[0], minlength=I).astype(np.float32)
        m = sp.csr_matrix(((np.float32(1.0) / cnt[w[0]]).astype(np.float32), (w[0], w[1])), shape=(I, N_WORDS))
        m.sum_duplicates()
        m.sort_indices()
        mt = m.T.tocsr()
        mt.sort_indices()
        self._word_half = CsrHalf(m.indptr, m.indices, m.data, N_WORDS, self.device_)
        self._word_half_t = CsrHalf(mt.indptr, mt.indices, mt.data, I, self.device_)

    @torch.no_grad()
    def rebuild_text_feature(self):
        """Re-derive t_feat from the CURRENT word_embedding.  The reference does this exactly once, inside its constructor
        (EliMRec.py:376-378); call it after loading a state_dict whose word_embedding differs from this model's init."""
        w = self.word_embedding.weight.detach().to(self.device_, torch.float32).contiguous()
        ops.spmm(self._word_half, w, self._feat["t"], WORD_DIM)
        self.__dict__.pop("_feat_tf32", None)
        if getattr(self, "linear", False):      # the constant tables of the linear schedule derive from the features
            self._lin_build_zbar()

    def _feat_tc(self, m):
        """Features pre-rounded (to nearest) to TF32 once; what the tensor-core projections stream."""
        cache = self.__dict__.setdefault("_feat_tf32", {})
        if m not in cache:
            cache[m] = torch.empty_like(self._feat[m])
            ops.round_tf32(self._feat[m], cache[m])
        return cache[m]

    # tables cached by the last training forward (EliMRec.py:98-99,109); materialised on demand when lazy_tables is on
    @property
    def all_users(self):
        self._materialize_tables()
        return self._all_users

    @all_users.setter
    def all_users(self, v):
        self._all_users = v

    @property
    def all_items(self):
        self._materialize_tables()
        return self._all_items

    @all_items.setter
    def all_items(self, v):
        self._all_items = v

    @property
    def all_s_embs(self):
        self._materialize_tables()
        return self._all_s_embs

    @all_s_embs.setter
    def all_s_embs(self, v):
        self._all_s_embs = v

    # parameters that take part in the computation, in a fixed order
    @property
    def _param_names(self):
        names = ["embedding_user.weight", "embedding_item.weight"]
        if self.tiktok and self.word_grad:
            names.append("word_embedding.weight")
        for m in self.mods:
            names += [f"{m}_dense.weight", f"{m}_dense.bias"]
        names += ["embedding_user_after_GCN.weight", "embedding_user_after_GCN.bias",
                  "embedding_item_after_GCN.weight", "embedding_item_after_GCN.bias"]
        for m in self.mods:
            names += [f"s_dense_{m}.weight", f"s_dense_{m}.bias"]
        return names

    def _params(self):
        d = dict(self.named_parameters())
        return {n: d[n] for n in self._param_names}

    # ------------------------------------------------------------------------------------------
    # workspace: every buffer of a step, allocated once (static addresses => CUDA-graph friendly)
    # ------------------------------------------------------------------------------------------
    # keys of the workspace that depend on the batch size (everything else - slabs, tables, gradient buffers, ~all of the
    # memory - is allocated once and shared by every batch size; PairwiseSamplerV2 has drop_last=False, so the last batch of
    # an epoch is short) and keys that describe the LAST forward
    _WS_PER_BATCH = ("B", "density", "terms", "inst_rows", "inst_grad", "O_inst", "dO_inst", "split_inst", "gemm_ws", "inst_ws",
                     "inst_dummy", "F_c", "S_c", "c_users", "c_pos", "c_neg", "Zg", "wg_splits", "wg_ws")
    _WS_PER_FORWARD = ("pre_last_layer", "last_layer", "seed_zeroed", "bw_pending")

    def _workspace(self, B):
        ws = self._ws
        if ws is not None and ws["B"] == B:
            return ws
        by_B = self.__dict__.setdefault("_ws_by_B", {})
        if B in by_B:
            self._ws = by_B[B]
            return self._ws
        if by_B:      # another batch size: share every batch-independent buffer, allocate only the 3B-row ones
            ws = {k: v for k, v in next(iter(by_B.values())).items()
                  if k not in self._WS_PER_BATCH and k not in self._WS_PER_FORWARD}
            self._workspace_batch(ws, B)
            by_B[B] = self._ws = ws
            return ws
        dev, U, I, L = self.device_, self.num_users, self.num_items, self.n_layers
        N, G = U + I, 1 + len(self.mods)
        Fw = D * G
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        ws = dict(G=G, F=Fw, nt=G, cache={})  # cache: tables derived from the last forward (normalised heads, fp16 splits, DP bucket)
        ws["mask"] = torch.zeros(N, dtype=torch.uint8, device=dev)        # 1 on the <= 3B instance rows of the step
        ws["need2"] = torch.zeros(N, dtype=torch.uint8, device=dev)       # instance rows + the rows they gather (2 hops)
        ws["O"] = e(N, Fw)
        ws["F_all"] = e(N, D)
        ws["S"] = [e(N, D) for _ in self.mods]
        ws["loss"] = e(1)
        if self.mm_fusion_mode == "mean":   # tied fusion weights [W/G | ... | W/G] and the gradient w.r.t. them
            ws["W_eff"] = {"u": e(D, Fw), "i": e(D, Fw)}
            ws["g_eff"] = {"u": e(D, Fw), "i": e(D, Fw)}
        # snapshot of the fusion / head weights and biases used by the last forward (what lazily built tables must use)
        ws["snap_names"] = (["embedding_user_after_GCN.bias", "embedding_item_after_GCN.bias"] +
                            [f"s_dense_{m}.bias" for m in self.mods])
        if self.fuse_precision != "x3":
            ws["snap_names"] += (["embedding_user_after_GCN.weight", "embedding_item_after_GCN.weight"] +
                                 [f"s_dense_{m}.weight" for m in self.mods])
        Pn = self._params()
        ws["snap"] = {n: (e(D, Fw) if n.endswith("after_GCN.weight") else torch.empty_like(Pn[n], device=dev))
                      for n in ws["snap_names"]}
        ws["snap_dst"] = [ws["snap"][n] for n in ws["snap_names"]]
        ws["W_split"] = {"u": (e(D, Fw), e(D, Fw)), "i": (e(D, Fw), e(D, Fw))}
        for m in self.mods:
            ws["W_split"][m] = (e(D, D), e(D, D))
        if self.linear:
            self._lin_workspace_shared(ws)
        else:
            self._workspace_shared(ws)
        self._workspace_batch(ws, B)
        by_B[B] = self._ws = ws
        return ws

    def _workspace_shared(self, ws):
        """batch-independent buffers of the slab schedules (reference / row-sparse / generic)"""
        dev, U, I, L = self.device_, self.num_users, self.num_items, self.n_layers
        N, Fw = U + I, ws["F"]
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        if not self._generic:
            ws["X0_i"] = e(I, Fw)
        ws["X0_u"] = e(U, D)                                              # E_u as of the last forward (lazy tables)
        rows = lambda side: U if side == "u" else I
        ws["XW"], ws["XN"] = {}, {}
        if self._generic:
            # self loops: every layer is a full [N x F] slab (users first); layer 0's item rows are the projection slab
            ws["GX"] = [e(N, Fw) for _ in range(L + 1)]
            ws["X0_i"] = ws["GX"][0][U:]
            ws["GdX"] = [e(N, Fw), e(N, Fw)]
            ws["dE_u"] = e(U, D)
        else:
            for k in range(1, L):  # layer L is consumed by the fused mean epilogue and never stored
                side = "u" if k % 2 == 1 else "i"
                ws["XW"][k] = e(rows(side), Fw)
                ws["XN"][k] = e(rows("i" if side == "u" else "u"), D)
        R = max(U, I)
        if not self._generic:
            ws["dW"] = [e(R, Fw), e(R, Fw)]
            ws["dN"] = [e(R, D), e(R, D)]
            if self.lazy_tables and self.fused_layer_grad:
                # G wide / G folded over the graph blocks, for all N nodes; only the instance rows are ever written or read
                ws["Gw"], ws["Gn"] = e(N, Fw), e(N, D)
        if self.tiktok and self.word_grad:
            ws["dT"] = e(I, WORD_DIM)
        # gradients of the small parameters
        # every small gradient is a view of ONE flat buffer (no packing before the data-parallel all-reduce); it starts
        # with the projection bias gradients, contiguous so that one column-sum pass writes all of them
        small = {n: p for n, p in self._params().items()
                 if not n.startswith("embedding_user.w") and not n.startswith("embedding_item.w")}
        nb = D * len(self.mods)
        ws["g_flat"] = torch.zeros(nb + sum(p.numel() for n, p in small.items() if not n.endswith("_dense.bias") or
                                            n.startswith("s_dense")), dtype=torch.float32, device=dev)
        ws["g_proj_bias"] = ws["g_flat"][:nb]
        ws["g"], o = {}, nb
        for j, m in enumerate(self.mods):
            ws["g"][f"{m}_dense.bias"] = ws["g_proj_bias"][D * j:D * (j + 1)]
        for n, p in small.items():
            if n not in ws["g"]:
                ws["g"][n] = ws["g_flat"][o:o + p.numel()].view(p.shape)
                o += p.numel()
        assert o == ws["g_flat"].numel()
        # split-K plan + workspaces
        dmax = max(self._feat[m].shape[1] for m in self.mods)
        ws["split_proj"] = max(1, min(256, (I + 1023) // 1024))
        ws["dmax"] = dmax
        ws["W_tf32"] = {m: e(D, self._feat[m].shape[1]) for m in self.mods}
        ws["wgrad_ws_m"] = {m: e(max(1, ops.linear_tf32_wgrad_ws_floats(I, self._feat[m].shape[1]))) for m in self.mods}
        ws["gemm_ws_m"] = {m: (e(ws["split_proj"] * self._feat[m].shape[1] * D) if self.proj_precision == "fp32"
                               or self._feat[m].shape[1] % 4 else None) for m in self.mods}
        ws["colsum_ws"] = e(ops.colsum_ws_floats(I, D * len(self.mods)))

    def _workspace_batch(self, ws, B):
        """the buffers whose size follows the batch: everything indexed by the 3B instance rows"""
        dev, U, I = self.device_, self.num_users, self.num_items
        Fw, nt = ws["F"], ws["nt"]
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        ws["B"] = B
        ws["density"] = {"u": min(100, 100 * B // max(U, 1) + 1), "i": min(100, 200 * B // max(I, 1) + 1)}   # % rows marked
        ws["terms"] = e(nt * B)
        ws["inst_rows"] = torch.empty(3 * B, dtype=torch.int32, device=dev)
        ws["inst_grad"] = e(3 * B, D * nt)
        ws["O_inst"] = e(3 * B, Fw)
        ws["dO_inst"] = e(3 * B, Fw)
        ws["split_inst"] = max(1, min(64, (3 * B + 127) // 128))
        ws["gemm_ws"] = e(max(ws["split_proj"] * ws["dmax"] * D, ws["split_inst"] * Fw * D))
        ws["inst_ws"] = e(ops.inst_backward_ws_floats(B, nt, Fw))
        ws["inst_dummy"] = torch.empty(3 * B, dtype=torch.int32, device=dev)
        ws["F_c"], ws["S_c"] = e(3 * B, D), [e(3 * B, D) for _ in self.mods]
        ar = torch.arange(B, dtype=torch.int64, device=dev)
        ws["c_users"], ws["c_pos"], ws["c_neg"] = ar.clone(), ar.clone(), ar + B
        if self.linear:
            self._lin_workspace_batch(ws, B)

    # ------------------------------------------------------------------------------------------
    # forward  (compute + gcn_cf + bpr losses; EliMRec.py:228-272,144-153,115-142)
    # ------------------------------------------------------------------------------------------
    def _forward(self, users, pos, neg):
        if self.linear:
            return self._lin_forward(users, pos, neg)
        P = self._params()
        U, I, L = self.num_users, self.num_items, self.n_layers
        B = int(users.numel())
        ws = self._workspace(B)
        G, Fw = ws["G"], ws["F"]
        g = self.graph
        if self._generic:
            self._forward_generic(P, ws, users, pos, neg)
            return self._loss(P, ws, users, pos, neg)
        Eu = P["embedding_user.weight"].detach()
        Ei = P["embedding_item.weight"].detach()
        X0_i = ws["X0_i"]
        # propagation: layer k has a WIDE side (distinct per graph) and a NARROW side (shared, 64 wide); the two
        # launches of a layer are independent and run on two streams
        O = ws["O"]
        Ou, Oi = O[:U], O[U:]
        prev_u = [(ws["X0_u"], D)]       # layers seen by user rows, in order (E_u as copied by this forward)
        prev_i = [(X0_i, Fw)]    # layers seen by item rows
        inv = 1.0 / (L + 1)
        lazy = self.lazy_tables
        mask, need2 = ws["mask"], ws["need2"]
        mask_of = {"u": mask[:U], "i": mask[U:]}
        need2_of = {"u": need2[:U], "i": need2[U:]}
        sL_w = "u" if L % 2 == 1 else "i"          # wide side of the last layer; layer L-1 is wide on the other side
        sL_o = "i" if sL_w == "u" else "u"
        self._prep_weights(P, ws)
        # The projections (HBM-bound, ~300 MB of features) must reach the SMs BEFORE the narrow layer-1 SpMM (L2-bound, fills
        # every register file if it arrives first): they start right after the weight prep, while the side stream first does
        # the small copies and only then the SpMM - the two then overlap instead of queueing.
        side = ops.fork_side()
        with torch.cuda.stream(side):
            if lazy:
                # the loss reads the layer-mean output at the instance rows only: mark them, and keep E_u of THIS forward
                # (Adam overwrites the parameter) for tables materialised later
                # Two hops: the last layer's wide SpMM reads layer L-1 only at the neighbours of its instance rows, the
                # mean epilogue at the instance rows themselves -> need2 = those rows of the other side.
                ops.inst_rows(users, pos, neg, U, ws["inst_rows"], mask, need2)
                if L >= 2:
                    ops.mark_neighbors(g.ui if sL_w == "u" else g.iu, mask_of[sL_w], need2_of[sL_o])
                # off the critical path, on a stream of their own: the weights this forward uses (for tables completed
                # later) and the zeroed instance rows of the backward's seed slabs
                aux = ops.fork_side(7)
                with torch.cuda.stream(aux):
                    self._snapshot(P, ws)
                    self._zero_seed_rows(ws)
            ops.copy_2d(Eu, ws["X0_u"], U, D)
            ops.copy_2d(Ei, X0_i, I, D)          # layer 0, item side: [E_i | P_v | P_a | P_t]
            ev_copy = torch.cuda.Event()
            ev_copy.record(side)
        self._proj_forward(P, ws, X0_i, 0, I)    # projections write straight into the slab
        torch.cuda.current_stream().wait_event(ev_copy)
        wide_in, narrow_in = X0_i, ws["X0_u"]
        for k in range(1, L + 1):
            users_wide = (k % 2 == 1)
            half_w, half_n = (g.ui, g.iu) if users_wide else (g.iu, g.ui)
            last = (k == L)
            if k > 1:
                side = ops.fork_side()
            if not last:
                Yw, Yn = ws["XW"][k], ws["XN"][k]
                with torch.cuda.stream(side):
                    ops.spmm(half_n, narrow_in, Yn, D)
                sparse2 = lazy and k == L - 1          # only the rows the last layer will read
                ops.spmm(half_w, wide_in, Yw, Fw, row_mask=need2_of[sL_o] if sparse2 else None)
                if sparse2:
                    ws["pre_last_layer"] = (half_w, wide_in, Yw)
                if users_wide:
                    prev_u.append((Yw, Fw)); prev_i.append((Yn, D))
                else:
                    prev_i.append((Yw, Fw)); prev_u.append((Yn, D))
                wide_in, narrow_in = Yw, Yn
            else:
                out_w, out_n = (Ou, Oi) if users_wide else (Oi, Ou)
                pw, pn = (prev_u, prev_i) if users_wide else (prev_i, prev_u)
                if L == 1:
                    ops.join_side(side)      # the narrow epilogue reads layer 0 of its side (X0_i when L is odd)
                    side = ops.fork_side()
                s_w, s_n = ("u", "i") if users_wide else ("i", "u")
                with torch.cuda.stream(side):
                    ops.spmm(half_n, narrow_in, None, D, ops.mean_epilogue(pn, out_n, Fw, inv),
                             row_mask=mask_of[s_n] if lazy else None, density=ws["density"][s_n])
                ops.spmm(half_w, wide_in, None, Fw, ops.mean_epilogue(pw, out_w, Fw, inv),
                         row_mask=mask_of[s_w] if lazy else None, density=ws["density"][s_w])
                if lazy:   # what completes the last layer on demand (every row, same inputs)
                    ws["last_layer"] = (half_n, narrow_in, pn, out_n, half_w, wide_in, pw, out_w, inv)
            ops.join_side(side)
        if lazy:
            ops.join_side(aux)
        self._tables_version = getattr(self, "_tables_version", 0) + 1
        return self._loss(P, ws, users, pos, neg)

    def _loss(self, P, ws, users, pos, neg, gathered=False):
        """fusion Linear + heads + the 1+M BPR losses on the layer-mean slab (EliMRec.py:261-272,144-153,115-142);
        ``gathered``: O[instance rows] is already in ws["O_inst"] (linear schedule)"""
        U, B, Fw, O = self.num_users, ws["B"], ws["F"], ws["O"]
        if self.kwai:
            self.modality = "v"  # EliMRec.py:133-134
        weights = self._loss_weights()
        if not self.lazy_tables:
            # fusion Linear (concat) and single-modal heads over ALL rows, as the reference does every step
            self._dense_tables(P, ws)
            ops.bpr([ws["F_all"]] + ws["S"], weights, users, pos, neg, U, ws["loss"], ws["inst_rows"], ws["inst_grad"],
                    ws["terms"])
            ops.gather_rows(ws["inst_rows"], O, ws["O_inst"], Fw)
        else:
            # only the sampled rows: gather O[inst], fusion + heads on 3B rows, BPR on the compact tables
            rows = ws["inst_rows"]
            if not gathered:
                ops.gather_rows(rows, O, ws["O_inst"], Fw)
            if self.inst_fuse == "ffma":
                # exact fp32 on 3B/16 CTAs (measured slower than the tensor-core form: instruction-issue bound)
                Wu, Wi = self._fusion_weights(P, ws)
                ops.inst_forward(B, 1 + len(self.mods), Fw, ws["O_inst"], Wu, Wi,
                                 [P[f"s_dense_{m}.weight"].detach() for m in self.mods],
                                 P["embedding_user_after_GCN.bias"].detach(), P["embedding_item_after_GCN.bias"].detach(),
                                 [P[f"s_dense_{m}.bias"].detach() for m in self.mods], ws["F_c"], ws["S_c"])
            else:
                su_ = ops.fork_side(6)
                with torch.cuda.stream(su_):
                    self._fuse_heads_rows(ws, ws["O_inst"][:B], ws["F_c"][:B], [s_[:B] for s_ in ws["S_c"]], "u")
                self._fuse_heads_rows(ws, ws["O_inst"][B:], ws["F_c"][B:], [s_[B:] for s_ in ws["S_c"]], "i")
                ops.join_side(su_)
            bpr_args = ([ws["F_c"]] + ws["S_c"], weights, ws["c_users"], ws["c_pos"], ws["c_neg"], B, ws["loss"], ws["inst_dummy"],
                        ws["inst_grad"], ws["terms"])
            if getattr(self, "_defer_loss", False):
                # train_step: nothing in the backward reads the loss scalar - its reduction leaves the critical path (joined
                # before train_step returns)
                ops.bpr(*bpr_args, part=1)
                sl = ops.fork_side(9)
                with torch.cuda.stream(sl):
                    ops.bpr(*bpr_args, part=2)
                self._loss_side = sl
            else:
                ops.bpr(*bpr_args)
            self._tables_pending = True
        return ws["loss"][0]

    def _loss_weights(self):
        """weight of the fused BPR term and of each single-modal term in the loss (EliMRec.py:125-142)"""
        if self.predict_type == "normal":
            return [1.0] + [0.0] * len(self.mods)
        alpha = float(self.config.alpha)
        return [1.0] + [alpha * self.modality.count(m) for m in self.mods]

    def _dead_params(self):
        """Heads whose loss term has weight 0 (predict_type='normal', modality ablations): the reference never calls them,
        their .grad stays None and Adam skips them - so no gradient is handed out for them here either."""
        dead = set()
        for m, w in zip(self.mods, self._loss_weights()[1:]):
            if w == 0:
                dead |= {f"s_dense_{m}.weight", f"s_dense_{m}.bias"}
        return dead

    def _fusion_weights(self, P, ws):
        """[64 x F] weights of the fusion Linear as the concat kernels see them: the parameters themselves, or for
        mm_fusion_mode='mean' the tied form [W/G | ... | W/G] refreshed by _prep_weights (EliMRec.py:224-225)."""
        if self.mm_fusion_mode == "mean":
            return ws["W_eff"]["u"], ws["W_eff"]["i"]
        return P["embedding_user_after_GCN.weight"].detach(), P["embedding_item_after_GCN.weight"].detach()

    def _snapshot(self, P, ws):
        """keep the fusion / head weights and biases THIS forward used (lazily built tables must use them)"""
        Wu, Wi = self._fusion_weights(P, ws)
        eff = {"embedding_user_after_GCN.weight": Wu, "embedding_item_after_GCN.weight": Wi}
        torch._foreach_copy_(ws["snap_dst"], [eff[n] if n in eff else P[n].detach() for n in ws["snap_names"]])

    def _fuse_heads_rows(self, ws, O_rows, Fout, Sout, who):
        """fusion Linear + heads on a block of rows, with the weights snapshotted / split by this forward"""
        sn = ws["snap"]
        name = "user" if who == "u" else "item"
        bf = sn[f"embedding_{name}_after_GCN.bias"]
        n = O_rows.shape[0]
        Fw = ws["F"]
        if n == 0:
            return
        if self.fuse_precision == "x3":
            sp = ws["W_split"]
            ops.fuse_heads_x3(O_rows, sp[who][0], sp[who][1], bf, [sp[m][0] for m in self.mods], [sp[m][1] for m in self.mods],
                              [sn[f"s_dense_{m}.bias"] for m in self.mods], Fout, Sout)
        else:
            ops.gemm(n, D, Fw, O_rows, Fw, 1, sn[f"embedding_{name}_after_GCN.weight"], 1, Fw, Fout, D, 1, bias=bf, tag="fuse_fwd")
            for j, m in enumerate(self.mods):
                ops.gemm(n, D, D, O_rows[:, D * (j + 1):], Fw, 1, sn[f"s_dense_{m}.weight"], 1, D, Sout[j], D, 1,
                         bias=sn[f"s_dense_{m}.bias"], tag="head_fwd")

    def _dense_tables(self, P, ws, from_snapshot=False):
        """all_users / all_items / all_s_embs over every row of the layer-mean slab"""
        U, I = self.num_users, self.num_items
        O, F_all = ws["O"], ws["F_all"]
        if not from_snapshot:
            self._snapshot(P, ws)
        if self.fuse_precision == "x3":
            sp, sn = ws["W_split"], ws["snap"]
            ops.fuse_heads_x3_all(U, I, O, sp["u"], sn["embedding_user_after_GCN.bias"], sp["i"],
                                  sn["embedding_item_after_GCN.bias"], [sp[m][0] for m in self.mods],
                                  [sp[m][1] for m in self.mods], [sn[f"s_dense_{m}.bias"] for m in self.mods], F_all, ws["S"])
        else:
            self._fuse_heads_rows(ws, O[:U], F_all[:U], [s_[:U] for s_ in ws["S"]], "u")
            self._fuse_heads_rows(ws, O[U:], F_all[U:], [s_[U:] for s_ in ws["S"]], "i")
        self._all_users, self._all_items = F_all[:U], F_all[U:]
        self._all_s_embs = {}
        for j, m in enumerate(self.mods):
            self._all_s_embs[f"pre_fusion_user_{m}"] = ws["S"][j][:U]
            self._all_s_embs[f"pre_fusion_item_{m}"] = ws["S"][j][U:]
        self._tables_pending = False

    def _materialize_tables(self):
        if self._tables_pending:
            with torch.no_grad():
                if self.linear:
                    return self._lin_materialize()
                ws = self._ws
                Fw = ws["F"]
                # the training step produced the last layer at its instance rows only: complete it (every row, from the
                # layer inputs of that same forward), then the fusion Linear + heads with that forward's weights
                if self.n_layers >= 2:
                    half_p, in_p, out_p = ws["pre_last_layer"]
                    ops.spmm(half_p, in_p, out_p, Fw)
                half_n, narrow_in, pn, out_n, half_w, wide_in, pw, out_w, inv = ws["last_layer"]
                ops.spmm(half_n, narrow_in, None, D, ops.mean_epilogue(pn, out_n, Fw, inv))
                ops.spmm(half_w, wide_in, None, Fw, ops.mean_epilogue(pw, out_w, Fw, inv))
                self._dense_tables(None, ws, from_snapshot=True)

    # ------------------------------------------------------------------------------------------
    # backward: instance rows -> fusion/head weights -> 2L SpMMs -> projection weights
    # ------------------------------------------------------------------------------------------
    def _backward(self, gscale=None, split=False, fuse_adam=False):
        """Returns {param name: gradient view}.  ``gscale``: 1-element device tensor (upstream grad) or None.
        ``split``: stop once the two embedding-table gradients are final and return only those; the weight gradients
        (fusion / heads / projections / word table) are then produced by ``_backward_weights()`` - data-parallel
        replicas put the table gradients' all-reduce on the wire in between (``make_graphed_step``)."""
        if self.linear:
            return self._lin_backward(gscale, split, fuse_adam)
        P = self._params()
        ws = self._ws
        B, G, Fw, nt = ws["B"], ws["G"], ws["F"], ws["nt"]
        ig, gr = ws["inst_grad"], ws["g"]
        Oin, dOin = ws["O_inst"], ws["dO_inst"]
        if gscale is not None:
            gscale = gscale.reshape(1)
        Wu, Wi = self._fusion_weights(P, ws)
        tied = self.mm_fusion_mode == "mean"
        gWu = ws["g_eff"]["u"] if tied else gr["embedding_user_after_GCN.weight"]
        gWi = ws["g_eff"]["i"] if tied else gr["embedding_item_after_GCN.weight"]
        # fusion Linear + heads, backward on the instance rows: dO[inst], all weight and bias gradients (3 launches)
        # part 1 (d O[inst]) seeds the propagation backward; part 2 (weight / bias gradients) only feeds Adam -> side stream
        ib = lambda part: ops.inst_backward(
            B, nt, Fw, ig, Oin, gscale, Wu, Wi, [P[f"s_dense_{m}.weight"].detach() for m in self.mods], dOin,
            gWu, gWi, gr["embedding_user_after_GCN.bias"], gr["embedding_item_after_GCN.bias"],
            [gr[f"s_dense_{m}.weight"] for m in self.mods], [gr[f"s_dense_{m}.bias"] for m in self.mods], ws["inst_ws"], part=part)
        def inst_weights():
            ib(2)
            if tied:    # d W = (1/G) * sum of the G column blocks of the tied weight's gradient
                ops.fold_blocks(gWu, gr["embedding_user_after_GCN.weight"], G, 1.0 / G)
                ops.fold_blocks(gWi, gr["embedding_item_after_GCN.weight"], G, 1.0 / G)

        ib(1)
        side_w = None
        if not split:
            side_w = ops.fork_side(5)     # forked after d O[inst]: the weight gradients must not delay it
            with torch.cuda.stream(side_w):
                inst_weights()
        if self._generic:
            dE_u, dX0_i = self._backward_generic(ws)
        else:
            dE_u, dX0_i = self._backward_prop(ws)
        grads = {"embedding_user.weight": dE_u, "embedding_item.weight": dX0_i[:, :D]}
        ws["bw_pending"] = (inst_weights, dX0_i, side_w)
        if split:
            return grads
        grads.update(self._backward_weights())
        return grads

    def _backward_weights(self):
        """second half of the backward: every gradient that is not an embedding table's"""
        if self.linear:
            return self._lin_backward_weights()
        ws = self._ws
        inst_weights, dX0_i, side_w = ws.pop("bw_pending")
        if side_w is None:      # split backward: the instance-row weight gradients run beside the projection ones
            side_w = ops.fork_side(5)
            with torch.cuda.stream(side_w):
                inst_weights()
        self._proj_wgrad(ws, dX0_i, 0, self.num_items)
        self._word_wgrad(ws, dX0_i)
        ops.join_side(side_w)
        dead = self._dead_params()
        return {n: g for n, g in ws["g"].items() if n not in dead} if dead else ws["g"]

    def _zero_seed_rows(self, ws):
        """zero the instance rows of d x_L's two slabs (the rest of those slabs is never read: column masks)"""
        U, I, L = self.num_users, self.num_items, self.n_layers
        N, Fw, rows = U + I, ws["F"], ws["inst_rows"]
        if "Gw" in ws:      # fused layer-mean gradient: the G slabs double as the seeds
            ops.zero_rows(rows, 0, N, 0, ws["Gw"], Fw)
            ops.zero_rows(rows, 0, N, 0, ws["Gn"], D)
        else:
            s_w = "u" if L % 2 == 1 else "i"
            for dst, sd, w in ((ws["dW"][0], s_w, Fw), (ws["dN"][0], "i" if s_w == "u" else "u", D)):
                a, b, off = (0, U, 0) if sd == "u" else (U, N, U)
                ops.zero_rows(rows, a, b, off, dst, w)
        ws["seed_zeroed"] = True

    def _backward_prop(self, ws):
        """propagation backward of the bipartite schedule: returns (d E_u, d x_0[item rows] = [dE_i | dP_v | dP_a | dP_t])"""
        U, I, L = self.num_users, self.num_items, self.n_layers
        N, Fw = U + I, ws["F"]
        g = self.graph
        rows, dOin = ws["inst_rows"], ws["dO_inst"]
        # layer-mean gradient G = dO / (L+1), row-sparse; it enters every layer of the chain
        inv = 1.0 / (L + 1)
        lo = {"u": (0, U, 0), "i": (U, N, U)}
        nrows = {"u": U, "i": I}

        def add_G(dst, side, wide):
            a, b, off = lo[side]
            ops.scatter_add_rows(rows, a, b, off, dOin, Fw, dst, Fw if wide else D, inv)

        s_w = "u" if L % 2 == 1 else "i"      # wide side of the last layer
        s_n = "i" if s_w == "u" else "u"
        dWc, dNc = ws["dW"][0][:nrows[s_w]], ws["dN"][0][:nrows[s_n]]
        lazy = self.lazy_tables
        mask_of = {"u": ws["mask"][:U], "i": ws["mask"][U:]}
        need2_of = {"u": ws["need2"][:U], "i": ws["need2"][U:]}
        fused = lazy and "Gw" in ws
        if lazy:     # d x_L is non-zero at the instance rows only: zero just those, the first SpMMs skip all other columns
            if not ws.pop("seed_zeroed", False):     # normally done by the forward, off the critical path
                self._zero_seed_rows(ws)
        else:
            dWc.zero_(); dNc.zero_()
        if fused:
            # G (wide) and its fold over the graph blocks (narrow), once, for both sides; every SpMM of the chain adds its
            # slice in the epilogue.  d x_L = G itself: the first SpMMs read the G slabs through their column masks.
            Gw, Gn = ws["Gw"], ws["Gn"]
            ops.scatter_add_rows(rows, 0, N, 0, dOin, Fw, Gw, Fw, inv)
            ops.scatter_add_rows(rows, 0, N, 0, dOin, Fw, Gn, D, inv)
            G_of = {"u": (Gw[:U], Gn[:U]), "i": (Gw[U:], Gn[U:])}
            dWc, dNc = G_of[s_w][0], G_of[s_n][1]
            add_G = lambda dst, side, wide: None
        else:
            add_G(dWc, s_w, True)
            add_G(dNc, s_n, False)
        gw = lambda sd: dict(addend=G_of[sd][0], add_mask=mask_of[sd]) if fused else {}
        gn = lambda sd: dict(addend=G_of[sd][1], add_mask=mask_of[sd]) if fused else {}
        flip = 1
        for k in range(L, 0, -1):
            s = "u" if k % 2 == 1 else "i"    # wide side of layer k
            o = "i" if s == "u" else "u"
            half_o, half_s = (g.iu_t, g.ui_t) if s == "u" else (g.ui_t, g.iu_t)     # blocks of A_hat^T
            nW, nN = ws["dW"][flip][:nrows[o]], ws["dN"][flip][:nrows[s]]
            side = ops.fork_side()
            sparse_in = lazy and k == L
            with torch.cuda.stream(side):
                # d x_{k-1}[s, narrow] = A[s,o] @ d x_k[o, narrow]
                ops.spmm(half_s, dNc, nN, D, col_mask=mask_of[o] if sparse_in else None, **gn(s))
                add_G(nN, s, False)
            # d x_{k-1}[o, wide]   = A[o,s] @ d x_k[s, wide].  Row-sparse step: d x_L lives on the instance rows, so
            # d x_{L-1}[o, wide] is non-zero only on need2 (their neighbours + the instance rows of that side, which get G):
            # layer L writes just those rows and layer L-1 reads just those columns.
            if sparse_in:
                ops.spmm(half_o, dWc, nW, Fw, col_mask=mask_of[s], row_mask=need2_of[o] if L >= 2 else None, **gw(o))
            elif lazy and k == L - 1:
                ops.spmm(half_o, dWc, nW, Fw, col_mask=need2_of[s], **gw(o))
            else:
                ops.spmm(half_o, dWc, nW, Fw, **gw(o))
            add_G(nW, o, True)
            ops.join_side(side)
            dWc, dNc, flip = nW, nN, flip ^ 1
        # now dWc = d x_0[item rows, wide] = [dE_i | dP_v | dP_a | dP_t], dNc = d x_0[user rows] = dE_u
        return dNc, dWc

    def _word_wgrad(self, ws, dX0_i):
        """literal 'tiktok': d t_feat = d P_t @ W_t, then d word_embedding = M^T @ d t_feat (the gradient the reference
        keeps sending into word_embedding through the t_feat it built at construction, EliMRec.py:376-378)."""
        if not (self.tiktok and self.word_grad):
            return
        Fw, I = ws["F"], self.num_items
        Wt = self.t_dense.weight.detach()
        c0 = D * (1 + self.mods.index("t"))
        ops.gemm(I, WORD_DIM, D, dX0_i, Fw, 1, Wt, WORD_DIM, 1, ws["dT"], WORD_DIM, 1, a_off=c0, tag="word_dgrad")
        ops.spmm(self._word_half_t, ws["dT"], ws["g"]["word_embedding.weight"], WORD_DIM)

    # ------------------------------------------------------------------------------------------
    # generic schedule: A_hat with a diagonal (adj_type 'norm' / 'mean', EliMRec.py:332-334,349-352)
    #   x_k = A_bip x_{k-1} + s * x_{k-1} on every row and every graph block; no dedup, no row sparsity
    # ------------------------------------------------------------------------------------------
    def _forward_generic(self, P, ws, users, pos, neg):
        U, I, L = self.num_users, self.num_items, self.n_layers
        N, G, Fw = U + I, ws["G"], ws["F"]
        g = self.graph
        X = ws["GX"]
        self._prep_weights(P, ws)
        side = ops.fork_side()
        with torch.cuda.stream(side):
            ops.broadcast_cols(P["embedding_user.weight"].detach(), X[0][:U], U, G)   # the same E_u feeds every graph
            ops.copy_2d(P["embedding_item.weight"].detach(), X[0][U:], I, D)
        self._proj_forward(P, ws, ws["X0_i"], 0, I)
        ops.join_side(side)
        for k in range(1, L + 1):
            side = ops.fork_side()
            with torch.cuda.stream(side):
                ops.spmm(g.iu, X[k - 1][:U], X[k][U:], Fw)
            ops.spmm(g.ui, X[k - 1][U:], X[k][:U], Fw)
            ops.join_side(side)
            ops.axpy_rows(g.self_all, X[k - 1], X[k], Fw)
        ops.layer_mean(X, ws["O"], Fw, 1.0 / (L + 1))
        self._tables_version = getattr(self, "_tables_version", 0) + 1

    def _backward_generic(self, ws):
        U, I, L = self.num_users, self.num_items, self.n_layers
        N, G, Fw = U + I, ws["G"], ws["F"]
        g = self.graph
        rows, dOin = ws["inst_rows"], ws["dO_inst"]
        inv = 1.0 / (L + 1)
        add_G = lambda dst: ops.scatter_add_rows(rows, 0, N, 0, dOin, Fw, dst, Fw, inv)
        cur, flip = ws["GdX"][0], 1
        cur.zero_()
        add_G(cur)                                   # d x_L = G (layer-mean gradient, on the instance rows)
        for k in range(L, 0, -1):                    # d x_{k-1} = A_hat^T d x_k + G
            nxt = ws["GdX"][flip]
            side = ops.fork_side()
            with torch.cuda.stream(side):
                ops.spmm(g.iu_t, cur[:U], nxt[U:], Fw)
            ops.spmm(g.ui_t, cur[U:], nxt[:U], Fw)
            ops.join_side(side)
            ops.axpy_rows(g.self_all, cur, nxt, Fw)
            add_G(nxt)
            cur, flip = nxt, flip ^ 1
        ops.fold_blocks(cur[:U], ws["dE_u"], G)      # the user table fed all G graphs
        return ws["dE_u"], cur[U:]

    # projections over item rows [r0, r1) - the whole table on one GPU, the owned block when row-sharded
    def _prep_weights(self, P, ws, proj=True, extra=()):
        """TF32 rounding (projections) and hi/lo split (fusion, heads) of the small weights, one launch."""
        prep = list(extra)
        if self.mm_fusion_mode == "mean":
            G = ws["G"]
            ops.tie_blocks(P["embedding_user_after_GCN.weight"].detach(), ws["W_eff"]["u"], G, 1.0 / G)
            ops.tie_blocks(P["embedding_item_after_GCN.weight"].detach(), ws["W_eff"]["i"], G, 1.0 / G)
        if proj and self.proj_precision == "tf32":
            prep += [(P[f"{m}_dense.weight"].detach(), ws["W_tf32"][m], None) for m in self.mods
                     if self._feat[m].shape[1] % 4 == 0]
        if self.fuse_precision == "x3":
            sp = ws["W_split"]
            Wu, Wi = self._fusion_weights(P, ws)
            prep += [(Wu, *sp["u"]), (Wi, *sp["i"])]
            prep += [(P[f"s_dense_{m}.weight"].detach(), *sp[m]) for m in self.mods]
        if prep:
            ops.prep_weights_tf32(prep)

    def _proj_forward(self, P, ws, X0_i, r0, r1):
        Fw = ws["F"]
        tc, sides = [], []
        for j, m in enumerate(self.mods):
            Wm, bm = P[f"{m}_dense.weight"].detach(), P[f"{m}_dense.bias"].detach()
            Dm = Wm.shape[1]
            if self.proj_precision == "tf32" and Dm % 4 == 0:
                tc.append((self._feat_tc(m)[r0:r1], ws["W_tf32"][m], bm, X0_i[r0:r1], D * (j + 1)))
            else:   # exact-fp32 FFMA path, one stream per modality (disjoint output columns)
                st = ops.fork_side(2 + j)
                with torch.cuda.stream(st):
                    ops.gemm(r1 - r0, D, Dm, self._feat[m], Dm, 1, Wm, 1, Dm, X0_i, Fw, 1, bias=bm, a_off=r0 * Dm,
                             c_off=r0 * Fw + D * (j + 1), tag="proj_fwd")
                sides.append(st)
        if tc:      # every tensor-core projection of the step in ONE persistent launch (one CTA per SM)
            ops.linear_tf32_fwd_multi(tc)
        for st in sides:
            ops.join_side(st)

    def _proj_wgrad(self, ws, dX0_i, r0, r1):
        """dW_m, db_m from rows [r0, r1) of d x_0[item rows] = [dE_i | dP_v | dP_a | dP_t]."""
        Fw, gr = ws["F"], ws["g"]
        sides = []
        for j, m in enumerate(self.mods):   # one stream (and one scratch buffer) per modality
            Xm = self._feat[m]
            Dm = Xm.shape[1]
            c0 = D * (j + 1)
            st = ops.fork_side(2 + j)
            with torch.cuda.stream(st):
                if self.proj_precision == "tf32" and Dm % 4 == 0:
                    ops.linear_tf32_wgrad(dX0_i[r0:r1], self._feat_tc(m)[r0:r1], gr[f"{m}_dense.weight"], ws["wgrad_ws_m"][m],
                                          col=c0)
                else:
                    ops.gemm(Dm, D, r1 - r0, Xm, 1, Dm, dX0_i, Fw, 1, gr[f"{m}_dense.weight"], 1, Dm, split_k=ws["split_proj"],
                             ws=ws["gemm_ws_m"][m], a_off=r0 * Dm, b_off=r0 * Fw + c0, tag="proj_wgrad")
            sides.append(st)
        ops.colsum(r1 - r0, D * len(self.mods), dX0_i, Fw, ws["g_proj_bias"], ws["colsum_ws"], a_off=r0 * Fw + D)
        for st in sides:
            ops.join_side(st)

    # ------------------------------------------------------------------------------------------
    # public training API
    # ------------------------------------------------------------------------------------------
    def _triples(self, users, pos, neg):
        dev = self.device_
        cv = lambda x: (x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))).to(dev, non_blocking=True).long().contiguous()
        return cv(users), cv(pos), cv(neg)

    def bpr_loss(self, users, pos_items, neg_items):
        """models/EliMRec.py:115-142.  Returns a scalar tensor whose backward fills ``.grad``."""
        users, pos, neg = self._triples(users, pos_items, neg_items)
        P = self._params()
        if next(iter(P.values())).device != self.device_:
            raise ElimrecError("model parameters are not on config.device; call .to(config.device)")
        return _StepFunction.apply(self, users, pos, neg, *[P[n] for n in self._param_names])

    @torch.no_grad()
    def getEmbedding(self, users, pos_items, neg_items):
        """models/EliMRec.py:274-289: runs the propagation (``compute()``), caches ``all_users`` / ``all_items`` and returns
        their rows at the batch plus the three raw ("ego") embedding lookups.  No autograd history: training goes through
        ``bpr_loss`` / ``train_step``, whose fused kernel does these gathers itself (elimrec_bpr_forward_backward)."""
        users, pos, neg = self._triples(users, pos_items, neg_items)
        self._forward(users, pos, neg)
        au, ai = self.all_users, self.all_items           # completed on demand (every row), as the reference caches them
        Eu, Ei = self.embedding_user.weight.detach(), self.embedding_item.weight.detach()
        return au[users], ai[pos], ai[neg], Eu[users], Ei[pos], Ei[neg]

    def make_optimizer(self, lr=None, weight_decay=None, betas=(0.9, 0.999), eps=1e-8):
        from .optim import FusedAdam
        self._adam = FusedAdam(self, lr if lr is not None else self.config.lr,
                               weight_decay if weight_decay is not None else self.config.weight_decay, betas, eps)
        return self._adam

    def train_step(self, users, pos_items, neg_items):
        """main.py:94-102 in one call: forward, backward, Adam.  Returns the loss (device scalar)."""
        if self._adam is None:
            self.make_optimizer()
        users, pos, neg = self._triples(users, pos_items, neg_items)
        with torch.no_grad():
            # linear schedule: the step counter ticks on a side stream of the forward, off the critical path
            early = self._tick_early = bool(self.linear)
            # ... and, on one GPU, Adam on the two embedding tables is the epilogue of the last backward hop (fused_adam=False
            # keeps the separate optimizer pass; data-parallel replicas must average the gradients first)
            fuse = self._fuse_adam_now = bool(self.linear and not getattr(self, "_dp", False) and _cfg(self.config, "fused_adam", True))
            self._defer_loss = bool(self.linear and self.lazy_tables and not getattr(self, "_dp", False))
            self._loss_side = None
            try:
                loss = self._forward(users, pos, neg)
            finally:
                self._tick_early = self._fuse_adam_now = self._defer_loss = False
            if fuse:      # moments of the tables must exist before the backward hands them to the kernel
                P = self._params()
                for n in ("embedding_user.weight", "embedding_item.weight"):
                    self._adam._st(n, P[n])
            grads = self._backward(None, fuse_adam=fuse)
            if getattr(self, "_dp", False):
                grads = self._allreduce_grads(grads)
            self._adam.apply(grads, tick=not early)
            if self._loss_side is not None:
                ops.join_side(self._loss_side)
                self._loss_side = None
        return loss

    # -- data-parallel replicas: each rank draws its own triples, gradients are averaged (NCCL) ----
    def enable_data_parallel(self):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise ElimrecError("enable_data_parallel needs an initialised torch.distributed process group")
        self._dp = dist.get_world_size() > 1
        for p in self.parameters():  # identical starting point on every rank
            dist.broadcast(p.data, src=0)

    def _allreduce_grads(self, grads):
        """One flat bucket (embedding tables + every small gradient), one all-reduce(AVG)."""
        from .dist import GradBucket
        ws = self._ws
        ws = ws["cache"]
        if "bucket" not in ws:
            P = self._params()
            head = self._param_names[:2]           # the two embedding tables: packed (dE_i is a strided slab view)
            ws["bucket"] = GradBucket({n: tuple(P[n].shape) for n in head}, self.device_, tail_flat=self._ws["g_flat"],
                                      tail_views={n: self._ws["g"][n] for n in self._param_names[2:]})
        ws["bucket"].pack(grads, ws["bucket"].head_names)
        views = ws["bucket"].all_reduce_mean()
        return {n: views[n] for n in grads}      # heads the loss does not use stay without a gradient

    # -- whole step as one CUDA graph (launch-bound otherwise: ~60 small launches per step) ------------
    def _extra_state(self):
        """trainable state kept outside the nn.Module parameters / FusedAdam (sharded modes): tensors that building a graphed
        step must leave unchanged"""
        return []

    def make_graphed_step(self, batch_size=None, device_sampler=None, host_loss=False):
        """The whole training step as a CUDA graph.  ``device_sampler`` (a ``PairwiseSamplerV2``): the graph starts with the
        Philox batch sampler reading its position from the optimizer's device step counter, so ``runner()`` with no
        arguments draws a new batch and trains on it - sampling + forward + backward + Adam in one replay.
        ``host_loss``: the graph ends with the copy of the loss into pinned host memory and ``runner(...)`` returns a host
        scalar tensor AFTER synchronising the stream - for loops that read the loss every step anyway (main.py:102) this
        replaces the separate device-to-host copy + sync of ``loss.item()``.  ``host_loss="deferred"``: ``runner(...)`` launches
        step i and returns the loss of step i-1 (``None`` on the first call; ``runner.flush()`` returns the last one): the host
        then prepares and launches the next step while this one runs, instead of idling the GPU for the launch + sync latency
        of every step."""
        B = int(batch_size or self.config["batch_size"])
        if self._adam is None:
            self.make_optimizer()
        dev = self.device_
        s3 = torch.zeros(3, B, dtype=torch.int64, device=dev)      # one buffer: a packed host batch arrives in ONE copy
        su, sp_, sn = s3[0], s3[1], s3[2]
        s3_flat = s3.view(-1)
        loss_host = torch.zeros((), dtype=torch.float32).pin_memory() if host_loss else None
        deferred = host_loss == "deferred"
        done_ev = [torch.cuda.Event(), torch.cuda.Event()] if deferred else None
        if device_sampler is not None:
            device_sampler.ensure_device(dev)
        model = self
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        # warm-up outside capture (allocates workspace + Adam state, loads kernels).  It is a real step on a dummy batch, so
        # parameters, Adam moments and the device step counter are put back afterwards: building the runner changes nothing.
        ad = self._adam
        saved_p = {n: p.detach().clone() for n, p in self.named_parameters()}
        saved_x = [t.clone() for t in self._extra_state()]
        saved_st = {n: (m.clone(), v.clone()) for n, (m, v) in ad.state.items()}
        saved_step, saved_consts = ad.step_dev.clone(), ad.consts.clone()
        with torch.cuda.stream(side):
            self.train_step(su, sp_, sn)
        torch.cuda.current_stream().wait_stream(side)
        with torch.no_grad():
            for n, p in self.named_parameters():
                p.copy_(saved_p[n])
            for n, (m, v) in ad.state.items():
                if n in saved_st:
                    m.copy_(saved_st[n][0]); v.copy_(saved_st[n][1])
                else:
                    m.zero_(); v.zero_()
            ad.step_dev.copy_(saved_step); ad.consts.copy_(saved_consts)
            for t, old in zip(self._extra_state(), saved_x):
                t.copy_(old)
        del saved_p, saved_st, saved_x
        torch.cuda.synchronize()
        dp = getattr(self, "_dp", False)
        dead = self._dead_params()
        before = CALLS["launches"]

        def draw():
            if device_sampler is not None:
                device_sampler.sample_batch_device(self._adam.step_dev, su, sp_, sn)

        single = bool(_cfg(self.config, "dp_single_graph", False))
        if not dp:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                if self.linear and self._lin_prefork and bool(_cfg(self.config, "prefork_propagation", True)):
                    self._prop_stream = ops.fork_side(11)      # before the batch exists: the dense layers start at t = 0
                draw()
                loss = self.train_step(su, sp_, sn)
                if host_loss:
                    loss_host.copy_(loss.reshape(()), non_blocking=True)
            graphs = (graph,)
        elif single:
            # EXPERIMENTAL (dp_single_graph=True, not yet run on hardware - DESIGN.md section 8 item 1): the same schedule with
            # the two all-reduces captured INSIDE one graph, so the four graph launches (and the ~85 us they cost) become one.
            graph = torch.cuda.CUDAGraph()
            bucket = self._ws["cache"]["bucket"]
            with torch.cuda.graph(graph):
                with torch.no_grad():
                    draw()
                    loss = self._forward(su, sp_, sn)
                    bucket.pack(self._backward(None, split=True), bucket.head_names)
                    w1 = bucket.all_reduce_mean_part(0, async_op=True)
                    self._backward_weights()
                    w2 = bucket.all_reduce_mean_part(1, async_op=True)
                    w1.wait()
                    self._adam.apply({n: bucket.views[n] for n in bucket.head_names})
                    w2.wait()
                    self._adam.apply({n: bucket.views[n] for n in bucket.tail_names if n not in dead}, tick=False)
            graphs, dp = (graph,), False       # the runner just replays it
        else:
            # data-parallel replicas, four graphs around two NCCL all-reduces:
            #   A = forward + backward down to the embedding-table gradients (29 of the 30 MB), packed -> all-reduce #1 goes
            #       on the wire (NCCL's own stream) ...
            #   B = ... while the weight gradients (instance rows, projections: one more pass over the features) are
            #       computed straight into the bucket's tail -> all-reduce #2 (1 MB)
            #   C = Adam on the averaged tables (only needs all-reduce #1, and hides #2), D = Adam on the small tensors.
            # Collectives stay outside the captures.
            gA, gB, gC, gD = (torch.cuda.CUDAGraph() for _ in range(4))
            bucket = self._ws["cache"]["bucket"]
            with torch.cuda.graph(gA):
                with torch.no_grad():
                    draw()
                    loss = self._forward(su, sp_, sn)
                    bucket.pack(self._backward(None, split=True), bucket.head_names)
            with torch.cuda.graph(gB, pool=gA.pool()):
                with torch.no_grad():
                    self._backward_weights()        # writes straight into the bucket's tail (ws["g_flat"])
            with torch.cuda.graph(gC, pool=gA.pool()):
                with torch.no_grad():
                    self._adam.apply({n: bucket.views[n] for n in bucket.head_names})
            with torch.cuda.graph(gD, pool=gA.pool()):
                with torch.no_grad():
                    self._adam.apply({n: bucket.views[n] for n in bucket.tail_names if n not in dead}, tick=False)
            graphs = (gA, gB, gC, gD)
        n_launch = CALLS["launches"] - before + 2  # + the two memsets of the backward seeds

        class _Runner:
            launches_per_step = n_launch

            triples = (su, sp_, sn)          # the batch the last replay trained on
            _n_calls = 0

            def flush(self):
                """deferred host loss: wait for the last launched step and return its loss (None if there is none)"""
                if not deferred or self._n_calls == 0:
                    return None
                done_ev[(self._n_calls - 1) & 1].synchronize()
                return loss_host.clone()

            def __call__(self, users=None, pos=None, neg=None):
                if users is None:
                    if device_sampler is None:
                        raise ElimrecError("runner() without a batch needs make_graphed_step(device_sampler=...)")
                else:
                    if device_sampler is not None:
                        raise ElimrecError("this runner samples its own batches; call it without arguments")
                    if all(torch.is_tensor(x) and x.dtype == torch.int64 and x.numel() == B for x in (users, pos, neg)):
                        # (pinned) host or device int64 tensors: straight into the graph's static inputs, no temporaries -
                        # and in ONE copy when the three are adjacent slices of one buffer (PairwiseSamplerV2 packs its
                        # pinned epochs [batch][users | pos | neg] for exactly this)
                        pu = users.data_ptr()
                        if (pos.data_ptr() == pu + 8 * B and neg.data_ptr() == pu + 16 * B and users.is_contiguous()
                                and pos.is_contiguous() and neg.is_contiguous()
                                and users.untyped_storage().data_ptr() == neg.untyped_storage().data_ptr()):
                            s3_flat.copy_(users.as_strided((3 * B,), (1,)), non_blocking=True)
                        else:
                            su.copy_(users, non_blocking=True); sp_.copy_(pos, non_blocking=True); sn.copy_(neg, non_blocking=True)
                    else:
                        u, p, n = model._triples(users, pos, neg)
                        su.copy_(u, non_blocking=True); sp_.copy_(p, non_blocking=True); sn.copy_(n, non_blocking=True)
                graphs[0].replay()
                if dp:
                    bucket = model._ws["cache"]["bucket"]
                    w1 = bucket.all_reduce_mean_part(0, async_op=True)
                    graphs[1].replay()
                    w2 = bucket.all_reduce_mean_part(1, async_op=True)
                    if w1 is not None:
                        w1.wait()          # stream-level waits: the launching thread never blocks
                    graphs[2].replay()
                    if w2 is not None:
                        w2.wait()
                    graphs[3].replay()
                # what _forward / _loss do on the host besides launching: a replay is a new training forward, so every
                # table cached from the previous one (all_users / all_items / all_s_embs, normalised heads, fp16 splits)
                # is stale
                model._tables_version = getattr(model, "_tables_version", 0) + 1
                if model.lazy_tables:
                    model._tables_pending = True
                if deferred and not dp:
                    # step i is on its way; hand back the loss of step i-1, which the graph left in pinned memory (step i
                    # overwrites it only at its very end, a full step after the event we wait for here)
                    k = self._n_calls
                    done_ev[k & 1].record()
                    self._n_calls = k + 1
                    if k == 0:
                        return None
                    done_ev[(k - 1) & 1].synchronize()
                    return loss_host.clone()
                if host_loss and not dp:
                    torch.cuda.current_stream().synchronize()
                    return loss_host
                return loss

        return _Runner()

    # ------------------------------------------------------------------------------------------
    # scoring (predict, EliMRec.py:96-113) - tables cached by the last training forward
    # ------------------------------------------------------------------------------------------
    def _active_mods(self):
        if self.predict_type == "normal":
            return []
        if self.fusion_mode != "rubi":      # hm / sum multiply in every modality, whatever `modality` says (EliMRec.py:190-210)
            return list(range(len(self.mods)))
        return [j for j, m in enumerate(self.mods) if m in self.modality]

    def rank_tables(self):
        """Descriptor of the cached tables for the rank kernels (normalised single-modal tables are
        refreshed once per training forward, not per batch)."""
        if self.all_users is None:
            raise TypeError("'NoneType' object is not subscriptable (predict before any bpr_loss, as in the reference)")
        ws = self._ws["cache"]
        if ws.get("S_norm_version") != self._tables_version:
            fu, fi, su, si = self._tables()
            if "S_norm" not in ws:      # allocated once, refreshed in place after every training forward
                ws["S_norm"] = [(torch.empty_like(a), torch.empty_like(b)) for a, b in zip(su, si)]
            for (a, b), (an, bn) in zip(zip(su, si), ws["S_norm"]):
                ops.row_normalize(a, an)
                ops.row_normalize(b, bn)
            ws["rank_f"] = (fu, fi)
            ws["S_norm_version"] = self._tables_version
        act = self._active_mods()
        fu, fi = ws["rank_f"]
        su = [ws["S_norm"][j][0] for j in act]
        si = [ws["S_norm"][j][1] for j in act]
        mode = PREDICT_MODE[self.predict_type] + 4 * FUSION_MODE[self.fusion_mode]
        return ops.rank_tables(self.num_users, self.num_items, mode, fu, fi, su, si)

    def rank_tc_tables(self):
        """fp16 hi/lo split of the cached tables for the tensor-core evaluator (rebuilt once per training forward)."""
        if self.fusion_mode != "rubi":
            raise ElimrecError("the tensor-core evaluator implements the 'rubi' score fusion; hm / sum run on the fp32 rank path")
        self.rank_tables()                      # refreshes the normalised single-modal tables
        ws = self._ws["cache"]
        key = (self._tables_version, self.predict_type, self.modality)
        if ws.get("rank_tc_key") != key:
            fu, fi = ws["rank_f"]
            act = self._active_mods()
            srcs = [(fu, fi)] + [ws["S_norm"][j] for j in act]
            out = []
            bufs = ws.setdefault("rank_tc_bufs", {})       # fp16 hi / lo tables, allocated once and refilled
            for k, (a, b) in enumerate(srcs):
                parts = []
                scales = []
                for side_, x in enumerate((a, b)):
                    amax = float(x.abs().max()) if k == 0 else 1.0      # heads are L2-normalised: |x| <= 1
                    sc = 2.0 ** np.floor(np.log2(4096.0 / max(amax, 1e-30)))
                    key_ = (k, side_, tuple(x.shape))
                    if key_ not in bufs:
                        bufs[key_] = (torch.empty(x.shape, dtype=torch.float16, device=x.device),
                                      torch.empty(x.shape, dtype=torch.float16, device=x.device))
                    hi, lo = bufs[key_]
                    ops.split_fp16(x.contiguous(), float(sc), hi, lo)
                    parts += [hi, lo]
                    scales.append(sc)
                out.append((parts[0], parts[1], parts[2], parts[3], float(1.0 / (scales[0] * scales[1]))))
            ws["rank_tc"] = ops.rank_tc_tables(self.num_users, self.num_items, PREDICT_MODE[self.predict_type], out)
            ws["rank_tc_key"] = key
        return ws["rank_tc"]

    def _tables(self):
        """(fused users, fused items, [single-modal users], [single-modal items]) cached by the last training forward."""
        U = self.num_users
        S = self._ws["S"]
        return (self.all_users.contiguous(), self.all_items.contiguous(), [s_[:U].contiguous() for s_ in S],
                [s_[U:].contiguous() for s_ in S])

    @torch.no_grad()
    def predict(self, user_ids, candidate_items=None):
        users = torch.as_tensor(np.asarray(user_ids)).to(self.device_).int().contiguous()
        t = self.rank_tables()
        mean = None
        if self.predict_type == "TIE":
            mean = torch.empty(users.numel(), dtype=torch.float32, device=self.device_)
            ops.rank_rowmean(t, users, mean)
        out = torch.empty(users.numel(), self.num_items, dtype=torch.float32, device=self.device_)
        ops.rank_scores(t, users, mean, out)
        return out.cpu()

    def forward(self, users, items):
        raise NotImplementedError("EliMRec.forward is not on the path main.py drives")


__all__ = ["BasicModel", "EliMRec", "CALLS"]
