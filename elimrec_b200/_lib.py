"""ctypes binding of ``libelimrec_b200.so`` (the C-ABI declared in ``include/elimrec_b200.h``).

There is NO fallback: if the shared library is missing or a call fails, this raises.  The library
is built in-tree by ``python -m elimrec_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libelimrec_b200.so")

MAX_LAYERS = 8
MAX_MODS = 3
D = 64

vp, i32, i64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double


class MeanEpilogue(C.Structure):
    _fields_ = [("n_prev", i32), ("prev", vp * MAX_LAYERS), ("prev_ld", i64 * MAX_LAYERS),
                ("prev_width", i32 * MAX_LAYERS), ("mean_out", vp), ("mean_ld", i64), ("mean_width", i32),
                ("mean_scale", f32)]


class AdamTensor(C.Structure):
    _fields_ = [("param", vp), ("grad", vp), ("exp_avg", vp), ("exp_avg_sq", vp), ("numel", i64), ("row_len", i64),
                ("grad_ld", i64)]


class PrepTensor(C.Structure):
    _fields_ = [("src", vp), ("hi", vp), ("lo", vp), ("numel", i64)]


class LinearDesc(C.Structure):
    _fields_ = [("M", i64), ("K", i64), ("X", vp), ("ldx", i64), ("W", vp), ("b", vp), ("Y", vp), ("ldy", i64)]


class Spmm64Half(C.Structure):
    _fields_ = [("n_item", i32), ("n_split_item", i32), ("item", vp), ("split_rows", vp), ("counter", vp), ("partial", vp),
                ("col", vp), ("val", vp), ("X", vp), ("ldx", i64), ("Y", vp), ("ldy", i64), ("row_mask", vp), ("col_mask", vp),
                ("addend", vp), ("ld_add", i64), ("add_mask", vp), ("adam_param", vp), ("adam_exp_avg", vp),
                ("adam_exp_avg_sq", vp), ("adam_old_out", vp)]


class AdamConsts(C.Structure):
    _fields_ = [("consts_dev", vp), ("beta1", f64), ("beta2", f64), ("eps", f32), ("weight_decay", f32)]


class WgradProblem(C.Structure):
    _fields_ = [("A", vp), ("lda", i64), ("B", vp), ("ldb", i64), ("K", i64), ("row_begin", i64), ("row_end", i64), ("out", vp),
                ("ldo", i64), ("bias_out", vp), ("scale_by_g", i32)]


class LinLayers(C.Structure):
    _fields_ = [("n", i32), ("user", vp * (MAX_LAYERS + 1)), ("item", vp * (MAX_LAYERS + 1)),
                ("user_ld", i64 * (MAX_LAYERS + 1)), ("item_ld", i64 * (MAX_LAYERS + 1))]


class PackProj(C.Structure):
    _fields_ = [("W", vp), ("b", vp), ("dst", vp), ("Dm", i64), ("Kp", i64)]


class RankTables(C.Structure):
    _fields_ = [("num_users", i32), ("num_items", i32), ("n_mod", i32), ("mode", i32), ("f_user", vp),
                ("f_item", vp), ("s_user", vp * MAX_MODS), ("s_item", vp * MAX_MODS)]


class RankTcTables(C.Structure):
    _fields_ = [("num_users", i32), ("num_items", i32), ("n_mod", i32), ("mode", i32), ("user_hi", vp * (1 + MAX_MODS)),
                ("user_lo", vp * (1 + MAX_MODS)), ("item_hi", vp * (1 + MAX_MODS)), ("item_lo", vp * (1 + MAX_MODS)),
                ("inv_scale", f32 * (1 + MAX_MODS))]


_SIGS = {
    "elimrec_spmm": [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i64, vp, i64, vp, C.POINTER(MeanEpilogue), vp],
    "elimrec_spmm_masked": [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i64, vp, i64, vp, C.POINTER(MeanEpilogue), vp, vp, i32,
                            vp, i64, vp, vp],
    "elimrec_spmm64_pair": [C.POINTER(Spmm64Half), C.POINTER(Spmm64Half), C.POINTER(AdamConsts), i32, i32, vp],
    "elimrec_mark_rows": [i32, vp, i64, vp, vp],
    "elimrec_inst_rows": [i32, vp, vp, vp, i32, vp, i64, vp, vp, vp],
    "elimrec_mark_neighbors": [i32, vp, vp, vp, vp, vp],
    "elimrec_zero_rows": [i32, vp, i32, i32, i32, vp, i64, i32, vp],
    "elimrec_scatter_add_rows": [i32, vp, i32, i32, i32, vp, i64, i32, vp, i64, i32, f32, vp],
    "elimrec_gather_rows": [i32, vp, vp, i64, vp, i64, i32, vp],
    "elimrec_broadcast_cols": [i64, vp, i64, vp, i64, i32, vp],
    "elimrec_copy_2d": [i64, i32, vp, i64, vp, i64, vp],
    "elimrec_tie_blocks": [i64, vp, i64, vp, i64, i32, f32, vp],
    "elimrec_fold_blocks": [i64, vp, i64, i32, f32, vp, i64, vp],
    "elimrec_axpy_rows": [i64, i32, vp, vp, i64, vp, i64, vp],
    "elimrec_layer_mean": [i64, i32, i32, C.POINTER(vp), C.POINTER(i64), f32, vp, i64, vp],
    "elimrec_wgrad_multi": [i32, C.POINTER(WgradProblem), i32, vp, vp, vp],
    "elimrec_wgrad_multi_x3": [i32, C.POINTER(WgradProblem), i32, vp, vp, vp],
    "elimrec_lin_assemble": [i64, vp, i32, C.POINTER(LinLayers), f32, i32, i32, vp, i64, vp],
    "elimrec_lin_seed": [i32, vp, i32, i32, vp, i64, i32, f32, vp, i64, vp],
    "elimrec_lin_seed2": [i32, vp, vp, i64, i32, f32, vp, vp, i64, vp],
    "elimrec_pack_proj_weights": [i32, C.POINTER(PackProj), i32, vp],
    "elimrec_split3_rows": [i64, i32, vp, i64, vp, i64, i32, vp],
    "elimrec_axpy_2d": [i64, i32, f32, vp, i64, vp, i64, i32, vp],
    "elimrec_gemm": [i64, i64, i64, vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, i32, i32, vp, vp, vp],
    "elimrec_colsum": [i64, i64, vp, i64, vp, vp, i32, vp, vp],
    "elimrec_linear_tf32_fwd": [i64, i64, vp, i64, vp, vp, vp, i64, vp],
    "elimrec_linear_tf32_fwd_multi": [i32, C.POINTER(LinearDesc), vp],
    "elimrec_linear_tf32_wgrad": [i64, i64, vp, i64, vp, i64, vp, vp, vp],
    "elimrec_round_tf32": [i64, vp, vp, vp],
    "elimrec_split_tf32": [i64, vp, vp, vp, vp],
    "elimrec_prep_weights_tf32": [i32, C.POINTER(PrepTensor), vp],
    "elimrec_fuse_heads_x3_all": [i64, i64, i32, vp, i64, vp, vp, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), vp,
                                  C.POINTER(vp), vp],
    "elimrec_fuse_heads_x3": [i64, i32, vp, i64, vp, vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), vp, C.POINTER(vp), vp],
    "elimrec_linear_x3_fwd": [i64, i64, vp, i64, vp, vp, vp, vp, i64, vp],
    "elimrec_bpr_forward_backward": [i32, i32, C.POINTER(vp), C.POINTER(f32), vp, vp, vp, i32, vp, vp, vp, vp, vp],
    "elimrec_bpr_forward_backward_part": [i32, i32, i32, C.POINTER(vp), C.POINTER(f32), vp, vp, vp, i32, vp, vp, vp, vp, vp],
    "elimrec_inst_backward": [i32, i32, i32, vp, vp, vp, vp, vp, C.POINTER(vp), vp, vp, vp, vp, vp, C.POINTER(vp),
                              C.POINTER(vp), vp, vp],
    "elimrec_inst_dout_seed": [i32, i32, i32, vp, vp, vp, vp, C.POINTER(vp), vp, vp, i32, f32, vp, vp, i64, vp],
    "elimrec_inst_forward": [i32, i32, i32, vp, vp, vp, C.POINTER(vp), vp, vp, C.POINTER(vp), vp, C.POINTER(vp), vp],
    "elimrec_inst_backward_part": [i32, i32, i32, i32, vp, vp, vp, vp, vp, C.POINTER(vp), vp, vp, vp, vp, vp, C.POINTER(vp),
                              C.POINTER(vp), vp, vp],
    "elimrec_adam_apply_multi": [i32, C.POINTER(AdamTensor), vp, f64, f64, f32, f32, vp],
    "elimrec_adam_tick": [vp, vp, f64, f64, f64, vp],
    "elimrec_adam_apply": [i64, vp, vp, i64, i64, vp, vp, vp, f64, f64, f32, f32, vp],
    "elimrec_sample_epoch_compat": [vp, i32, vp, vp, vp, i32, i64, vp, vp, vp],
    "elimrec_sample_triples_device": [C.c_uint64, C.c_uint64, i64, i32, vp, vp, vp, i32, vp, vp, vp, vp],
    "elimrec_sample_batch_device": [C.c_uint64, C.c_uint64, vp, i64, i32, vp, vp, vp, i32, vp, vp, vp, vp],
    "elimrec_row_normalize": [i64, vp, vp, vp],
    "elimrec_rank_rowmean": [C.POINTER(RankTables), i32, vp, vp, vp],
    "elimrec_rank_scores": [C.POINTER(RankTables), i32, vp, vp, vp, vp],
    "elimrec_rank_topk": [C.POINTER(RankTables), i32, vp, vp, vp, vp, i32, vp, vp, vp],
    "elimrec_topk_matrix": [i32, i32, vp, i32, vp, vp, vp],
    "elimrec_mask_train": [i32, i32, vp, vp, vp, vp, vp],
    "elimrec_split_fp16": [i64, vp, f32, vp, vp, vp],
    "elimrec_rank_tc": [C.POINTER(RankTcTables), i32, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp],
    "elimrec_metric_rows": [i32, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp],
    "elimrec_cs_inst_rows": [i32, i32, vp, i32, vp, i64, vp, vp, vp],
    "elimrec_cs_pack": [i64, vp, i32, C.POINTER(LinLayers), f32, i32, vp, vp],
    "elimrec_cs_unpack": [i32, i64, i32, vp, i32, vp, i64, vp],
    "elimrec_cs_seed_pack": [i64, i32, i32, vp, i64, i32, f32, vp, vp],
    "elimrec_cs_seed_scatter": [i64, vp, i32, vp, vp, vp, i64, vp],
    "elimrec_comm_unique_id": [vp],
    "elimrec_comm_init": [vp, i32, i32, C.POINTER(vp)],
    "elimrec_comm_destroy": [vp],
    "elimrec_comm_allreduce": [vp, vp, vp, i64, i32, vp],
    "elimrec_comm_allgather": [vp, vp, vp, i64, vp],
    "elimrec_comm_alltoall": [vp, vp, vp, i64, vp],
}
_I64_RET = {
    "elimrec_gemm_workspace_floats": [i64, i64, i32],
    "elimrec_colsum_workspace_floats": [i64, i64],
    "elimrec_linear_tf32_wgrad_workspace_floats": [i64, i64],
    "elimrec_inst_backward_workspace_floats": [i32, i32, i32],
    "elimrec_rank_tc_workspace_bytes": [i32],
    "elimrec_wgrad_multi_workspace_floats": [i32, C.POINTER(WgradProblem), i32],
}
# every symbol include/elimrec_b200.h declares (tests/test_abi.py checks the header against this)
EXPORTS = sorted(list(_SIGS) + list(_I64_RET) + ["elimrec_last_error", "elimrec_abi_version",
                                                   "elimrec_compat_rng_seed", "elimrec_compat_rng_next"])

_lib = None


class ElimrecError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ElimrecError(f"{LIB_PATH} is missing - build it with `python -m elimrec_b200.build` "
                               "(there is no CPU / eager fallback)")
        l = C.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(l, name)
            fn.argtypes, fn.restype = args, C.c_int
        for name, args in _I64_RET.items():
            fn = getattr(l, name)
            fn.argtypes, fn.restype = args, i64
        l.elimrec_last_error.restype = C.c_char_p
        l.elimrec_abi_version.restype = C.c_int
        l.elimrec_compat_rng_seed.argtypes = [vp, C.c_uint32]
        l.elimrec_compat_rng_seed.restype = None
        l.elimrec_compat_rng_next.argtypes = [vp]
        l.elimrec_compat_rng_next.restype = C.c_uint32
        _lib = l
    return _lib


# kernels launched through the C-ABI (bench.py reports `gpu_launches` from this) and an optional
# per-family CUDA-event profile (bench.py --profile-kernels; events sit on the launching stream)
CALLS = {"n": 0, "launches": 0}
_LAUNCHES = {"elimrec_rank_tc": 2, "elimrec_colsum": 2, "elimrec_bpr_forward_backward": 2, "elimrec_metric_rows": 2, "elimrec_inst_backward": 3, "elimrec_inst_backward_part": 0,
             "elimrec_bpr_forward_backward_part": 0}
PROFILE = {"on": False, "events": []}


def call(name: str, *args, launches=None, tag=None):
    if PROFILE["on"]:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib(), name)(*args)
        e1.record()
        PROFILE["events"].append((tag or name, e0, e1))
    else:
        rc = getattr(lib(), name)(*args)
    CALLS["n"] += 1
    CALLS["launches"] += launches if launches is not None else _LAUNCHES.get(name, 1)
    if rc != 0:
        raise ElimrecError(f"{name} failed ({rc}): {lib().elimrec_last_error().decode()}")


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t, dtype=None, allow_none=False):
    """Device pointer of a CUDA tensor (no copy, no sync).  ``None`` -> NULL if allowed."""
    if t is None:
        if allow_none:
            return None
        raise ElimrecError("NULL tensor passed where a device buffer is required")
    if not t.is_cuda:
        raise ElimrecError("expected a CUDA tensor - this path has no CPU implementation")
    if dtype is not None and t.dtype != dtype:
        raise ElimrecError(f"expected dtype {dtype}, got {t.dtype}")
    return t.data_ptr()
