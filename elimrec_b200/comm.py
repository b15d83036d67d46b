"""Collectives of the multi-GPU modes (one process per GPU).

On CUDA the collectives go through the C-ABI ``comm`` family (``csrc/comm.cu``): an NCCL communicator owned by the library,
created from a unique id that is broadcast ONCE through ``torch.distributed``; every call is enqueued on the current torch
stream, in order with the kernels around it - no host synchronisation, and capturable into the step's CUDA graph (collectives
issued through torch's own process group are not, the one-graph data-parallel step of round 1 hung on exactly that).
On CPU tensors (the ``gloo`` world-size-2 tests of the host logic) the same three operations map to ``torch.distributed``.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from ._lib import ElimrecError, call, ptr, stream


class Comm:
    def __init__(self, device):
        if not (dist.is_available() and dist.is_initialized()):
            raise ElimrecError("Comm needs an initialised torch.distributed process group (one process per GPU)")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = torch.device(device)
        self.native = self.device.type == "cuda"
        self._h = None
        if self.native:
            idt = torch.zeros(128, dtype=torch.uint8, device=self.device)
            if self.rank == 0:
                buf = (C.c_uint8 * 128)()
                call("elimrec_comm_unique_id", buf, launches=0)
                idt.copy_(torch.frombuffer(bytearray(buf), dtype=torch.uint8))
            dist.broadcast(idt, src=0)
            raw = (C.c_uint8 * 128).from_buffer_copy(bytes(idt.cpu().numpy().tobytes()))
            h = C.c_void_p()
            call("elimrec_comm_init", raw, self.rank, self.world, C.byref(h), launches=0)
            self._h = h

    def close(self):
        """destroy the communicator.  Call it explicitly, after every CUDA graph that captured its collectives is gone and
        the device is idle; at interpreter exit the communicator is simply left to the process teardown (destroying it from a
        finaliser, in whatever order the garbage collector picks, can block on the peer)."""
        if self._h is not None:
            _lib.lib().elimrec_comm_destroy(self._h)
            self._h = None

    # ---- the three collectives -----------------------------------------------------------------------------------------
    def all_reduce(self, t: torch.Tensor, average: bool = False):
        """in place: t = sum (or mean) over ranks of t"""
        if self.world == 1 or t.numel() == 0:
            return
        if self.native:
            call("elimrec_comm_allreduce", self._h, ptr(t, torch.float32), ptr(t, torch.float32), t.numel(), int(average), stream(),
                 launches=1, tag="comm_allreduce")
        else:
            dist.all_reduce(t)
            if average:
                t.div_(self.world)

    def all_gather(self, out: torch.Tensor, inp: torch.Tensor):
        """out [world * n] (flat) = every rank's inp [n], in rank order"""
        if not (out.is_contiguous() and inp.is_contiguous()) or out.numel() != self.world * inp.numel() or out.dtype != inp.dtype:
            raise ElimrecError("all_gather: out must be a contiguous [world x inp] buffer of the same dtype")
        if self.native:
            call("elimrec_comm_allgather", self._h, ptr(inp), ptr(out), inp.numel() * inp.element_size(), stream(), launches=1,
                 tag="comm_allgather")
        else:
            dist.all_gather_into_tensor(out.view(-1), inp.view(-1))

    def all_to_all(self, out: torch.Tensor, inp: torch.Tensor):
        """out[q] = rank q's inp[me], both [world x n] contiguous"""
        if not (out.is_contiguous() and inp.is_contiguous()) or out.numel() != inp.numel() or inp.numel() % self.world:
            raise ElimrecError("all_to_all: contiguous [world x n] buffers of equal size required")
        if self.native:
            call("elimrec_comm_alltoall", self._h, ptr(inp), ptr(out), inp.numel() * inp.element_size() // self.world, stream(),
                 launches=1, tag="comm_alltoall")
        else:
            dist.all_to_all_single(out.view(-1), inp.view(-1))
