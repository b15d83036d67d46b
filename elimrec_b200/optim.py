"""Fused Adam - replaces ``torch.optim.Adam(params, lr, weight_decay)`` of ``main.py:49,99-101``.

Same arithmetic as torch's (coupled L2, betas (0.9, 0.999), eps 1e-8, bias correction), but
  * the step counter lives on the device (``elimrec_adam_tick``), so a whole training step can be
    replayed from a CUDA graph;
  * gradients are read in place from the backward workspace (strided views allowed), there is no
    ``.grad`` materialisation, no ``zero_grad`` pass.
Also usable as a plain optimizer over ``.grad`` (``zero_grad()`` / ``step()``) behind ``bpr_loss``.
"""
from __future__ import annotations

import torch

from . import ops


class FusedAdam:
    def __init__(self, model, lr=1e-3, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8):
        self.model = model
        self.lr, self.wd, self.betas, self.eps = float(lr), float(weight_decay), betas, float(eps)
        dev = model.device_
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.consts = torch.zeros(2, dtype=torch.float64, device=dev)
        self.state = {}

    def _st(self, name, p):
        st = self.state.get(name)
        if st is None:
            st = (torch.zeros_like(p.data), torch.zeros_like(p.data))
            self.state[name] = st
        return st

    def tick(self):
        """advance the device step counter and refresh the bias corrections (once per step, any time before apply)"""
        ops.adam_tick(self.step_dev, self.consts, self.lr, self.betas[0], self.betas[1])

    def apply(self, grads: dict, tick: bool = True):
        """One Adam step from a {name: gradient view} dict (EliMRec._backward output).  ``tick=False``: a second group of
        tensors of the SAME step (the step counter and bias corrections were advanced by the first call)."""
        P = self.model._params()
        if tick:
            self.tick()
        items = []
        for name, g in grads.items():
            p = P[name]
            m, v = self._st(name, p)
            items.append((p.data, g, m, v))
        ops.adam_apply_multi(items, self.consts, self.betas[0], self.betas[1], self.eps, self.wd)

    # torch.optim-like surface for the autograd path
    def zero_grad(self, set_to_none=True):
        for p in self.model.parameters():
            p.grad = None

    def step(self):
        P = self.model._params()
        self.apply({n: p.grad.contiguous() for n, p in P.items() if p.grad is not None})
