"""Flat fused optimizers for the torch-executed model families (BERT / KWT / ViT / MobileNet on CUDA).

All trainable parameters of a stage are re-homed into ONE flat fp32 buffer (each ``param.data`` becomes a view),
gradients accumulate into a matching flat buffer (``param.grad`` views), and a step is a single launch of the
hand-written ``sgd_momentum_kernel`` / ``adamw_kernel`` — which also zeroes the gradient buffer, so
``zero_grad`` is free.  Semantics match ``torch.optim.SGD(lr, momentum)`` (reference src/train/VGG16.py:62) and
``torch.optim.AdamW(lr, weight_decay)`` (src/train/BERT.py:69, src/train/KWT.py:62) with their default betas/eps.
"""
from __future__ import annotations

from typing import Iterable, List

import torch

from . import native as N


def _align(n: int, a: int = 128) -> int:
    return (n + a - 1) // a * a


class FlatFusedOptimizer(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], kind: str = "sgd", lr: float = 1e-3, momentum: float = 0.0,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01, shadow: bool = False):
        """``shadow``: keep a flat bf16 copy of the parameters, refreshed inside the update kernel; the token-model
        kernels (``ops.nn``) read weights through ``param._slb_bf16`` views of it."""
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("no trainable parameters")
        dev = params[0].device
        if dev.type != "cuda" or any(p.dtype != torch.float32 or p.device != dev for p in params):
            raise ValueError("FlatFusedOptimizer needs fp32 CUDA parameters on one device")
        super().__init__(params, dict(lr=lr, momentum=momentum, betas=betas, eps=eps, weight_decay=weight_decay))
        N.require()
        self.kind = kind
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += _align(p.numel())
        self.flat_p = torch.zeros(total, device=dev)
        self.flat_g = torch.zeros(total, device=dev)
        self.flat_m = torch.zeros(total, device=dev)
        self.flat_v = torch.zeros(total, device=dev) if kind == "adamw" else None
        with torch.no_grad():
            for p, o in zip(params, offs):
                view = self.flat_p[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_g[o:o + p.numel()].view_as(p)
        self._params: List[torch.nn.Parameter] = params
        self._offs = offs
        self.steps = 0
        self.dev_step = torch.zeros(1, dtype=torch.int32, device=dev)     # AdamW bias correction reads it on the device
        self.flat_pb = None
        if shadow:
            self.flat_pb = torch.empty(total, dtype=torch.bfloat16, device=dev)
            for p, o in zip(params, offs):
                p._slb_bf16 = self.flat_pb[o:o + p.numel()].view_as(p)
            self.refresh_shadow()

    def refresh_shadow(self) -> None:
        if self.flat_pb is not None:
            N.cast_f32_bf16(self.flat_p, self.flat_pb)
            for p in self._params:
                p._slb_ver = p._version

    def zero_grad(self, set_to_none: bool = True) -> None:      # the fused kernels zero the flat gradient buffer
        for p, o in zip(self._params, self._offs):
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * o:
                p.grad = self.flat_g[o:o + p.numel()].view_as(p)

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        for p, o in zip(self._params, self._offs):               # a grad that autograd re-allocated: fold it back in
            if p.grad is not None and p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * o:
                self.flat_g[o:o + p.numel()].view_as(p).add_(p.grad)
                p.grad = self.flat_g[o:o + p.numel()].view_as(p)
        self.steps += 1                     # host mirror only (not advanced by graph replays)
        if self.kind == "adamw":
            b1, b2 = g["betas"]
            N.counter_inc(self.dev_step)
            N.adamw(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.flat_pb, g["lr"], b1, b2, g["eps"], g["weight_decay"],
                    self.steps, self.dev_step)
        else:
            N.sgd_momentum(self.flat_p, self.flat_g, self.flat_m, self.flat_pb, g["lr"], g["momentum"])
        return None
