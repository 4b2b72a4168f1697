"""Minimal LoRA (no ``peft`` dependency).

The reference wraps BERT stages with peft's LoRA (r=8, alpha=16, dropout=0.1, target
modules query/key/value/dense, bias none) and merges before uploading weights
(src/RpcClient.py:61-66,99-103,121-122).  ``peft`` is not installable here, so this is a
self-contained equivalent: ``apply_lora`` swaps matching ``nn.Linear`` modules for
``LoRALinear`` in place (base weights frozen), ``merge_lora`` folds ``B@A * alpha/r`` back
and restores plain Linears, so the uploaded state-dict has the reference's key layout.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Iterable, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class LoraConfig:
    r: int = 8
    lora_alpha: int = 16
    lora_dropout: float = 0.1
    target_modules: Sequence[str] = ("query", "key", "value", "dense")
    bias: str = "none"
    task_type: str = "SEQ_CLS"


class LoRALinear(nn.Module):
    def __init__(self, base: nn.Linear, cfg: LoraConfig):
        super().__init__()
        self.base = base
        for p in self.base.parameters():
            p.requires_grad = False
        self.scaling = cfg.lora_alpha / cfg.r
        self.lora_A = nn.Parameter(torch.empty(cfg.r, base.in_features, device=base.weight.device))
        self.lora_B = nn.Parameter(torch.zeros(base.out_features, cfg.r, device=base.weight.device))
        nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
        self.drop = nn.Dropout(cfg.lora_dropout)

    def forward(self, x):
        return self.base(x) + F.linear(F.linear(self.drop(x), self.lora_A), self.lora_B) * self.scaling

    @torch.no_grad()
    def merged(self) -> nn.Linear:
        self.base.weight += (self.lora_B @ self.lora_A) * self.scaling
        for p in self.base.parameters():
            p.requires_grad = True
        return self.base


def _walk(module: nn.Module, prefix=""):
    for name, child in module.named_children():
        yield module, name, child, f"{prefix}{name}"
        yield from _walk(child, f"{prefix}{name}.")


def apply_lora(model: nn.Module, cfg: LoraConfig = LoraConfig(), keep_trainable: Iterable[str] = ()) -> nn.Module:
    """Freeze everything, then LoRA-wrap target Linears.  ``keep_trainable``: name prefixes
    left fully trainable (the reference keeps ``layer15.classifier`` trainable on stage 2)."""
    for p in model.parameters():
        p.requires_grad = False
    targets = [(parent, name) for parent, name, child, _ in list(_walk(model))
               if isinstance(child, nn.Linear) and name in cfg.target_modules]
    for parent, name in targets:
        setattr(parent, name, LoRALinear(getattr(parent, name), cfg))
    for n, p in model.named_parameters():
        if any(n.startswith(k) for k in keep_trainable):
            p.requires_grad = True
    return model


def merge_lora(model: nn.Module) -> nn.Module:
    for parent, name, child, _ in list(_walk(model)):
        if isinstance(child, LoRALinear):
            setattr(parent, name, child.merged())
    for p in model.parameters():
        p.requires_grad = True
    return model
