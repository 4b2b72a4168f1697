"""MobileNetv1-style CNN (84 indexed layers) for CIFAR-10 / MNIST.

Layer vocabulary matches other/Vanilla_SL/src/model/MobileNetv1_CIFAR10.py:5-185
(note: the reference's "depthwise" 3x3 convs are dense, groups=1).
Generated from a compact (kind, channels, stride) program.
"""
from __future__ import annotations

from typing import List

import torch

from .base import LayerSpec, SplitModel

# sequence of (kernel, out_channels, stride) after the stem
_PROGRAM = [
    (3, 32, 1),                      # stem
    (3, 32, 1), (1, 64, 1),
    (3, 64, 2), (1, 128, 1),
    (3, 128, 1), (1, 128, 1),
    (3, 128, 2), (1, 256, 1),
    (3, 256, 1), (1, 256, 1),
    (3, 256, 2), (1, 512, 1),
] + [(3, 512, 1), (1, 512, 1)] * 5 + [
    (3, 512, 2), (1, 1024, 1),
    (3, 1024, 1), (1, 1024, 1),
]


def _table(in_ch: int) -> List[LayerSpec]:
    t: List[LayerSpec] = []
    c = in_ch
    for k, out, s in _PROGRAM:
        if k == 3:
            t.append(LayerSpec("conv", (c, out, 3, s, 1)))
        else:
            t.append(LayerSpec("conv", (c, out, 1)))
        t += [LayerSpec("bn2d", (out,)), LayerSpec("relu")]
        c = out
    t += [LayerSpec("maxpool2"), LayerSpec("flatten", (1, -1)), LayerSpec("linear", (1024, 10))]
    return t


class MobileNetv1_CIFAR10(SplitModel):
    LAYERS = _table(3)
    MODEL_NAME, DATA_NAME = "MobileNetv1", "CIFAR10"

    @classmethod
    def example_input(cls, batch, device="cpu"):
        return torch.randn(batch, 3, 32, 32, device=device)

    @classmethod
    def num_classes(cls):
        return 10


class MobileNetv1_MNIST(SplitModel):
    LAYERS = _table(1)
    MODEL_NAME, DATA_NAME = "MobileNetv1", "MNIST"

    @classmethod
    def example_input(cls, batch, device="cpu"):
        return torch.randn(batch, 1, 32, 32, device=device)

    @classmethod
    def num_classes(cls):
        return 10


assert len(MobileNetv1_CIFAR10.LAYERS) == 84
