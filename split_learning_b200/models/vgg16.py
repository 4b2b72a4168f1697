"""VGG16 for CIFAR-10 / MNIST as layer tables.

Index vocabulary and state-dict keys match the reference
(src/model/VGG16_CIFAR10.py:9-117: 52 indexed layers;
other/Vanilla_SL/src/model/VGG16_MNIST.py: 51 layers, no last pool).
The table is generated from the classic VGG configuration string rather than
spelled out layer by layer.
"""
from __future__ import annotations

from typing import List

import torch

from .base import LayerSpec, SplitModel

# (channels..., 'M' = 2x2 max-pool)
_VGG16_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]


def _vgg_table(in_ch: int, cfg, classifier_in: int, num_classes: int) -> List[LayerSpec]:
    t: List[LayerSpec] = []
    c = in_ch
    for item in cfg:
        if item == "M":
            t.append(LayerSpec("maxpool2"))
        else:
            t += [LayerSpec("conv3x3", (c, item)), LayerSpec("bn2d", (item,)), LayerSpec("relu")]
            c = item
    t += [
        LayerSpec("flatten", (1, -1)),
        LayerSpec("dropout", (0.5,)),
        LayerSpec("linear", (classifier_in, 4096)),
        LayerSpec("relu"),
        LayerSpec("dropout", (0.5,)),
        LayerSpec("linear", (4096, 4096)),
        LayerSpec("relu"),
        LayerSpec("linear", (4096, num_classes)),
    ]
    return t


class VGG16_CIFAR10(SplitModel):
    LAYERS = _vgg_table(3, _VGG16_CFG, 512, 10)
    MODEL_NAME, DATA_NAME = "VGG16", "CIFAR10"

    @classmethod
    def example_input(cls, batch, device="cpu"):
        return torch.randn(batch, 3, 32, 32, device=device)

    @classmethod
    def num_classes(cls):
        return 10


class VGG16_MNIST(SplitModel):
    # 28x28 input: four pools bring it to 1x1, so the fifth pool is dropped (51 layers)
    LAYERS = _vgg_table(1, _VGG16_CFG[:-1], 512, 10)
    MODEL_NAME, DATA_NAME = "VGG16", "MNIST"

    @classmethod
    def example_input(cls, batch, device="cpu"):
        return torch.randn(batch, 1, 28, 28, device=device)

    @classmethod
    def num_classes(cls):
        return 10


assert len(VGG16_CIFAR10.LAYERS) == 52 and len(VGG16_MNIST.LAYERS) == 51
