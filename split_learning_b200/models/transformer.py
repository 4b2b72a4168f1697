"""Small pre-LN transformer classifiers: KWT (speech commands) and ViT (CIFAR/MNIST).

Cut vocabulary / state-dict keys follow the reference:
  KWT  — src/model/KWT_SPEECHCOMMANDS.py:26-109 (17 layers; top-level ``cls_token`` /
         ``pos_embed`` parameters belong to layers 2 and 3)
  ViT  — other/Vanilla_SL/src/model/ViT_CIFAR10.py:27-116 (12 layers; ``cls_token`` is
         layer 3, ``pos_embed`` layer 4)
Both share one encoder block and one token-sequence head implementation here.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .base import LayerSpec, SplitModel


class EncoderBlock(nn.Module):
    """Pre-LN block: x + MHA(LN(x)); x + MLP(LN(x)). Keys: ln1, mha, ln2, mlp.{0,2}."""

    def __init__(self, embed_dim: int, num_heads: int, mlp_dim: int):
        super().__init__()
        self.ln1 = nn.LayerNorm(embed_dim)
        self.mha = nn.MultiheadAttention(embed_dim, num_heads, batch_first=True)
        self.ln2 = nn.LayerNorm(embed_dim)
        self.mlp = nn.Sequential(nn.Linear(embed_dim, mlp_dim), nn.GELU(), nn.Linear(mlp_dim, embed_dim))

    def forward(self, x):
        h = self.ln1(x)
        x = x + self.mha(h, h, h, need_weights=False)[0]
        return x + self.mlp(self.ln2(x))


class _TokenModel(SplitModel):
    """Shared machinery: top-level cls_token / pos_embed parameters tied to layer indices."""

    CLS_LAYER = 0
    POS_LAYER = 0
    EMBED = 0
    TOKENS = 0          # sequence length incl. CLS
    HEAD_NORM_LAYER = 0  # LayerNorm applied to x[:, 0]
    POS_DROPOUT: Optional[float] = None

    def __init__(self, start_layer=0, end_layer=None):
        super().__init__(start_layer, end_layer)
        if self._owns(self.CLS_LAYER):
            self.cls_token = nn.Parameter(torch.randn(1, 1, self.EMBED))
            nn.init.trunc_normal_(self.cls_token, std=0.02)
        if self._owns(self.POS_LAYER):
            self.pos_embed = nn.Parameter(torch.randn(1, self.TOKENS, self.EMBED))
            nn.init.trunc_normal_(self.pos_embed, std=0.02)
            if self.POS_DROPOUT is not None:
                self.dropout = nn.Dropout(self.POS_DROPOUT)

    def _owns(self, i: int) -> bool:
        return self.start_layer < i <= self.end_layer

    def _pre(self, i: int, x):
        return x

    def forward(self, x, **kwargs):
        for i in self.owned_indices():
            x = self._pre(i, x)
            if i == self.CLS_LAYER:
                x = torch.cat([self.cls_token.expand(x.size(0), -1, -1), x], dim=1)
                continue
            if i == self.POS_LAYER:
                x = x + self.pos_embed
                if self.POS_DROPOUT is not None:
                    x = self.dropout(x)
                continue
            if i == self.HEAD_NORM_LAYER:
                x = x[:, 0]
            x = getattr(self, f"layer{i}")(x)
            x = self._post(i, x)
        return x

    def _post(self, i: int, x):
        return x


class _Null(nn.Module):
    """Placeholder for indices whose state lives in top-level parameters."""

    def forward(self, x):
        return x


def _null():
    return LayerSpec("module", (_Null,))


def _block(embed, heads, mlp):
    return LayerSpec("module", (EncoderBlock, embed, heads, mlp))


class KWT_SPEECHCOMMANDS(_TokenModel):
    MODEL_NAME, DATA_NAME = "KWT", "SPEECHCOMMANDS"
    EMBED, TOKENS = 64, 99
    CLS_LAYER, POS_LAYER, HEAD_NORM_LAYER, POS_DROPOUT = 2, 3, 16, 0.1
    LAYERS = (
        [LayerSpec("linear", (40, 64)), _null(), _null()]
        + [_block(64, 1, 256) for _ in range(12)]
        + [LayerSpec("module", (nn.LayerNorm, 64)), LayerSpec("linear", (64, 10))]
    )

    def __init__(self, start_layer=0, end_layer=None):
        super().__init__(start_layer, end_layer)
        for i in (2, 3):  # reference has no layer2/layer3 attributes
            if hasattr(self, f"layer{i}"):
                delattr(self, f"layer{i}")

    def _pre(self, i, x):
        if i == 1:  # (B, n_mfcc, T) -> (B, T, n_mfcc)
            x = x.transpose(1, 2)
        return x

    @classmethod
    def example_input(cls, batch, device="cpu"):
        return torch.randn(batch, 40, 98, device=device)

    @classmethod
    def num_classes(cls):
        return 10


class _ViT(_TokenModel):
    EMBED = 128
    CLS_LAYER, POS_LAYER, HEAD_NORM_LAYER = 3, 4, 11
    IN_CH, IMG = 3, 32

    def __init__(self, start_layer=0, end_layer=None):
        super().__init__(start_layer, end_layer)
        if hasattr(self, "layer3"):
            delattr(self, "layer3")  # reference: cls_token only; layer4 is nn.Identity (kept)

    def _post(self, i, x):
        if i == 2:  # after Flatten(2): (B, E, P) -> (B, P, E)
            x = x.transpose(1, 2)
        return x

    @classmethod
    def example_input(cls, batch, device="cpu"):
        return torch.randn(batch, cls.IN_CH, cls.IMG, cls.IMG, device=device)

    @classmethod
    def num_classes(cls):
        return 10


def _vit_table(in_ch):
    return (
        [LayerSpec("conv", (in_ch, 128), {"kernel_size": 4, "stride": 4}), LayerSpec("flatten", (2,)),
         _null(), LayerSpec("module", (nn.Identity,))]
        + [_block(128, 4, 256) for _ in range(6)]
        + [LayerSpec("module", (nn.LayerNorm, 128)), LayerSpec("linear", (128, 10))]
    )


class ViT_CIFAR10(_ViT):
    MODEL_NAME, DATA_NAME = "ViT", "CIFAR10"
    IN_CH, IMG, TOKENS = 3, 32, 65
    LAYERS = _vit_table(3)


class ViT_MNIST(_ViT):
    MODEL_NAME, DATA_NAME = "ViT", "MNIST"
    IN_CH, IMG, TOKENS = 1, 28, 50
    LAYERS = _vit_table(1)
