"""BERT-base encoders as cut-able tables.

* ``BERT_AGNEWS``  — 15 indexed layers (1 embeddings, 2-13 encoder blocks, 14 pooler,
  15 classifier/4 labels); key layout as src/model/BERT_AGNEWS.py:167-219.
* ``BERT_EMOTION`` — 27 indexed layers where each encoder block is cut-able between its
  attention half (even index) and its FFN half (odd index); key layout as
  other/Vanilla_SL/src/model/BERT_EMOTION.py:183-428 (``layer{i}.0.*`` / ``layer{i}.1.*``).

Implementation is our own: attention uses one fused QKV projection view over the three
reference-named Linear modules and ``scaled_dot_product_attention`` (dense, S<=128), which
is numerically the reference's softmax(QK^T/sqrt(d))V with dropout on the probabilities.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .base import LayerSpec, SplitModel


class AttrDict(dict):
    """dict with attribute access — PEFT-style code reads ``model.config.<field>``."""

    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


class Embeddings(nn.Module):
    def __init__(self, vocab, hidden, max_pos, type_vocab, p):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, hidden, padding_idx=0)
        self.position_embeddings = nn.Embedding(max_pos, hidden)
        self.token_type_embeddings = nn.Embedding(type_vocab, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=1e-12)
        self.dropout = nn.Dropout(p)

    def forward(self, input_ids, token_type_ids=None):
        if input_ids.dtype != torch.long:
            input_ids = input_ids.long()
        pos = torch.arange(input_ids.size(1), device=input_ids.device).unsqueeze(0)
        tok = torch.zeros_like(input_ids) if token_type_ids is None else token_type_ids
        e = self.word_embeddings(input_ids) + self.position_embeddings(pos) + self.token_type_embeddings(tok)
        return self.dropout(self.LayerNorm(e))


class SelfAttention(nn.Module):
    """Multi-head self attention; parameter names query/key/value as in HF BERT."""

    def __init__(self, hidden, heads, p):
        super().__init__()
        self.heads = heads
        self.head_dim = hidden // heads
        self.query = nn.Linear(hidden, hidden)
        self.key = nn.Linear(hidden, hidden)
        self.value = nn.Linear(hidden, hidden)
        self.dropout = nn.Dropout(p)

    def forward(self, x, attention_mask=None):
        b, s, h = x.shape
        split = lambda t: t.view(b, s, self.heads, self.head_dim).transpose(1, 2)
        q, k, v = split(self.query(x)), split(self.key(x)), split(self.value(x))
        bias = None
        if attention_mask is not None:
            bias = (1.0 - attention_mask[:, None, None, :].to(q.dtype)) * -10000.0
        ctx = F.scaled_dot_product_attention(
            q, k, v, attn_mask=bias, dropout_p=self.dropout.p if self.training else 0.0)
        return ctx.transpose(1, 2).reshape(b, s, h)


class ResidualDenseNorm(nn.Module):
    """dense -> dropout -> LayerNorm(residual + .)   (BertSelfOutput / BertOutput)."""

    def __init__(self, d_in, d_out, p):
        super().__init__()
        self.dense = nn.Linear(d_in, d_out)
        self.LayerNorm = nn.LayerNorm(d_out, eps=1e-12)
        self.dropout = nn.Dropout(p)

    def forward(self, h, residual):
        return self.LayerNorm(self.dropout(self.dense(h)) + residual)


class Intermediate(nn.Module):
    def __init__(self, hidden, inter):
        super().__init__()
        self.dense = nn.Linear(hidden, inter)
        self.intermediate_act_fn = nn.GELU()

    def forward(self, x):
        return self.intermediate_act_fn(self.dense(x))


class Attention(nn.Module):
    def __init__(self, hidden, heads, p):
        super().__init__()
        self.self = SelfAttention(hidden, heads, p)
        self.output = ResidualDenseNorm(hidden, hidden, p)

    def forward(self, x, attention_mask=None):
        return self.output(self.self(x, attention_mask), x)


class EncoderLayer(nn.Module):
    def __init__(self, hidden, heads, inter, p):
        super().__init__()
        self.attention = Attention(hidden, heads, p)
        self.intermediate = Intermediate(hidden, inter)
        self.output = ResidualDenseNorm(inter, hidden, p)

    def forward(self, x):
        a = self.attention(x)
        return self.output(self.intermediate(a), a)


class Pooler(nn.Module):
    def __init__(self, hidden):
        super().__init__()
        self.dense = nn.Linear(hidden, hidden)
        self.activation = nn.Tanh()

    def forward(self, x):
        return self.activation(self.dense(x[:, 0]))


class Classifier(nn.Module):
    def __init__(self, hidden, labels, p=0.1):
        super().__init__()
        self.dropout = nn.Dropout(p)
        self.classifier = nn.Linear(hidden, labels)

    def forward(self, x):
        return self.classifier(self.dropout(x))


class AttentionHalf(nn.ModuleList):
    """[SelfAttention, ResidualDenseNorm] — the attention half of a block (BERT_EMOTION)."""

    def __init__(self, hidden, heads, p):
        super().__init__([SelfAttention(hidden, heads, p), ResidualDenseNorm(hidden, hidden, p)])

    def forward(self, x, attention_mask=None):
        return self[1](self[0](x, attention_mask), x)


class FfnHalf(nn.ModuleList):
    def __init__(self, hidden, inter, p):
        super().__init__([Intermediate(hidden, inter), ResidualDenseNorm(inter, hidden, p)])

    def forward(self, x):
        return self[1](self[0](x), x)


_H, _NH, _I, _P = 768, 12, 3072, 0.1


class BERT_AGNEWS(SplitModel):
    MODEL_NAME, DATA_NAME = "BERT", "AGNEWS"
    VOCAB = 28996
    LAYERS = (
        [LayerSpec("module", (Embeddings, VOCAB, _H, 512, 2, _P))]
        + [LayerSpec("module", (EncoderLayer, _H, _NH, _I, _P)) for _ in range(12)]
        + [LayerSpec("module", (Pooler, _H)), LayerSpec("module", (Classifier, _H, 4))]
    )

    def __init__(self, start_layer=0, end_layer=None, **_):
        super().__init__(start_layer, end_layer)
        self.config = AttrDict(
            model_type="bert", vocab_size=self.VOCAB, hidden_size=_H, num_attention_heads=_NH,
            intermediate_size=_I, max_position_embeddings=512, bos_token_id=101, eos_token_id=102,
            pad_token_id=0, is_encoder_decoder=False, tie_word_embeddings=False,
            use_return_dict=True, output_attentions=False, output_hidden_states=False)

    def forward(self, input_ids=None, token_type_ids=None, **kwargs):
        x = input_ids
        for i in self.owned_indices():
            layer = getattr(self, f"layer{i}")
            x = layer(x, token_type_ids) if i == 1 else layer(x)
        return x

    @classmethod
    def example_input(cls, batch, device="cpu"):
        return torch.randint(1, cls.VOCAB, (batch, 128), device=device)

    @classmethod
    def num_classes(cls):
        return 4


class BERT_EMOTION(SplitModel):
    MODEL_NAME, DATA_NAME = "BERT", "EMOTION"
    VOCAB = 30522
    LAYERS = (
        [LayerSpec("module", (Embeddings, VOCAB, _H, 512, 2, _P))]
        + [spec for _ in range(12) for spec in (
            LayerSpec("module", (AttentionHalf, _H, _NH, _P)),
            LayerSpec("module", (FfnHalf, _H, _I, _P)))]
        + [LayerSpec("module", (Pooler, _H)), LayerSpec("module", (Classifier, _H, 4))]
    )

    def forward(self, x=None, attention_mask=None, token_type_ids=None, input_ids=None, **kwargs):
        if x is None:
            x = input_ids
        for i in self.owned_indices():
            layer = getattr(self, f"layer{i}")
            if i == 1:
                x = layer(x, token_type_ids)
            elif i <= 25 and i % 2 == 0:
                x = layer(x, attention_mask)
            else:
                x = layer(x)
        return x

    @classmethod
    def example_input(cls, batch, device="cpu"):
        return torch.randint(1, cls.VOCAB, (batch, 128), device=device)

    @classmethod
    def num_classes(cls):
        return 4


assert len(BERT_AGNEWS.LAYERS) == 15 and len(BERT_EMOTION.LAYERS) == 27
