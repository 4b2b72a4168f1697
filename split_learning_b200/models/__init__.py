"""Model registry: ``get_model_class(model_name, data_name)``.

The main reference tree hard-codes VGG16/BERT/KWT (src/RpcClient.py:78-92); the variants
resolve ``globals()[f"{model}_{data}"]`` (other/Vanilla_SL/src/Server.py:192).  We support
both: a (model, data) lookup with the main tree's defaults when ``data`` is omitted.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple, Type

from .base import LayerSpec, SplitModel
from .bert import BERT_AGNEWS, BERT_EMOTION
from .mobilenet import MobileNetv1_CIFAR10, MobileNetv1_MNIST
from .transformer import KWT_SPEECHCOMMANDS, ViT_CIFAR10, ViT_MNIST
from .vgg16 import VGG16_CIFAR10, VGG16_MNIST

_REGISTRY: Dict[Tuple[str, str], Type[SplitModel]] = {}
_DEFAULT_DATA = {"VGG16": "CIFAR10", "BERT": "AGNEWS", "KWT": "SPEECHCOMMANDS",
                 "ViT": "CIFAR10", "MobileNetv1": "CIFAR10"}


def register_model(cls: Type[SplitModel]) -> Type[SplitModel]:
    _REGISTRY[(cls.MODEL_NAME.upper(), cls.DATA_NAME.upper())] = cls
    return cls


for _c in (VGG16_CIFAR10, VGG16_MNIST, MobileNetv1_CIFAR10, MobileNetv1_MNIST, ViT_CIFAR10, ViT_MNIST,
           KWT_SPEECHCOMMANDS, BERT_AGNEWS, BERT_EMOTION):
    register_model(_c)


def get_model_class(model_name: str, data_name: Optional[str] = None) -> Type[SplitModel]:
    key_m = model_name.upper()
    if data_name is None:
        canon = {k.upper(): v for k, v in _DEFAULT_DATA.items()}
        data_name = canon.get(key_m)
    try:
        return _REGISTRY[(key_m, str(data_name).upper())]
    except KeyError:
        raise ValueError(f"unknown model {model_name}_{data_name}; known: {sorted(_REGISTRY)}") from None


def build_stage(model_name: str, data_name: Optional[str], layers) -> SplitModel:
    """Instantiate the stage described by a START message's ``layers=[a, b]``.

    Conventions (src/Server.py:222-228, src/RpcClient.py:86-92): ``b == -1`` → to the end,
    ``[0, 0]`` → whole model."""
    klass = get_model_class(model_name, data_name)
    a, b = int(layers[0]), int(layers[1])
    if b == 0:
        return klass()
    return klass(start_layer=a, end_layer=None if b == -1 else b)


__all__ = ["LayerSpec", "SplitModel", "get_model_class", "build_stage", "register_model",
           "VGG16_CIFAR10", "VGG16_MNIST", "MobileNetv1_CIFAR10", "MobileNetv1_MNIST", "ViT_CIFAR10",
           "ViT_MNIST", "KWT_SPEECHCOMMANDS", "BERT_AGNEWS", "BERT_EMOTION"]
