"""Import-time stand-in for ``peft`` (not installable offline).  The reference imports it at
module top (src/RpcClient.py:14) but only *calls* it for BERT; the VGG16 benchmark path never
does, so these names only need to exist."""


class LoraConfig:
    def __init__(self, *a, **k):
        raise RuntimeError("peft is not available in this image (BERT+LoRA reference arm unsupported)")


class TaskType:
    SEQ_CLS = "SEQ_CLS"


def get_peft_model(*a, **k):
    raise RuntimeError("peft is not available in this image")
