"""Checkpoint layout: one flat full-model fp32 state-dict ``./{model}_{data}.pth`` with
``layer{i}.*`` keys (reference src/Server.py:190,193,230-254).  Loadable by the reference's
own model classes, and vice versa.  ``slice_for_stage`` is the resume path: copy exactly the
keys the stage module owns.  Extra (not in the reference): an optional sidecar
``.meta.json`` with round counter for exact resume."""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Sequence

import torch

from .models import build_stage


def checkpoint_path(model_name: str, data_name: str, directory: str = ".") -> str:
    return os.path.join(directory, f"{model_name}_{data_name}.pth")


def save_checkpoint(state_dict: Dict[str, torch.Tensor], path: str, meta: Optional[dict] = None) -> None:
    import threading
    cpu = {k: v.detach().to("cpu") for k, v in state_dict.items()}
    # unique temporary name: two writer threads (asynchronous checkpoints of consecutive rounds) must not share one
    tmp = f"{path}.tmp.{os.getpid()}.{threading.get_ident()}"
    torch.save(cpu, tmp)
    os.replace(tmp, path)               # atomic: a crash never leaves a torn checkpoint
    if meta is not None:
        with open(path + ".meta.json", "w") as f:
            json.dump(meta, f)


def load_checkpoint(path: str) -> Optional[Dict[str, torch.Tensor]]:
    if not os.path.exists(path):
        return None
    return torch.load(path, map_location="cpu", weights_only=True)


def load_meta(path: str) -> dict:
    p = path + ".meta.json"
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def slice_for_stage(full_state_dict: Dict[str, torch.Tensor], model_name: str, data_name: Optional[str],
                    layers: Sequence[int]) -> Dict[str, torch.Tensor]:
    stage = build_stage(model_name, data_name, layers)
    keys = stage.state_dict().keys()
    return {k: full_state_dict[k] for k in keys}


def merge_stages(stage_dicts: Sequence[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    full: Dict[str, torch.Tensor] = {}
    for sd in stage_dicts:
        full.update(sd)
    return full
