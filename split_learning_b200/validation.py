"""Server-side validation of the merged full model (reference src/val/*.py, get_val.py).

Full model on the test split, batch 100 (images) / 20 (text, audio), CE as mean of batch
means + accuracy, logged through the server logger.  ``get_val`` returns ``(ok, metrics)``;
``strict=True`` adds the Vanilla_SL gate that fails the round on NaN / exploding loss
(other/Vanilla_SL/src/Validation.py:46,55-56).  ``pump`` is called every 5 batches — the
hook DCSL uses to keep its AMQP heartbeat alive (other/DCSL/src/Validation.py:48-52); here it
lets the broker thread breathe / watchdogs tick.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn as nn

from .data import data_loader
from .models import get_model_class


@torch.no_grad()
def evaluate(model: nn.Module, loader, device, pump: Optional[Callable[[], None]] = None) -> Dict[str, float]:
    criterion = nn.CrossEntropyLoss()
    model.eval()
    correct = total = 0
    total_loss, batches = 0.0, 0
    for i, batch in enumerate(loader):
        if isinstance(batch, dict):
            x, y = batch["input_ids"].to(device), batch["labels"].to(device)
            out = model(input_ids=x) if "input_ids" in model.forward.__code__.co_varnames else model(x)
        else:
            x, y = batch[0].to(device), torch.as_tensor(batch[1]).to(device)
            out = model(x)
        total_loss += float(criterion(out, y))
        correct += int((out.argmax(1) == y).sum())
        total += int(y.numel())
        batches += 1
        if pump is not None and i % 5 == 4:
            pump()
    return {"val_loss": total_loss / max(batches, 1), "val_acc": 100.0 * correct / max(total, 1),
            "val_correct": correct, "val_total": total}


def get_val(model_name: str, data_name: str, state_dict_full, logger=None, strict: bool = False,
            device: Optional[str] = None, pump=None, synthetic: Optional[bool] = None) -> Tuple[bool, Dict[str, float]]:
    try:
        klass = get_model_class(model_name, data_name)
    except ValueError:
        return False, {}
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    model = klass()
    model.load_state_dict(state_dict_full)
    model.to(device)
    loader = data_loader(data_name=data_name, train=False, synthetic=synthetic)
    m = evaluate(model, loader, device, pump)
    line = "Test set:Loss: {:.4f}; Accuracy: {}/{} ({:.2f}%)\n".format(
        m["val_loss"], m["val_correct"], m["val_total"], m["val_acc"])
    if logger is not None:
        logger.log_info(line)
    if strict and (math.isnan(m["val_loss"]) or abs(m["val_loss"]) > 1e6):
        return False, m
    return True, m
