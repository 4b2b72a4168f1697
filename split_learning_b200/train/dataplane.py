"""Host data plane: activations / gradients as pickled CPU arrays through the broker.

This is the *compatibility* path (CPU, gloo-less tests, heterogeneous devices): same
message fields as the reference (``data_id, data, label, trace`` — src/train/VGG16.py:20-53)
and the queue grammar of every variant.  On B200 the data plane is
``parallel.mailbox.PeerDataPlane`` instead: tiles are stored into the consumer's HBM from
inside the producing kernel and no host code touches the payload.
"""
from __future__ import annotations

import pickle
from typing import List, Optional

import numpy as np
import torch

from ..transport import Channel


class QueueGrammar:
    """Queue names per algorithm variant."""

    def __init__(self, variant: str = "main"):
        self.variant = variant

    def forward_queue(self, layer_id: int, cluster=None, target=None) -> str:
        v = self.variant
        if v in ("vanilla_sl", "cluster_fsl"):
            return f"intermediate_queue_{layer_id}"
        if v == "dcsl":
            return f"intermediate_queue_{target}" if target is not None else f"intermediate_queue_{layer_id}"
        if v == "2ls":
            return f"intermediate_queue_{layer_id}_{target}"
        return f"intermediate_queue_{layer_id}_{cluster}"

    def gradient_queue(self, layer_id: int, client_id) -> str:
        return f"gradient_queue_{layer_id}_{client_id}"


def _to_numpy(t: torch.Tensor) -> np.ndarray:
    t = t.detach()
    if t.dtype == torch.bfloat16:
        t = t.float()
    return t.cpu().numpy()


class HostDataPlane:
    def __init__(self, channel: Channel, client_id, layer_id: int, cluster=None, grammar: Optional[QueueGrammar] = None,
                 device="cpu"):
        self.ch, self.client_id, self.layer_id, self.cluster = channel, client_id, layer_id, cluster
        self.grammar = grammar or QueueGrammar("main")
        self.device = device
        self.my_grad_q = self.grammar.gradient_queue(layer_id, client_id)
        self.ch.queue_declare(self.my_grad_q)

    # ---- producer side -------------------------------------------------
    def send_forward(self, data_id, output: torch.Tensor, labels, trace: Optional[List] = None, target=None) -> None:
        trace = list(trace) + [self.client_id] if trace else [self.client_id]
        q = self.grammar.forward_queue(self.layer_id, self.cluster, target)
        lab = labels.detach().cpu() if isinstance(labels, torch.Tensor) else labels
        self.ch.basic_publish(q, pickle.dumps(
            {"data_id": data_id, "data": _to_numpy(output), "label": lab, "trace": trace},
            protocol=pickle.HIGHEST_PROTOCOL))

    def send_gradient(self, data_id, gradient: torch.Tensor, trace: List) -> None:
        trace = list(trace)
        to_client = trace.pop(-1)
        q = self.grammar.gradient_queue(self.layer_id - 1, to_client)
        self.ch.basic_publish(q, pickle.dumps(
            {"data_id": data_id, "data": _to_numpy(gradient), "trace": trace}, protocol=pickle.HIGHEST_PROTOCOL))

    # ---- consumer side -------------------------------------------------
    def recv_forward(self, timeout: float = 0.0, source=None):
        q = self.grammar.forward_queue(self.layer_id - 1, self.cluster, source)
        body = self.ch.basic_get(q, timeout)
        if body is None:
            return None
        m = pickle.loads(body)
        m["data"] = torch.from_numpy(np.ascontiguousarray(m["data"])).to(self.device)
        if isinstance(m.get("label"), torch.Tensor):
            m["label"] = m["label"].to(self.device)
        return m

    def recv_gradient(self, timeout: float = 0.0):
        body = self.ch.basic_get(self.my_grad_q, timeout)
        if body is None:
            return None
        m = pickle.loads(body)
        m["data"] = torch.from_numpy(np.ascontiguousarray(m["data"])).to(self.device)
        return m
