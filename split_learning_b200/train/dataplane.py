"""Host data plane: activation / gradient *messages* through the broker.

This is the general path (any topology, any variant, any executor, CPU or GPU): same message
fields as the reference (``data_id, data, label, trace`` — src/train/VGG16.py:20-53) and the
queue grammar of every variant.  The payload travels as
  * a pickled CPU array (default; heterogeneous devices, the reference's format), or
  * with ``b200.wire: cuda`` — stays on the GPU: the message carries only a handle (the tensor
    object for a consumer thread of this process, a CUDA-IPC handle made by torch's reducer for a
    consumer process on this box) and the consumer pulls the bytes GPU->GPU over NVLink.  This is
    what the token models (KWT / ViT / BERT) and odd topologies use between GPUs.
For VGG-family stages with dividing replica counts the *device* data plane (``parallel/``) replaces
all of this: tiles are stored into the consumer's HBM from inside the producing kernel.
"""
from __future__ import annotations

import io
import itertools
import os
import pickle
import threading
from typing import Dict, List, Optional

import numpy as np
import torch

from ..transport import Channel, codec


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


_LOCAL: Dict[int, tuple] = {}            # same-process hand-off: token -> (tensor, ready event)
_LOCAL_LOCK = threading.Lock()
_TOKENS = itertools.count(1)


def _to_numpy(t: torch.Tensor) -> np.ndarray:
    t = t.detach()
    if t.dtype == torch.bfloat16:
        t = t.float()
    return t.cpu().numpy()


class HostDataPlane:
    sent = {"host": 0, "cuda_local": 0, "cuda_ipc": 0}     # payload encodings used by this process (tests / logs)

    def __init__(self, channel: Channel, client_id, layer_id: int, cluster=None, grammar: Optional[QueueGrammar] = None,
                 device="cpu", wire: str = "host"):
        self.ch, self.client_id, self.layer_id, self.cluster = channel, client_id, layer_id, cluster
        self.grammar = grammar or QueueGrammar("main")
        self.device = device
        self.cuda_wire = wire == "cuda" and torch.device(device).type == "cuda"
        self.my_grad_q = self.grammar.gradient_queue(layer_id, client_id)
        self.ch.queue_declare(self.my_grad_q)

    # ---- payload encoding ------------------------------------------------
    def _pack(self, t: torch.Tensor):
        if not (self.cuda_wire and t.is_cuda):
            HostDataPlane.sent["host"] += 1
            return _to_numpy(t)
        from ..transport.broker import InProcBroker
        t = t.detach()
        if isinstance(self.ch, InProcBroker):           # consumer is a thread of this process: hand the tensor over
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(t.device))
            tok = next(_TOKENS)
            with _LOCAL_LOCK:
                _LOCAL[tok] = (t, ev)
            HostDataPlane.sent["cuda_local"] += 1
            return {"__cuda_local__": tok, "pid": os.getpid()}
        # consumer is another process on this box: CUDA IPC handle (torch's reducer keeps the block alive until the
        # consumer has released it); producer-side kernels must be complete before the handle is usable
        from multiprocessing.reduction import ForkingPickler
        import torch.multiprocessing as _tmp  # noqa: F401  (registers the CUDA tensor reducers)
        torch.cuda.current_stream(t.device).synchronize()
        buf = io.BytesIO()
        ForkingPickler(buf, pickle.HIGHEST_PROTOCOL).dump(t.contiguous())
        if HostDataPlane.sent["cuda_ipc"] == 0:
            print(f"[data plane] client {self.client_id}: payloads stay on the GPU (CUDA IPC handles through the broker)", flush=True)
        HostDataPlane.sent["cuda_ipc"] += 1
        return {"__cuda_ipc__": buf.getvalue()}

    def _unpack(self, d) -> torch.Tensor:
        if isinstance(d, dict) and "__cuda_local__" in d:
            if d["pid"] != os.getpid():
                raise RuntimeError("cuda wire: in-process handle received by another process")
            with _LOCAL_LOCK:
                t, ev = _LOCAL.pop(d["__cuda_local__"])
            torch.cuda.current_stream(torch.device(self.device)).wait_event(ev)
            return t.to(self.device, non_blocking=True)
        if isinstance(d, dict) and "__cuda_ipc__" in d:
            src = codec.loads_cuda_ipc(d["__cuda_ipc__"])  # maps the producer's block (peer memory if another GPU)
            out = src.to(self.device, non_blocking=False)
            if out.data_ptr() == src.data_ptr():           # same GPU: detach from the producer's allocation
                out = src.clone()
            del src
            return out
        return torch.from_numpy(np.ascontiguousarray(d)).to(self.device)

    # ---- producer side -------------------------------------------------
    def send_forward(self, data_id, output: torch.Tensor, labels, trace: Optional[List] = None, target=None) -> None:
        trace = list(trace) + [self.client_id] if trace else [self.client_id]
        q = self.grammar.forward_queue(self.layer_id, self.cluster, target)
        lab = labels.detach().cpu() if isinstance(labels, torch.Tensor) else labels
        self.ch.basic_publish(q, codec.dumps({"data_id": data_id, "data": self._pack(output), "label": lab, "trace": trace}))

    def send_gradient(self, data_id, gradient: torch.Tensor, trace: List) -> None:
        trace = list(trace)
        to_client = trace.pop(-1)
        q = self.grammar.gradient_queue(self.layer_id - 1, to_client)
        self.ch.basic_publish(q, codec.dumps({"data_id": data_id, "data": self._pack(gradient), "trace": trace}))

    # ---- consumer side -------------------------------------------------
    def recv_forward(self, timeout: float = 0.0, source=None):
        q = self.grammar.forward_queue(self.layer_id - 1, self.cluster, source)
        body = self.ch.basic_get(q, timeout)
        if body is None:
            return None
        m = codec.loads(body)
        m["data"] = self._unpack(m["data"])
        if isinstance(m.get("label"), torch.Tensor):
            m["label"] = m["label"].to(self.device)
        return m

    def recv_gradient(self, timeout: float = 0.0):
        body = self.ch.basic_get(self.my_grad_q, timeout)
        if body is None:
            return None
        m = codec.loads(body)
        m["data"] = self._unpack(m["data"])
        return m
