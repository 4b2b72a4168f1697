"""Wire codec of the control plane and the host data plane.

The reference ships raw ``pickle`` over RabbitMQ (src/RpcClient.py:138-146, src/train/VGG16.py:26-38) and relies on the
broker's credentials for safety.  Here every message is still a pickled dict with the reference's field names, but

* ``dumps`` writes tensors as (raw little-endian bytes, dtype name, shape) — never through ``torch.storage._load_from_bytes``
  (which is a nested, unrestricted ``torch.load``) — and
* ``loads`` is a *restricted* unpickler: the only globals it resolves are containers, ``uuid.UUID``, NumPy array / dtype
  reconstruction and this module's tensor rebuild function.  Anything else (``os.system``, ``subprocess.Popen``, ``eval`` ...)
  raises ``UnsafePayload``, so a peer that can reach the broker port cannot execute code in a server or client process.

``loads_cuda_ipc`` is the one wider gate: it additionally resolves the handful of torch functions that torch's own CUDA-IPC
reducer emits (``b200.wire: cuda``); the payload still cannot name anything outside that list.
"""
from __future__ import annotations

import collections
import io
import pickle
import uuid
from typing import Any

import numpy as np
import torch


class UnsafePayload(pickle.UnpicklingError):
    pass


def _rebuild_tensor(raw: np.ndarray, dtype: str, shape) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(raw)).view(getattr(torch, dtype))
    return t.view(tuple(shape))


class _Pickler(pickle.Pickler):
    def reducer_override(self, obj):
        if isinstance(obj, torch.Tensor):
            t = obj.detach()
            if t.is_cuda:
                t = t.cpu()
            t = t.contiguous()
            raw = t.reshape(-1).view(torch.uint8).numpy() if t.numel() else np.zeros(0, dtype=np.uint8)
            return _rebuild_tensor, (raw, str(t.dtype).split(".")[-1], tuple(t.shape))
        return NotImplemented


def dumps(obj: Any) -> bytes:
    buf = io.BytesIO()
    _Pickler(buf, protocol=pickle.HIGHEST_PROTOCOL).dump(obj)
    return buf.getvalue()


_NUMPY_MODS = ("numpy.core.multiarray", "numpy._core.multiarray", "numpy.core.numeric", "numpy._core.numeric", "numpy")
_NUMPY_NAMES = {"_reconstruct", "scalar", "_frombuffer", "ndarray", "dtype"}
_BUILTINS = {"set", "frozenset", "slice", "complex", "range", "bytearray", "bytes", "dict", "list", "tuple", "int", "float",
             "bool", "str"}


def _safe_class(module: str, name: str):
    if module == "builtins" and name in _BUILTINS:
        return getattr(__import__("builtins"), name)
    if module == "collections" and name == "OrderedDict":
        return collections.OrderedDict
    if module == "uuid" and name == "UUID":
        return uuid.UUID
    if module in _NUMPY_MODS and name in _NUMPY_NAMES:
        mod = __import__(module, fromlist=[name])
        return getattr(mod, name)
    if module == "numpy.dtypes" and name.endswith("DType"):
        import numpy.dtypes as nd
        return getattr(nd, name)
    if module == __name__ and name == "_rebuild_tensor":
        return _rebuild_tensor
    if module == "torch" and name == "Size":
        return torch.Size
    return None


class _Unpickler(pickle.Unpickler):
    extra = None

    def find_class(self, module, name):
        c = _safe_class(module, name)
        if c is None and self.extra is not None:
            c = self.extra(module, name)
        if c is None:
            raise UnsafePayload(f"refusing to unpickle global {module}.{name}")
        return c


def loads(data: bytes) -> Any:
    return _Unpickler(io.BytesIO(data)).load()


def _cuda_ipc_class(module: str, name: str):
    """Globals emitted by ``torch.multiprocessing.reductions`` for a CUDA tensor handle (and nothing else)."""
    if module == "torch.multiprocessing.reductions" and name in ("rebuild_cuda_tensor", "rebuild_tensor", "rebuild_storage_empty"):
        import torch.multiprocessing.reductions as r
        return getattr(r, name)
    if module == "torch" and (name in ("Tensor", "device", "Size") or isinstance(getattr(torch, name, None), torch.dtype)):
        return getattr(torch, name)
    if module == "torch.storage" and name in ("UntypedStorage", "TypedStorage"):
        import torch.storage as s
        return getattr(s, name)
    if module == "torch.nn.parameter" and name == "Parameter":
        return torch.nn.Parameter
    return None


class _IpcUnpickler(_Unpickler):
    extra = staticmethod(_cuda_ipc_class)


def loads_cuda_ipc(data: bytes) -> Any:
    return _IpcUnpickler(io.BytesIO(data)).load()
