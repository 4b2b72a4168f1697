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


def _rebuild_tensor(raw, dtype: str, shape) -> torch.Tensor:
    if 0 in tuple(shape):
        return torch.empty(tuple(shape), dtype=getattr(torch, dtype))
    if not isinstance(raw, np.ndarray):                  # out-of-band buffer: a view of the receive buffer (zero copy)
        mv = memoryview(raw)
        raw = np.frombuffer(mv, dtype=np.uint8) if not mv.readonly else np.frombuffer(mv, dtype=np.uint8).copy()
    elif not raw.flags.writeable:                        # small in-band array rebuilt from immutable bytes
        raw = raw.copy()
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


# Large payloads (state-dicts: 134 MB for VGG16) travel as *segments*: a small pickle plus the tensors' memory as
# out-of-band buffers (pickle protocol 5).  The sender hands the segments to ``socket.sendall`` one by one and the receiver
# rebuilds tensors as views of the receive buffer — no byte of tensor data is copied while holding the GIL, so a checkpoint
# upload running in a background thread does not stall the threads that are launching the next round's kernels.
_MAGIC = b"SLB5"
_SEG_MIN = 1 << 16           # tensors smaller than this stay in-band


class _OobPickler(_Pickler):
    def reducer_override(self, obj):
        r = super().reducer_override(obj)
        if r is not NotImplemented:
            fn, (raw, dtype, shape) = r
            # small tensors stay in-band as bytes (an ndarray would also be sent out of band under protocol 5)
            return fn, (pickle.PickleBuffer(raw) if raw.nbytes >= _SEG_MIN else raw.tobytes(), dtype, shape)
        return r


def dump_segments(obj: Any) -> list:
    """[bytes-like, ...] whose concatenation ``loads`` understands; one element (a plain pickle) when nothing is large."""
    import struct
    buf = io.BytesIO()
    oob: list = []
    _OobPickler(buf, protocol=5, buffer_callback=oob.append).dump(obj)
    if not oob:
        return [buf.getvalue()]
    views = [b.raw() for b in oob]
    pick = buf.getvalue()
    head = _MAGIC + struct.pack(f"<IQ{len(views)}Q", len(views), len(pick), *[v.nbytes for v in views])
    return [head + pick] + views


def _loads_segments(data, unpickler) -> Any:
    import struct
    mv = memoryview(data)
    n, plen = struct.unpack_from("<IQ", mv, 4)
    lens = struct.unpack_from(f"<{n}Q", mv, 16)
    off = 16 + 8 * n
    pick = mv[off: off + plen]
    off += plen
    bufs = []
    for ln in lens:
        bufs.append(mv[off: off + ln])
        off += ln
    if off != len(mv):
        raise UnsafePayload("segmented payload: lengths do not add up")
    return unpickler(io.BytesIO(pick), buffers=bufs).load()


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


def loads(data) -> Any:
    if bytes(data[:4]) == _MAGIC:
        return _loads_segments(data, _Unpickler)
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


def loads_cuda_ipc(data) -> Any:
    if bytes(data[:4]) == _MAGIC:
        return _loads_segments(data, _IpcUnpickler)
    return _IpcUnpickler(io.BytesIO(data)).load()
