"""Cut-edge mailboxes in consumer HBM (the device data plane).

One ``Mailbox`` = a ring of ``depth`` (= control-count) slots that lives in the *consumer's*
memory.  The producer's final kernel of a pass stores the payload tiles straight into the
slot — a local pointer when both stages share a GPU, an NVLink-peer pointer (CUDA IPC
mapping) when they do not — and its last CTA publishes ``flag[slot] = seq`` with
``st.release.sys``; the consumer's graph starts with a one-thread ``ld.acquire.sys`` wait.
No host code, NCCL call, pickle or D2H/H2D copy touches activations or gradients
(reference path being replaced: src/train/VGG16.py:20-53).

Slot reuse is safe without "empty" acknowledgements because at most ``depth`` microbatches
are in flight (control-count): slot ``s`` is rewritten for microbatch ``i + depth`` only
after the gradient of microbatch ``i`` came back.

Layout of one exported allocation (all offsets 256-byte aligned):
    payload[depth][bytes_per_slot] | labels[depth][B] int64 | flags[depth] u32 | header[16] u32
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from ..ops import native as N


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class _RawCuda:
    """Expose a raw device pointer to torch through ``__cuda_array_interface__``."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def tensor_from_ptr(ptr: int, nbytes: int, device) -> torch.Tensor:
    return torch.as_tensor(_RawCuda(ptr, nbytes), device=device)


class DevPtr:
    """A typed view of raw device memory that is NOT a torch tensor (peer-mapped IPC memory must never be touched
    by torch ops — they would run on, or copy from, the exporting device).  Kernels only need ``data_ptr()``."""

    def __init__(self, ptr: int, shape, itemsize: int):
        self._ptr, self.shape, self.itemsize = int(ptr), tuple(shape), itemsize

    def data_ptr(self) -> int:
        return self._ptr

    def numel(self) -> int:
        n = 1
        for d in self.shape:
            n *= d
        return n

    def element_size(self) -> int:
        return self.itemsize

    def reshape(self, *shape):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape
        if -1 in shape:
            known = 1
            for d in shape:
                if d != -1:
                    known *= d
            shape = tuple(self.numel() // known if d == -1 else d for d in shape)
        return DevPtr(self._ptr, shape, self.itemsize)


@dataclass
class MailboxSpec:
    depth: int
    batch: int
    payload_shape: Tuple[int, ...]      # per-microbatch tensor shape
    with_labels: bool = True
    itemsize: int = 4                   # 4 = fp32 payload (tf32 parity mode, the reference's wire format), 2 = bf16

    @property
    def dtype(self):
        return torch.float32 if self.itemsize == 4 else torch.bfloat16

    @property
    def payload_bytes(self) -> int:
        n = self.itemsize
        for d in self.payload_shape:
            n *= d
        return _align(n)

    @property
    def labels_off(self) -> int:
        return self.depth * self.payload_bytes

    @property
    def flags_off(self) -> int:
        return self.labels_off + _align(self.depth * self.batch * 8)

    @property
    def header_off(self) -> int:
        return self.flags_off + 256

    @property
    def total_bytes(self) -> int:
        return self.header_off + 256


_SAME_PROCESS = {}          # IPC handle -> (raw pointer, owner Mailbox), for partners that live in this very process
_GATE_LOCK = __import__("threading").Lock()   # several same-process peers may open one mailbox at the same time (competing
#                                               consumers): exactly one HostGate per mailbox, or posts get lost


class HostGate:
    """Host-side companion of a mailbox when both ends are threads of ONE process sharing a GPU: the consumer thread
    does not enqueue its program before the producer thread has *enqueued* the matching publish.  This keeps flag-wait
    kernels from spinning on work that is not even submitted yet — which on a shared GPU can dead-lock against CUDA's
    lazy kernel loading (a first launch may wait for running kernels).  Cross-process edges need no gate."""

    def __init__(self):
        import threading
        self._cv = threading.Condition()
        self._n = 0

    def post(self) -> None:
        with self._cv:
            self._n += 1
            self._cv.notify_all()

    def reset(self) -> None:
        """New round (owner side, while every role is parked between UPDATE and SYN): waits count from zero again.  Without
        it the previous round's posts satisfy every wait of the next round at once and the gate stops gating — harmless
        while every program of the round is already captured, but with competing consumers a replica may have to capture a
        (lane, slot) program it has not served before, i.e. instantiate a CUDA graph while the producer's wait kernel for
        exactly that program is already spinning on the shared GPU."""
        with self._cv:
            self._n = 0

    def wait(self, count: int, timeout: float = 120.0) -> None:
        with self._cv:
            if not self._cv.wait_for(lambda: self._n >= count, timeout):
                raise TimeoutError("producer thread never enqueued the matching publish")


_OPENED: dict = {}         # IPC handle -> mapped pointer (cudaIpcOpenMemHandle may be called once per handle and process)


def alloc_exportable(nbytes: int, device) -> Tuple[torch.Tensor, bytes, int]:
    """cudaMalloc'ed (not caching-allocator) zeroed memory + its 64-byte CUDA-IPC handle: (uint8 tensor view, handle, ptr)."""
    lib = N.lib()
    nbytes = _align(max(int(nbytes), 256))
    with torch.cuda.device(device):
        ptr = ctypes.c_void_p()
        rc = lib.slb_malloc(ctypes.byref(ptr), ctypes.c_longlong(nbytes))
        if rc != 0:
            raise N.NativeError(f"slb_malloc({nbytes}) failed: {rc}")
        handle = (ctypes.c_uint8 * 64)()
        rc = lib.slb_ipc_get_handle(ptr, handle)
        if rc != 0:
            raise N.NativeError(f"cudaIpcGetMemHandle failed: {rc}")
    _SAME_PROCESS[bytes(handle)] = (ptr.value, None)
    return tensor_from_ptr(ptr.value, nbytes, device), bytes(handle), ptr.value


def open_exported(handle: bytes, device) -> int:
    """Device pointer of a peer's exported allocation (the raw pointer itself when the exporter lives in this process)."""
    if handle in _SAME_PROCESS:
        return _SAME_PROCESS[handle][0]
    if handle in _OPENED:
        return _OPENED[handle]
    lib = N.lib()
    with torch.cuda.device(device):
        q = ctypes.c_void_p()
        buf = (ctypes.c_uint8 * 64).from_buffer_copy(handle)
        rc = lib.slb_ipc_open(buf, ctypes.byref(q))
        if rc != 0:
            raise N.NativeError(f"cudaIpcOpenMemHandle failed: {rc}")
    _OPENED[handle] = q.value
    return q.value


class Mailbox:
    """View of a mailbox allocation (owner side or a peer-mapped producer side)."""

    def __init__(self, spec: MailboxSpec, base: Optional[torch.Tensor], owner: bool, raw_ptr: Optional[int] = None,
                 gate: Optional[HostGate] = None):
        self.spec, self.base, self.owner = spec, base, owner
        self.gate = gate
        self.raw_ptr = raw_ptr if raw_ptr is not None else base.data_ptr()
        d = spec.depth
        if base is None:                     # peer-mapped: raw pointer views only
            self.payload = [DevPtr(self.raw_ptr + s * spec.payload_bytes, spec.payload_shape, spec.itemsize) for s in range(d)]
            self.labels = [DevPtr(self.raw_ptr + spec.labels_off + s * spec.batch * 8, (spec.batch,), 8) for s in range(d)]
            self.flags = None
            self.header = None
            return
        self.payload: List[torch.Tensor] = []
        for s in range(d):
            nb = spec.itemsize
            for x in spec.payload_shape:
                nb *= x
            t = base[s * spec.payload_bytes: s * spec.payload_bytes + nb].view(spec.dtype).view(spec.payload_shape)
            self.payload.append(t)
        lab = base[spec.labels_off: spec.labels_off + d * spec.batch * 8].view(torch.int64).view(d, spec.batch)
        self.labels = [lab[s] for s in range(d)]
        self.flags = base[spec.flags_off: spec.flags_off + 4 * d].view(torch.int32)
        self.header = base[spec.header_off: spec.header_off + 64].view(torch.int32)

    def flag_ptr(self, slot: int) -> int:
        return self.raw_ptr + self.spec.flags_off + 4 * slot

    def view(self, batch: int) -> "Mailbox":
        """The same ring seen with a smaller microbatch (the trailing partial batch of an epoch): payload and label slots
        are prefixes of the full slots (batch is the outermost dimension), flags / gate / sequence numbers are shared."""
        if batch == self.spec.batch:
            return self
        assert 0 < batch < self.spec.batch
        v = object.__new__(Mailbox)
        v.spec, v.base, v.owner, v.gate, v.raw_ptr = self.spec, self.base, self.owner, self.gate, self.raw_ptr
        shape = (batch,) + tuple(self.spec.payload_shape[1:])
        if self.base is None:
            v.payload = [DevPtr(p.data_ptr(), shape, self.spec.itemsize) for p in self.payload]
            v.labels = [DevPtr(l.data_ptr(), (batch,), 8) for l in self.labels]
        else:
            v.payload = [p[:batch] for p in self.payload]
            v.labels = [l[:batch] for l in self.labels]
        v.flags, v.header = self.flags, self.header
        return v

    # ---- allocation / export --------------------------------------------------------
    @staticmethod
    def allocate_local(spec: MailboxSpec, device) -> "Mailbox":
        base = torch.zeros(spec.total_bytes, dtype=torch.uint8, device=device)
        return Mailbox(spec, base, owner=True)

    @staticmethod
    def allocate_exportable(spec: MailboxSpec, device) -> Tuple["Mailbox", bytes]:
        """cudaMalloc'ed (not caching-allocator) memory + its 64-byte IPC handle."""
        lib = N.lib()
        ptr = ctypes.c_void_p()
        rc = lib.slb_malloc(ctypes.byref(ptr), ctypes.c_longlong(spec.total_bytes))
        if rc != 0:
            raise N.NativeError(f"slb_malloc failed: {rc}")
        handle = (ctypes.c_uint8 * 64)()
        rc = lib.slb_ipc_get_handle(ptr, handle)
        if rc != 0:
            raise N.NativeError(f"cudaIpcGetMemHandle failed: {rc}")
        base = tensor_from_ptr(ptr.value, spec.total_bytes, device)
        mb = Mailbox(spec, base, owner=True, raw_ptr=ptr.value)
        _SAME_PROCESS[bytes(handle)] = (ptr.value, mb)
        return mb, bytes(handle)

    @staticmethod
    def open_peer(spec: MailboxSpec, handle: bytes, device) -> "Mailbox":
        if handle in _SAME_PROCESS:            # cudaIpcOpenMemHandle refuses handles exported by the same process
            raw, owner_mb = _SAME_PROCESS[handle]
            with _GATE_LOCK:
                if owner_mb.gate is None:
                    owner_mb.gate = HostGate()
                gate = owner_mb.gate
            return Mailbox(spec, None, owner=False, raw_ptr=raw, gate=gate)
        return Mailbox(spec, None, owner=False, raw_ptr=open_exported(handle, device))     # one mapping per handle and process


class EdgeCounters:
    """Device-resident sequence counters of one end of an edge: ``seq[s]`` (next value the
    producer publishes for slot s) or ``expect[s]`` (last value the consumer has consumed)."""

    def __init__(self, depth: int, device):
        self.ctr = torch.zeros(max(depth, 1) * 4, dtype=torch.int32, device=device)

    def at(self, slot: int) -> torch.Tensor:
        return self.ctr[slot * 4: slot * 4 + 1]

    def reset(self) -> None:
        self.ctr.zero_()
