"""End-of-round FedAvg over NVLink peer memory.

Reference: every client pickles its state-dict to the server, the server averages on the CPU
and ships the result back inside START (src/Server.py:398-434, src/Utils.py:35-66).  Here the
replicas of a stage keep their flat fp32 parameter buffer on the GPU; each one copies it into
an IPC-exported staging buffer, then runs ONE fused kernel that pulls every replica's staging
buffer through NVLink peer loads and writes

        p[i] = sum_r coef_r * nan_to_num(p_r[i])            coef_r = w_r / sum(w)   (x 1/#clusters)

in place — fp32 master and bf16 shadow in the same pass — so the averaged parameters are
resident on every replica and the next round needs no START payload.  Integer state
(``num_batches_tracked``) is averaged-and-rounded on the host (13 scalars for VGG16).  A round
in which any replica saw a NaN loss is skipped, like the reference (src/Server.py:162-170).

``torch.distributed`` is used for the tiny control messages only (weights / NaN votes /
barriers around the staging copy).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from ..ops import native as N
from .mailbox import tensor_from_ptr


class PeerFedAvg:
    def __init__(self, n_elems: int, device, group_ranks: Sequence[int], group=None):
        """Collective over ``group_ranks`` (all must call).  ``n_elems`` fp32 elements, multiple of 4."""
        self.n = (n_elems + 3) // 4 * 4
        self.device = torch.device(device)
        self.ranks = list(group_ranks)
        self.group = group
        self.me = dist.get_rank()
        lib = N.lib()
        ptr = ctypes.c_void_p()
        rc = lib.slb_malloc(ctypes.byref(ptr), ctypes.c_longlong(self.n * 4))
        if rc != 0:
            raise N.NativeError(f"slb_malloc failed: {rc}")
        handle = (ctypes.c_uint8 * 64)()
        rc = lib.slb_ipc_get_handle(ptr, handle)
        if rc != 0:
            raise N.NativeError(f"cudaIpcGetMemHandle failed: {rc}")
        self.staging = tensor_from_ptr(ptr.value, self.n * 4, self.device).view(torch.float32)
        gathered: List[Optional[bytes]] = [None] * len(self.ranks)
        dist.all_gather_object(gathered, bytes(handle), group=group)
        self.peer_ptrs: List[int] = []
        for r, h in zip(self.ranks, gathered):
            if r == self.me:
                self.peer_ptrs.append(ptr.value)
                continue
            q = ctypes.c_void_p()
            buf = (ctypes.c_uint8 * 64).from_buffer_copy(h)
            rc = lib.slb_ipc_open(buf, ctypes.byref(q))
            if rc != 0:
                raise N.NativeError(f"cudaIpcOpenMemHandle failed: {rc}")
            self.peer_ptrs.append(q.value)

    def average(self, flat_fp32: torch.Tensor, flat_bf16: Optional[torch.Tensor], weight: float, ok: bool = True,
                cluster_scale: float = 1.0) -> bool:
        """In-place weighted average of ``flat_fp32`` across the group.  Returns False (and leaves the
        parameters untouched) when any member reports ``ok == False``."""
        info = torch.tensor([float(weight), 1.0 if ok else 0.0], device=self.device)
        allinfo = [torch.zeros_like(info) for _ in self.ranks]
        dist.all_gather(allinfo, info, group=self.group)
        ws = [float(t[0]) for t in allinfo]
        if not all(float(t[1]) > 0.5 for t in allinfo) or sum(ws) <= 0:
            return False
        n = flat_fp32.numel()
        assert n <= self.n
        self.staging[:n].copy_(flat_fp32)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)                       # every staging buffer is complete
        coefs = [w / sum(ws) * cluster_scale for w in ws]
        n4 = n // 4 * 4
        N.fedavg(flat_fp32, flat_bf16, self.peer_ptrs, coefs, n4)
        if n4 != n:                                          # tail (never happens with 128-aligned layouts)
            flat_fp32[n4:] = sum(c * torch.nan_to_num(tensor_from_ptr(p, self.n * 4, self.device).view(torch.float32)[n4:n])
                                 for c, p in zip(coefs, self.peer_ptrs))
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)                       # nobody overwrites staging before all have read
        return True


def average_int_state(tensors: Dict[str, torch.Tensor], weight: float, group=None) -> None:
    """``num_batches_tracked`` & co: weighted mean, rounded back to the integer dtype (src/Utils.py:59-60)."""
    if not tensors:
        return
    keys = sorted(tensors)
    dev = tensors[keys[0]].device
    vals = torch.stack([tensors[k].to(torch.float64) for k in keys]) * weight
    w = torch.tensor([weight], dtype=torch.float64, device=dev)
    dist.all_reduce(vals, group=group)
    dist.all_reduce(w, group=group)
    avg = (vals / w).round()
    for k, v in zip(keys, avg):
        tensors[k].copy_(v.to(tensors[k].dtype))


def local_fedavg(flats: Sequence[torch.Tensor], weights: Sequence[float], out: torch.Tensor,
                 out_bf16: Optional[torch.Tensor] = None) -> None:
    """Same kernel with all sources on one device (server-side aggregation of uploaded stages)."""
    tot = float(sum(weights))
    N.fedavg(out, out_bf16, [f.data_ptr() for f in flats], [w / tot for w in weights], out.numel() // 4 * 4)
