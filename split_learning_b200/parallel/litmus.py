"""Cross-GPU memory-ordering litmus test of the mailbox idiom (payload stores -> fence.sys -> st.release.sys flag on the
producer, ld.acquire.sys -> barrier -> payload loads on the consumer), run between two ranks over CUDA-IPC peer memory.
Collective over ``rank_a`` / ``rank_b`` (every other rank returns None)."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.distributed as dist

from ..ops import native as N
from .mailbox import alloc_exportable, open_exported


def pingpong(rank: int, rank_a: int, rank_b: int, device, iters: int = 100_000, n_words: int = 1024) -> Optional[dict]:
    world = dist.get_world_size()
    raw, handle, ptr = alloc_exportable(n_words * 4 + 4096, device)
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    if rank not in (rank_a, rank_b):
        return None
    peer = open_exported(handles[rank_b if rank == rank_a else rank_a], device)
    res = torch.zeros(4, dtype=torch.int32, device=device)
    flag_off = n_words * 4
    st = torch.cuda.current_stream(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    rc = N.lib().slb_litmus_pingpong(ctypes.c_void_p(peer), ctypes.c_void_p(peer + flag_off), ctypes.c_void_p(ptr),
                                     ctypes.c_void_p(ptr + flag_off), ctypes.c_int(n_words), ctypes.c_uint32(iters),
                                     ctypes.c_int(0 if rank == rank_a else 1), ctypes.c_uint64(1 << 24), ctypes.c_void_p(res.data_ptr()),
                                     ctypes.c_void_p(st.cuda_stream))
    N._check(rc, "litmus_pingpong")
    e1.record(st)
    st.synchronize()
    r = res.tolist()
    return {"iters": iters, "payload_words": n_words, "errors": int(r[0]), "timeout": bool(r[1]),
            "round_trip_us": e0.elapsed_time(e1) * 1e3 / iters}
