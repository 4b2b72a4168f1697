"""Ticket ring: the device-side shared queue of a cut edge whose consumers compete for work.

Reference behaviour being reproduced (src/train/VGG16.py:143-154 and :40-53): all replicas of stage i+1 in a cluster
``basic_get`` the same ``intermediate_queue_{i+1}_{cluster}``, so a free replica takes the next activation whoever produced
it, and the gradient goes back to ``trace[-1]``.  On the device plane the activation stays in the producer's *outbox* (its
own HBM, exported over CUDA IPC); what is queued is a 32-byte ticket in a ring that lives in the memory of the edge's first
consumer (``ops/csrc/ticket.cu``): producers append with ``atom.add.sys`` on ``tail`` + a release store of the entry's
sequence word, consumers claim with ``atom.add.sys`` on ``head`` and an acquire spin on that word.  The claim lands in
mapped pinned host memory, because the host must know the origin to enqueue the matching program (copy-in from that origin's
outbox, gradient store + flag into that origin's gradient mailbox) and to learn that the round is drained.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from ..ops import native as N
from .mailbox import alloc_exportable, open_exported

ENTRIES = 1024


class TicketRing:
    def __init__(self, ptr: int, device, base: Optional[torch.Tensor] = None, handle: Optional[bytes] = None):
        self.ptr, self.device, self.base, self.handle = int(ptr), torch.device(device), base, handle
        self.entries = ENTRIES
        self._out: Optional[torch.Tensor] = None
        self._claim_stream: Optional[torch.cuda.Stream] = None

    # ---- construction -------------------------------------------------------------------
    @staticmethod
    def allocate(device) -> "TicketRing":
        base, handle, ptr = alloc_exportable(N.ticket_ring_bytes(ENTRIES), device)
        return TicketRing(ptr, device, base=base, handle=handle)

    @staticmethod
    def open(handle: bytes, device) -> "TicketRing":
        return TicketRing(open_exported(handle, device), device)

    def reset(self) -> None:
        """Owner only, between rounds (every participant is parked between UPDATE and the next SYN)."""
        assert self.base is not None, "only the owner resets the ring"
        self.base.zero_()
        torch.cuda.current_stream(self.device).synchronize()

    # ---- producer -------------------------------------------------------------------------
    def publish(self, origin: int, it: int, gseq_ctr: Optional[torch.Tensor], batch: int) -> None:
        """Stream-ordered (current stream): runs after the pass that filled outbox slot ``it % depth``.  ``gseq_ctr``: the
        producer's device counter of gradients consumed from that slot; the ticket carries ``*gseq_ctr + 1``."""
        N.ticket_publish(self.ptr, self.entries, origin, it, gseq_ctr, batch)

    # ---- consumer -------------------------------------------------------------------------
    def claim(self, total: int, max_spins: int = 1 << 22, timeout: float = 120.0,
              alive=None, stop=None) -> Optional[Tuple[int, int, int, int, int]]:
        """Blocks (on the host) until this replica owns the next ticket: (ticket, origin, it, gseq, batch), or None when all
        ``total`` tickets of the round have been handed out.  Each attempt is one tiny kernel on a private stream that never
        waits for a ticket to appear; while the ring is empty the host backs off (20 us .. 1 ms).  ``alive``: optional
        callable polled while waiting — returning False aborts (the run was stopped).  ``stop``: optional callable for
        producers whose number of tickets is not known up front (sequential variants: ``total`` = 2^32 - 1): once it
        returns True *and* the ring is empty, the queue is drained (the caller saw PAUSE: every producer has finished and
        published all its tickets)."""
        import time
        if self._out is None:
            with torch.cuda.device(self.device):
                self._out = torch.zeros(8, dtype=torch.int32).pin_memory()
                self._claim_stream = N.new_stream(self.device)
        t0 = time.monotonic()
        pause = 2e-5
        while True:
            self._out.zero_()
            with torch.cuda.stream(self._claim_stream):
                N.ticket_claim(self.ptr, self.entries, total, max_spins, self._out.data_ptr())
            self._claim_stream.synchronize()
            status, ticket, origin, it, gseq, batch = (int(v) & 0xFFFFFFFF for v in self._out[:6].tolist())
            if status == 1:
                return ticket, origin, it, gseq, batch
            if status == 2:
                return None
            if status == 4:                                   # nothing published yet
                if stop is not None and stop():
                    return None
                if time.monotonic() - t0 > timeout or (alive is not None and not alive()):
                    raise TimeoutError(f"ticket ring: no ticket for {timeout:.0f}s (producers dead?)")
                time.sleep(pause)
                pause = min(pause * 1.5, 1e-3)
                continue
            raise TimeoutError(f"ticket ring: ticket {ticket} was allocated but never published (producer dead?)")

    def abort(self) -> None:
        """Owner only: wake every spinning claim (a role failed)."""
        if self.base is not None:                           # torch ops only ever touch memory this process owns
            self.base.view(torch.int32)[2] = 1
