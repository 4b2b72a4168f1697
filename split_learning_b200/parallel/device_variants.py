"""Device data plane for the sequential variants (Vanilla_SL, Cluster_FSL — reference other/Vanilla_SL, other/Cluster_FSL).

In these variants groups of first-stage clients train one after another while the later stages stay up for the whole round
(``algorithms/variants.SequentialServer``); the last stage neither knows who will talk to it next nor how many microbatches
are coming — it serves its queue until the server's PAUSE.  That is exactly what the ticket ring offers
(``parallel/ticket.py``): first-stage clients append tickets as they go, the last stage claims them until it has seen PAUSE
*and* the ring is empty.  Differences from ``DeviceRpcClient`` (main algorithm):

* the edge into the last stage is always a ticket ring;
* the last stage wires a lane *lazily*, when the first ticket of that lane arrives (the lane's client may not even have
  received its START when the round begins), and needs no batch-count announcements;
* parameters travel through the server as in the reference (sequential hand-off of the first-stage weights, FedAvg at the
  end of the round) — no device all-reduce, the groups are not alive at the same time.

Supported: two stages, ``local-round`` 1, no ``limited-time`` (anything else keeps the host data plane, decided identically
by every client from the shared config).  DCSL (per-device queues + SDA batch concatenation), FLEX and 2LS stay on the host
plane with the native executors.

STATUS: experimental and OFF by default (``b200.device-variants: true`` opts in).  The producer side and the ticket ring are
the verified code of the main algorithm; the lazily wired, PAUSE-terminated last stage below completed its first hardware
run only through the host-plane fallback and hung in the second (round-2 GPU budget exhausted before it could be debugged),
so ``tests/test_executor_gpu.py::test_sequential_variants_on_device_plane`` is opt-in too (``SLB200_TEST_DEVICE_VARIANTS=1``).
Without the flag the variants run exactly as before: host data plane, native executors."""
from __future__ import annotations

import time
import types
from typing import Dict, List

import torch

from ..algorithms.variants import SequentialClient
from ..log import print_with_color
from .device_client import DeviceRpcClient
from .mailbox import EdgeCounters, Mailbox, MailboxSpec
from .pipeline import DeviceStage

UNBOUNDED = 0xFFFFFFFF


class SequentialDeviceClient(DeviceRpcClient, SequentialClient):
    """MRO: device plane first; whenever it declines (``dstage is None``: unsupported topology / options) the call falls
    through to ``SequentialClient`` — the host-plane loops with local rounds, limited time and strict ordering."""

    VARIANT = "vanilla_sl"

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.opts["dynamic-consumers"] = True
        self.opts["device-fedavg"] = False               # hand-off / FedAvg through the server, as in the reference

    # ------------------------------------------------------------------ wiring
    def _wire(self, msg: dict) -> None:
        from ..train.b200_executor import B200Executor
        if not isinstance(self.executor, B200Executor) or self.num_layers != 2:
            raise RuntimeError("sequential variants: the device plane needs two native VGG stages")
        if int(msg.get("local_round", 1) or 1) != 1 or (msg.get("config_time") or {}).get("enable"):
            raise RuntimeError("sequential variants: local-round > 1 / limited-time run on the host plane")
        if self.layer_id == 1:
            return super()._wire(msg)                    # producer: outbox + gradient mailbox + ticket offers (base class)
        self._wire_last(msg)

    def _wire_last(self, msg: dict) -> None:
        from .ticket import TicketRing
        ex = self.executor
        B = int(self.learning["batch-size"])
        depth = int(self.learning.get("control-count", 3))
        lanes = self._lanes(msg)                         # every lane: (lane, producer id, None)
        dev = ex.device
        isz = 4 if ex.fp32 else 2
        self.channel.queue_declare(f"ipc_{self.client_id}")
        cache = self.__dict__.setdefault("_mailbox_cache", {})
        c, h, w = ex.in_shape
        spec_in = MailboxSpec(depth, B, (B, h, w, c), itemsize=isz)
        fwd_in: Dict[int, Mailbox] = {}
        for lane, _, _ in lanes:
            key = ("inbox", lane, spec_in.depth, spec_in.batch, spec_in.payload_shape, spec_in.itemsize)
            if key not in cache:
                cache[key] = Mailbox.allocate_exportable(spec_in, dev)
            mb = cache[key][0]
            mb.base[spec_in.flags_off: spec_in.flags_off + 4 * spec_in.depth].zero_()
            fwd_in[lane] = mb
        consumers = self._stage_ids[self.num_layers]
        self._outboxes: Dict[int, Mailbox] = {}
        grad_out: Dict[int, Mailbox] = {}
        self._handles: Dict[tuple, dict] = {}            # (kind, lane) -> handle message, filled as producers come up
        self._ring = None
        if consumers[0] == str(self.client_id):
            if "ring" not in cache:
                cache["ring"] = TicketRing.allocate(dev)
            self._ring = cache["ring"]
            self._ring.reset()
            for cid in self._stage_ids[self.num_layers - 1] + consumers[1:]:
                self.channel.publish_obj(f"ipc_{cid}", {"kind": "ring", "handle": self._ring.handle})
        else:
            t0 = time.monotonic()
            while self._ring is None:
                self._drain_handles(0.25)
                if time.monotonic() - t0 > self.watchdog:
                    raise TimeoutError("the ring owner never posted the ticket ring")
        multi = len(lanes) > 1
        if multi:
            ex.plan(B).bind_inputs([t for lane, _, _ in lanes for t in fwd_in[lane].payload])
        self.dstages: Dict[int, DeviceStage] = {}
        self._in_seq = {lane: EdgeCounters(depth, dev) for lane, _, _ in lanes}
        self._lane_info = lanes
        self._lane_index = {lane: k for k, (lane, _, _) in enumerate(lanes)}
        self._edges = (fwd_in, {}, {}, grad_out, multi, depth)
        self._tail_stages: Dict[tuple, DeviceStage] = {}
        import os
        self.dstage = types.SimpleNamespace(stream=ex.stream, B=B, depth=depth,
                                            wait_spins=int(os.environ.get("SLB200_WAIT_SPINS", str(1 << 28))))

    def _drain_handles(self, timeout: float) -> None:
        from .ticket import TicketRing
        m = self.channel.get_obj(f"ipc_{self.client_id}", timeout)
        while m is not None:
            if m["kind"] == "ring":
                self._ring = TicketRing.open(m["handle"], self.executor.device)
            else:
                self._handles[(m["kind"], int(m["lane"]))] = m
            m = self.channel.get_obj(f"ipc_{self.client_id}", 0.0)

    def _ensure_lane(self, lane: int) -> DeviceStage:
        """The lane's stage programs exist once its producer has come up (posted its outbox + gradient mailbox)."""
        if lane in self.dstages:
            return self.dstages[lane]
        fwd_in, _, _, grad_out, multi, depth = self._edges
        ex, B = self.executor, self.dstage.B
        t0 = time.monotonic()
        while ("outbox", lane) not in self._handles or ("grad", lane) not in self._handles:
            self._drain_handles(0.05)
            if time.monotonic() - t0 > self.watchdog:
                raise TimeoutError(f"lane {lane}: its producer never posted its mailboxes")
        isz = 4 if ex.fp32 else 2
        for kind, store in (("outbox", self._outboxes), ("grad", grad_out)):
            m = self._handles[(kind, lane)]
            store[lane] = Mailbox.open_peer(MailboxSpec(depth, B, tuple(m["shape"]), itemsize=isz), m["handle"], ex.device)
        k = self._lane_index[lane]
        self.dstages[lane] = DeviceStage(ex, B, depth, fwd_in=fwd_in[lane], grad_out=grad_out[lane],
                                         slot_offset=k * depth if multi else 0, bind_inputs=not multi)
        return self.dstages[lane]

    # ------------------------------------------------------------------ last stage: serve the ring until PAUSE
    def run_stage(self):
        if self.dstage is None or not self.is_last:
            return super().run_stage()
        from ..ops import native as N
        ex = self.executor
        depth = self.dstage.depth
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self._device_busy = True
        ev0.record(ex.stream)
        if not hasattr(self, "_loss_pool"):
            self._loss_pool = [torch.zeros(4).pin_memory() for _ in range(512)]
        loss_log: List[torch.Tensor] = []
        ahead = max(1, int(self.opts.get("claim-ahead", 2)))
        done: List[torch.cuda.Event] = []
        self.claimed: List[tuple] = []
        def paused() -> bool:
            # idle moment of the claim loop: wire lanes whose client has come up in the meantime (its START arrives mid-round)
            # before its first ticket — opening its gradient mailbox is also what creates the host gate a same-process
            # producer waits on before it enqueues a backward pass
            self._drain_handles(0.0)
            for lane, _, _ in self._lane_info:
                if lane not in self.dstages and ("outbox", lane) in self._handles and ("grad", lane) in self._handles:
                    self._ensure_lane(lane)
            return self.trainer._poll_pause(0.0)             # PAUSE (or STOP) is kept in trainer.pause_msg once seen
        while True:
            if len(done) >= ahead:
                done[len(done) - ahead].synchronize()
            got = self._ring.claim(UNBOUNDED, timeout=self.watchdog * 4, stop=paused)
            if got is None:
                break
            _, lane, it, gseq, b = got
            self._ensure_lane(lane)
            st = self.dstages[lane] if b == self.dstage.B else self._tail_stage(lane, b)
            slot = it % depth
            src, box = self._outboxes[lane], st.fwd_in
            with torch.cuda.stream(st.stream):
                N.memcpy_async(box.payload[slot].data_ptr(), src.payload[slot].data_ptr(), b * self._payload_bytes_per_sample(src))
                N.memcpy_async(box.labels[slot].data_ptr(), src.labels[slot].data_ptr(), b * 8)
                N.set_flag(box.flag_ptr(slot), 0, seq=self._in_seq[lane].at(slot))
                N.store_u32(st.seq_grad.at(slot).data_ptr(), (gseq - 1) & 0xFFFFFFFF)
            st.last(it)
            if self.opts.get("loss-every-step", True):
                hbuf = self._loss_pool[len(loss_log)] if len(loss_log) < len(self._loss_pool) else torch.zeros(4).pin_memory()
                with torch.cuda.stream(st.stream):
                    hbuf.copy_(ex.loss_buf, non_blocking=True)
                loss_log.append(hbuf)
            ev = torch.cuda.Event()
            ev.record(st.stream)
            done.append(ev)
            self.claimed.append((lane, it))
        ev1.record(ex.stream)
        ex.stream.synchronize()
        self._device_busy = False
        self.last_device_ms = ev0.elapsed_time(ev1)
        self.last_loss = (sum(float(hh[0]) for hh in loss_log) / len(loss_log)) if loss_log else None
        for st in list(self.dstages.values()) + list(self._tail_stages.values()):
            st.check()
        total = len(self.claimed)
        self._device_steps = int(self.__dict__.get("_device_steps", 0)) + total
        self.trainer._wait_pause()                            # already seen: returns at once
        print_with_color(f"[device plane] served {total} microbatches from the ticket ring", "green")
        return (not ex.nan_detected()), total


class VanillaSLDeviceClient(SequentialDeviceClient):
    VARIANT = "vanilla_sl"


class ClusterFSLDeviceClient(SequentialDeviceClient):
    VARIANT = "cluster_fsl"


DEVICE_CLIENTS = {"vanilla_sl": VanillaSLDeviceClient, "cluster_fsl": ClusterFSLDeviceClient}
