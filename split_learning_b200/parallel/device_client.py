"""Client FSM with the device data plane: the same START / READY / SYN / NOTIFY / PAUSE / UPDATE
conversation as ``client.RpcClient`` (the user-facing API), but activations and gradients move
through peer-memory mailboxes written by the stage kernels instead of broker queues.

Wiring happens while handling START: every client allocates the mailboxes it *consumes*
(activations from upstream, gradients from downstream), exports them with CUDA IPC and posts
the 64-byte handles to its chain partner's ``ipc_{client_id}`` queue; partners are derived
from the START ``peers`` table.  Every first-stage client defines a *lane*; lane i is served at
stage s by member ``i % n_s`` — the static counterpart of the reference's competing-consumer
queue — so many-clients-few-servers topologies ([4,2], [2,1], [4,2,1]) run on the device plane
with the downstream stage multiplexing its lanes on one executor (fan-in).  If a stage does not
divide its predecessor, or the stage has no native plan, the client keeps the host data plane —
a *topology* fallback (logged), never a kernel fallback.

Competing consumers (``b200.dynamic-consumers: true``; this is also what keeps a last stage that does not divide its
predecessor, e.g. clients [4, 3], on the device plane): the edge into the last stage becomes a *ticket ring*
(``parallel/ticket.py``).  Producers
keep each microbatch in their own outbox and append a ticket; every last-stage replica claims the next ticket whenever it
has fewer than ``b200.claim-ahead`` programs in flight, copies the payload in over NVLink and returns the gradient to the
ticket's origin — the reference's shared ``intermediate_queue`` + ``trace`` routing (src/train/VGG16.py:40-53,143-154)
without a broker hop, and a slow replica simply claims fewer tickets.

A trailing partial batch of the loader (``num-sample`` not a multiple of the batch size) is not dropped: it runs
through a second program set compiled for its size on the *same* mailboxes (prefix views of the slots, shared sequence
counters), so microbatch counts — the FedAvg weights, src/train/VGG16.py:109 — match the reference's.
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional

import torch

from .. import messages as M
from ..client import RpcClient
from ..log import print_with_color
from .mailbox import Mailbox, MailboxSpec
from .pipeline import DeviceStage


class DeviceRpcClient(RpcClient):
    def on_start(self, msg: dict) -> None:
        t_start = time.perf_counter()
        self.timing: Dict[str, float] = {}
        self.dstage: Optional[DeviceStage] = None
        super_ready = self.send_to_server
        sent: List[dict] = []
        self.send_to_server = lambda m: sent.append(m)       # hold READY until the mailboxes are wired
        try:
            super().on_start(msg)
        finally:
            self.send_to_server = super_ready
        try:
            # resident round (same executor, same peers): the stages, graphs and mailboxes of the previous round stay
            key = (id(self.executor), repr(msg.get("peers")), int(self.learning["batch-size"]), int(self.learning.get("control-count", 3)))
            if getattr(self, "_wired_key", None) == key and getattr(self, "dstages", None):
                self.dstage = self.dstages[self._lane_info[0][0]]
                for st in self.dstages.values():
                    st._posted = {"F": 0, "B": 0, "L": 0}
                ring = getattr(self, "_ring", None)
                if ring is not None and ring.base is not None:
                    ring.reset()                             # I own the edge's ticket ring: a fresh queue for this round
            else:
                self._wired_key = None
                self._wire(msg)
                self._wired_key = key
        except Exception as e:          # topology not supported on the device plane → host plane
            print_with_color(f"[device plane] falling back to host data plane: {e}", "yellow")
            self.dstage = None
        # competing consumers: the host gates of the mailboxes I own count per round (same-process partners only; see
        # mailbox.HostGate.reset).  Static lanes keep the free-running later rounds: all their programs exist after round 1.
        for entry in (self.__dict__.get("_mailbox_cache", {}).values() if self.__dict__.get("_dynamic") else ()):
            mb = entry[0] if isinstance(entry, tuple) else None
            if mb is not None and getattr(mb, "gate", None) is not None:
                mb.gate.reset()
        self.timing["on_start"] = (time.perf_counter() - t_start) * 1e3
        for m in sent:
            self.send_to_server(m)

    PIN_RING = 32

    def _pinned(self, it: int, x: torch.Tensor, y: torch.Tensor):
        """Page-locked staging for a microbatch that is not pinned yet: a ring of ``PIN_RING`` slots (bounded, whatever the
        length of the epoch).  A slot is reused only after the H2D copy that read it has completed (event recorded by
        ``_pin_mark`` right after the copy was enqueued).  Loaders that already yield pinned tensors pass through."""
        self._pin_last = None
        if x.is_cuda or (x.is_pinned() and y.is_pinned() and x.dtype == torch.float32 and y.dtype == torch.long):
            return x, y
        pool = self.__dict__.setdefault("_pin_pool", {})
        key = (it % self.PIN_RING, tuple(x.shape))
        if key not in pool:
            pool[key] = [torch.empty(x.shape, dtype=torch.float32).pin_memory(), torch.empty(y.shape, dtype=torch.long).pin_memory(), None]
        px, py, ev = pool[key]
        if ev is not None:
            ev.synchronize()                                 # PIN_RING microbatches ago: long done
        px.copy_(x)
        py.copy_(y)
        self._pin_last = key
        return px, py

    def _pin_mark(self, stream) -> None:
        key = self.__dict__.get("_pin_last")
        if key is not None:
            ev = torch.cuda.Event()
            ev.record(stream)
            self._pin_pool[key][2] = ev
            self._pin_last = None

    # ------------------------------------------------------------------
    def _lanes(self, msg: dict):
        """A *lane* is one first-stage client's chain through the stages.  Lane ``i`` is served at stage ``s`` by member
        ``i % n_s`` (the deterministic counterpart of the reference's competing-consumer queue), which needs every
        stage to divide its predecessor: [4,4], [4,2], [2,1], [4,2,1] ... A stage member serving several lanes
        multiplexes them on its one executor (fan-in).  Returns [(lane, upstream id | None, downstream id | None)]."""
        members: Dict[int, list] = msg["peers"]["members"]
        ids = {s: [cid for cid, _ in members[s]] for s in members}
        n = {s: len(ids[s]) for s in ids}
        L = self.num_layers
        # the edge into the last stage is dynamic (ticket ring) on request — including topologies static lanes cannot cover
        # ([4, 3]); without the flag those keep the host data plane (the ring's hardware tests still fail intermittently
        # when several clients share ONE GPU, so it does not switch itself on)
        self._dynamic = L >= 2 and n.get(L, 0) > 0 and bool(self.opts.get("dynamic-consumers", False))
        for s in range(2, L + 1):
            if s == L and self._dynamic:
                continue
            if n[s] == 0 or n[s - 1] % n[s] != 0:
                raise RuntimeError(f"replica counts {[n[k] for k in sorted(n)]}: stage {s} does not divide stage {s - 1}")
        self._stage_ids = ids
        me = ids[self.layer_id].index(str(self.client_id))
        lanes = []
        for lane in range(n[1]):
            if self._dynamic and self.layer_id == L:         # a competing consumer serves whichever lane it claims
                lanes.append((lane, ids[L - 1][lane % n[L - 1]], None))
                continue
            if lane % n[self.layer_id] != me:
                continue
            up = ids[self.layer_id - 1][lane % n[self.layer_id - 1]] if self.layer_id > 1 else None
            down = ids[self.layer_id + 1][lane % n[self.layer_id + 1]] if self.layer_id < self.num_layers else None
            if self._dynamic and self.layer_id == L - 1:
                down = None                                  # no fixed partner: the outbox + ticket ring take its place
            lanes.append((lane, up, down))
        return lanes

    def _wire(self, msg: dict) -> None:
        from ..train.b200_executor import B200Executor
        if not isinstance(self.executor, B200Executor) or self.num_layers < 2:
            raise RuntimeError("stage has no native plan")
        ex = self.executor
        B = int(self.learning["batch-size"])
        depth = int(self.learning.get("control-count", 3))
        lanes = self._lanes(msg)
        dev = ex.device
        isz = 4 if ex.fp32 else 2                                  # payload element size (all stages share b200.precision)
        my_q = f"ipc_{self.client_id}"
        self.channel.queue_declare(my_q)
        fwd_in: Dict[int, Mailbox] = {}
        grad_in: Dict[int, Mailbox] = {}
        # Mailboxes I own are allocated once per (edge, geometry) and re-used by every later round (a START re-wires the
        # stages — the executor may be new — but must not leak cudaMalloc'ed rings or IPC mappings round after round).
        cache = self.__dict__.setdefault("_mailbox_cache", {})

        def own(kind: str, lane: int, spec: MailboxSpec):
            key = (kind, lane, spec.depth, spec.batch, spec.payload_shape, spec.itemsize)
            if key not in cache:
                cache[key] = Mailbox.allocate_exportable(spec, dev)
            mb, hdl = cache[key]
            mb.base[spec.flags_off: spec.flags_off + 4 * spec.depth].zero_()      # fresh sequence numbers for fresh stages
            return mb, hdl
        L = self.num_layers
        dyn_consumer = self._dynamic and self.layer_id == L
        dyn_producer = self._dynamic and self.layer_id == L - 1
        consumers = self._stage_ids[L] if self._dynamic else []
        fwd_out: Dict[int, Mailbox] = {}
        grad_out: Dict[int, Mailbox] = {}
        self._outboxes: Dict[int, Mailbox] = {}
        self._ring = None
        need = 0
        for lane, up, down in lanes:
            if dyn_consumer:                                       # local inbox: filled by my own copy-in from the origin's outbox
                c, h, w = ex.in_shape
                fwd_in[lane], _ = own("inbox", lane, MailboxSpec(depth, B, (B, h, w, c), itemsize=isz))
                need += 2                                          # the origin's outbox + its gradient mailbox
            elif up is not None:                                   # I consume this lane's activations
                c, h, w = ex.in_shape
                spec_in = MailboxSpec(depth, B, (B, h, w, c), itemsize=isz)
                fwd_in[lane], hdl = own("act", lane, spec_in)
                self.channel.publish_obj(f"ipc_{up}", {"kind": "act", "lane": lane, "handle": hdl,
                                                       "shape": spec_in.payload_shape})
                need += 1
            if dyn_producer:                                       # outbox + gradient mailbox, offered to every consumer
                c, h, w = ex.out_shape
                spec_out = MailboxSpec(depth, B, (B, h, w, c), itemsize=isz)
                fwd_out[lane], h_out = own("outbox", lane, spec_out)
                grad_in[lane], h_grad = own("grad", lane, spec_out)
                for cid in consumers:
                    self.channel.publish_obj(f"ipc_{cid}", {"kind": "outbox", "lane": lane, "handle": h_out, "shape": spec_out.payload_shape})
                    self.channel.publish_obj(f"ipc_{cid}", {"kind": "grad", "lane": lane, "handle": h_grad, "shape": spec_out.payload_shape})
            elif down is not None:                                 # I consume the gradients of this lane's output
                c, h, w = ex.out_shape
                spec_out = MailboxSpec(depth, B, (B, h, w, c), itemsize=isz)
                grad_in[lane], hdl = own("grad", lane, spec_out)
                self.channel.publish_obj(f"ipc_{down}", {"kind": "grad", "lane": lane, "handle": hdl,
                                                         "shape": spec_out.payload_shape})
                need += 1
        if self._dynamic and self.layer_id >= L - 1:
            from .ticket import TicketRing
            if dyn_consumer and consumers[0] == str(self.client_id):   # the edge's first consumer hosts the ring
                if "ring" not in cache:
                    cache["ring"] = TicketRing.allocate(dev)
                self._ring = cache["ring"]
                self._ring.reset()
                for cid in self._stage_ids[L - 1] + consumers[1:]:
                    self.channel.publish_obj(f"ipc_{cid}", {"kind": "ring", "handle": self._ring.handle})
            else:
                need += 1
        t0 = time.monotonic()
        while need:
            m = self.channel.get_obj(my_q, 0.25)
            if m is None:
                if time.monotonic() - t0 > self.watchdog:
                    raise TimeoutError("peer never posted its IPC handle")
                continue
            if m["kind"] == "ring":
                from .ticket import TicketRing
                self._ring = TicketRing.open(m["handle"], dev)
                need -= 1
                continue
            spec = MailboxSpec(depth, B, tuple(m["shape"]), itemsize=isz)
            mb = Mailbox.open_peer(spec, m["handle"], dev)
            if m["kind"] == "act":                                 # downstream's activation ring: I produce into it
                fwd_out[m["lane"]] = mb
            elif m["kind"] == "outbox":                            # an origin's outbox: I copy claimed microbatches out of it
                self._outboxes[m["lane"]] = mb
            else:                                                  # upstream's gradient ring
                grad_out[m["lane"]] = mb
            need -= 1
        multi = len(lanes) > 1
        if multi:                                                  # all lanes' mailbox slots become the plan's input slots
            ex.plan(B).bind_inputs([t for lane, _, _ in lanes for t in fwd_in[lane].payload])
        self.dstages: Dict[int, DeviceStage] = {}
        for k, (lane, up, down) in enumerate(lanes):
            self.dstages[lane] = DeviceStage(ex, B, depth, fwd_in=fwd_in.get(lane), grad_in=grad_in.get(lane),
                                             fwd_out=fwd_out.get(lane), grad_out=grad_out.get(lane),
                                             slot_offset=k * depth if multi else 0, bind_inputs=not multi)
        if dyn_consumer:
            from .mailbox import EdgeCounters
            self._in_seq = {lane: EdgeCounters(depth, dev) for lane, _, _ in lanes}      # copy-in publishes of my inboxes
        self._lane_info = lanes
        self.dstage = self.dstages[lanes[0][0]]
        self._edges = (fwd_in, grad_in, fwd_out, grad_out, multi, depth)
        self._tail_stages: Dict[tuple, DeviceStage] = {}

    def _tail_stage(self, lane: int, b: int) -> DeviceStage:
        """Program set of lane ``lane`` for a microbatch of ``b`` < batch-size samples, on the lane's own mailboxes."""
        key = (lane, b)
        if key not in self._tail_stages:
            fwd_in, grad_in, fwd_out, grad_out, multi, depth = self._edges
            ex = self.executor
            lanes = self._lane_info
            k = [l for l, _, _ in lanes].index(lane)
            view = lambda d: d[lane].view(b) if lane in d else None
            if multi and not any(kk[1] == b for kk in self._tail_stages):
                ex.plan(b).bind_inputs([t for l, _, _ in lanes for t in fwd_in[l].view(b).payload])
            self._tail_stages[key] = DeviceStage(ex, b, depth, fwd_in=view(fwd_in), grad_in=view(grad_in), fwd_out=view(fwd_out),
                                                 grad_out=view(grad_out), slot_offset=k * depth if multi else 0,
                                                 bind_inputs=not multi, counters_from=self.dstages[lane])
        return self._tail_stages[key]

    # ------------------------------------------------------------------
    def _collect_plans(self, lanes) -> Dict[int, int]:
        """Batch counts announced by the first-stage client of every lane I serve."""
        plan_q = f"plan_{self.client_id}"
        counts: Dict[int, int] = {}
        self._tails: Dict[int, int] = {}
        t0 = time.monotonic()
        while len(counts) < len(lanes):
            m = self.channel.get_obj(plan_q, 0.25)
            if m is None:
                if time.monotonic() - t0 > self.watchdog:
                    raise TimeoutError("upstream never announced its batch count")
                continue
            counts[int(m["lane"])] = int(m["batches"])
            self._tails[int(m["lane"])] = int(m.get("tail", 0))
        return counts

    def progress(self) -> int:
        # while the device loop runs, host counters stand still (the host only enqueues, then blocks in a stream synchronize);
        # the device-side waits are bounded by their own spin limits, so "busy on the device" counts as moving
        if self.__dict__.get("_device_busy"):
            return time.monotonic_ns() // 1_000_000
        return super().progress() + int(self.__dict__.get("_device_steps", 0))

    def _announce(self, down, lane: int, batches: int, tail: int) -> None:
        """Batch count of a lane to whoever consumes it: the fixed partner, or every competing consumer."""
        targets = [down] if down is not None else (self._stage_ids[self.num_layers] if self._dynamic else [])
        for cid in targets:
            self.channel.publish_obj(f"plan_{cid}", {"lane": lane, "batches": batches, "tail": tail})

    def _offer(self, st: DeviceStage, lane: int, it: int) -> None:
        """Producer side of the ticket ring: microbatch ``it`` of ``lane`` sits in my outbox — queue a ticket for it."""
        if self._dynamic and self.layer_id == self.num_layers - 1:
            with torch.cuda.stream(st.stream):
                self._ring.publish(lane, it, st.exp_grad.at(it % st.depth), st.B)

    def _run_competing_consumer(self, lanes, counts, tails, log_loss) -> int:
        """Last stage behind a ticket ring: claim → copy the origin's outbox slot in → L program → gradient to the origin."""
        from ..ops import native as N
        depth = self.dstage.depth
        total = sum(counts.values()) + sum(1 for l in tails if tails[l])
        ahead = max(1, int(self.opts.get("claim-ahead", 2)))
        slow = float((self.opts.get("debug-slow-ms") or {}).get(self.rank, 0.0)) if isinstance(self.opts.get("debug-slow-ms"), dict) else 0.0
        done: List[torch.cuda.Event] = []
        mine = 0
        self.claimed: List[tuple] = []
        for lane, b in tails.items():                        # compile the trailing-batch programs before the first claim,
            if b:                                            # not while an origin is already waiting for its gradient
                self._tail_stage(lane, b)
        while True:
            if len(done) >= ahead:
                done[len(done) - ahead].synchronize()            # a busy replica does not hoard tickets
            got = self._ring.claim(total, timeout=self.watchdog)
            if got is None:
                break
            _, lane, it, gseq, b = got
            st = self.dstages[lane] if b == self.dstage.B else self._tail_stage(lane, b)
            slot = it % depth
            src, box = self._outboxes[lane], st.fwd_in
            nbytes = b * self._payload_bytes_per_sample(src)
            with torch.cuda.stream(st.stream):
                N.memcpy_async(box.payload[slot].data_ptr(), src.payload[slot].data_ptr(), nbytes)
                N.memcpy_async(box.labels[slot].data_ptr(), src.labels[slot].data_ptr(), b * 8)
                N.set_flag(box.flag_ptr(slot), 0, seq=self._in_seq[lane].at(slot))
                N.store_u32(st.seq_grad.at(slot).data_ptr(), (gseq - 1) & 0xFFFFFFFF)
                if slow > 0:
                    torch.cuda._sleep(int(slow * 1.9e6))         # test hook: an artificially slow replica
            st.last(it)
            log_loss(st)
            ev = torch.cuda.Event()
            ev.record(st.stream)
            done.append(ev)
            self.claimed.append((lane, it))
            mine += 1
        return mine

    @staticmethod
    def _payload_bytes_per_sample(mb: Mailbox) -> int:
        n = mb.spec.itemsize
        for d in mb.spec.payload_shape[1:]:
            n *= d
        return n

    def run_stage(self):
        if self.dstage is None:
            return super().run_stage()
        lanes = self._lane_info
        B = self.dstage.B
        ex = self.executor
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self._device_busy = True
        ev0.record(self.dstage.stream)
        self.timing["at_run_stage"] = time.monotonic()          # at_*: CLOCK_MONOTONIC stamps (one box: comparable across processes)
        # the reference prints the loss of every microbatch (src/train/VGG16.py:168, a host sync per step); here every step
        # copies it to pinned host memory asynchronously and the round reports the mean
        loss_log: List[torch.Tensor] = []

        def log_loss(st):
            if self.is_last and self.opts.get("loss-every-step", True):
                h = self._loss_pool[len(loss_log)] if len(loss_log) < len(self._loss_pool) else torch.zeros(4).pin_memory()
                with torch.cuda.stream(st.stream):
                    h.copy_(ex.loss_buf, non_blocking=True)
                loss_log.append(h)
        if not hasattr(self, "_loss_pool"):
            self._loss_pool = [torch.zeros(4).pin_memory() for _ in range(512)] if self.is_last else []
        if self.is_first:
            lane, _, down = lanes[0]
            st = self.dstages[lane]
            loader = self.train_loader
            t_l = time.perf_counter()
            n_samples = None
            if getattr(loader, "batch_size", None) == B and not getattr(loader, "drop_last", False) and hasattr(loader, "dataset"):
                try:
                    n_samples = len(loader.dataset)
                except TypeError:
                    n_samples = None
            if n_samples is not None:
                # batch count known up front: announce it, then stream — the loader assembles microbatch i+1 on the host
                # while the device runs microbatch i (the short trailing batch, if any, comes last)
                n, tb = n_samples // B, n_samples % B
                source = loader
            else:                                                   # opaque iterable: materialise it to learn the counts
                full, tail = [], None
                for batch in loader:
                    x = batch[0] if not isinstance(batch, dict) else batch["input_ids"]
                    if x.shape[0] == B:
                        full.append(batch)
                    elif 0 < x.shape[0] < B:
                        tail = batch
                n, tb = len(full), (int((tail[0] if not isinstance(tail, dict) else tail["input_ids"]).shape[0]) if tail is not None else 0)
                source = full + ([tail] if tail is not None else [])
            self._announce(down, lane, n, tb)
            total = n + (1 if tb else 0)
            stage_of = lambda it: st if it < n else self._tail_stage(lane, tb)
            it = it_b = 0
            prof = {"loader": 0.0, "stage": 0.0, "fwd": 0.0, "bwd": 0.0} if self.opts.get("profile-host") else None
            pc = time.perf_counter
            t_prev = pc()
            for batch in source:
                if prof is not None:
                    prof["loader"] += pc() - t_prev
                x, y = batch if not isinstance(batch, dict) else (batch["input_ids"], batch["labels"])
                if it >= total or int(x.shape[0]) != (B if it < n else tb):
                    raise RuntimeError(f"loader produced microbatch {it} of {int(x.shape[0])} samples; announced {n} x {B} + {tb}")
                if not x.is_cuda:
                    x, y = x.float(), torch.as_tensor(y).long()
                x, y = self._pinned(it, x, y)
                t0 = pc()
                if it - it_b >= st.depth:
                    stage_of(it_b).backward(it_b)
                    it_b += 1
                t1 = pc()
                stage_of(it).stage_input(it, x, y)
                self._pin_mark(stage_of(it).stream)
                t2 = pc()
                stage_of(it).forward(it)
                self._offer(stage_of(it), lane, it)
                it += 1
                t_prev = pc()
                if prof is not None:
                    prof["bwd"] += t1 - t0
                    prof["stage"] += t2 - t1
                    prof["fwd"] += t_prev - t2
            if it != total:
                raise RuntimeError(f"loader produced {it} microbatches; announced {total}")
            while it_b < total:
                stage_of(it_b).backward(it_b)
                it_b += 1
            self.timing["loader_and_launch"] = (time.perf_counter() - t_l) * 1e3
            if prof is not None:
                self.timing.update({f"host_{k}": v * 1e3 for k, v in prof.items()})
        else:
            counts = self._collect_plans(lanes)
            tails = self._tails
            if not self.is_last:
                for lane, _, down in lanes:
                    self._announce(down, lane, counts[lane], tails.get(lane, 0))
            depth = self.dstage.depth
            n_of = {lane: counts[lane] + (1 if tails.get(lane, 0) else 0) for lane, _, _ in lanes}
            most = max(n_of.values()) if n_of else 0

            def stage_of(lane, it):
                return self.dstages[lane] if it < counts[lane] else self._tail_stage(lane, tails[lane])
            # lanes are interleaved microbatch by microbatch (round-robin over the upstream replicas)
            t_l = time.perf_counter()
            dyn_last = self._dynamic and self.is_last
            for it in range(0 if dyn_last else most + (0 if self.is_last else depth)):
                for lane, _, _ in lanes:
                    n = n_of[lane]
                    if self.is_last:
                        if it < n:
                            stage_of(lane, it).last(it)
                            log_loss(stage_of(lane, it))
                    else:
                        if 0 <= it - depth < n:
                            stage_of(lane, it - depth).backward(it - depth)
                        if it < n:
                            stage_of(lane, it).forward(it)
                            self._offer(stage_of(lane, it), lane, it)
            total = self._run_competing_consumer(lanes, counts, tails, log_loss) if dyn_last else sum(n_of.values())
            self.timing["launch_downstream"] = (time.perf_counter() - t_l) * 1e3
        ev1.record(self.dstage.stream)
        t_s = time.perf_counter()
        self.dstage.stream.synchronize()
        self.timing["final_sync"] = (time.perf_counter() - t_s) * 1e3
        self.timing["at_device_done"] = time.monotonic()
        self._device_busy = False
        self._device_steps = int(self.__dict__.get("_device_steps", 0)) + int(total)
        self.last_device_ms = ev0.elapsed_time(ev1)
        t_c = time.perf_counter()
        self.last_loss = (sum(float(h[0]) for h in loss_log) / len(loss_log)) if loss_log else None
        for st in list(self.dstages.values()) + list(self._tail_stages.values()):
            st.check()
        self.timing["check"] = (time.perf_counter() - t_c) * 1e3
        if self.is_first:
            self.send_to_server(M.notify(self.client_id, self.layer_id, self.cluster))
        t_p = time.perf_counter()
        self.timing["at_notify"] = time.monotonic()
        self.trainer._wait_pause()
        self.timing["at_pause"] = time.monotonic()
        self.timing["wait_pause"] = (time.perf_counter() - t_p) * 1e3
        return (not ex.nan_detected()), total

    # ------------------------------------------------------------------ round end
    def _replicas(self) -> List[str]:
        return [cid for cid, _ in self.start_msg["peers"]["members"][self.layer_id]]

    def upload(self, result: bool, size: int, send: bool = True) -> None:
        """Round end on the device plane: every trainable client of every cluster joins ONE device all-reduce
        (``parallel/allreduce.py``) that applies the reference's per-cluster weighted FedAvg and the cross-cluster mean
        (src/Server.py:398-434) in place over NVLink — also between stages of clusters cut at different layers.  Every
        replica ends the round holding the global model for its layers, so the next START carries no parameters; only
        the stage leaders of the first cluster upload a state-dict (validation / checkpoint on the server)."""
        everyone = (self.start_msg.get("peers") or {}).get("all") or []
        members = sorted(str(cid) for cid, _, _ in everyone)
        if self.dstage is None or len(members) < 2 or not self.opts.get("device-fedavg", True):
            return super().upload(result, size, send)
        from .allreduce import DeviceFedAvg
        from .fedavg import BrokerComm
        ex = self.executor
        me = str(self.client_id)
        key = (tuple(members), id(ex))
        if getattr(self, "_fa_key", None) != key:
            comm = BrokerComm(self.channel, "fa_all", members, me, timeout=self.watchdog)
            self._fa = DeviceFedAvg(ex, me, int(self.cluster or 0), comm)
            self._fa.setup()                                 # one handle exchange; later rounds are device-only
            self._fa_key = key
        t_f = time.perf_counter()
        done = self._fa.run(float(size), ok=bool(result))
        self.timing["fedavg"] = (time.perf_counter() - t_f) * 1e3
        mine = sorted(str(cid) for cid, cl, lid in everyone if int(cl) == int(self.cluster or 0) and int(lid) == self.layer_id)
        first_cluster = min(int(cl) for _, cl, _ in everyone)
        leader = bool(mine) and mine[0] == me and int(self.cluster or 0) == first_cluster
        want = leader and send and done and bool(self.start_msg.get("save_parameters", True))
        rnd = int(self.start_msg.get("round", self.rounds_done + 1))
        self.timing["at_update"] = time.monotonic()
        extra = dict(resident=True, device_ms=getattr(self, "last_device_ms", None), loss=getattr(self, "last_loss", None),
                     timing=dict(self.timing))
        if want and self.start_msg.get("async_checkpoint"):
            # the checkpoint copy leaves the round's critical path: snapshot on the device now (the next round may already
            # be training when the bytes cross PCIe), UPDATE without a payload, CHECKPOINT from a background thread
            t_s = time.perf_counter()
            snap = ex.state_dict()                           # detached device clones in the reference's key order
            extra["timing"]["snapshot"] = (time.perf_counter() - t_s) * 1e3
            self.send_to_server(M.update(self.client_id, self.layer_id, bool(result) and done, size, self.cluster, None,
                                         checkpoint_follows=rnd, **extra))
            self._ship_checkpoint(snap, rnd)
            return
        sd = {k: v.detach().to("cpu") for k, v in ex.state_dict().items()} if want else None
        self.send_to_server(M.update(self.client_id, self.layer_id, bool(result) and done, size, self.cluster, sd, **extra))

    def _ship_checkpoint(self, snap: dict, rnd: int) -> None:
        import threading
        ch = self.__dict__.get("_ckpt_channel")
        if ch is None:
            ch = self._ckpt_channel = self.channel.clone()
        prev = self.__dict__.get("_ckpt_thread")

        def ship():
            if prev is not None:
                prev.join()                                   # rounds arrive in order
            with torch.cuda.device(self.executor.device):
                host = {k: v.to("cpu") for k, v in snap.items()}
            ch.publish_obj(M.CKPT_QUEUE, M.checkpoint(self.client_id, self.layer_id, self.cluster, rnd, host))
        self._ckpt_thread = threading.Thread(target=ship, name=f"slb200-ckpt-upload-{rnd}")     # non-daemon: finishes before exit
        self._ckpt_thread.start()
