"""Per-stage training loops (first / middle / last).

Behavioural contract = reference ``Train_VGG16`` (src/train/VGG16.py:61-190) generalised:
  * first stage: gradient-first scheduling, at most ``control-count`` microbatches in
    flight, epoch ends when the loader is exhausted and every forward got its gradient,
    then NOTIFY and wait for PAUSE; returns ``(ok, data_count)`` with ``data_count`` in
    *batches* (the FedAvg weight, src/train/VGG16.py:109);
  * last stage: fwd + loss + bwd + step per received microbatch, gradient sent to
    ``trace[-1]``; leaves on PAUSE once its queue is drained;
  * middle stage (absent in the reference, contract derived in SURVEY §3.5): forward and
    re-publish with own id appended to ``trace``; on gradient: (re)compute, step, send the
    input gradient upstream.
Variant knobs: ``local_round`` epochs per round (DCSL), ``limited_time`` wall-clock budget
(Vanilla_SL), ``strict`` 1-in-flight (DCSL/FLEX), ``sda_size`` batch-concat on the last stage
(DCSL SDA, other/DCSL/src/Scheduler.py:152-221), ``targets`` round-robin per-device queues.
A watchdog turns the reference's silent deadlock into a ``TimeoutError``.
"""
from __future__ import annotations

import time
import uuid
from typing import Any, List, Optional, Sequence, Tuple

import torch

from .. import messages as M
from ..log import print_with_color
from ..transport import Channel
from .executor import StageExecutor


def _split_batch(batch) -> Tuple[Any, torch.Tensor]:
    if isinstance(batch, dict):
        return batch, batch["labels"]
    x, y = batch
    return x, (y if isinstance(y, torch.Tensor) else torch.as_tensor(y))


class StageTrainer:
    def __init__(self, client_id, layer_id: int, channel: Channel, executor: StageExecutor, dataplane,
                 watchdog: float = 120.0, verbose: bool = False):
        self.client_id, self.layer_id, self.ch = client_id, layer_id, channel
        self.ex, self.dp = executor, dataplane
        self.data_count = 0
        self.watchdog = watchdog
        self.verbose = verbose
        self.reply_q = M.reply_queue(client_id)
        self.pause_msg: Optional[dict] = None
        self.alive_at = time.monotonic()          # last heartbeat relayed by the server: idle timers restart from here

    # ------------------------------------------------------------------
    def _send_to_server(self, msg) -> None:
        self.ch.publish_obj(M.RPC_QUEUE, msg)

    def _poll_pause(self, timeout: float = 0.0) -> bool:
        if self.pause_msg is not None:
            return True
        m = self.ch.get_obj(self.reply_q, timeout)
        if m is None:
            return False
        if m.get("action") == M.HEARTBEAT:       # aged by its send time: a backlog of old beacons proves nothing
            age = max(0.0, time.time() - float(m.get("t", time.time())))
            prog = m.get("progress")
            if prog is None or prog != self.__dict__.get("_seen_progress"):
                # the server is alive AND somebody (a peer, the server, me) has done work since the last beacon; beacons
                # over a system that has stopped moving do not extend a wait
                self._seen_progress = prog
                self.alive_at = max(self.alive_at, time.monotonic() - age)
            return False
        if m.get("action") == M.PAUSE:
            self.pause_msg = m
            return True
        if m.get("action") == M.STOP:
            self.pause_msg = m
            return True
        return False

    def _wait_pause(self) -> None:
        # PAUSE arrives when the *slowest* first-stage peer of the cluster has finished (src/Server.py:137-153): that can
        # take arbitrarily long on a healthy run, so the wait is bounded by server silence (heartbeats), not by wall time
        t0 = time.monotonic()
        while not self._poll_pause(0.05):
            if time.monotonic() - max(t0, self.alive_at) > self.watchdog:
                raise TimeoutError(f"client {self.client_id}: no PAUSE and no server heartbeat for {self.watchdog}s")

    def _idle_too_long(self, since: float) -> bool:
        return time.monotonic() - max(since, self.alive_at) > self.watchdog

    def _check_stop(self) -> None:
        """While blocked on the data plane, keep reading the control queue: heartbeats refresh the watchdog, STOP (the
        server aborting the run, e.g. a dead peer) ends the loop instead of waiting for data that will never come."""
        for _ in range(64):                       # drain the beacons that piled up while this loop was busy
            if self._poll_pause(0.0):
                if self.pause_msg.get("action") == M.STOP:
                    raise RuntimeError(f"client {self.client_id}: server stopped the run")
                return
            if self.ch.queue_depth(self.reply_q) == 0:
                return

    # ------------------------------------------------------------------
    def train_on_first_layer(self, learning: dict, train_loader, cluster=None, local_round: int = 1,
                             limited_time: Optional[dict] = None, strict: bool = False,
                             targets: Optional[Sequence] = None) -> Tuple[bool, int]:
        cc = 1 if strict else int(learning.get("control-count", 3))
        budget = float(limited_time["time"]) if limited_time and limited_time.get("enable") else None
        max_epochs = int(limited_time.get("epoch", 100)) if budget is not None else local_round
        t_start = time.monotonic()
        nf = nb = 0
        rr = 0
        for _epoch in range(max_epochs):
            it = iter(train_loader)
            end = False
            last_progress = time.monotonic()
            while True:
                g = self.dp.recv_gradient(0.0)
                if g is not None:
                    self.ex.backward(g["data_id"], g["data"])
                    nb += 1
                    last_progress = time.monotonic()
                elif self.ex.in_flight() >= cc or end:
                    g = self.dp.recv_gradient(0.02)       # block briefly instead of busy-spinning
                    if g is not None:
                        self.ex.backward(g["data_id"], g["data"])
                        nb += 1
                        last_progress = time.monotonic()
                    elif self._idle_too_long(last_progress):
                        self._check_stop()            # read the control queue (heartbeats / STOP) before giving up
                        if not self._idle_too_long(last_progress):
                            continue
                        raise TimeoutError(f"stage-1 client {self.client_id}: gradient never arrived "
                                           f"({nf} sent / {nb} received)")
                else:
                    try:
                        x, labels = _split_batch(next(it))
                    except StopIteration:
                        end = True
                        continue
                    data_id = uuid.uuid4()
                    out = self.ex.forward_only(data_id, x)
                    nf += 1
                    self.data_count += 1
                    tgt = None
                    if targets:
                        tgt = targets[rr % len(targets)]
                        rr += 1
                    self.dp.send_forward(data_id, out, labels, trace=None, target=tgt)
                    last_progress = time.monotonic()
                if end and nf == nb:
                    break
                if budget is not None and time.monotonic() - t_start > budget and nf == nb:
                    end = True
                    break
            if budget is not None and time.monotonic() - t_start > budget:
                break
        self._send_to_server(M.notify(self.client_id, self.layer_id, cluster))
        self._wait_pause()
        return (not self.ex.nan_detected()), self.data_count

    # ------------------------------------------------------------------
    def train_on_last_layer(self, learning: dict, cluster=None, sda_size: int = 1,
                            source=None) -> Tuple[bool, int]:
        pending: List[dict] = []
        idle_since = time.monotonic()
        while True:
            m = self.dp.recv_forward(0.0 if pending else 0.005, source=source)
            if m is not None:
                idle_since = time.monotonic()
                if sda_size <= 1:
                    grad = self.ex.forward_backward_last(m["data"], m["label"])
                    self.data_count += 1
                    if grad is not None:
                        self.dp.send_gradient(m["data_id"], grad, m["trace"])
                    if self.verbose:
                        print_with_color(f"Loss: {self.ex.last_loss()}", "end")
                else:
                    pending.append(m)
                    if len(pending) >= sda_size:
                        self._sda_step(pending)
                        pending = []
                continue
            if self._poll_pause(0.0):
                if pending:                       # flush a partial SDA group before leaving
                    self._sda_step(pending)
                break
            if self._idle_too_long(idle_since):
                raise TimeoutError(f"last-stage client {self.client_id}: idle for {self.watchdog}s without PAUSE or heartbeat")
        return (not self.ex.nan_detected()), self.data_count

    def _sda_step(self, group: List[dict]) -> None:
        sizes = [g["data"].shape[0] for g in group]
        x = torch.cat([g["data"] for g in group], dim=0)
        y = torch.cat([g["label"] for g in group], dim=0)
        grad = self.ex.forward_backward_last(x, y)
        self.data_count += 1
        if grad is not None:
            for g, piece in zip(group, torch.split(grad, sizes, dim=0)):
                self.dp.send_gradient(g["data_id"], piece.contiguous(), g["trace"])

    # ------------------------------------------------------------------
    def train_on_middle_layer(self, learning: dict, cluster=None) -> Tuple[bool, int]:
        traces = {}
        idle_since = time.monotonic()
        while True:
            g = self.dp.recv_gradient(0.0)
            if g is not None:
                idle_since = time.monotonic()
                gin = self.ex.backward(g["data_id"], g["data"])
                up_trace = traces.pop(g["data_id"])
                if gin is not None:
                    self.dp.send_gradient(g["data_id"], gin, up_trace)
                continue
            m = self.dp.recv_forward(0.002)
            if m is not None:
                idle_since = time.monotonic()
                out = self.ex.forward_only(m["data_id"], m["data"])
                traces[m["data_id"]] = list(m["trace"])
                self.data_count += 1
                self.dp.send_forward(m["data_id"], out, m["label"], trace=m["trace"])
                continue
            if not traces and self._poll_pause(0.0):
                break
            if self._idle_too_long(idle_since):
                raise TimeoutError(f"middle-stage client {self.client_id}: idle for {self.watchdog}s")
        return (not self.ex.nan_detected()), self.data_count
