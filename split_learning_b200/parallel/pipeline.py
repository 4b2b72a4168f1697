"""Device-resident split pipeline: stage programs as CUDA graphs chained by mailbox flags.

``DeviceStage`` binds a ``B200Executor`` to its edges:

    first stage   F(slot): forward; the fused BN/ReLU/pool kernel of the cut block stores
                           straight into the next stage's mailbox slot (+labels) and flags it
                  B(slot): wait(grad flag) → recompute forward → backward → SGD
    middle stage  F(slot): wait(act flag) → forward → store+flag downstream
                  B(slot): wait(grad flag) → recompute → backward (cut-head dgrad stores dX
                           into the upstream gradient mailbox) → flag → SGD
    last stage    L(slot): wait(act flag) → forward → CE → backward (dgrad stores dX upstream)
                           → flag → SGD

Each program is captured once per slot; an epoch is a static 1F1B schedule (``control-count``
forwards of warm-up, then gradient-first alternation — the steady state of the reference's
loop, src/train/VGG16.py:76-119) that the host merely enqueues.  ``LocalPipeline`` runs all
stages of one replica chain in a single process/stream (N = 1 GPU); the multi-process runner
(``parallel/runner.py``) gives every stage its own GPU and wires the same objects to
IPC-mapped peer mailboxes.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ..ops import native as N
from ..train.b200_executor import B200Executor
from ..utils.timing import capture_graph
from .mailbox import EdgeCounters, Mailbox, MailboxSpec


def act_spec(ex: B200Executor, batch: int, depth: int) -> MailboxSpec:
    """Mailbox geometry of the activation edge *leaving* ``ex`` (== gradient edge entering it)."""
    if ex.out_kind != "image":
        raise NotImplementedError("device data plane supports cuts inside the convolutional trunk")
    c, h, w = ex.out_shape
    return MailboxSpec(depth, batch, (batch, h, w, c), with_labels=True, itemsize=4 if ex.fp32 else 2)


class DeviceStage:
    def __init__(self, ex: B200Executor, batch: int, depth: int,
                 fwd_in: Optional[Mailbox] = None, grad_in: Optional[Mailbox] = None,
                 fwd_out: Optional[Mailbox] = None, grad_out: Optional[Mailbox] = None,
                 stream: Optional[torch.cuda.Stream] = None, wait_spins: Optional[int] = None,
                 slot_offset: int = 0, bind_inputs: bool = True, counters_from: Optional["DeviceStage"] = None):
        """``slot_offset`` / ``bind_inputs=False``: one of several *lanes* multiplexed on the same executor (fan-in: a
        stage fed by several upstream replicas).  The caller binds the plan's input slots to the concatenated mailbox
        payloads of all lanes once; this lane's slot ``s`` is plan slot ``slot_offset + s``."""
        self.ex, self.B, self.depth = ex, batch, depth
        self.off = int(slot_offset)
        self.fwd_in, self.grad_in, self.fwd_out, self.grad_out = fwd_in, grad_in, fwd_out, grad_out
        self.stream = stream or ex.stream
        self.plan = ex.plan(batch)
        dev = ex.device
        if counters_from is not None:
            # a second program set on the same edges (another batch size: the trailing partial microbatch of an epoch) keeps
            # counting where the main stage stands — mailbox flags are monotonic per edge, not per program
            self.seq_fwd, self.seq_grad = counters_from.seq_fwd, counters_from.seq_grad
            self.exp_fwd, self.exp_grad = counters_from.exp_fwd, counters_from.exp_grad
        else:
            self.seq_fwd = EdgeCounters(depth, dev)       # values I publish downstream
            self.seq_grad = EdgeCounters(depth, dev)      # values I publish upstream
            self.exp_fwd = EdgeCounters(depth, dev)       # values I have consumed from upstream activations
            self.exp_grad = EdgeCounters(depth, dev)      # values I have consumed from downstream gradients
        self.status = torch.zeros(4, dtype=torch.int32, device=dev)
        import os
        self.wait_spins = wait_spins if wait_spins is not None else int(os.environ.get("SLB200_WAIT_SPINS", str(1 << 28)))
        self._posted = {"F": 0, "B": 0, "L": 0}
        self.labels_slots = [torch.zeros(batch, dtype=torch.int64, device=dev) for _ in range(depth)]
        if fwd_in is not None and bind_inputs:        # consume activations in place from my mailbox
            self.plan.bind_inputs(fwd_in.payload)
        elif fwd_in is None and self.plan.n_slots < depth:
            raise RuntimeError("executor has fewer input slots than the pipeline depth")
        self.graphs: Dict[Tuple[str, int], torch.cuda.CUDAGraph] = {}
        self.use_graphs = ex.use_graphs
        self._warmed: set = set()
        self.launches_per: Dict[str, int] = {}

    # ---- program bodies ----------------------------------------------------------------
    def _wait(self, mb: Mailbox, ctr: EdgeCounters, slot: int):
        """Flag acquisition of a slot, fused into the first kernel of the consuming pass (see ``_Plan._forward``)."""
        return (mb.flag_ptr(slot), ctr.at(slot), self.wait_spins, self.status)

    def _F(self, slot: int) -> None:
        wait = None
        if self.fwd_in is not None:
            wait = self._wait(self.fwd_in, self.exp_fwd, slot)
            labels = self.fwd_in.labels[slot]
        else:
            labels = self.labels_slots[slot]
        pub = None
        out = None
        if self.fwd_out is not None:
            N.memcpy_async(self.fwd_out.labels[slot].data_ptr(), labels.data_ptr(), self.B * 8)
            out = self.fwd_out.payload[slot]
            pub = {"flag": self.fwd_out.flag_ptr(slot), "seq": self.seq_fwd.at(slot)}
        if wait is not None and self.fwd_out is not None:
            # a middle stage copies the labels downstream before its first kernel: the slot must be acquired first
            N.wait_flag(*wait[:1], 0, wait[1], wait[2], wait[3])
            wait = None
        self.plan._forward(self.off + slot, out_ptr_override=out, publish=pub, wait=wait)

    def _publish_grad(self, slot: int):
        return (self.grad_out.flag_ptr(slot), self.seq_grad.at(slot)) if self.grad_out is not None else None

    def _B(self, slot: int) -> None:
        wait = self._wait(self.grad_in, self.exp_grad, slot)
        if self.ex.recompute:
            self.plan._forward(self.off + slot, wait=wait)   # faithful recompute with current weights, no publish
        else:
            N.wait_flag(wait[0], 0, wait[1], wait[2], wait[3])
        gout = self.grad_out.payload[slot] if self.grad_out is not None else None
        self.plan._backward(self.grad_in.payload[slot], grad_out_override=gout, publish_grad=self._publish_grad(slot))
        if self.grad_out is not None and not self.plan.grad_published:
            N.set_flag(self.grad_out.flag_ptr(slot), 0, self.seq_grad.at(slot))

    def _L(self, slot: int) -> None:
        labels = self.labels_slots[slot]
        wait = None
        if self.fwd_in is not None:
            wait = self._wait(self.fwd_in, self.exp_fwd, slot)
            labels = self.fwd_in.labels[slot]
        gout = self.grad_out.payload[slot] if self.grad_out is not None else None
        self.plan._last(self.off + slot, labels=labels, grad_out_override=gout, wait=wait, publish_grad=self._publish_grad(slot))
        if self.grad_out is not None and not self.plan.grad_published:
            N.set_flag(self.grad_out.flag_ptr(slot), 0, self.seq_grad.at(slot))

    # ---- execution ----------------------------------------------------------------------
    def _exec(self, kind: str, slot: int) -> None:
        body = {"F": self._F, "B": self._B, "L": self._L}[kind]
        key = (kind, slot)
        with torch.cuda.stream(self.stream):
            if not self.use_graphs:
                body(slot)
                return
            g = self.graphs.get(key)
            if g is None:
                if kind not in self._warmed:          # first ever call of this program: eager (loads modules)
                    before = N.LAUNCHES
                    body(slot)
                    self.launches_per[kind] = N.LAUNCHES - before
                    self._warmed.add(kind)
                    return
                g = capture_graph(self.stream, lambda: body(slot))
                self.graphs[key] = g
            g.replay()

    # Host gates exist only for same-process partners (see mailbox.HostGate); they order *enqueueing*, not execution.
    def _gate_wait(self, mb: Optional[Mailbox], it: int) -> None:
        if mb is not None and mb.gate is not None:
            mb.gate.wait(it + 1)

    def _gate_post(self, mb: Optional[Mailbox]) -> None:
        if mb is not None and mb.gate is not None:
            mb.gate.post()

    def forward(self, it: int) -> None:
        self._gate_wait(self.fwd_in, it)
        self._exec("F", it % self.depth)
        self._posted["F"] += 1
        self._gate_post(self.fwd_out)

    def backward(self, it: int) -> None:
        self._gate_wait(self.grad_in, it)
        self._exec("B", it % self.depth)
        self._gate_post(self.grad_out)

    def last(self, it: int) -> None:
        self._gate_wait(self.fwd_in, it)
        self._exec("L", it % self.depth)
        self._posted["L"] += 1
        self._gate_post(self.grad_out)

    def stage_input(self, it: int, x_host: torch.Tensor, y_host: torch.Tensor) -> None:
        """First stage: H2D copy of a microbatch (pinned host memory) into input slot ``it % depth``."""
        slot = it % self.depth
        with torch.cuda.stream(self.stream):
            self.plan.x_in[slot].copy_(x_host, non_blocking=True)
            self.labels_slots[slot].copy_(y_host, non_blocking=True)

    def check(self) -> None:
        if int(self.status[0].item()) != 0:
            def flags(mb):
                return None if mb is None or mb.flags is None else mb.flags.tolist()
            raise TimeoutError(
                "device pipeline: a mailbox flag never arrived (peer stage dead?) "
                f"[stage first={self.ex.is_first} last={self.ex.is_last} seq_fwd={self.seq_fwd.ctr[::4].tolist()} "
                f"seq_grad={self.seq_grad.ctr[::4].tolist()} exp_fwd={self.exp_fwd.ctr[::4].tolist()} "
                f"exp_grad={self.exp_grad.ctr[::4].tolist()} fwd_in.flags={flags(self.fwd_in)} grad_in.flags={flags(self.grad_in)}]")

    def reset_counters(self) -> None:
        for c in (self.seq_fwd, self.seq_grad, self.exp_fwd, self.exp_grad):
            c.reset()


class LocalPipeline:
    """All stages of one chain on one GPU (the N = 1 configuration), one CUDA stream per stage.

    Mailboxes are plain local allocations (producer pointer == consumer pointer), i.e. exactly the
    multi-GPU kernels with ``peer = self``; the mailbox flags are the only synchronisation between
    the stage streams, so stage 1's recompute+backward of microbatch i-3 overlaps stage 2's work on
    microbatch i.  The host enqueues producers before their consumers, so every flag wait is
    released by work that is already submitted.  ``overlap=False`` puts all stages on one stream.
    """

    def __init__(self, executors: Sequence[B200Executor], batch: int, depth: int, overlap: bool = True):
        assert executors[0].is_first and executors[-1].is_last
        self.depth, self.B = depth, batch
        dev = executors[0].device
        stream = executors[0].stream
        self.overlap = overlap
        self.stages: List[DeviceStage] = []
        n = len(executors)
        acts = [Mailbox.allocate_local(act_spec(executors[i], batch, depth), dev) for i in range(n - 1)]
        grads = [Mailbox.allocate_local(act_spec(executors[i], batch, depth), dev) for i in range(n - 1)]
        for i, ex in enumerate(executors):
            self.stages.append(DeviceStage(
                ex, batch, depth,
                fwd_in=acts[i - 1] if i > 0 else None, grad_in=grads[i] if i < n - 1 else None,
                fwd_out=acts[i] if i < n - 1 else None, grad_out=grads[i - 1] if i > 0 else None,
                stream=(ex.stream if overlap else stream)))
        self.stream = stream
        self.streams = [st.stream for st in self.stages]
        self.it_f = 0
        self.it_b = 0

    def feed(self, x_host: torch.Tensor, y_host: torch.Tensor) -> None:
        self.stages[0].stage_input(self.it_f, x_host, y_host)

    def step_forward(self) -> None:
        """Forward microbatch ``it_f`` through every non-last stage, then the last stage's
        fused forward/backward, i.e. everything that depends only on this microbatch's input."""
        it = self.it_f
        for st in self.stages[:-1]:
            st.forward(it)
        self.stages[-1].last(it)
        self.it_f += 1

    def step_backward(self) -> None:
        it = self.it_b
        for st in reversed(self.stages[:-1]):
            st.backward(it)
        self.it_b += 1

    def run(self, batches, steps: Optional[int] = None) -> int:
        """1F1B over an iterable of (x_host, y_host): warm-up ``depth`` forwards, then alternate."""
        n = 0
        for x, y in batches:
            if self.it_f - self.it_b >= self.depth:
                self.step_backward()
            self.feed(x, y)
            self.step_forward()
            n += 1
            if steps is not None and n >= steps:
                break
        while self.it_b < self.it_f:
            self.step_backward()
        return n

    def loss(self) -> torch.Tensor:
        return self.stages[-1].ex.loss_buf

    @property
    def loss_stream(self) -> torch.cuda.Stream:
        return self.stages[-1].stream

    def join(self) -> None:
        """Make ``self.stream`` (stage 1's) wait for every other stage stream (for event timing)."""
        for st in self.streams:
            if st is not self.stream:
                self.stream.wait_stream(st)

    def synchronize(self) -> None:
        for st in self.streams:
            st.synchronize()
        for s in self.stages:
            s.check()
