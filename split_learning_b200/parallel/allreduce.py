"""Round-end FedAvg as a device-resident all-reduce over NVSwitch peer memory.

Reference: every client pickles its stage state-dict to the server (UPDATE), the server runs a weighted mean per stage
and cluster on the CPU, then an unweighted mean of the per-cluster full models, and ships the result back inside the
next START (src/Server.py:173-210,398-434, src/Utils.py:35-66).  Here the replicas keep their flat fp32 masters, BN
running statistics and integer counters in IPC-exported buffers (``B200Executor.P / S / I``); at round end every
replica launches ``fedavg_allreduce_kernel`` (ops/csrc/allreduce.cu) once per *segment*:

* a segment is a maximal run of parameters that the same set of replicas holds contiguously — with one cut point per
  cluster that is simply "the whole stage"; with clusters cut at different layers (BASELINE config #4: cuts 7 / 14)
  layers 8-14 are averaged between cluster 0's second stage and cluster 1's first stage, etc.;
* coefficient of replica q = w_q / (sum of w over the replicas of q's cluster that hold the segment) / #clusters — the
  reference's two-level mean in one pass; w = microbatch count (src/train/VGG16.py:109);
* two-shot schedule: replica r reduces slice r (peer loads) and stores it into every holder's buffer (peer stores);
* weights, NaN votes and barriers are device words in each replica's exported sync block: after the one-time handle
  exchange (``setup``) a round involves no host message at all.

Every replica therefore ends the round holding the global model for its own layers, in place — the next START carries
no parameters ("resident").
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ..ops import native as N
from .mailbox import alloc_exportable, open_exported

KINDS = ("P", "S", "I")          # parameters, BN running statistics, integer counters (float mirrors, rounded)
MAX_PARTICIPANTS = 16


@dataclass
class Member:
    """One replica (client) taking part in the aggregation."""
    uid: str
    cluster: int
    start_layer: int
    end_layer: int
    handles: Dict[str, bytes]      # "P" / "S" / "I" / "sync" -> CUDA IPC handle
    n: Dict[str, int]              # floats in each buffer
    gpu: str = ""                  # physical GPU (uuid): replicas sharing a GPU must all fit on it at once (see run)

    def wire(self) -> dict:
        return {"uid": self.uid, "cluster": self.cluster, "start": self.start_layer, "end": self.end_layer,
                "handles": self.handles, "n": self.n, "gpu": self.gpu}

    @staticmethod
    def from_wire(d: dict) -> "Member":
        return Member(str(d["uid"]), int(d["cluster"]), int(d["start"]), int(d["end"]), dict(d["handles"]), dict(d["n"]),
                      str(d.get("gpu", "")))


@dataclass
class Segment:
    kind: str
    first_key: str
    holders: List[int]             # global participant indices
    offs: List[int]                # float offset of the segment in every holder's buffer
    n: int                         # floats
    index: int = 0                 # position in the global segment order (epoch numbering)


def build_segments(members: Sequence[Member], layouts: Sequence[Dict[str, Dict[str, Tuple[int, int]]]],
                   key_order: Sequence[str]) -> List[Segment]:
    """Global, deterministic segment list (identical on every replica): walk the model's keys in layer order and merge
    consecutive keys held contiguously by the same replicas."""
    segs: List[Segment] = []
    for kind in KINDS:
        cur: Optional[Segment] = None
        for key in key_order:
            holders = [q for q in range(len(members)) if key in layouts[q][kind]]
            if not holders:
                continue
            offs = [layouts[q][kind][key][0] for q in holders]
            n = layouts[holders[0]][kind][key][1]
            if any(layouts[q][kind][key][1] != n for q in holders):
                raise ValueError(f"replicas disagree on the size of {key}")
            if (cur is not None and cur.holders == holders and all(o == co + cur.n for o, co in zip(offs, cur.offs))):
                cur.n += n
            else:
                cur = Segment(kind, key, holders, offs, n)
                segs.append(cur)
    for i, s in enumerate(segs):
        s.index = i
    return segs


def model_key_order(model_cls) -> List[str]:
    """Keys of the full model in layer order (parameters, running statistics, counters)."""
    from ..train.b200_executor import flat_layouts
    full = flat_layouts(model_cls, 0, model_cls.num_layers())
    return list(full["P"]) + list(full["S"]) + list(full["I"])


class DeviceFedAvg:
    """The aggregation group of one replica.  ``setup`` is collective (one all-gather of handles through ``comm``: a
    ``TorchDistComm`` / ``BrokerComm`` from parallel/fedavg.py, or any object with ``all_gather_object``); ``run`` is a
    pure device operation."""

    def __init__(self, ex, uid: str, cluster: int, comm, spin_limit: Optional[int] = None):
        self.ex, self.uid, self.cluster, self.comm = ex, str(uid), int(cluster), comm
        self.device = ex.device
        import os
        # a participant that died before the round end must not park the survivors for minutes: same bound as the mailbox waits
        self.spin_limit = int(spin_limit if spin_limit is not None else os.environ.get("SLB200_WAIT_SPINS", str(1 << 28)))
        raw, self.sync_handle, self.sync_ptr = alloc_exportable(4096, self.device)
        self.sync = raw.view(torch.int32)
        self.round = 0
        self.members: List[Member] = []
        self.segments: List[Segment] = []
        self.mine: List[Segment] = []
        self.ptrs: Dict[Tuple[int, str], int] = {}
        self.last_ms: Optional[float] = None

    # ------------------------------------------------------------------ collective setup
    def setup(self) -> None:
        from ..train.b200_executor import flat_layouts
        ex = self.ex
        me = Member(self.uid, self.cluster, ex.start_layer, ex.end_layer,
                    {"P": ex.P_handle, "S": ex.S_handle, "I": ex.I_handle, "sync": self.sync_handle},
                    {"P": ex.n_params, "S": ex.n_stats, "I": ex.n_ints},
                    gpu=str(torch.cuda.get_device_properties(self.device).uuid))
        wires = self.comm.all_gather_object(me.wire())
        self.members = sorted((Member.from_wire(w) for w in wires), key=lambda m: m.uid)
        if len(self.members) > MAX_PARTICIPANTS:
            raise RuntimeError(f"device FedAvg supports up to {MAX_PARTICIPANTS} replicas, got {len(self.members)}")
        self.me = [m.uid for m in self.members].index(self.uid)
        layouts = [flat_layouts(ex.model_cls, m.start_layer, m.end_layer) for m in self.members]   # end_layer is resolved (never -1)
        mine = layouts[self.me]
        if (sum(n for _, n in mine["P"].values()) > ex.n_params or len(mine["P"]) != len(ex.entries)):
            raise RuntimeError("layout mismatch between flat_layouts() and the executor")
        self.segments = build_segments(self.members, layouts, model_key_order(ex.model_cls))
        self.mine = [s for s in self.segments if self.me in s.holders]
        for q, m in enumerate(self.members):
            for kind in KINDS + ("sync",):
                self.ptrs[(q, kind)] = (self.sync_ptr if kind == "sync" else getattr(ex, kind).data_ptr()) if q == self.me \
                    else open_exported(m.handles[kind], self.device)
        self.clusters = sorted({m.cluster for m in self.members})
        # Every kernel of a round spins (thread 0 of each CTA) until its peers' kernels have started.  Replicas that share
        # one physical GPU (threads / processes of a single-GPU run) must therefore be resident together: split the SM's
        # CTA slots (4 x 512 threads) between them, with slack.  One replica per GPU gets the full 2 CTAs per SM.
        coloc = max(1, sum(1 for m in self.members if m.gpu == self.members[self.me].gpu))
        sms = N.num_sms(self.device)
        self.grid = 2 * sms if coloc == 1 else max(1, int(4 * sms / 1.5 / coloc))

    # ------------------------------------------------------------------ one round
    def run(self, weight: float, ok: bool = True, timed: bool = False) -> bool:
        """Average in place.  Returns False when the round was skipped (a replica voted ``ok=False`` — NaN loss — or a
        peer never showed up); the parameters are then untouched."""
        ex = self.ex
        lib = N.lib()
        # The replica's own (non-blocking) stream: it already orders the round's last optimizer step before us, and —
        # unlike the legacy default stream, which every thread of a process shares — a kernel spinning here for its peers
        # can never sit in front of a co-located peer's kernel in the same queue.
        stream = ex.stream
        with torch.cuda.stream(stream):
            return self._run_on(stream, lib, weight, ok, timed)

    def _run_on(self, stream, lib, weight: float, ok: bool, timed: bool) -> bool:
        ex = self.ex
        for bn, st in ex.bn_state.items():                              # integer counters -> float mirrors
            o = ex.int_entries[f"layer{bn}.num_batches_tracked"][0]
            ex.I[o:o + 1].copy_(st["num_batches_tracked"].to(torch.float32).reshape(1))
        e0 = e1 = None
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        nseg = max(len(self.segments), 1)
        for s in self.mine:
            k = len(s.holders)
            bufs = (ctypes.c_void_p * k)(*[ctypes.c_void_p(self.ptrs[(q, s.kind)] + 4 * o) for q, o in zip(s.holders, s.offs)])
            syncs = (ctypes.c_void_p * k)(*[ctypes.c_void_p(self.ptrs[(q, "sync")]) for q in s.holders])
            gids = (ctypes.c_int * k)(*s.holders)
            cl = (ctypes.c_int * k)(*[self.members[q].cluster for q in s.holders])
            ncl = len({self.members[q].cluster for q in s.holders})
            epoch = self.round * nseg + s.index + 1
            rc = lib.slb_fedavg_allreduce(bufs, syncs, gids, cl, ctypes.c_int(k), ctypes.c_int(s.holders.index(self.me)),
                                          ctypes.c_int(ncl), ctypes.c_longlong(s.n), ctypes.c_longlong(0),
                                          ctypes.c_longlong(s.n if s.kind == "I" else 0), ctypes.c_uint32(epoch),
                                          ctypes.c_float(float(weight)), ctypes.c_int(int(bool(ok))),
                                          ctypes.c_uint64(self.spin_limit), ctypes.c_int(self.grid),
                                          ctypes.c_void_p(stream.cuda_stream))
            N._check(rc, "fedavg_allreduce", 2)
        self.round += 1
        if ex.PB is not None:
            N.cast_f32_bf16(ex.P, ex.PB)                                   # bf16 mode: refresh the weight shadow
        if timed:
            e1.record(stream)
        stream.synchronize()
        if timed:
            self.last_ms = e0.elapsed_time(e1)
        flags = self.sync[34:38].tolist()
        if flags[2]:                                                       # [36]: a peer never arrived
            self.sync[36] = 0
            raise TimeoutError("device FedAvg: a replica never reached the all-reduce (dead peer?)")
        done = bool(flags[1])
        if done:
            for bn, st in ex.bn_state.items():
                o = ex.int_entries[f"layer{bn}.num_batches_tracked"][0]
                st["num_batches_tracked"].copy_(ex.I[o].round().to(torch.int64))
        return done

    def link_bytes(self) -> int:
        """Bytes this replica moves over NVLink per round (in + out), for roofline accounting."""
        tot = 0
        for s in self.mine:
            k = len(s.holders)
            tot += 2 * (k - 1) * (4 * s.n // k)
        return tot
