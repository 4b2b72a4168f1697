"""One-process-per-GPU device pipeline: wiring of peer mailboxes over CUDA IPC + the
multi-GPU arm of ``bench.py``.

Rank layout for ``clients: [n, n]`` (N = 2n GPUs): rank r < n is stage 1 of chain r, rank
r + n is its stage 2 — a 1:1 pairing of the reference's competing-consumer queue
(``intermediate_queue_{layer}_{cluster}``, src/train/VGG16.py:143-154) which is what a
round-robin consumer assignment degenerates to when both layers have the same replica count.
``torch.distributed`` (NCCL) is used only to bootstrap (exchange 64-byte IPC handles,
barriers, the max-over-ranks of the measured time); activations and gradients never go
through it.
"""
from __future__ import annotations

import os
import time
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from ..models import VGG16_CIFAR10
from ..train.b200_executor import B200Executor
from .mailbox import Mailbox, MailboxSpec
from .pipeline import DeviceStage, act_spec


def init_dist(device: torch.device) -> None:
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)


def exchange_handles(my: Dict[str, bytes]) -> List[Dict[str, bytes]]:
    out: List[Optional[Dict[str, bytes]]] = [None] * dist.get_world_size()
    dist.all_gather_object(out, my)
    return out  # type: ignore


def build_chain_stage(rank: int, world: int, cut: int, batch: int, depth: int, learning: dict, device,
                      use_graphs: bool = True) -> DeviceStage:
    """Create this rank's stage of a 2-stage chain and wire it to its peer."""
    n = world // 2
    first = rank < n
    peer = rank + n if first else rank - n
    torch.manual_seed(1000 + (rank % n))                 # both ends of a chain derive weights from the chain id
    if first:
        ex = B200Executor(VGG16_CIFAR10(0, cut), "VGG16", learning, device, is_first=True, use_graphs=use_graphs)
    else:
        ex = B200Executor(VGG16_CIFAR10(cut, 52), "VGG16", learning, device, is_last=True, use_graphs=use_graphs)
    # geometry of the cut edge is defined by the producing (first) stage; the last stage derives it from its input
    if first:
        spec = act_spec(ex, batch, depth)
    else:
        c, h, w = ex.in_shape
        spec = MailboxSpec(depth, batch, (batch, h, w, c), itemsize=4 if ex.fp32 else 2)
    own, handle = Mailbox.allocate_exportable(spec, device)          # grads (stage 1) / activations (stage 2)
    handles = exchange_handles({"mb": handle})
    remote = Mailbox.open_peer(spec, handles[peer]["mb"], device)
    if first:
        st = DeviceStage(ex, batch, depth, fwd_in=None, grad_in=own, fwd_out=remote, grad_out=None)
    else:
        st = DeviceStage(ex, batch, depth, fwd_in=own, grad_in=None, fwd_out=None, grad_out=remote)
    dist.barrier()
    return st


def run_steps(st: DeviceStage, n_steps: int, batches=None, loss_host: Optional[torch.Tensor] = None) -> None:
    """Static 1F1B schedule for one rank (``batches``: iterable of pinned (x, y) for stage 1)."""
    first = st.ex.is_first
    if first:
        it_b = 0
        src = iter(batches) if batches is not None else None
        for it in range(n_steps):
            if it - it_b >= st.depth:
                st.backward(it_b)
                it_b += 1
            if src is not None:
                x, y = next(src)
                st.stage_input(it, x, y)
            st.forward(it)
        while it_b < n_steps:
            st.backward(it_b)
            it_b += 1
    else:
        for it in range(n_steps):
            st.last(it)
            if loss_host is not None:
                with torch.cuda.stream(st.stream):
                    loss_host.copy_(st.ex.loss_buf, non_blocking=True)


def fedavg_round(st: DeviceStage, rank: int, world: int, dev) -> Optional[dict]:
    """End-of-round aggregation as the reference does it once per round (src/Server.py:398-434) — here in place over
    NVLink: every replica of a stage averages its flat fp32 parameters (+ BN running statistics) with its peers,
    weights = microbatch counts.  Not part of the timed training steps (the reference aggregates outside the epoch
    loop as well); its device time is reported separately.  Collective: every rank must call."""
    from .fedavg import PeerFedAvg, average_int_state
    n = world // 2
    first_ranks, last_ranks = list(range(n)), list(range(n, world))
    g_first = dist.new_group(first_ranks)
    g_last = dist.new_group(last_ranks)
    if n < 2:
        return None
    ex = st.ex
    mine, grp = (first_ranks, g_first) if ex.is_first else (last_ranks, g_last)
    stats = [t for bn in ex.bn_state.values() for t in (bn["running_mean"], bn["running_var"])]
    n_stats = sum(t.numel() for t in stats)
    flat_stats = torch.cat([t.reshape(-1) for t in stats]) if stats else torch.zeros(0, device=dev)
    pad = (-n_stats) % 4
    flat_stats = torch.cat([flat_stats, torch.zeros(pad, device=dev)]) if pad else flat_stats
    fa = PeerFedAvg(ex.n_params, dev, mine, group=grp)
    fs = PeerFedAvg(max(flat_stats.numel(), 4), dev, mine, group=grp)
    weight = float(st._posted["F"] + st._posted["L"] or 1)      # microbatches processed (the reference's FedAvg weight)
    ok = not ex.nan_detected()
    done = fa.average(ex.P, ex.PB, weight, ok=ok)         # the aggregation itself (first call also pays NCCL sub-group setup)
    torch.cuda.synchronize()
    dist.barrier(group=grp)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fa.average(ex.P, ex.PB, weight, ok=ok)                # timed repeat (idempotent: replicas are already equal)
    if flat_stats.numel():
        fs.average(flat_stats, None, weight, ok=ok)
        o = 0
        for t in stats:
            t.copy_(flat_stats[o:o + t.numel()].view_as(t))
            o += t.numel()
    average_int_state({f"nbt{i}": bn["num_batches_tracked"] for i, bn in ex.bn_state.items()}, weight, group=grp)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # replicas of a stage must now hold identical parameters
    chk = ex.P[:1024].clone()
    ref = chk.clone()
    dist.broadcast(ref, src=mine[0], group=grp)
    same = torch.tensor([float(torch.equal(chk, ref))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    return {"ms_max_over_ranks": float(ms.item()), "replicas_per_stage": n, "aggregated": bool(done),
            "stage2_param_bytes": 4 * ex.n_params if not ex.is_first else None, "replicas_identical": bool(same.item() > 0.5),
            "note": "in-place NVLink peer-load FedAvg (params + BN statistics) incl. staging copy, weight exchange and 2 barriers "
                    "per buffer; steady-state repeat; outside the timed training steps"}


def bench_multi_gpu(args) -> dict:
    from bench import DTYPE_LABEL, ClockSampler, synthetic_batches     # bench.py is the entry script (repo root on sys.path)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus or world % 2:
        raise SystemExit(f"bench: launch with torchrun --nproc-per-node {args.gpus} (even), got WORLD_SIZE={world}")
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    init_dist(dev)
    W, K, B, depth = args.warmup, args.steps, args.batch, args.depth
    learning = {"learning-rate": 0.0005, "momentum": 0.5, "batch-size": B, "control-count": depth,
                "precision": getattr(args, "precision", "tf32")}
    st = build_chain_stage(rank, world, args.cut, B, depth, learning, dev, use_graphs=not args.no_graphs)
    n = world // 2
    pool = synthetic_batches(16, B, seed=1 + rank) if st.ex.is_first else None
    loss_host = torch.zeros(4).pin_memory()

    def batches(k):
        for i in range(k):
            yield pool[i % len(pool)]

    # setup (captures every slot graph) + warm-up
    run_steps(st, 2 * depth + 2, batches(2 * depth + 2) if pool else None)
    torch.cuda.synchronize()
    dist.barrier()
    run_steps(st, W, batches(W) if pool else None)
    torch.cuda.synchronize()
    dist.barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    results = {}
    t0 = time.perf_counter()
    for mode in ("device", "e2e"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        with torch.cuda.stream(st.stream):
            e0.record()
        if mode == "device":
            run_steps(st, K, None)                      # inputs stay resident in the device slots
        else:
            run_steps(st, K, batches(K) if pool else None, loss_host if not st.ex.is_first else None)
        with torch.cuda.stream(st.stream):
            e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        results[mode] = float(ms.item())
    t1 = time.perf_counter()
    st.check()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    fed = None
    if os.environ.get("SLB200_BENCH_FEDAVG", "1") != "0":
        fed = fedavg_round(st, rank, world, dev)          # round end: NVLink FedAvg among the replicas of each stage
    per = torch.tensor([float(sum(st.launches_per.values()))], device=dev)
    dist.all_reduce(per, op=dist.ReduceOp.SUM)
    loss = torch.tensor([float(loss_host[0]) if not st.ex.is_first else 0.0], device=dev)
    dist.all_reduce(loss, op=dist.ReduceOp.SUM)
    dist.barrier()
    dist.destroy_process_group()
    if rank != 0:
        return {}
    images = n * K * B
    return {
        "metric": "VGG16/CIFAR10 split images/sec", "value": images / (results["device"] / 1e3), "unit": "images/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": results["device"] / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_LABEL[getattr(args, "precision", "tf32")], "data": "synthetic",
        "config": {"model": "VGG16_CIFAR10", "global_batch": B * n, "microbatch": B, "seq_len": None, "cut_layers": [args.cut],
                   "clients": [n, n], "control_count": depth, "parallelism": f"pp2 x dp{n} (one GPU per stage replica)",
                   "optimizer": "SGD lr=5e-4 momentum=0.5, step per microbatch", "recompute": True,
                   "cuda_graphs": not args.no_graphs, "cut_transport": "in-kernel P2P store into peer HBM + st.release.sys flag",
                   "l2": "per-step working set ~470 MB on stage-2 GPUs > 126 MB L2; no flush needed"},
        "e2e": {"value": images / (results["e2e"] / 1e3), "unit": "images/s", "ms_per_step": results["e2e"] / K,
                "h2d_bytes_per_step": n * (B * 3 * 32 * 32 * 4 + B * 8), "d2h_bytes_per_step": n * 16},
        "gpu_launches": int(per.item()) * K, "launches_per_step": int(per.item()), "clocks": clocks,
        "final_loss": float(loss.item()) / n, "impl": "ours", "fedavg_round": fed,
    }
