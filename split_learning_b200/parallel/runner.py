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


def chain_models(chain: int, cut: int):
    """(stage-1 model, stage-2 model) of chain ``chain`` — both ends of a chain, wherever they live, and the single-GPU
    self-check replica derive their initial weights from the chain id alone (fixed construction order)."""
    torch.manual_seed(1000 + chain)
    m1 = VGG16_CIFAR10(0, cut)
    m2 = VGG16_CIFAR10(cut, 52)
    return m1, m2


def build_ring_stages(rank: int, world: int, cut: int, batch: int, depth: int, learning: dict, device,
                      use_graphs: bool = True):
    """Balanced placement for the scaling run: N GPUs host N chains; GPU r runs stage 1 of chain r *and* stage 2 of chain
    r-1, so every GPU carries one full chain's worth of work (the N = 1 configuration replicated) while every cut edge
    crosses NVLink: stage 1 on GPU r stores its activation tiles into the mailbox in GPU r+1's HBM from inside the
    cut-tail kernel, and the cut-head dgrad on GPU r+1 stores the gradient back into GPU r's mailbox.  Returns
    (first-stage DeviceStage of chain r, last-stage DeviceStage of chain r-1)."""
    nxt, prv = (rank + 1) % world, (rank - 1) % world
    m1, _ = chain_models(rank, cut)
    _, m2 = chain_models(prv, cut)
    ex1 = B200Executor(m1, "VGG16", learning, device, is_first=True, use_graphs=use_graphs)
    ex2 = B200Executor(m2, "VGG16", learning, device, is_last=True, use_graphs=use_graphs)
    spec = act_spec(ex1, batch, depth)
    grad_in, h_grad = Mailbox.allocate_exportable(spec, device)        # gradients of chain r come back here
    act_in, h_act = Mailbox.allocate_exportable(spec, device)          # activations of chain r-1 arrive here
    handles = exchange_handles({"grad": h_grad, "act": h_act})
    fwd_out = Mailbox.open_peer(spec, handles[nxt]["act"], device)     # chain r's stage 2 lives on GPU r+1
    grad_out = Mailbox.open_peer(spec, handles[prv]["grad"], device)   # chain r-1's stage 1 lives on GPU r-1
    a = DeviceStage(ex1, batch, depth, fwd_in=None, grad_in=grad_in, fwd_out=fwd_out, grad_out=None)
    b = DeviceStage(ex2, batch, depth, fwd_in=act_in, grad_in=None, fwd_out=None, grad_out=grad_out)
    dist.barrier()
    return a, b


def run_ring_steps(a: DeviceStage, b: DeviceStage, n_steps: int, batches=None, loss_host: Optional[torch.Tensor] = None,
                   start: int = 0, loss_log: Optional[List[torch.Tensor]] = None) -> None:
    """1F1B schedule of one rank of the ring: iterations ``start .. start + n_steps`` of chain r's first stage and of
    chain r-1's last stage, each on its own stream; the mailbox flags are the only synchronisation between GPUs."""
    it_b = start
    src = iter(batches) if batches is not None else None
    for it in range(start, start + n_steps):
        if it - it_b >= a.depth:
            a.backward(it_b)
            it_b += 1
        if src is not None:
            x, y = next(src)
            a.stage_input(it, x, y)
        a.forward(it)
        b.last(it)
        if loss_host is not None or loss_log is not None:
            with torch.cuda.stream(b.stream):
                if loss_log is not None:
                    h = torch.zeros(4).pin_memory()
                    h.copy_(b.ex.loss_buf, non_blocking=True)
                    loss_log.append(h)
                else:
                    loss_host.copy_(b.ex.loss_buf, non_blocking=True)
    while it_b < start + n_steps:
        a.backward(it_b)
        it_b += 1


def ring_selfcheck(rank: int, world: int, args, learning: dict, dev, ring_losses: List[float], n: int) -> Optional[dict]:
    """Cross-GPU correctness of the mailbox / flag path, exercised by every multi-GPU bench run: the first ``n`` losses of
    chain 0 (stage 1 on GPU 0, stage 2 on GPU 1, tiles and flags over NVLink) must reproduce on a single-GPU replica of the
    same chain (same initial weights, same batches, same dropout counters).  Collective."""
    from bench import synthetic_batches
    from .pipeline import LocalPipeline
    owner = 1 % world                                    # chain 0's last stage lives on GPU 1
    t = torch.zeros(n, device=dev)
    if rank == owner:
        t.copy_(torch.tensor(ring_losses[:n]))
    dist.broadcast(t, src=owner)
    if rank != 0:
        return None
    m1, m2 = chain_models(0, args.cut)
    exs = [B200Executor(m1, "VGG16", learning, dev, is_first=True, use_graphs=not args.no_graphs),
           B200Executor(m2, "VGG16", learning, dev, is_last=True, use_graphs=not args.no_graphs)]
    pipe = LocalPipeline(exs, args.batch, args.depth)
    pool = synthetic_batches(16, args.batch, seed=1)
    local = []
    for i in range(n):
        if pipe.it_f - pipe.it_b >= pipe.depth:
            pipe.step_backward()
        pipe.feed(*pool[i % len(pool)])
        pipe.step_forward()
        h = torch.zeros(4).pin_memory()
        with torch.cuda.stream(pipe.loss_stream):
            h.copy_(pipe.loss(), non_blocking=True)
        local.append(h)
    while pipe.it_b < pipe.it_f:
        pipe.step_backward()
    pipe.synchronize()
    loc = torch.tensor([float(h[0]) for h in local])
    ring = t.cpu()
    rel = float(((loc - ring).abs() / loc.abs().clamp_min(1e-6)).max())
    return {"chain": 0, "steps": n, "max_rel_loss_diff_vs_single_gpu": rel, "ok": bool(rel < 2e-3),
            "ring_losses": [round(float(v), 5) for v in ring[:6]], "single_gpu_losses": [round(float(v), 5) for v in loc[:6]]}


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
    """Split placement (clients [n, n], one stage replica per GPU): round-end aggregation among the replicas of each
    stage with the device all-reduce.  Collective: every rank must call."""
    n = world // 2
    first_ranks, last_ranks = list(range(n)), list(range(n, world))
    g_first = dist.new_group(first_ranks)
    g_last = dist.new_group(last_ranks)
    if n < 2:
        return None
    ex = st.ex
    mine, grp = (first_ranks, g_first) if ex.is_first else (last_ranks, g_last)
    weight = float(st._posted["F"] + st._posted["L"] or 1)      # microbatches processed (the reference's FedAvg weight)
    before = ex.P[:4096].clone()
    r = fedavg_stage(ex, weight, mine, dev, group=grp, tag="s1-" if ex.is_first else "s2-")
    ms = torch.tensor([r["ms"]], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    mean = before.clone()
    dist.all_reduce(mean, group=grp)
    mean /= n
    err = torch.tensor([float((ex.P[:4096] - mean).abs().max())], device=dev)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    link = torch.tensor([float(r["link_bytes"])], device=dev)
    dist.all_reduce(link, op=dist.ReduceOp.MAX)
    return {"ms_max_over_ranks": float(ms.item()), "replicas_per_stage": n, "aggregated": bool(r["aggregated"]),
            "nvlink_bytes_per_gpu_stage2": int(link.item()), "max_abs_err_vs_nccl_mean": float(err.item()),
            "algorithm": "two-shot reduce-scatter + all-gather over peer memory, in place, device-side weights/votes/barriers"}


def fedavg_stage(ex: B200Executor, weight: float, ranks: List[int], dev, group=None, tag: str = "") -> Optional[dict]:
    """Round-end aggregation of one stage among ``ranks`` with the device all-reduce (parallel/allreduce.py): one
    untimed round (handle exchange + first touch of the peer mappings), a barrier, then a timed round — CUDA events
    around the all-reduce kernels *including* both device barriers.  Collective over ``ranks``."""
    from .allreduce import DeviceFedAvg
    from .fedavg import TorchDistComm
    if len(ranks) < 2:
        return None
    fa = DeviceFedAvg(ex, f"{tag}{dist.get_rank():03d}", 0, TorchDistComm(ranks, group))
    fa.setup()
    ok = not ex.nan_detected()
    done = fa.run(weight, ok=ok)
    torch.cuda.synchronize()
    dist.barrier(group=group)
    done = fa.run(weight, ok=ok, timed=True) and done        # idempotent: the replicas are already equal
    return {"ms": fa.last_ms, "aggregated": bool(done), "link_bytes": fa.link_bytes(), "segments": len(fa.mine)}


def bench_ring(args) -> dict:
    """``bench.py --gpus N`` (default placement): N chains on N GPUs, every GPU hosting stage 1 of one chain and stage 2 of
    its neighbour's (``build_ring_stages``).  Per-GPU work is constant in N (weak scaling)."""
    from bench import DTYPE_LABEL, ClockSampler, synthetic_batches
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"bench: launch with torchrun --nproc-per-node {args.gpus}, got WORLD_SIZE={world}")
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    init_dist(dev)
    W, K, B, depth = args.warmup, args.steps, args.batch, args.depth
    learning = {"learning-rate": 0.0005, "momentum": 0.5, "batch-size": B, "control-count": depth,
                "precision": getattr(args, "precision", "tf32")}
    a, b = build_ring_stages(rank, world, args.cut, B, depth, learning, dev, use_graphs=not args.no_graphs)
    pool = synthetic_batches(16, B, seed=1 + rank)
    loss_host = torch.zeros(4).pin_memory()

    def batches(k, first=0):
        for i in range(first, first + k):
            yield pool[i % len(pool)]

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()

    # (0) setup = the first iterations (captures every slot graph) — also the cross-GPU self-check
    n_check = max(12, 2 * depth + 2)
    log: List[torch.Tensor] = []
    run_ring_steps(a, b, n_check, batches(n_check), loss_log=log)
    sync_all()
    check = None
    if not getattr(args, "no_selfcheck", False):
        check = ring_selfcheck(rank, world, args, learning, dev, [float(h[0]) for h in log], n_check)
        sync_all()
    litmus = None
    if not getattr(args, "no_selfcheck", False) and world >= 2:
        # payload-then-flag message passing between GPU 0 and GPU 1, 10^5 round trips in both directions
        from .litmus import pingpong
        lit = pingpong(rank, 0, 1, dev, iters=100_000)
        t = torch.tensor([float(lit["errors"]) if lit else 0.0, float(lit["timeout"]) if lit else 0.0,
                          float(lit["round_trip_us"]) if lit else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        litmus = {"pattern": "stores -> fence.sys -> st.release.sys | ld.acquire.sys -> loads, GPU0 <-> GPU1", "iters": 100_000,
                  "payload_words": 1024, "errors": int(t[0].item()), "timeout": bool(t[1].item()),
                  "round_trip_us": round(float(t[2].item()), 3), "ok": bool(t[0].item() == 0 and t[1].item() == 0)}
        sync_all()
    it0 = n_check
    run_ring_steps(a, b, W, batches(W, it0), start=it0)
    it0 += W
    sync_all()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    results = {}
    t0 = time.perf_counter()
    for mode in ("device", "e2e"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        torch.cuda.synchronize()
        with torch.cuda.stream(a.stream):
            e0.record()
        if mode == "device":
            run_ring_steps(a, b, K, None, start=it0)                  # inputs stay resident in the device slots
        else:
            run_ring_steps(a, b, K, batches(K, it0), loss_host, start=it0)
        it0 += K
        a.stream.wait_stream(b.stream)
        with torch.cuda.stream(a.stream):
            e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        results[mode] = float(ms.item())
    t1 = time.perf_counter()
    a.check()
    b.check()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    fed = None
    if os.environ.get("SLB200_BENCH_FEDAVG", "1") != "0" and world >= 2:
        ranks = list(range(world))
        before = b.ex.P[:4096].clone()
        r1 = fedavg_stage(a.ex, float(a._posted["F"] or 1), ranks, dev, tag="s1-")
        r2 = fedavg_stage(b.ex, float(b._posted["L"] or 1), ranks, dev, tag="s2-")
        t = torch.tensor([r1["ms"] + r2["ms"], r2["ms"]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # correctness: every replica now holds the mean of what the replicas held before (equal weights here)
        mean = before.clone()
        dist.all_reduce(mean)
        mean /= world
        err = torch.tensor([float((b.ex.P[:4096] - mean).abs().max())], device=dev)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        link = r2["link_bytes"]
        fed = {"ms_max_over_ranks": float(t[0].item()), "stage2_ms": float(t[1].item()), "replicas_per_stage": world,
               "stage2_param_bytes": 4 * b.ex.n_params, "nvlink_bytes_per_gpu_stage2": link,
               "stage2_fraction_of_770GBs_per_direction": (link / 2 / 770e9) / (float(t[1].item()) / 1e3) if t[1].item() > 0 else None,
               "max_abs_err_vs_nccl_mean": float(err.item()), "aggregated": bool(r1["aggregated"] and r2["aggregated"]),
               "algorithm": "two-shot reduce-scatter + all-gather over peer memory, in place, device-side weights/votes/barriers"}
        # the number that matters for a training round: K steps + the aggregation
        fed["round_images_per_s"] = world * K * B / ((results["device"] + fed["ms_max_over_ranks"]) / 1e3)
    per = float(sum(a.launches_per.values()) + sum(b.launches_per.values()))
    loss = torch.tensor([float(loss_host[0])], device=dev)
    dist.all_reduce(loss, op=dist.ReduceOp.SUM)
    dist.barrier()
    if rank != 0:
        return {}
    images = world * K * B
    out = {
        "metric": "VGG16/CIFAR10 split images/sec", "value": images / (results["device"] / 1e3), "unit": "images/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": results["device"] / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_LABEL[getattr(args, "precision", "tf32")], "data": "synthetic",
        "config": {"model": "VGG16_CIFAR10", "global_batch": B * world, "microbatch": B, "seq_len": None, "cut_layers": [args.cut],
                   "clients": [world, world], "control_count": depth,
                   "parallelism": f"pp2 x dp{world}: GPU r hosts stage 1 of chain r and stage 2 of chain r-1 (every cut edge crosses NVLink)",
                   "placement": "ring", "optimizer": "SGD lr=5e-4 momentum=0.5, step per microbatch", "recompute": True,
                   "cuda_graphs": not args.no_graphs, "cut_transport": "in-kernel P2P store into peer HBM + st.release.sys flag",
                   "l2": "per-step working set (fp32 master + momentum, 268 MB per stage-2 replica) > 126 MB L2; no flush needed"},
        "e2e": {"value": images / (results["e2e"] / 1e3), "unit": "images/s", "ms_per_step": results["e2e"] / K,
                "h2d_bytes_per_step": world * (B * 3 * 32 * 32 * 4 + B * 8), "d2h_bytes_per_step": world * 16},
        "gpu_launches": int(per) * world * K, "launches_per_step": int(per) * world, "clocks": clocks,
        "final_loss": float(loss.item()) / world, "impl": "ours", "fedavg_round": fed, "selfcheck": check, "litmus": litmus,
    }
    return out


def bench_multi_gpu(args) -> dict:
    if getattr(args, "placement", "ring") == "ring":
        return bench_ring(args)
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
    if rank != 0:
        return {}
    images = n * K * B
    return {
        "metric": "VGG16/CIFAR10 split images/sec", "value": images / (results["device"] / 1e3), "unit": "images/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": results["device"] / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_LABEL[getattr(args, "precision", "tf32")], "data": "synthetic",
        "config": {"model": "VGG16_CIFAR10", "global_batch": B * n, "microbatch": B, "seq_len": None, "cut_layers": [args.cut],
                   "clients": [n, n], "control_count": depth, "parallelism": f"pp2 x dp{n} (one GPU per stage replica)",
                   "placement": "split",
                   "optimizer": "SGD lr=5e-4 momentum=0.5, step per microbatch", "recompute": True,
                   "cuda_graphs": not args.no_graphs, "cut_transport": "in-kernel P2P store into peer HBM + st.release.sys flag",
                   "l2": "per-step working set ~470 MB on stage-2 GPUs > 126 MB L2; no flush needed"},
        "e2e": {"value": images / (results["e2e"] / 1e3), "unit": "images/s", "ms_per_step": results["e2e"] / K,
                "h2d_bytes_per_step": n * (B * 3 * 32 * 32 * 4 + B * 8), "d2h_bytes_per_step": n * 16},
        "gpu_launches": int(per.item()) * K, "launches_per_step": int(per.item()), "clocks": clocks,
        "final_loss": float(loss.item()) / n, "impl": "ours", "fedavg_round": fed,
    }
