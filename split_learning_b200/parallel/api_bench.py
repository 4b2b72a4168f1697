"""The benchmark through the public API: a real coordinator (``Server``) and real clients (``DeviceRpcClient``: REGISTER →
START → READY → SYN → training → NOTIFY → PAUSE → FedAvg → UPDATE) talking over the native broker daemon — exactly what
``python server.py`` + ``python client.py --layer_id k`` (or ``launch.py``) run, started here from the torchrun ranks.

Placement = the ring of ``bench.py``: rank r hosts a first-stage client and a last-stage client on GPU r; the REGISTER
``rank`` fields make the server pair stage-1 client r with the stage-2 client on GPU r+1.  Every first-stage client owns
K microbatches of pinned host memory per round (H2D copy of every microbatch inside the training loop), the last stage
copies every step's loss back to pinned host memory.  Timed per round: CUDA events around each client's training loop
(``device_ms``, max over clients) and the server's wall clock for the whole round (START … last UPDATE, FedAvg included).
"""
from __future__ import annotations

import os
import threading
import time
import traceback
import uuid
from typing import List, Optional

import torch


def api_config(n_chains: int, steps: int, batch: int, depth: int, cut, precision: str, port: int, rounds: int, workdir: str,
               clusters: Optional[dict] = None):
    from ..config import normalize
    cuts = list(cut) if isinstance(cut, (list, tuple)) else [cut]
    clients = [n_chains] * (len(cuts) + 1)
    raw = {
        "name": "Split Learning",
        "server": {
            "global-round": rounds, "clients": clients, "auto-mode": False, "model": "VGG16", "data-name": "CIFAR10",
            "parameters": {"load": False, "save": True}, "validation": False,
            "data-distribution": {"non-iid": False, "num-sample": steps * batch, "num-label": 10, "dirichlet": {"alpha": 1},
                                  "refresh": False},
            "random-seed": 1,
            "manual": {"cluster-mode": False, "no-cluster": {"cut-layers": cuts},
                       "cluster": {"num-cluster": 1, "cut-layers": [cuts], "infor-cluster": [clients]}},
            "cluster-selection": {"num-cluster": 1, "algorithm-cluster": "KMeans", "selection-mode": False},
        },
        "rabbit": {"address": "127.0.0.1", "username": "admin", "password": "admin", "virtual-host": "/"},
        "log_path": workdir, "debug_mode": False,
        "learning": {"learning-rate": 0.0005, "weight-decay": 0.01, "momentum": 0.5, "batch-size": batch, "control-count": depth,
                     "precision": precision},
        "b200": {"synthetic-data": True, "data-plane": "device", "port": port, "watchdog-seconds": 180, "precision": precision,
                 "profile-host": os.environ.get("SLB200_PROFILE_HOST", "0") == "1"},
    }
    if clusters:
        raw["server"]["manual"].update(clusters["manual"])
        raw["server"]["clients"] = clusters["clients"]
        raw["server"]["data-distribution"].update(clusters.get("data-distribution", {}))
    return normalize(raw)


def scenario(name: str, world: int, rank: int, args, port: int, workdir: str, K: int):
    """(config, roles of this rank) of a named BASELINE.json configuration on ``world`` GPUs, one client per GPU:
      split        #2 / #3  two stages cut 7, clients [world/2, world/2], server FedAvg per stage
      clusters     #4       two clusters, cut 7 and cut 14, (world/4 + world/4) clients each
      three-stage  #5       three stages cut [5, 10], clients [world/2, world/4, world/4], non-IID rate 0.5"""
    from ..plan import rank_assignment
    rounds = int(getattr(args, "rounds", 3))
    if name == "split":
        n = world // 2
        cfg = api_config(n, K, args.batch, args.depth, 7, args.precision, port, rounds, workdir)
        ranks = rank_assignment([n, n], None)
    elif name == "clusters":
        q = max(1, world // 4)
        cl = {"clients": [2 * q, 2 * q],
              "manual": {"cluster-mode": True, "cluster": {"num-cluster": 2, "cut-layers": [[7], [14]], "infor-cluster": [[q, q], [q, q]]}}}
        cfg = api_config(2 * q, K, args.batch, args.depth, 7, args.precision, port, rounds, workdir, clusters=cl)
        ranks = rank_assignment([2 * q, 2 * q], [[q, q], [q, q]])
    elif name == "three-stage":
        a, b = max(1, world // 2), max(1, world // 4)
        cl = {"clients": [a, b, b], "manual": {"cluster-mode": False, "no-cluster": {"cut-layers": [5, 10]}},
              "data-distribution": {"non-iid": True, "non-iid-rate": 0.5}}
        cfg = api_config(a, K, args.batch, args.depth, [5, 10], args.precision, port, rounds, workdir, clusters=cl)
        ranks = rank_assignment([a, b, b], None)
    else:
        raise ValueError(name)
    layer_id, cluster, _ = ranks[rank]
    clustered = name == "clusters"
    return cfg, [(layer_id, rank, cluster if clustered else -1)]


def run_api(args, roles=None, cfg=None) -> dict:
    """Collective over the torchrun ranks (or a single process for N = 1).  ``roles``: [(layer_id, register-rank, cluster)] of
    the clients this rank hosts; default = ring placement of a two-stage chain."""
    import tempfile
    from ..algorithms import client_class, server_class
    from ..runner import DEFAULT_PROFILE
    from ..transport import connect, make_broker
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = f"cuda:{local_rank}"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device(dev))
    # microbatches per client and round (IID split over 10 labels -> a multiple of 5).  A round is a user-level unit of work:
    # it is never shorter than 100 microbatches here, so that a short ``--steps`` (the kernel-pipeline loop times exactly
    # that many) does not turn the end-to-end number into a measurement of the round's pipeline fill and drain.
    K = max(int(args.steps), int(os.environ.get("SLB200_API_MIN_STEPS", "100")))
    K = K - K % 5 or 5
    rounds = int(getattr(args, "rounds", 3))
    port = 29800 + (int(os.environ.get("MASTER_PORT", "0")) % 150)
    workdir = tempfile.mkdtemp(prefix="slb200_api_")
    name = getattr(args, "scenario", None) or "ring"
    if cfg is None and name != "ring":
        cfg, roles = scenario(name, world, rank, args, port, workdir, K)
    if cfg is None:
        cfg = api_config(world, K, args.batch, args.depth, args.cut, args.precision, port, rounds, workdir)
    if roles is None:
        roles = [(1, rank, -1), (2, world + (rank - 1) % world, -1)]
    errors: List[BaseException] = []
    server = broker = None
    threads = []

    def guard(fn):
        def run():
            try:
                fn()
            except BaseException as e:          # noqa
                traceback.print_exc()
                errors.append(e)
        return run
    if rank == 0:
        broker = make_broker("127.0.0.1", port, str(cfg.b200.get("broker", "native")))
        server = server_class("main")(cfg, broker.channel(), workdir=workdir)
        threads.append(threading.Thread(target=guard(lambda: server.start(idle_timeout=args.timeout)), daemon=True, name="server"))
    if dist is not None:
        dist.barrier()
    clients = []
    for layer_id, reg_rank, cluster in roles:
        ch = connect("127.0.0.1", port)
        cli = client_class("main", cfg.b200)(str(uuid.uuid4()), layer_id, ch, device=dev, b200_opts=cfg.b200, rank=reg_rank)
        clients.append(cli)

        def body(cli=cli, cluster=cluster):
            torch.cuda.set_device(local_rank)
            cli.register(dict(DEFAULT_PROFILE), cluster)
            cli.wait_response(idle_timeout=args.timeout)
        threads.append(threading.Thread(target=guard(body), daemon=True, name=f"client-l{layer_id}"))
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    limit = min(float(args.timeout), float(os.environ.get("SLB200_API_BENCH_TIMEOUT", "420")))      # for the whole public-API run
    for t in threads:
        t.join(max(1.0, limit - (time.perf_counter() - t0)))
    wall = time.perf_counter() - t0
    for cli in clients:                      # background checkpoint uploads of the last round: finish before the broker goes
        th = cli.__dict__.get("_ckpt_thread")
        if th is not None:
            th.join(60.0)
    ok = not errors and not any(t.is_alive() for t in threads)
    flag = torch.tensor([float(ok)], device=dev)
    if dist is not None:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.barrier()
    if broker is not None:
        broker.close()
    if rank != 0:
        return {}
    if flag.item() < 0.5:
        return {"error": (repr(errors[0]) if errors else "a role did not finish")[:300]}
    hist = server.history
    n_first = cfg.clients[0]
    per_round = []
    for h in hist:
        images = (h.get("first_stage_microbatches") or n_first * K) * args.batch
        dms = h.get("device_ms")
        per_round.append({"round": h["round"], "ok": h["ok"], "wall_ms": h["seconds"] * 1e3, "device_ms": dms,
                          "overhead_ms": (h["seconds"] * 1e3 - dms) if dms else None, "train_loss": h.get("train_loss"),
                          "phases_ms": h.get("phases_ms"), "client_timing_ms": h.get("client_timing_ms"),
                          "images_per_s_device": images / (dms / 1e3) if dms else None,
                          "images_per_s_round": images / h["seconds"]})
    steady = per_round[1:] or per_round
    best = min(steady, key=lambda r: r["wall_ms"])
    ckpt = os.path.exists(os.path.join(workdir, "VGG16_CIFAR10.pth"))
    from ..checkpoint import load_meta
    ckpt_round = load_meta(os.path.join(workdir, "VGG16_CIFAR10.pth")).get("round") if ckpt else None
    return {"scenario": name, "cut_layers": cfg.cluster_cut_layers if cfg.cluster_mode else cfg.no_cluster_cut_layers,
            "non_iid_rate": cfg.non_iid_rate, "path": "native broker daemon + Server + DeviceRpcClient FSMs (REGISTER/START/READY/SYN/NOTIFY/PAUSE/UPDATE), device data plane",
            "clients": list(cfg.clients), "microbatches_per_client_per_round": K, "rounds": per_round, "steady_round": best,
            "checkpoint_written": ckpt, "checkpoint_round": ckpt_round, "total_wall_s": wall,
            "h2d_bytes_per_step": n_first * (args.batch * 3 * 32 * 32 * 4 + args.batch * 8), "d2h_bytes_per_step": n_first * 16}
