#!/usr/bin/env python
"""``python launch.py [--config config.yaml] [--cpu]`` — one box, one command.

Spawns the coordinator (``server.py``) and one ``client.py`` per entry of ``server.clients``
(``clients: [n1, n2, ...]``, sum <= number of GPUs), client k pinned to GPU k (SURVEY §7.1:
"client == GPU").  Cluster membership follows ``manual.cluster.infor-cluster`` when given.
"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from split_learning_b200.config import load_config      # noqa: E402
from split_learning_b200.plan import rank_assignment    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="config.yaml")
    ap.add_argument("--cpu", action="store_true", help="run every client on the CPU (plumbing / debugging)")
    ap.add_argument("--algorithm", default=None)
    ap.add_argument("--timeout", type=float, default=0.0)
    ap.add_argument("--dry-run", action="store_true", help="print the rank -> (stage, cluster, device) plan and exit")
    args = ap.parse_args()
    cfg = load_config(args.config)
    info = cfg.infor_cluster if (cfg.cluster_mode and cfg.infor_cluster_given) else None
    ranks = rank_assignment(cfg.clients, info)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    extra = ["--algorithm", args.algorithm] if args.algorithm else []
    if args.dry_run:
        try:
            import torch
            ngpu = torch.cuda.device_count()
        except Exception:
            ngpu = 0
        print(f"model {cfg.model}/{cfg.data_name}  algorithm {args.algorithm or cfg.b200.get('algorithm', 'main')}  "
              f"data-plane {cfg.b200.get('data-plane', 'host')}  clients {list(cfg.clients)}  cuts {cfg.cluster_cut_layers if cfg.cluster_mode else cfg.no_cluster_cut_layers}")
        for r, (layer_id, cluster, i) in enumerate(ranks):
            dev = "cpu" if (args.cpu or ngpu == 0) else f"cuda:{r % ngpu}"
            print(f"  rank {r}: stage {layer_id}  cluster {cluster}  member {i}  device {dev}")
        return
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "server.py"), "--config", args.config] + extra, env=env)]
    time.sleep(0.5)
    try:
        import torch
        ngpu = torch.cuda.device_count()
    except Exception:
        ngpu = 0
    for r, (layer_id, cluster, _i) in enumerate(ranks):
        dev = "cpu" if (args.cpu or ngpu == 0) else f"cuda:{r % ngpu}"
        cmd = [sys.executable, os.path.join(ROOT, "client.py"), "--layer_id", str(layer_id), "--device", dev,
               "--config", args.config] + extra
        if info is not None:
            cmd += ["--cluster", str(cluster)]
        procs.append(subprocess.Popen(cmd, env=dict(env, LOCAL_RANK=str(r))))
    rc = 0
    deadline = time.time() + args.timeout if args.timeout > 0 else None
    try:
        for p in procs:
            left = None if deadline is None else max(1.0, deadline - time.time())
            rc |= p.wait(timeout=left)
    except (KeyboardInterrupt, subprocess.TimeoutExpired):
        rc = 1
    finally:
        for p in procs:
            if p.poll() is None:
                p.terminate()
    sys.exit(rc)


if __name__ == "__main__":
    main()
