#!/usr/bin/env python
"""``python client.py --layer_id N [--device D] [--cluster C]`` (reference client.py:1-62).

Variant flags are accepted too: ``--idx/--incluster/--outcluster`` (2LS), ``--c/--s`` (FLEX).
Unlike the reference (quirk C6) a missing ``profiling.json`` is not fatal in manual mode: a
neutral profile is registered instead.
"""
import argparse
import json
import os
import uuid

import torch

from split_learning_b200.algorithms import client_class
from split_learning_b200.config import load_config
from split_learning_b200.log import print_with_color
from split_learning_b200.runner import DEFAULT_PROFILE
from split_learning_b200.transport import connect

parser = argparse.ArgumentParser(description="Split learning framework")
parser.add_argument("--layer_id", type=int, required=True, help="ID of layer, start from 1")
parser.add_argument("--device", type=str, required=False, help="Device of client")
parser.add_argument("--cluster", "--c", dest="cluster", type=int, required=False, help="ID cluster by device")
parser.add_argument("--s", dest="select", type=int, required=False, help="FLEX: 1 = selected device")
parser.add_argument("--idx", type=int, required=False, help="2LS: device index")
parser.add_argument("--incluster", type=int, required=False, help="2LS: in-cluster id")
parser.add_argument("--outcluster", type=int, required=False, help="2LS: out-cluster id")
parser.add_argument("--config", default="config.yaml")
parser.add_argument("--algorithm", default=None)
args = parser.parse_args()


def main():
    cfg = load_config(args.config)
    if args.algorithm:
        cfg.b200["algorithm"] = args.algorithm
    if args.device is None:
        rank = int(os.environ.get("LOCAL_RANK", 0))
        device = f"cuda:{rank}" if torch.cuda.is_available() else "cpu"
    else:
        device = args.device
    print(f"Using device: {device}")
    from split_learning_b200.transport.broker import broker_token
    channel = connect(cfg.raw.get("rabbit", {}).get("address", "127.0.0.1"), int(cfg.b200.get("port", 29777)),
                      token=broker_token(cfg))
    client_id = uuid.uuid4()
    profile = dict(DEFAULT_PROFILE)
    if os.path.exists("profiling.json"):
        print_with_color("Exists profiling.json.", "green")
        with open("profiling.json", "r", encoding="utf-8") as f:
            profile = json.load(f)
    elif cfg.auto_mode:
        print_with_color("[>>>] Profiling file is not existing, break", "yellow")
        return
    extra = {k: v for k, v in (("idx", args.idx), ("in_cluster", args.incluster),
                               ("out_cluster", args.outcluster), ("select", args.select)) if v is not None}
    client = client_class(cfg.b200.get("algorithm", "main"), cfg.b200)(client_id, args.layer_id, channel, device, b200_opts=cfg.b200)
    print_with_color("[>>>] Client sending registration message to server...", "red")
    client.register(profile, -1 if args.cluster is None else args.cluster, **extra)
    client.wait_response()


if __name__ == "__main__":
    main()
