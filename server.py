#!/usr/bin/env python
"""``python server.py`` — coordinator entry point (reference server.py:1-31).

Reads ``config.yaml`` from the CWD, starts the in-box broker on ``b200.port`` (the native ``slb_broker`` daemon; this
replaces the external RabbitMQ server), purges stale queues, installs the SIGINT handler and serves.
"""
import argparse
import signal
import sys

from split_learning_b200.algorithms import server_class
from split_learning_b200.config import load_config
from split_learning_b200.transport import make_broker
from split_learning_b200.transport.broker import broker_token, delete_old_queues

parser = argparse.ArgumentParser(description="Split learning framework with controller.")
parser.add_argument("--config", default="config.yaml")
parser.add_argument("--algorithm", default=None, help="main|vanilla_sl|cluster_fsl|dcsl|flex|2ls")
args = parser.parse_args()


def main():
    cfg = load_config(args.config)
    if args.algorithm:
        cfg.b200["algorithm"] = args.algorithm
    # a non-loopback ``rabbit.address`` needs ``b200.broker-token`` (or SLB200_BROKER_TOKEN): the broker refuses otherwise
    broker = make_broker(cfg.raw.get("rabbit", {}).get("address", "127.0.0.1"), int(cfg.b200.get("port", 29777)),
                         str(cfg.b200.get("broker", "native")), token=broker_token(cfg))
    channel = broker.channel()

    def on_sigint(sig, frame):
        print("\nCatch stop signal Ctrl+C. Stop the program.")
        delete_old_queues(channel)
        broker.close()
        sys.exit(0)

    signal.signal(signal.SIGINT, on_sigint)
    delete_old_queues(channel)
    server = server_class(cfg.b200.get("algorithm", "main"))(cfg, channel)
    server.start()
    broker.close()


if __name__ == "__main__":
    main()
