"""Peer-memory data plane: mailboxes (cut edges), device pipelines, NVLink FedAvg, multi-GPU runner."""
