"""Algorithm variants: the main tree plus the five ``other/*`` forks, as (Server, Client)
policy pairs over one shared skeleton."""
from __future__ import annotations


def server_class(name: str = "main"):
    from ..server import Server
    name = (name or "main").lower()
    if name == "main":
        return Server
    from . import variants
    return variants.SERVERS[name]


def client_class(name: str = "main"):
    from ..client import RpcClient
    name = (name or "main").lower()
    if name == "main":
        return RpcClient
    from . import variants
    return variants.CLIENTS[name]


ALGORITHMS = ("main", "vanilla_sl", "cluster_fsl", "dcsl", "flex", "2ls")
