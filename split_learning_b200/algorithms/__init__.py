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


def client_class(name: str = "main", opts=None):
    """``opts``: the ``b200`` config section; ``data-plane: device`` selects the peer-memory client for the main
    algorithm (activations/gradients through NVLink mailboxes instead of broker queues)."""
    from ..client import RpcClient
    name = (name or "main").lower()
    if name == "main":
        if opts and str(opts.get("data-plane", "host")).lower() == "device":
            from ..parallel.device_client import DeviceRpcClient
            return DeviceRpcClient
        return RpcClient
    if (opts and str(opts.get("data-plane", "host")).lower() == "device" and opts.get("device-variants", False)
            and name in ("vanilla_sl", "cluster_fsl")):
        # EXPERIMENTAL, opt-in (b200.device-variants: true): not yet verified on hardware — see parallel/device_variants.py
        from ..parallel.device_variants import DEVICE_CLIENTS      # sequential variants over the ticket ring
        return DEVICE_CLIENTS[name]
    from . import variants
    return variants.CLIENTS[name]


ALGORITHMS = ("main", "vanilla_sl", "cluster_fsl", "dcsl", "flex", "2ls")
