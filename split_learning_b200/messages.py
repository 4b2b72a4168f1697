"""Control-plane verbs and payload builders (reference client.py:57,
src/Server.py:140-153,262-296, src/RpcClient.py:128-130, src/train/VGG16.py:121-122).

Wire format: pickled dicts, field names identical to the reference.  Additions are
strictly extra keys: ``READY`` (client → server: START has been applied; replaces the
reference's fixed ``time.sleep(25)``), ``rank`` in REGISTER (GPU ordinal), and
``resident`` in UPDATE (parameters stayed on the GPU and were averaged in place)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

RPC_QUEUE = "rpc_queue"
CKPT_QUEUE = "rpc_queue_ckpt"      # CHECKPOINT bodies (a stage state-dict, up to 134 MB) must not sit in front of a NOTIFY

REGISTER, START, SYN, NOTIFY, PAUSE, UPDATE, STOP, READY, HEARTBEAT, CHECKPOINT = (
    "REGISTER", "START", "SYN", "NOTIFY", "PAUSE", "UPDATE", "STOP", "READY", "HEARTBEAT", "CHECKPOINT")


def reply_queue(client_id) -> str:
    return f"reply_{client_id}"


def register(client_id, layer_id: int, profile: Optional[dict], cluster: int = -1, **extra) -> Dict[str, Any]:
    m = {"action": REGISTER, "client_id": client_id, "layer_id": layer_id, "profile": profile,
         "cluster": cluster, "message": "Hello from Client!"}
    m.update(extra)
    return m


def start(parameters, layers: List[int], model_name: str, data_name: str, learning: dict,
          label_count: list, refresh: bool, cluster: int, **extra) -> Dict[str, Any]:
    m = {"action": START, "message": "Server accept the connection!", "parameters": parameters,
         "layers": layers, "model_name": model_name, "data_name": data_name, "learning": learning,
         "label_count": label_count, "refresh": refresh, "cluster": cluster}
    m.update(extra)
    return m


def syn() -> Dict[str, Any]:
    return {"action": SYN, "message": "Synchronize client devices"}


def notify(client_id, layer_id: int, cluster) -> Dict[str, Any]:
    return {"action": NOTIFY, "client_id": client_id, "layer_id": layer_id,
            "message": "Finish training!", "cluster": cluster}


def pause(**extra) -> Dict[str, Any]:
    m = {"action": PAUSE, "message": "Pause training and please send your parameters", "parameters": None}
    m.update(extra)
    return m


def update(client_id, layer_id: int, result: bool, size: int, cluster, parameters, **extra) -> Dict[str, Any]:
    m = {"action": UPDATE, "client_id": client_id, "layer_id": layer_id, "result": result, "size": size,
         "cluster": cluster, "message": "Sent parameters to Server", "parameters": parameters}
    m.update(extra)
    return m


def stop(message: str = "Stop training!") -> Dict[str, Any]:
    return {"action": STOP, "message": message, "parameters": None}


def heartbeat(client_id=None, progress=None) -> Dict[str, Any]:
    """Liveness beacon (no counterpart in the reference, where a dead peer is a silent deadlock): clients publish it to
    ``rpc_queue`` every few seconds, the server relays one to every ``reply_{id}``.  Receivers only refresh their idle
    timers — waits are bounded by *silence of the peer*, not by how long a healthy round takes.  ``progress``: a counter of
    work done (messages handled + microbatches trained); the server relays the sum over all clients, and a blocked client
    keeps waiting only while that sum moves — a peer that still beacons but no longer trains (hung, or alive-but-deaf) is a
    dead-lock like any other and ends in the watchdog."""
    import time
    return {"action": HEARTBEAT, "client_id": client_id, "message": "alive", "t": time.time(), "progress": progress}


def checkpoint(client_id, layer_id: int, cluster, round_no: int, parameters) -> Dict[str, Any]:
    """Device plane, ``resident`` rounds: the stage leader ships the (already aggregated) stage state-dict *after* its
    UPDATE, off the round's critical path; the server assembles the full checkpoint of ``round_no`` in a writer thread."""
    return {"action": CHECKPOINT, "client_id": client_id, "layer_id": layer_id, "cluster": cluster, "round": round_no,
            "parameters": parameters, "message": "stage parameters for the checkpoint"}


def ready(client_id, layer_id: int) -> Dict[str, Any]:
    return {"action": READY, "client_id": client_id, "layer_id": layer_id, "message": "ready"}
