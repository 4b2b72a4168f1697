"""Run a whole topology from one ``Config``: server + clients as threads (single process,
``InProcBroker``) or as one process per GPU (``launch.py``; ``TcpBroker``).

This is the convenience layer SURVEY §7.1 asks for on top of the reference-compatible
``server.py`` / ``client.py --layer_id`` entry points."""
from __future__ import annotations

import threading
import traceback
import uuid
from typing import List, Optional, Sequence

from .algorithms import client_class, server_class
from .config import Config
from .log import Logger
from .plan import rank_assignment
from .transport import InProcBroker

DEFAULT_PROFILE = {"exe_time": [1.0], "size_data": [1.0], "speed": 1.0, "network": 1.0}


def run_inproc(cfg: Config, devices: Optional[Sequence[str]] = None, profiles: Optional[List[dict]] = None,
               workdir: str = ".", logger: Optional[Logger] = None, client_kwargs: Optional[dict] = None,
               timeout: float = 600.0):
    """Run every role in this process.  Returns the finished ``Server`` (history, topology)."""
    broker = InProcBroker()
    algo = cfg.b200.get("algorithm", "main")
    server = server_class(algo)(cfg, broker, logger=logger, workdir=workdir)
    errors: List[BaseException] = []

    def guard(fn):
        def run():
            try:
                fn()
            except BaseException as e:          # surfaced to the caller below
                traceback.print_exc()
                errors.append(e)
                server.done = True
                try:                            # a failed role must not leave the others waiting out their watchdogs
                    from . import messages as M
                    for c in list(server.clients):
                        server.send_to_response(c.client_id, M.stop("run aborted"))
                except Exception:
                    pass
        return run

    info = cfg.infor_cluster if (cfg.cluster_mode and cfg.infor_cluster_given) else None
    ranks = rank_assignment(cfg.clients, info)
    threads = [threading.Thread(target=guard(lambda: server.start(idle_timeout=timeout)), daemon=True, name="server")]
    clients = []
    for r, (layer_id, cluster, _i) in enumerate(ranks):
        dev = devices[r % len(devices)] if devices else "cpu"
        cli = client_class(algo, cfg.b200)(str(uuid.uuid4()), layer_id, broker, device=dev, b200_opts=cfg.b200, rank=r,
                                 **(client_kwargs or {}))
        prof = profiles[r] if profiles else dict(DEFAULT_PROFILE)
        clients.append(cli)

        def body(cli=cli, prof=prof, cluster=cluster):
            cli.register(prof, cluster if info is not None else -1)
            cli.wait_response(idle_timeout=timeout)
        threads.append(threading.Thread(target=guard(body), daemon=True, name=f"client{r}"))
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout)
    if errors:
        raise errors[0]
    if any(t.is_alive() for t in threads):
        raise TimeoutError("run_inproc: roles still alive after timeout")
    server.clients_objs = clients
    return server


def run_variant(cfg: Config, client_specs: Sequence[dict], workdir: str = ".", timeout: float = 600.0, devices=None):
    """Like ``run_inproc`` but with explicit per-client registration fields (``layer_id``, ``cluster``,
    ``idx``/``in_cluster``/``out_cluster``, ``select``) — the CLI flags of the variant ``client.py`` scripts."""
    broker = InProcBroker()
    algo = cfg.b200.get("algorithm", "main")
    server = server_class(algo)(cfg, broker, workdir=workdir)
    errors: List[BaseException] = []

    def guard(fn):
        def run():
            try:
                fn()
            except BaseException as e:
                traceback.print_exc()
                errors.append(e)
                server.done = True
                try:                            # a failed role must not leave the others waiting out their watchdogs
                    from . import messages as M
                    for c in list(server.clients):
                        server.send_to_response(c.client_id, M.stop("run aborted"))
                except Exception:
                    pass
        return run
    threads = [threading.Thread(target=guard(lambda: server.start(idle_timeout=timeout)), daemon=True)]
    clients = []
    for r, spec in enumerate(client_specs):
        spec = dict(spec)
        layer_id = spec.pop("layer_id")
        cluster = spec.pop("cluster", -1)
        dev = devices[r % len(devices)] if devices else "cpu"
        cli = client_class(algo, cfg.b200)(str(uuid.uuid4()), layer_id, broker, device=dev, b200_opts=cfg.b200, rank=r)
        clients.append(cli)

        def body(cli=cli, cluster=cluster, spec=spec):
            cli.register(dict(DEFAULT_PROFILE), cluster, **spec)
            cli.wait_response(idle_timeout=timeout)
        threads.append(threading.Thread(target=guard(body), daemon=True))
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout)
    if errors:
        raise errors[0]
    if any(t.is_alive() for t in threads):
        raise TimeoutError("run_variant: roles still alive after timeout")
    server.clients_objs = clients
    return server
