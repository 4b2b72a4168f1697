from __future__ import annotations

import collections
import os
import pickle
import shutil
import socket
import struct
import subprocess
import threading
import time
from typing import Deque, Dict, Optional


class Channel:
    """Minimal AMQP-like surface used by the control plane and the host data plane."""

    def queue_declare(self, queue: str, durable: bool = False) -> None: ...
    def basic_publish(self, routing_key: str, body: bytes, exchange: str = "") -> None: ...
    def basic_get(self, queue: str, timeout: float = 0.0) -> Optional[bytes]: ...
    def queue_delete(self, queue: str) -> None: ...
    def queue_purge(self, queue: str) -> None: ...
    def queue_depth(self, queue: str) -> int: ...
    def list_queues(self): ...
    def close(self) -> None: ...

    # convenience: messages are pickled dicts (the reference's wire format) written / read by ``codec`` — tensors as raw
    # bytes and a *restricted* unpickler, so bytes arriving on a queue can never execute code (see transport/codec.py)
    def publish_obj(self, routing_key: str, obj) -> None:
        from . import codec
        segs = codec.dump_segments(obj)
        if len(segs) == 1:
            self.basic_publish(routing_key, segs[0])
        else:
            self.basic_publish_segments(routing_key, segs)

    def basic_publish_segments(self, routing_key: str, segments) -> None:
        """One message whose body is the concatenation of ``segments`` (bytes-like); socket channels send them without
        joining."""
        self.basic_publish(routing_key, b"".join(segments))

    def get_obj(self, queue: str, timeout: float = 0.0):
        from . import codec
        body = self.basic_get(queue, timeout)
        return None if body is None else codec.loads(body)

    def clone(self) -> "Channel":
        """A channel usable concurrently from another thread (heartbeats): the in-process broker is thread-safe as is."""
        return self


class InProcBroker(Channel):
    """Thread-safe named FIFO queues with blocking get. Also the storage engine of TcpBroker."""

    def __init__(self):
        self._q: Dict[str, Deque[bytes]] = collections.defaultdict(collections.deque)
        self._cv = threading.Condition()

    def queue_declare(self, queue, durable=False):
        with self._cv:
            self._q[queue]

    def basic_publish(self, routing_key, body, exchange=""):
        with self._cv:
            self._q[routing_key].append(body)
            self._cv.notify_all()

    def basic_get(self, queue, timeout=0.0):
        deadline = time.monotonic() + timeout
        with self._cv:
            while True:
                q = self._q[queue]
                if q:
                    return q.popleft()
                remaining = deadline - time.monotonic()
                if remaining <= 0:
                    return None
                self._cv.wait(remaining)

    def queue_delete(self, queue):
        with self._cv:
            self._q.pop(queue, None)

    def queue_purge(self, queue):
        with self._cv:
            self._q[queue].clear()

    def queue_depth(self, queue):
        with self._cv:
            return len(self._q[queue])

    def list_queues(self):
        with self._cv:
            return list(self._q.keys())

    def close(self):
        pass

    def channel(self) -> "InProcBroker":
        return self


# ---------------------------------------------------------------------------
# TCP brokers.  One binary wire protocol (little endian), spoken by the native daemon
# (transport/csrc/slb_broker.cpp) and by the Python fallback below:
#   request : u8 op | u32 queue_len | u64 arg_len | queue bytes | arg bytes
#   reply   : u8 status | u64 len | payload            (no reply for PUB)
# ---------------------------------------------------------------------------
#   OP_AUTH (arg = shared token) must be the first request of a connection when the broker was started with a token;
#   OP_SHUTDOWN is honoured for loopback peers only.
OP_PUB, OP_GET, OP_DECLARE, OP_DELETE, OP_PURGE, OP_DEPTH, OP_LIST, OP_PING, OP_SHUTDOWN, OP_AUTH = range(1, 11)


def is_loopback(host: str) -> bool:
    return host in ("127.0.0.1", "localhost", "::1") or host.startswith("127.")


def broker_token(cfg) -> str:
    """Shared secret of the broker: ``b200.broker-token`` (or ``rabbit.password`` when a non-loopback ``rabbit.address``
    is configured — the reference's RabbitMQ credential doubles as ours), overridden by ``SLB200_BROKER_TOKEN``."""
    env = os.environ.get("SLB200_BROKER_TOKEN")
    if env:
        return env
    tok = str(cfg.b200.get("broker-token") or "")
    rabbit = cfg.raw.get("rabbit", {}) or {}
    if not tok and not is_loopback(str(rabbit.get("address", "127.0.0.1"))):
        tok = str(rabbit.get("password") or "")
    return tok


def check_bind(host: str, token: str) -> None:
    """The broker carries the control plane and (host plane) the training tensors: never expose it unauthenticated."""
    if not is_loopback(host) and not token:
        raise PermissionError(f"refusing to bind the broker to {host!r} without a shared token: set b200.broker-token "
                              "(or SLB200_BROKER_TOKEN) on the server and every client, or bind to 127.0.0.1")
_REQ = struct.Struct("<BIQ")
_REP = struct.Struct("<BQ")


def _recv_exact(sock: socket.socket, n: int) -> bytes:
    buf = bytearray(n)
    view = memoryview(buf)
    got = 0
    while got < n:
        r = sock.recv_into(view[got:], n - got)
        if r == 0:
            raise ConnectionError("peer closed")
        got += r
    return bytes(buf) if n < (1 << 20) else buf          # large bodies stay in the (writable) receive buffer: no second copy


def _send_request(sock: socket.socket, op: int, queue: str, arg: bytes = b"") -> None:
    q = queue.encode()
    sock.sendall(_REQ.pack(op, len(q), len(arg)) + q)
    if arg:
        sock.sendall(arg)


def _recv_reply(sock: socket.socket):
    status, n = _REP.unpack(_recv_exact(sock, _REP.size))
    return status, (_recv_exact(sock, n) if n else b"")


class TcpBroker:
    """Python fallback broker: thread(s) serving an InProcBroker over loopback TCP (same protocol as slb_broker)."""

    kind = "python"

    def __init__(self, host: str = "127.0.0.1", port: int = 29777, token: Optional[str] = None):
        self.token = token if token is not None else os.environ.get("SLB200_BROKER_TOKEN", "")
        check_bind(host, self.token)
        self.store = InProcBroker()
        self._srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self._srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self._srv.bind((host, port))
        self._srv.listen(64)
        self.host, self.port = host, self._srv.getsockname()[1]
        self._stop = False
        self._threads = []
        t = threading.Thread(target=self._accept_loop, daemon=True, name="slb200-broker")
        t.start()
        self._threads.append(t)

    def _accept_loop(self):
        while not self._stop:
            try:
                conn, peer = self._srv.accept()
            except OSError:
                return
            conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            t = threading.Thread(target=self._serve, args=(conn, peer[0]), daemon=True)
            t.start()

    def _serve(self, conn: socket.socket, peer_host: str = "127.0.0.1"):
        import hmac
        s = self.store
        authed = not self.token

        def reply(status: int, payload: bytes = b""):
            conn.sendall(_REP.pack(status, len(payload)) + payload)
        try:
            while True:
                op, qlen, alen = _REQ.unpack(_recv_exact(conn, _REQ.size))
                queue = _recv_exact(conn, qlen).decode() if qlen else ""
                arg = _recv_exact(conn, alen) if alen else b""
                if op == OP_AUTH:
                    authed = (not self.token) or hmac.compare_digest(arg, self.token.encode())
                    reply(1 if authed else 0)
                    if not authed:
                        return
                    continue
                if not authed:                            # unauthenticated peer: drop the connection
                    return
                if op == OP_PUB:
                    s.basic_publish(queue, arg)           # fire-and-forget, like basic_publish
                elif op == OP_GET:
                    body = s.basic_get(queue, struct.unpack("<d", arg)[0] if len(arg) == 8 else 0.0)
                    reply(0 if body is None else 1, body or b"")
                elif op == OP_DECLARE:
                    s.queue_declare(queue)
                    reply(1)
                elif op == OP_DELETE:
                    s.queue_delete(queue)
                    reply(1)
                elif op == OP_PURGE:
                    s.queue_purge(queue)
                    reply(1)
                elif op == OP_DEPTH:
                    reply(1, struct.pack("<Q", s.queue_depth(queue)))
                elif op == OP_LIST:
                    reply(1, "\n".join(s.list_queues()).encode())
                elif op == OP_PING:
                    reply(1)
                elif op == OP_SHUTDOWN:
                    if not is_loopback(peer_host):        # only the box itself may stop the broker
                        return
                    reply(1)
                    self.close()
                    return
                else:
                    return
        except (ConnectionError, OSError, EOFError, struct.error):
            pass
        finally:
            conn.close()

    def channel(self) -> InProcBroker:
        """Local (same-process) channel: no socket hop."""
        return self.store

    def close(self):
        self._stop = True
        try:
            self._srv.close()
        except OSError:
            pass


class TcpChannel(Channel):
    def __init__(self, host: str = "127.0.0.1", port: int = 29777, retry_seconds: float = 60.0, token: Optional[str] = None):
        deadline = time.monotonic() + retry_seconds
        last = None
        while True:
            try:
                self._sock = socket.create_connection((host, port), timeout=5.0)
                break
            except OSError as e:       # broker not up yet
                last = e
                if time.monotonic() > deadline:
                    raise ConnectionError(f"cannot reach broker {host}:{port}: {last}")
                time.sleep(0.05)
        self._sock.settimeout(None)
        self._sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        self._lock = threading.Lock()
        self._addr = (host, port)
        self._token = token if token is not None else os.environ.get("SLB200_BROKER_TOKEN", "")
        if self._token:
            status, _ = self._call(OP_AUTH, "", self._token.encode())
            if status != 1:
                raise ConnectionError("broker rejected the shared token (b200.broker-token / SLB200_BROKER_TOKEN)")

    def clone(self) -> "TcpChannel":
        return TcpChannel(self._addr[0], self._addr[1], retry_seconds=5.0, token=self._token)

    def _call(self, op: int, queue: str, arg: bytes = b"", reply: bool = True):
        with self._lock:
            _send_request(self._sock, op, queue, arg)
            if reply:
                return _recv_reply(self._sock)

    def queue_declare(self, queue, durable=False):
        self._call(OP_DECLARE, queue)

    def basic_publish(self, routing_key, body, exchange=""):
        self._call(OP_PUB, routing_key, body if isinstance(body, (bytes, bytearray, memoryview)) else bytes(body), reply=False)

    def basic_publish_segments(self, routing_key, segments):
        q = routing_key.encode()
        total = sum(memoryview(s).nbytes for s in segments)
        with self._lock:
            self._sock.sendall(_REQ.pack(OP_PUB, len(q), total) + q)
            for seg in segments:                          # sendall releases the GIL; tensor memory is never joined / copied
                self._sock.sendall(seg)

    def basic_get(self, queue, timeout=0.0):
        status, body = self._call(OP_GET, queue, struct.pack("<d", float(timeout)))
        return body if status == 1 else None

    def queue_delete(self, queue):
        self._call(OP_DELETE, queue)

    def queue_purge(self, queue):
        self._call(OP_PURGE, queue)

    def queue_depth(self, queue):
        return struct.unpack("<Q", self._call(OP_DEPTH, queue)[1])[0]

    def list_queues(self):
        names = self._call(OP_LIST, "")[1].decode()
        return names.split("\n") if names else []

    def ping(self) -> bool:
        return self._call(OP_PING, "")[0] == 1

    def shutdown_broker(self) -> None:
        try:
            self._call(OP_SHUTDOWN, "")
        except (ConnectionError, OSError, struct.error):
            pass

    def close(self):
        try:
            self._sock.close()
        except OSError:
            pass


# ---------------------------------------------------------------------------
# Native broker daemon (C++): built in-tree with g++, started as a child process
# ---------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
BROKER_SRC = os.path.join(_HERE, "csrc", "slb_broker.cpp")
BROKER_BIN = os.path.join(_HERE, "slb_broker")


def build_native_broker(force: bool = False) -> str:
    """g++ -O2 the daemon into transport/slb_broker (about a second); returns the binary path."""
    if not force and os.path.exists(BROKER_BIN) and os.path.getmtime(BROKER_BIN) >= os.path.getmtime(BROKER_SRC):
        return BROKER_BIN
    cxx = os.environ.get("CXX") or shutil.which("g++") or shutil.which("c++")
    if not cxx:
        raise RuntimeError("no C++ compiler for slb_broker")
    res = subprocess.run([cxx, "-O2", "-std=c++17", "-pthread", "-o", BROKER_BIN, BROKER_SRC], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building slb_broker failed:\n" + res.stderr[-2000:])
    return BROKER_BIN


class NativeBroker:
    """``slb_broker`` as a child process (dies with this process).  ``channel()`` is a TCP channel like any client's."""

    kind = "native"

    def __init__(self, host: str = "127.0.0.1", port: int = 29777, token: Optional[str] = None):
        self.token = token if token is not None else os.environ.get("SLB200_BROKER_TOKEN", "")
        check_bind(host, self.token)
        exe = build_native_broker()
        env = dict(os.environ)
        env["SLB200_BROKER_TOKEN"] = self.token           # the token travels in the environment, not on the command line
        self._proc = subprocess.Popen([exe, "--host", host, "--port", str(port)], stdout=subprocess.PIPE, text=True, env=env)
        line = self._proc.stdout.readline()
        if not line.startswith("SLB_BROKER_READY"):
            self._proc.kill()
            raise RuntimeError(f"slb_broker did not start on {host}:{port} (port in use?)")
        self.host, self.port = host, int(line.split()[1])
        self._channels = []

    def channel(self) -> TcpChannel:
        ch = TcpChannel(self.host, self.port, retry_seconds=5.0, token=self.token)
        self._channels.append(ch)
        return ch

    def close(self):
        if self._proc.poll() is None:
            try:
                TcpChannel("127.0.0.1" if self.host == "0.0.0.0" else self.host, self.port, retry_seconds=1.0,
                           token=self.token).shutdown_broker()
            except (ConnectionError, OSError):
                pass
            try:
                self._proc.wait(2.0)
            except subprocess.TimeoutExpired:
                self._proc.kill()
        for ch in self._channels:
            ch.close()


def make_broker(host: str = "127.0.0.1", port: int = 29777, kind: str = "native", token: Optional[str] = None):
    """The box-local broker replacing RabbitMQ: the C++ daemon, or the Python thread broker when asked for / when no
    compiler is available (both speak the same protocol, clients do not care)."""
    tok = token if token is not None else os.environ.get("SLB200_BROKER_TOKEN", "")
    check_bind(host, tok)
    if kind == "native":
        try:
            return NativeBroker(host, port, token=tok)
        except (RuntimeError, OSError) as e:
            print(f"[broker] native daemon unavailable ({e}); using the Python broker", flush=True)
    return TcpBroker(host, port, token=tok)


def connect(address: str = "127.0.0.1", port: int = 29777, retry_seconds: float = 60.0, token: Optional[str] = None) -> TcpChannel:
    return TcpChannel(address, port, retry_seconds, token=token)


def delete_old_queues(channel: Channel) -> bool:
    """Queue hygiene of reference src/Utils.py:8-32: delete reply*/intermediate_queue*/
    gradient_queue*/rpc_queue*, purge everything else."""
    for name in list(channel.list_queues() or []):
        if name.startswith(("reply", "intermediate_queue", "gradient_queue", "rpc_queue")):
            channel.queue_delete(name)
        else:
            channel.queue_purge(name)
    return True
