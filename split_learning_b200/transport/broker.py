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

    # convenience
    def publish_obj(self, routing_key: str, obj) -> None:
        self.basic_publish(routing_key, pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL))

    def get_obj(self, queue: str, timeout: float = 0.0):
        body = self.basic_get(queue, timeout)
        return None if body is None else pickle.loads(body)


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
OP_PUB, OP_GET, OP_DECLARE, OP_DELETE, OP_PURGE, OP_DEPTH, OP_LIST, OP_PING, OP_SHUTDOWN = range(1, 10)
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
    return bytes(buf)


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

    def __init__(self, host: str = "127.0.0.1", port: int = 29777):
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
                conn, _ = self._srv.accept()
            except OSError:
                return
            conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            t = threading.Thread(target=self._serve, args=(conn,), daemon=True)
            t.start()

    def _serve(self, conn: socket.socket):
        s = self.store

        def reply(status: int, payload: bytes = b""):
            conn.sendall(_REP.pack(status, len(payload)) + payload)
        try:
            while True:
                op, qlen, alen = _REQ.unpack(_recv_exact(conn, _REQ.size))
                queue = _recv_exact(conn, qlen).decode() if qlen else ""
                arg = _recv_exact(conn, alen) if alen else b""
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
    def __init__(self, host: str = "127.0.0.1", port: int = 29777, retry_seconds: float = 60.0):
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

    def _call(self, op: int, queue: str, arg: bytes = b"", reply: bool = True):
        with self._lock:
            _send_request(self._sock, op, queue, arg)
            if reply:
                return _recv_reply(self._sock)

    def queue_declare(self, queue, durable=False):
        self._call(OP_DECLARE, queue)

    def basic_publish(self, routing_key, body, exchange=""):
        self._call(OP_PUB, routing_key, bytes(body), reply=False)

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

    def __init__(self, host: str = "127.0.0.1", port: int = 29777):
        exe = build_native_broker()
        self._proc = subprocess.Popen([exe, "--host", host, "--port", str(port)], stdout=subprocess.PIPE, text=True)
        line = self._proc.stdout.readline()
        if not line.startswith("SLB_BROKER_READY"):
            self._proc.kill()
            raise RuntimeError(f"slb_broker did not start on {host}:{port} (port in use?)")
        self.host, self.port = host, int(line.split()[1])
        self._channels = []

    def channel(self) -> TcpChannel:
        ch = TcpChannel(self.host, self.port, retry_seconds=5.0)
        self._channels.append(ch)
        return ch

    def close(self):
        if self._proc.poll() is None:
            try:
                TcpChannel(self.host, self.port, retry_seconds=1.0).shutdown_broker()
            except (ConnectionError, OSError):
                pass
            try:
                self._proc.wait(2.0)
            except subprocess.TimeoutExpired:
                self._proc.kill()
        for ch in self._channels:
            ch.close()


def make_broker(host: str = "127.0.0.1", port: int = 29777, kind: str = "native"):
    """The box-local broker replacing RabbitMQ: the C++ daemon, or the Python thread broker when asked for / when no
    compiler is available (both speak the same protocol, clients do not care)."""
    if kind == "native":
        try:
            return NativeBroker(host, port)
        except (RuntimeError, OSError) as e:
            print(f"[broker] native daemon unavailable ({e}); using the Python broker", flush=True)
    return TcpBroker(host, port)


def connect(address: str = "127.0.0.1", port: int = 29777, retry_seconds: float = 60.0) -> TcpChannel:
    return TcpChannel(address, port, retry_seconds)


def delete_old_queues(channel: Channel) -> bool:
    """Queue hygiene of reference src/Utils.py:8-32: delete reply*/intermediate_queue*/
    gradient_queue*/rpc_queue*, purge everything else."""
    for name in list(channel.list_queues() or []):
        if name.startswith(("reply", "intermediate_queue", "gradient_queue", "rpc_queue")):
            channel.queue_delete(name)
        else:
            channel.queue_purge(name)
    return True
