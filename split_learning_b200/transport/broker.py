from __future__ import annotations

import collections
import pickle
import socket
import struct
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
# TCP broker: frame = u32 length | pickle((op, queue, arg))
# ---------------------------------------------------------------------------
def _send_frame(sock: socket.socket, payload: bytes) -> None:
    sock.sendall(struct.pack("<Q", len(payload)) + payload)


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


def _recv_frame(sock: socket.socket) -> bytes:
    (n,) = struct.unpack("<Q", _recv_exact(sock, 8))
    return _recv_exact(sock, n)


class TcpBroker:
    """Broker thread(s) serving an InProcBroker over loopback TCP."""

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
        try:
            while True:
                op, queue, arg = pickle.loads(_recv_frame(conn))
                if op == "pub":
                    s.basic_publish(queue, arg)
                    continue                      # fire-and-forget, like basic_publish
                if op == "get":
                    res = s.basic_get(queue, arg)
                elif op == "declare":
                    res = s.queue_declare(queue)
                elif op == "delete":
                    res = s.queue_delete(queue)
                elif op == "purge":
                    res = s.queue_purge(queue)
                elif op == "depth":
                    res = s.queue_depth(queue)
                elif op == "list":
                    res = s.list_queues()
                else:
                    res = None
                _send_frame(conn, pickle.dumps(res, protocol=pickle.HIGHEST_PROTOCOL))
        except (ConnectionError, OSError, EOFError):
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

    def _call(self, op, queue, arg=None, reply=True):
        with self._lock:
            _send_frame(self._sock, pickle.dumps((op, queue, arg), protocol=pickle.HIGHEST_PROTOCOL))
            if reply:
                return pickle.loads(_recv_frame(self._sock))

    def queue_declare(self, queue, durable=False):
        self._call("declare", queue)

    def basic_publish(self, routing_key, body, exchange=""):
        self._call("pub", routing_key, body, reply=False)

    def basic_get(self, queue, timeout=0.0):
        return self._call("get", queue, timeout)

    def queue_delete(self, queue):
        self._call("delete", queue)

    def queue_purge(self, queue):
        self._call("purge", queue)

    def queue_depth(self, queue):
        return self._call("depth", queue)

    def list_queues(self):
        return self._call("list", "")

    def close(self):
        try:
            self._sock.close()
        except OSError:
            pass


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
