"""In-box stand-in for the ``pika`` AMQP client so the UNMODIFIED reference (baseline/_ref)
can run without a RabbitMQ server.  Implements exactly the API surface the reference touches
(BlockingConnection/channel: queue_declare, basic_qos, basic_consume, start_consuming,
basic_publish, basic_get, basic_ack, queue_delete/purge; PlainCredentials,
ConnectionParameters, BasicProperties).  Queues live in one broker: in-process when every
role runs in this process, else a loopback TCP broker hosted by rank 0 (``serve``).
Nothing of split_learning_b200 is imported here.

Message-completion hook: ``on_get`` callbacks let the benchmark harness observe deliveries
(to time K pipeline steps) without touching the reference's code.
"""
import collections
import os
import pickle
import socket
import struct
import threading
import time

_EMPTY_WAIT = 2e-4          # a fast broker round trip; keeps busy-polling loops from hogging the GIL


class _Store:
    def __init__(self):
        self.q = collections.defaultdict(collections.deque)
        self.cv = threading.Condition()

    def publish(self, key, body):
        with self.cv:
            self.q[key].append(body)
            self.cv.notify_all()

    def get(self, key, wait=_EMPTY_WAIT):
        with self.cv:
            d = self.q[key]
            if not d and wait:
                self.cv.wait(wait)
            return d.popleft() if d else None

    def delete(self, key):
        with self.cv:
            self.q.pop(key, None)


_LOCAL = _Store()
_REMOTE = None              # (host, port) when a TCP broker is used
on_get = []                 # callbacks(queue_name, body) fired on every successful basic_get


def _send(sock, obj):
    b = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    sock.sendall(struct.pack("<Q", len(b)) + b)


def _recv(sock):
    def exact(n):
        buf = bytearray()
        while len(buf) < n:
            chunk = sock.recv(min(1 << 20, n - len(buf)))
            if not chunk:
                raise ConnectionError("closed")
            buf += chunk
        return bytes(buf)
    (n,) = struct.unpack("<Q", exact(8))
    return pickle.loads(exact(n))


def serve(host="127.0.0.1", port=29655):
    """Start the TCP broker (rank 0)."""
    srv = socket.socket()
    srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    srv.bind((host, port))
    srv.listen(64)

    def client(conn):
        conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        try:
            while True:
                op, key, arg = _recv(conn)
                if op == "pub":
                    _LOCAL.publish(key, arg)
                elif op == "get":
                    _send(conn, _LOCAL.get(key))
                elif op == "del":
                    _LOCAL.delete(key)
        except (ConnectionError, OSError, EOFError):
            pass

    def loop():
        while True:
            try:
                c, _ = srv.accept()
            except OSError:
                return
            threading.Thread(target=client, args=(c,), daemon=True).start()
    threading.Thread(target=loop, daemon=True).start()
    return srv


def use_remote(host="127.0.0.1", port=29655):
    global _REMOTE
    _REMOTE = (host, port)


class PlainCredentials:
    def __init__(self, *a, **k): pass


class ConnectionParameters:
    def __init__(self, *a, **k): pass


class BasicProperties:
    def __init__(self, reply_to=None, **k):
        self.reply_to = reply_to


class _Method:
    def __init__(self, tag):
        self.delivery_tag = tag


class _Channel:
    def __init__(self, conn):
        self._conn = conn
        self._consumers = []
        self._tag = 0

    # -- transport ---------------------------------------------------------
    def _pub(self, key, body):
        if self._conn.sock is None:
            _LOCAL.publish(key, body)
        else:
            with self._conn.lock:
                _send(self._conn.sock, ("pub", key, body))

    def _get(self, key):
        if self._conn.sock is None:
            return _LOCAL.get(key)
        with self._conn.lock:
            _send(self._conn.sock, ("get", key, None))
            return _recv(self._conn.sock)

    # -- pika surface ------------------------------------------------------
    def queue_declare(self, queue=None, durable=False, **k): return None
    def basic_qos(self, prefetch_count=0, **k): return None
    def queue_purge(self, queue=None): return None

    def queue_delete(self, queue=None):
        if self._conn.sock is None:
            _LOCAL.delete(queue)

    def basic_publish(self, exchange="", routing_key="", body=b"", properties=None, **k):
        self._pub(routing_key, body)

    def basic_get(self, queue=None, auto_ack=False):
        body = self._get(queue)
        if body is None:
            return None, None, None
        for cb in on_get:
            cb(queue, body)
        self._tag += 1
        return _Method(self._tag), BasicProperties(), body

    def basic_ack(self, delivery_tag=0, **k): return None

    def basic_consume(self, queue=None, on_message_callback=None, **k):
        self._consumers.append((queue, on_message_callback))

    def start_consuming(self):
        while self._consumers:
            for q, cb in list(self._consumers):
                body = self._get(q)
                if body is not None:
                    self._tag += 1
                    cb(self, _Method(self._tag), BasicProperties(), body)

    def stop_consuming(self):
        self._consumers = []


class BlockingConnection:
    def __init__(self, parameters=None):
        self.lock = threading.Lock()
        self.sock = None
        if _REMOTE is not None:
            deadline = time.time() + 60
            while True:
                try:
                    self.sock = socket.create_connection(_REMOTE, timeout=5)
                    break
                except OSError:
                    if time.time() > deadline:
                        raise
                    time.sleep(0.05)
            self.sock.settimeout(None)
            self.sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)

    def channel(self):
        return _Channel(self)

    def process_data_events(self, *a, **k): return None

    def close(self):
        if self.sock is not None:
            self.sock.close()
