"""In-box message transport replacing RabbitMQ/pika (reference L1 layer).

Queue grammar is kept (``rpc_queue``, ``reply_{id}``, ``intermediate_queue_{layer}_{cluster}``,
``gradient_queue_{layer}_{client}``) but the broker is box-local: ``InProcBroker`` for
single-process runs/tests; for one-process-per-GPU runs the native C++ daemon ``slb_broker``
(``NativeBroker``; ``csrc/slb_broker.cpp``, started by ``server.py``) or the Python ``TcpBroker``
fallback — one binary length-prefixed protocol, ``TcpChannel`` on the client side.  On CUDA the *data plane* does not go
through here at all (see ``parallel/mailbox.py``); only the seven control verbs do.
"""
from . import codec
from .broker import Channel, InProcBroker, NativeBroker, TcpBroker, TcpChannel, build_native_broker, connect, make_broker

__all__ = ["codec", "Channel", "InProcBroker", "NativeBroker", "TcpBroker", "TcpChannel", "build_native_broker", "connect", "make_broker"]
