"""In-box message transport replacing RabbitMQ/pika (reference L1 layer).

Queue grammar is kept (``rpc_queue``, ``reply_{id}``, ``intermediate_queue_{layer}_{cluster}``,
``gradient_queue_{layer}_{client}``) but the broker is a thread in the server process:
``InProcBroker`` for single-process runs/tests, ``TcpBroker`` + ``TcpChannel`` (loopback TCP,
length-prefixed frames) for one-process-per-GPU runs.  On CUDA the *data plane* does not go
through here at all (see ``parallel/mailbox.py``); only the seven control verbs do.
"""
from .broker import Channel, InProcBroker, TcpBroker, TcpChannel, connect

__all__ = ["Channel", "InProcBroker", "TcpBroker", "TcpChannel", "connect"]
