"""Wire hygiene of the control plane: restricted codec, zero-copy segments for state-dict sized messages, shared-token
authentication of both brokers, loopback-only binds without a token, liveness beacons."""
import os
import pickle
import threading
import time
import uuid

import pytest
import torch

from split_learning_b200 import messages as M
from split_learning_b200.transport import NativeBroker, TcpBroker, TcpChannel, codec
from split_learning_b200.transport.broker import check_bind


class _Evil:
    def __reduce__(self):
        return (os.system, ("echo pwned > /tmp/slb200_pwned",))


def test_codec_refuses_code_execution_payloads():
    evil = pickle.dumps({"action": "UPDATE", "parameters": _Evil()})
    with pytest.raises(codec.UnsafePayload):
        codec.loads(evil)
    with pytest.raises(codec.UnsafePayload):
        codec.loads_cuda_ipc(evil)
    nested = pickle.dumps({"x": [1, (2, {"y": _Evil()})]})
    with pytest.raises(codec.UnsafePayload):
        codec.loads(nested)
    assert not os.path.exists("/tmp/slb200_pwned")


def test_codec_round_trip_reference_message_shapes():
    sd = {"layer1.weight": torch.randn(64, 3, 3, 3), "layer2.num_batches_tracked": torch.tensor(7), "h": torch.randn(5).half(),
          "empty": torch.zeros(0)}
    msg = M.update(uuid.uuid4(), 2, True, 12, 0, sd, resident=False)
    for wire in (codec.dumps(msg), b"".join(codec.dump_segments(msg))):
        out = codec.loads(wire)
        assert out["action"] == M.UPDATE and out["client_id"] == msg["client_id"] and out["size"] == 12
        for k, v in sd.items():
            assert out["parameters"][k].dtype == v.dtype and out["parameters"][k].shape == v.shape and torch.equal(out["parameters"][k], v)


def test_codec_segments_are_zero_copy_and_writable():
    big = {"p": {"w": torch.randn(1024, 1024), "small": torch.arange(8)}, "tag": "ckpt"}
    segs = codec.dump_segments(big)
    assert len(segs) == 2 and memoryview(segs[1]).nbytes == 4 << 20           # one out-of-band buffer: the big tensor's memory
    assert memoryview(segs[1]).obj is not None and len(segs[0]) < 4096        # header + pickle stay tiny
    wire = bytearray(b"".join(segs))
    out = codec.loads(wire)
    assert torch.equal(out["p"]["w"], big["p"]["w"]) and torch.equal(out["p"]["small"], big["p"]["small"])
    out["p"]["w"][0, 0] = 123.0                                              # a view of the receive buffer, and writable
    assert torch.frombuffer(wire, dtype=torch.float32, count=1, offset=len(segs[0]))[0] == 123.0
    ro = codec.loads(bytes(wire))                                            # immutable input: tensors are copied, never aliased
    ro["p"]["w"][0, 0] = 1.0
    bad = bytearray(wire)
    bad[8:16] = (1 << 40).to_bytes(8, "little")                              # pickle length beyond the payload
    with pytest.raises(Exception):
        codec.loads(bad)


@pytest.mark.parametrize("kind", ["native", "python"])
def test_broker_ships_state_dict_sized_segments(kind):
    srv = NativeBroker(port=0) if kind == "native" else TcpBroker(port=0)
    try:
        a, b = TcpChannel(port=srv.port), TcpChannel(port=srv.port)
        sd = {f"layer{i}.weight": torch.randn(512, 1024) for i in range(6)}  # 12 MB in six out-of-band buffers
        a.publish_obj("rpc_queue", M.checkpoint("c", 1, 0, 3, sd))
        a.publish_obj("rpc_queue", {"action": "after"})                      # framing stays intact behind a segmented body
        got = b.get_obj("rpc_queue", 5.0)
        assert got["action"] == M.CHECKPOINT and got["round"] == 3
        assert all(torch.equal(got["parameters"][k], v) for k, v in sd.items())
        assert b.get_obj("rpc_queue", 5.0) == {"action": "after"}
        a.close(), b.close()
    finally:
        srv.close()


@pytest.mark.parametrize("kind", ["native", "python"])
def test_broker_shared_token(kind):
    srv = NativeBroker(port=0, token="s3cret") if kind == "native" else TcpBroker(port=0, token="s3cret")
    try:
        good = TcpChannel(port=srv.port, token="s3cret")
        good.publish_obj("q", {"ok": 1})
        assert good.get_obj("q", 2.0) == {"ok": 1}
        with pytest.raises((ConnectionError, OSError)):
            TcpChannel(port=srv.port, token="wrong", retry_seconds=1.0)
        with pytest.raises((ConnectionError, OSError, Exception)):
            anon = TcpChannel(port=srv.port, token="", retry_seconds=1.0)     # no AUTH at all: the first request is dropped
            anon.publish_obj("q", {"intruder": 1})
            assert anon.ping()
        assert good.queue_depth("q") == 0                                     # nothing the intruder sent was queued
        good.close()
    finally:
        srv.close()


def test_broker_refuses_public_bind_without_token():
    with pytest.raises(PermissionError):
        check_bind("0.0.0.0", "")
    with pytest.raises(PermissionError):
        TcpBroker(host="0.0.0.0", port=0, token="")
    check_bind("127.0.0.1", "")
    check_bind("0.0.0.0", "tok")


def test_heartbeat_refreshes_server_liveness_and_is_relayed():
    """Clients beacon to rpc_queue; the server records them and relays a beacon to every reply queue."""
    from split_learning_b200.client import RpcClient
    from split_learning_b200.transport import InProcBroker
    br = InProcBroker()
    cli = RpcClient("hb-client", 1, br, b200_opts={"watchdog-seconds": 2, "heartbeat-seconds": 0.05})
    cli.start_heartbeat()
    try:
        deadline = time.monotonic() + 5
        seen = None
        while time.monotonic() < deadline and seen is None:
            m = br.get_obj(M.RPC_QUEUE, 0.5)
            if m and m.get("action") == M.HEARTBEAT:
                seen = m
        assert seen is not None and str(seen["client_id"]) == "hb-client" and seen.get("t")
    finally:
        cli.stop_heartbeat()
