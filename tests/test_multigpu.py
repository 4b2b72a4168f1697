"""Multi-GPU paths (need >= 2 visible GPUs; skipped on a 1-GPU box): peer mailboxes over CUDA IPC,
the 2-stage device pipeline across two processes, and the NVLink FedAvg kernel."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _torchrun(n, script_args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533"] + script_args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout,
                          env=dict(os.environ, PYTHONPATH=ROOT))


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_bench_two_gpus():
    r = _torchrun(2, ["bench.py", "--gpus", "2", "--steps", "20", "--warmup", "5"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["n_gpus"] == 2 and out["value"] > 0 and out["e2e"]["value"] > 0
    assert 0.5 < out["final_loss"] < 5.0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_peer_fedavg_two_gpus():
    r = _torchrun(2, ["tools/check_fedavg_peer.py"])
    assert r.returncode == 0 and "FEDAVG_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
