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


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_device_allreduce_two_gpus():
    """Two-shot FedAvg all-reduce kernel vs the reference aggregation done with torch (weights, NaN scrub, int rounding, NaN vote)."""
    r = _torchrun(2, ["tools/check_allreduce.py"])
    assert r.returncode == 0 and "ALLREDUCE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.skipif(_ngpu() < 4, reason="needs 4 GPUs")
def test_device_allreduce_two_clusters_four_gpus():
    """Clusters cut at 7 / 14 (BASELINE config #4): layers 8-14 are averaged between stage 2 of one cluster and stage 1 of the other."""
    r = _torchrun(4, ["tools/check_allreduce.py"])
    assert r.returncode == 0 and "ALLREDUCE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_ring_bench_selfcheck_and_litmus():
    """bench.py --gpus 2 (ring placement): cross-GPU loss trajectory == single-GPU replica, 10^5-iteration payload/flag litmus clean,
    FedAvg all-reduce matches the NCCL mean."""
    r = _torchrun(2, ["bench.py", "--gpus", "2", "--steps", "20", "--warmup", "5"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["selfcheck"]["ok"], out["selfcheck"]
    assert out["litmus"]["ok"], out["litmus"]
    assert out["fedavg_round"]["max_abs_err_vs_nccl_mean"] < 1e-5 and out["fedavg_round"]["aggregated"]


@pytest.mark.skipif(_ngpu() < 4, reason="needs 4 GPUs")
@pytest.mark.parametrize("scenario,clients", [("split", [2, 2]), ("clusters", [2, 2]), ("three-stage", [2, 1, 1])])
def test_baseline_scenarios_through_public_api_four_gpus(scenario, clients):
    """BASELINE.json configs #3 / #4 / #5 scaled to four GPUs, one client per GPU, through the public API
    (``bench.py --scenario``): server + client FSMs over the broker, device data plane, device FedAvg all-reduce (also across
    clusters cut at 7 and at 14), asynchronous checkpoint — every round must finish and report device-timed throughput."""
    r = _torchrun(4, ["bench.py", "--gpus", "4", "--steps", "20", "--warmup", "5", "--scenario", scenario])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["config"]["clients"] == clients and out["value"] > 0
    assert all(rd["ok"] for rd in out["api"]["rounds"]) and out["api"]["checkpoint_written"]
    assert out["e2e"]["round_overhead_ms"] is not None and out["e2e"]["round_overhead_ms"] < 200
