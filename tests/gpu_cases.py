"""GPU test cases that run in a fresh interpreter (spawned by tests/test_executor_gpu.py): many clients sharing one GPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)


def competing(tmp: str, n1: int, n2: int, slow: int) -> dict:
    import torch
    import yaml
    os.environ.setdefault("SLB200_WAIT_SPINS", str(1 << 26))
    from split_learning_b200.checkpoint import load_checkpoint
    from split_learning_b200.config import normalize
    from split_learning_b200.runner import run_inproc
    clients = [n1, n2]
    raw = yaml.safe_load(open("config.yaml"))
    raw["server"].update({"clients": clients, "global-round": 2, "validation": False})
    raw["server"]["data-distribution"]["num-sample"] = 400 if slow else 208        # 208 = 6 x 32 + a trailing 16
    raw["server"]["manual"]["no-cluster"]["cut-layers"] = [7]
    raw["server"]["manual"]["cluster"] = {"num-cluster": 1, "cut-layers": [[7]], "infor-cluster": [clients]}
    raw["log_path"] = tmp
    raw["learning"].update({"batch-size": 32, "control-count": 2, "learning-rate": 0.01})
    raw["b200"] = {"synthetic-data": True, "data-plane": "device", "watchdog-seconds": 120, "dynamic-consumers": True, "claim-ahead": 1}
    if slow:
        raw["b200"]["debug-slow-ms"] = {n1: 4.0}              # REGISTER rank n1 = the first last-stage replica
    srv = run_inproc(normalize(raw), devices=["cuda:0"], workdir=tmp, timeout=240)
    cl = srv.clients_objs
    assert all(c.dstage is not None and c._dynamic for c in cl)
    last = [c for c in cl if c.layer_id == 2]
    per_lane = (400 // 32 + (1 if 400 % 32 else 0)) if slow else 7
    everything = sorted((lane, it) for lane in range(n1) for it in range(per_lane))
    sd = load_checkpoint(os.path.join(tmp, "VGG16_CIFAR10.pth"))
    assert all(torch.isfinite(v.float()).all() for v in sd.values())
    first = [c for c in cl if c.layer_id == 1]
    assert all(c.rounds_done == 2 for c in first)
    return {"rounds_ok": [h["ok"] for h in srv.history],
            "exactly_once": sorted(t for c in last for t in c.claimed) == everything,
            "all_claimed_something": all(len(c.claimed) > 0 for c in last),
            "n_slow": len(last[0].claimed), "n_fast": len(last[1].claimed), "slow_rank_ok": last[0].rank == n1,
            "ckpt_entries": len(sd)}


if __name__ == "__main__":
    kind = sys.argv[1]
    if kind == "competing":
        res = competing(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
    else:
        raise SystemExit(f"unknown case {kind}")
    print("CASE_OK " + json.dumps(res), flush=True)
    os._exit(0)                      # daemon threads of finished roles must not delay / break interpreter shutdown
