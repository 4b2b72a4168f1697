"""End-to-end demo through the user-facing entry points on N GPUs: writes a config (clients [n, n], device data plane,
synthetic CIFAR-10), runs ``launch.py`` (server + one client process per GPU), prints the round history from app.log.

    python tools/launch_demo.py --per-stage 2 --rounds 2
"""
import argparse
import os
import subprocess
import sys
import tempfile
import time

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--per-stage", type=int, default=2)
ap.add_argument("--rounds", type=int, default=2)
ap.add_argument("--samples", type=int, default=640)
ap.add_argument("--plane", default="device")
a = ap.parse_args()
tmp = tempfile.mkdtemp(prefix="slb200_demo_")
raw = yaml.safe_load(open(os.path.join(ROOT, "config.yaml")))
raw["server"].update({"clients": [a.per_stage, a.per_stage], "global-round": a.rounds, "validation": True})
raw["server"]["data-distribution"]["num-sample"] = a.samples
raw["server"]["manual"]["no-cluster"]["cut-layers"] = [7]
raw["log_path"] = tmp
raw["learning"].update({"batch-size": 32, "control-count": 3, "learning-rate": 0.01})
raw["b200"] = {"synthetic-data": True, "data-plane": a.plane, "watchdog-seconds": 180, "port": 29955}
cfg = os.path.join(tmp, "config.yaml")
yaml.safe_dump(raw, open(cfg, "w"))
t0 = time.time()
r = subprocess.run([sys.executable, os.path.join(ROOT, "launch.py"), "--config", cfg, "--timeout", "500"], cwd=tmp,
                   env=dict(os.environ, SLB200_QUIET="1"), capture_output=True, text=True)
print("launch.py rc", r.returncode, "seconds", round(time.time() - t0, 1))
if r.returncode != 0:
    print(r.stdout[-1500:], r.stderr[-3000:])
log = os.path.join(tmp, "app.log")
if os.path.exists(log):
    print("".join(open(log).readlines()[-8:]))
print("checkpoint:", os.path.exists(os.path.join(tmp, "VGG16_CIFAR10.pth")))
