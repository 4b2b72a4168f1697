"""Markdown summary of the multi-GPU measurement set written by tools/run_multi.sh (gpurun_out/r2_n{N}_*.json)."""
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, sys.argv[1] if len(sys.argv) > 1 else "gpurun_out")


def load(path):
    try:
        return json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])
    except Exception:
        return None


rows = []
for f in sorted(glob.glob(os.path.join(D, "r2_n*_*.json"))):
    m = re.match(r"r2_n(\d+)_(.+)\.json", os.path.basename(f))
    d = load(f)
    if not m or not d:
        continue
    n, kind = int(m.group(1)), m.group(2)
    e = d.get("e2e") or {}
    api = d.get("api") or {}
    st = api.get("steady_round") or {}
    fa = d.get("fedavg_round") or {}
    rows.append((n, kind, d.get("impl"), d.get("value"), d.get("ms_per_step"), e.get("value"), st.get("wall_ms") or e.get("round_wall_ms"),
                 st.get("overhead_ms") or e.get("round_overhead_ms"), api.get("clients") or (d.get("config") or {}).get("clients"),
                 fa.get("ms_max_over_ranks"), (d.get("clocks") or {}).get("sm_mhz"), (d.get("selfcheck") or {}).get("ok"),
                 (d.get("litmus") or {}).get("errors")))
print("| GPUs | run | impl | clients | device-timed img/s | ms/step | public-API e2e img/s | round wall ms | round overhead ms | FedAvg ms | SM MHz | selfcheck | litmus errors |")
print("|---:|---|---|---|---:|---:|---:|---:|---:|---:|---:|---|---|")
f = lambda v, p=0: "–" if v is None else (f"{v:,.{p}f}" if isinstance(v, (int, float)) else str(v))
for r in sorted(rows):
    print(f"| {r[0]} | {r[1]} | {r[2]} | {r[8]} | {f(r[3])} | {f(r[4], 3)} | {f(r[5])} | {f(r[6], 1)} | {f(r[7], 1)} | {f(r[9], 3)} | {f(r[10])} | {r[11]} | {r[12]} |")
