#!/bin/bash
# Multi-GPU measurement set of one box: tools/run_multi.sh N [ref]   (inside `gpurun --gpus N`)
N=$1; REF=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
O=gpurun_out/r2_n${N}
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    keep = {k: d.get(k) for k in ("impl", "value", "ms_per_step", "n_gpus", "unavailable", "error")}
    e = d.get("e2e") or {}
    keep["e2e"] = {k: e.get(k) for k in ("value", "round_wall_ms", "round_overhead_ms", "images_per_s_whole_round")}
    for k in ("selfcheck", "litmus", "fedavg", "round_images_per_s"):
        if k in d: keep[k] = d[k]
    api = d.get("api") or {}
    if api.get("steady_round"):
        keep["steady"] = {k: api["steady_round"].get(k) for k in ("wall_ms", "device_ms", "phases_ms", "client_timing_ms")}
    if api.get("error"): keep["api_error"] = api["error"]
    print(json.dumps(keep))
except Exception as ex:
    print("unreadable", sys.argv[1], ex)
PY
}
timeout 420 $TR --master-port 29711 bench.py --gpus $N --steps 100 --warmup 10 > ${O}_ring.json 2> ${O}_ring.err; echo "ring rc=$?"; show ${O}_ring.json; tail -3 ${O}_ring.err
timeout 200 $TR --master-port 29712 tools/check_allreduce.py > ${O}_allreduce.log 2>&1; echo "allreduce rc=$?"; tail -3 ${O}_allreduce.log
for S in split clusters three-stage; do
  if [ "$S" != "split" ] && [ "$N" -lt 4 ]; then continue; fi
  timeout 300 $TR --master-port 29713 bench.py --gpus $N --steps 100 --warmup 10 --scenario $S > ${O}_$S.json 2> ${O}_$S.err; echo "$S rc=$?"; show ${O}_$S.json; tail -3 ${O}_$S.err
done
if [ -n "$REF" ]; then
  timeout 400 $TR --master-port 29714 bench.py --impl reference --gpus $N --steps 30 --warmup 5 --placement split > ${O}_ref_split.json 2> ${O}_ref_split.err; echo "ref split rc=$?"; show ${O}_ref_split.json; tail -3 ${O}_ref_split.err
  timeout 400 $TR --master-port 29715 bench.py --impl reference --gpus $N --steps 30 --warmup 5 > ${O}_ref_ring.json 2> ${O}_ref_ring.err; echo "ref ring rc=$?"; show ${O}_ref_ring.json; tail -3 ${O}_ref_ring.err
fi
