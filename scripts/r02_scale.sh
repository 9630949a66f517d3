#!/bin/bash
# view-parallel scaling on one box: bench.py at N = $1 GPUs (compact exchange, gradient accumulation 2 and 1) -> gpurun_out/r02_scale_n$1*.json
N=${1:-2}
mkdir -p gpurun_out
run() { # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 60 --warmup 10 --no-cpu-baseline $2 2> gpurun_out/r02_scale_n${N}_$1.err | tail -1 > gpurun_out/r02_scale_n${N}_$1.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_scale_n${N}_$1.json"))
    print("N=$N $1:", round(d["value"],1), "frames/s", round(d["ms_per_step"],4), "ms/view-step", "e2e", round(d["e2e"]["value"],1), "exchange", json.dumps(d.get("exchange")), "c3", d.get("c3",{}).get("value"), d.get("c3",{}).get("exchange"))
except Exception as e:
    print("N=$N $1: failed", e); print(open("gpurun_out/r02_scale_n${N}_$1.err").read()[-1500:])
PY
}
if [ "${2:-all}" = "default" ]; then
run acc4 ""
else
run acc4 ""
run acc2 "--accumulate 2 --no-sub-records"
run acc1 "--accumulate 1 --no-sub-records"
fi
