#!/bin/bash
# call O: the rgb-only (training default) backward: quarter vs half sub-blocks; ncu of the quarter variant
mkdir -p gpurun_out
for m in 7 23; do
  GUTB200_SUBTILE_CULLING=$m python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-reference-gpu --sub-records train_default > gpurun_out/r02_o_bench_$m.json 2> gpurun_out/r02_o_bench_$m.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_o_bench_$m.json").read().strip().splitlines()[-1])
print("mode $m:", round(d["value"],1), "frames/s", "e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["stage_ms"].items()})
t=d["train_default"]; print("   train_default:", round(t["value"],1), {k:round(v,3) for k,v in t["stage_ms"].items()})
PY
done
tail -3 gpurun_out/r02_o_bench_7.err
