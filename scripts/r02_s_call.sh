#!/bin/bash
# call S: polling vs sleeping wait for the list total -- spread of the end-to-end number
mkdir -p gpurun_out
python -m pytest tests/test_gut_parity_gpu.py -m gpu -q -x -k "c1 or empty or error" 2>&1 | tail -2
for rep in 1 2 3; do
for m in spin sync; do
  GUTB200_TOTAL_WAIT=$m python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-reference-gpu --no-sub-records > gpurun_out/r02_s_bench_${m}_$rep.json 2> gpurun_out/r02_s_bench_${m}_$rep.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_s_bench_${m}_$rep.json").read().strip().splitlines()[-1])
print("$m $rep:", round(d["value"],1), "frames/s  e2e", round(d["e2e"]["value"],1), "host issue ms/step", round(d["e2e"]["host_issue_ms_per_step"],3))
PY
done; done
