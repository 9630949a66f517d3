#!/bin/bash
# call J: what does the forward's wait on the list total cost end to end?  (GUTB200_EXP_SKIP_WAIT: experiment-only switch)
mkdir -p gpurun_out
for rep in 1 2 3; do
for m in wait skip; do
  if [ $m = skip ]; then export GUTB200_EXP_SKIP_WAIT=1; else unset GUTB200_EXP_SKIP_WAIT; fi
  python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-reference-gpu --no-sub-records > gpurun_out/r02_j_bench_${m}_$rep.json 2> gpurun_out/r02_j_bench_${m}_$rep.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_j_bench_${m}_$rep.json").read().strip().splitlines()[-1])
print("$m $rep:", round(d["value"],1), "frames/s  e2e", round(d["e2e"]["value"],1), "host issue ms/step", round(d["e2e"]["host_issue_ms_per_step"],3))
PY
done; done
