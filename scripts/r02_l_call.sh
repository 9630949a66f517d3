#!/bin/bash
# call L: host side of the end-to-end step trimmed (tracer.py, one pinned slab per step); both wait modes once more
mkdir -p gpurun_out
python -m pytest tests/test_gut_parity_gpu.py tests/test_train_step_gpu.py tests/test_kbuffer_gpu.py -m gpu -x -q > gpurun_out/r02_l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_l_pytest.log
tail -3 gpurun_out/r02_l_pytest.log
for rep in 1 2 3; do
for m in wait skip; do
  if [ $m = skip ]; then export GUTB200_EXP_SKIP_WAIT=1; else unset GUTB200_EXP_SKIP_WAIT; fi
  python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-reference-gpu --no-sub-records > gpurun_out/r02_l_bench_${m}_$rep.json 2> gpurun_out/r02_l_bench_${m}_$rep.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_l_bench_${m}_$rep.json").read().strip().splitlines()[-1])
print("$m $rep:", round(d["value"],1), "frames/s  e2e", round(d["e2e"]["value"],1), "host issue ms/step", round(d["e2e"]["host_issue_ms_per_step"],3))
PY
done; done
unset GUTB200_EXP_SKIP_WAIT
python bench.py --steps 400 --warmup 10 --no-cpu-baseline --no-reference-gpu --no-sub-records --profile-host gpurun_out/r02_l_host_profile_wait.txt > /dev/null 2>&1
