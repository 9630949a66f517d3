#!/bin/bash
# call I: backward stages 512 entries per barrier (dynamic shared memory) vs 256, + the camera-mismatch test
mkdir -p gpurun_out
python -m pytest tests/test_gut_parity_gpu.py tests/test_gut_headline_parity_gpu.py -m gpu -x -q > gpurun_out/r02_i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_i_pytest.log
tail -3 gpurun_out/r02_i_pytest.log
for b in 512 256 512 256; do
  GUTB200_BWD_BATCH=$b python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-reference-gpu --no-sub-records > gpurun_out/r02_i_bench_$b.json 2> gpurun_out/r02_i_bench_$b.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_i_bench_$b.json").read().strip().splitlines()[-1])
print("batch $b:", round(d["value"],1), "frames/s  e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["stage_ms"].items()})
PY
done
