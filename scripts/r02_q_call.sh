#!/bin/bash
# call Q: tile_scan lays out the sub-buckets in its per-tile pass (no third pass)
mkdir -p gpurun_out
python -m pytest tests/test_gut_parity_gpu.py tests/test_gut_headline_parity_gpu.py tests/test_ref_cuda_gpu.py tests/test_kbuffer_gpu.py -m gpu -q -x > gpurun_out/r02_q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_q_pytest.log
tail -3 gpurun_out/r02_q_pytest.log
python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-reference-gpu --sub-records c3 > gpurun_out/r02_q_bench.json 2> gpurun_out/r02_q_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_q_bench.json").read().strip().splitlines()[-1])
print("bench:", round(d["value"],1), "frames/s", {k:round(v,3) for k,v in d["stage_ms"].items()})
t=d["c3"]; print("   c3:", round(t["value"],1), {k:round(v,3) for k,v in t["stage_ms"].items()})
PY
