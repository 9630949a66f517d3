#!/bin/bash
# round 2, call F: reference kernels on the GPU (debug + parity tests), refactored bench (sub-records, reference_gpu block)
mkdir -p gpurun_out
timeout 300 python scripts/refcuda_debug.py > gpurun_out/r02_f_refdebug.log 2>&1
grep -v "Warning\|frame #" gpurun_out/r02_f_refdebug.log | tail -8
timeout 1200 python -m pytest tests/test_ref_cuda_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | grep -E "ref-gpu|parity|passed|failed|Error|assert" | cut -c1-400 | tail -80 > gpurun_out/r02_f_refcuda.log
cat gpurun_out/r02_f_refcuda.log
timeout 900 python bench.py --steps 100 --warmup 10 2> gpurun_out/r02_f_bench.err | tail -1 > gpurun_out/r02_f_bench.json
tail -3 gpurun_out/r02_f_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_f_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["stage_ms"])
print("reference_gpu", json.dumps(d.get("reference_gpu")))
print("vs", json.dumps(d.get("vs_reference_gpu")))
print("c3", json.dumps(d.get("c3")))
print("c4", json.dumps(d.get("c4")))
PY
