#!/bin/bash
# call M: G7 accumulates W = sum grduGrd (x) d instead of the per-pair quaternion / scale contractions (closed-form normalize adjoint without a distance gradient)
mkdir -p gpurun_out
python -m pytest tests/test_gut_parity_gpu.py tests/test_gut_headline_parity_gpu.py tests/test_ref_cuda_gpu.py tests/test_kbuffer_gpu.py tests/test_train_step_gpu.py -m gpu -q > gpurun_out/r02_m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_m_pytest.log
tail -25 gpurun_out/r02_m_pytest.log
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-reference-gpu --sub-records train_default,c3 > gpurun_out/r02_m_bench.json 2> gpurun_out/r02_m_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_m_bench.json").read().strip().splitlines()[-1])
print("bench:", round(d["value"],1), "frames/s  e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["stage_ms"].items()}, "c3", d.get("c3",{}).get("value"), d.get("c3",{}).get("stage_ms"), "train_default", d["train_default"]["value"], d["train_default"]["stage_ms"])
PY
