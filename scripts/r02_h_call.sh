#!/bin/bash
# round 2, call H: tile sort v3 check (parity subset) + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gut_parity_gpu.py tests/test_gut_headline_parity_gpu.py tests/test_grt_parity_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -5
timeout 900 python bench.py --steps 100 --warmup 10 2> gpurun_out/r02_h_bench.err | tail -1 > gpurun_out/r02_h_bench.json
tail -3 gpurun_out/r02_h_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_h_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["stage_ms"])
print("c3", d["c3"]["value"], d["c3"]["stage_ms"])
print("c4", d["c4"]["value"], d["c4"]["stage_ms"], json.dumps(d["c4"].get("work"))[:1500])
PY
