#!/bin/bash
# round 2, call G: tile-binned front-end (per-tile histogram + on-chip tile sort, no global sort, no host stall): parity suite + bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r02_g_pytest.log
cat gpurun_out/r02_g_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 100 --warmup 10 2> gpurun_out/r02_g_bench.err | tail -1 > gpurun_out/r02_g_bench.json
tail -3 gpurun_out/r02_g_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_g_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["stage_ms"])
print("ref", d.get("reference_gpu",{}).get("value"), d.get("vs_reference_gpu",{}).get("speedup_device_timed"))
print("c3", d["c3"]["value"], d["c3"]["stage_ms"])
PY
