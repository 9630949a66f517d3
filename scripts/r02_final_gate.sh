#!/bin/bash
# last call of the round: the whole GPU suite, smoke and the default bench line on the final commit (logs kept under profiles/)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/r02_final_pytest.log
cat gpurun_out/r02_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_final_smoke.log
timeout 900 python bench.py 2> gpurun_out/r02_final_bench.err | tail -1 > gpurun_out/r02_final_bench.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_final_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], {k:round(v,3) for k,v in d["stage_ms"].items()}, "ref", d.get("reference_gpu",{}).get("value"), "c3", d["c3"]["value"], "c4", d["c4"]["value"], "train_default", d["train_default"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["traffic"], "cpu", d["cpu_baseline"]["value"])
PY
