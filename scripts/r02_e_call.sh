#!/bin/bash
# round 2, call E: reference-kernel library debug, GPU suite on the sub-block backward walk, A/B of the sub-block width
mkdir -p gpurun_out
timeout 300 python scripts/refcuda_debug.py > gpurun_out/r02_e_refdebug.log 2>&1
grep -v "Warning\|frame #" gpurun_out/r02_e_refdebug.log | head -30
for m in 7 23 39; do
GUTB200_SUBTILE_CULLING=$m timeout 300 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-reference-gpu 2>/dev/null | tail -1 > gpurun_out/r02_e_sub$m.json
python - <<PY
import json
d=json.load(open("gpurun_out/r02_e_sub$m.json")); w=d["roofline_fp32"]["work"]
print("mode $m", round(d["value"],1), round(d["e2e"]["value"],1), {k: round(v,4) for k,v in d["stage_ms"].items() if k.startswith("render")}, {k: w[k] for k in ("hit_iters","iters16","iters8","sub16_hits","sub8_hits","hits")})
PY
done
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_ref_cuda_gpu.py -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r02_e_pytest.log
cat gpurun_out/r02_e_pytest.log
for m in 23 39; do
GUTB200_SUBTILE_CULLING=$m timeout 900 python -m pytest tests/test_gut_parity_gpu.py tests/test_gut_headline_parity_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4
done
