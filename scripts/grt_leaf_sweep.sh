# sweep of the LBVH leaf size at C4
for l in 1 2 4 8; do
GRTB200_LEAF=$l timeout 300 python bench.py --workload c4 --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/c4_l.json
python - <<PY
import json; d=json.load(open("gpurun_out/c4_l.json")); print("leaf=$l", round(d["value"],1), {k: round(v,3) for k,v in d["stage_ms"].items()})
PY
done
