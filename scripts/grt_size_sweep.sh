# sweep of the LBVH size-class thresholds (denominators of radius / scene diagonal) at C4
for t in "48" "24" "32" "64" "96" "128,32" "96,32" "64,24" "128,48,16"; do
GRTB200_SIZE_T=$t timeout 300 python bench.py --workload c4 --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/c4_t.json
python - <<PY
import json; d=json.load(open("gpurun_out/c4_t.json")); print("T=$t", round(d["value"],1), {k: round(v,3) for k,v in d["stage_ms"].items()})
PY
done
