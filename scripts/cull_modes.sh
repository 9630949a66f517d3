# A/B of the sub-tile culling switch (GUTB200_SUBTILE_CULLING: bit 0 renderBackward, bit 1 render); prints fps and stage ms
W=${1:-c2}; S=${2:-100}
for m in 0 1 2 3; do
GUTB200_SUBTILE_CULLING=$m timeout 300 python bench.py --workload $W --steps $S --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${W}_cull$m.json
python - <<PY
import json
d=json.load(open("gpurun_out/${W}_cull$m.json")); print("mode $m", round(d["value"],1), round(d["ms_per_step"],4), round(d["e2e"]["value"],1), round(d["stage_ms"]["render"],4), round(d["stage_ms"]["render_backward"],4))
PY
done
