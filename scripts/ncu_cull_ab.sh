set -x
for m in 0 3; do
GUTB200_SUBTILE_CULLING=$m timeout 300 ncu --set full --clock-control none -k regex:render_ -s 6 -c 2 -o gpurun_out/cull_$m -f python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_$m.log 2>&1
ncu -i gpurun_out/cull_$m.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread > gpurun_out/cull_$m.csv 2>&1
done
GUTB200_SUBTILE_CULLING=1 timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/c2_cull1.json
GUTB200_SUBTILE_CULLING=0 timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/c2_cull0.json
python - <<'PY'
import json
for f in ("gpurun_out/c2_cull1.json","gpurun_out/c2_cull0.json"):
    d=json.load(open(f)); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["stage_ms"]["render"], d["stage_ms"]["render_backward"])
PY
cat gpurun_out/cull_0.csv | cut -c1-600
cat gpurun_out/cull_3.csv | cut -c1-600
