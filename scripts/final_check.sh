# end-of-round verification on one B200: GPU tests, smoke, default bench line, ncu launch list of the same command
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --steps 100 --warmup 10 2>&1 | tail -1 > gpurun_out/bench_c2.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c2.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d.get("cpu_baseline",{}).get("value"), d["optimizer_step"])
PY
