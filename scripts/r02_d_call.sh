#!/bin/bash
# round 2, call D: memcheck of the reference-kernel library on C1, GPU suite on the hit-word / canonical-sum backward, bench, ncu counters
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/refcuda_debug.py > gpurun_out/r02_d_memcheck.log 2>&1
grep -v "^=========     Host Frame\|^=========         Host Frame" gpurun_out/r02_d_memcheck.log | head -60
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_ref_cuda_gpu.py -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|FAILED|Error|\[hit words\]|\[work\]|assert" | tail -40 > gpurun_out/r02_d_pytest.log
cat gpurun_out/r02_d_pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-reference-gpu 2> gpurun_out/r02_d_bench.err | tail -1 > gpurun_out/r02_d_bench.json
tail -3 gpurun_out/r02_d_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_d_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["stage_ms"])
print("fp32", json.dumps(d.get("roofline_fp32")))
PY
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_elapsed.max,lts__t_sectors_op_red.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:"render_(forward|backward)_kernel|project_backward" -s 12 -c 6 --csv --log-file gpurun_out/r02_d_ncu.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-reference-gpu > gpurun_out/r02_d_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r02_d_ncu.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"]
if h:
    for r in rows[h[0]+1:]:
        if len(r)>14: print(r[4][:60], r[-3], r[-1])
PY
