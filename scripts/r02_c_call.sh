#!/bin/bash
# round 2, call C: full GPU suite (log kept), reference-kernels-on-GPU parity, default bench with the reference_gpu block,
# ncu --set full of our render kernels and of the reference's render / renderBackward (same tensors, same process)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_ref_cuda_gpu.py -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r02_c_pytest.log
tail -5 gpurun_out/r02_c_pytest.log
timeout 900 python -m pytest tests/test_ref_cuda_gpu.py -q -m gpu -s -p no:cacheprovider 2>&1 | grep -v "^$" | tail -120 > gpurun_out/r02_c_refcuda.log
tail -30 gpurun_out/r02_c_refcuda.log
timeout 600 python bench.py --steps 100 --warmup 10 2> gpurun_out/r02_c_bench.err | tail -1 > gpurun_out/r02_c_bench.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_c_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["stage_ms"])
print("reference_gpu", d.get("reference_gpu"))
print("vs", d.get("vs_reference_gpu"))
PY
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_elapsed.max,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"render_(forward|backward)_kernel|^render$|renderBackward" -s 8 -c 4 -o gpurun_out/r02_c_render -f python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r02_c_ncu.log 2>&1
ncu -i gpurun_out/r02_c_render.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_c_render.csv 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r02_c_render.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"]
if h:
    hd=rows[h[0]]
    for r in rows[h[0]+2:]:
        print(r[4][:70])
        for k,v in zip(hd[11:],r[11:]): print("   ",k,v)
PY
