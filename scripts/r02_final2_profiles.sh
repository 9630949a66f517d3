#!/bin/bash
# end-of-round evidence on one B200: full GPU suite (log kept), smoke, default bench line, ncu launch list of the same command,
# one ncu --set full capture of the top kernels (summaries extracted to CSV; the .ncu-rep stays in gpurun_out/)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/r02_zz_pytest.log
cat gpurun_out/r02_zz_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 100 --warmup 10 2> gpurun_out/r02_zz_bench.err | tail -1 > gpurun_out/r02_zz_bench.json
tail -2 gpurun_out/r02_zz_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_zz_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d["stage_ms"], "ref", d.get("reference_gpu",{}).get("value"), "c3", d["c3"]["value"], "c4", d["c4"]["value"], "train_default", d["train_default"]["value"], d["train_default"]["stage_ms"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/r02_zz_launches.csv python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-reference-gpu --no-sub-records > gpurun_out/r02_zz_ncu_launch.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"render_(forward|backward)_kernel|project_kernel|project_backward_kernel|expand_place|tile_sort|tile_scan" -s 14 -c 7 -o gpurun_out/r02_zz_full -f python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-reference-gpu --no-sub-records > gpurun_out/r02_zz_ncu_full.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
ncu -i gpurun_out/r02_zz_full.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_zz_full.csv 2>&1
ncu -i gpurun_out/r02_zz_full.ncu-rep --page source --csv --kernel-name regex:render_backward > gpurun_out/r02_zz_bwd_source.csv 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r02_zz_full.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"]
if h:
    hd=rows[h[0]]
    for r in rows[h[0]+2:]:
        print(r[4][:60], {k.split("__")[-1][:28]:v for k,v in zip(hd[11:],r[11:])})
PY
