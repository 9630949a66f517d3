# one ncu --set full capture of the forward and backward trace kernels at C4 + the raw metrics used in profiles/
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -s 6 -c 2 -o gpurun_out/grt_c4 -f python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_grt.log 2>&1
ncu -i gpurun_out/grt_c4.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio > gpurun_out/grt_c4.csv 2>&1
cat gpurun_out/grt_c4.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    print(r[4][:60])
    for k,v in zip(h[11:],r[11:]): print('   ',k,v)
"
