#!/bin/bash
mkdir -p gpurun_out
GUTB200_EXP_SKIP_WAIT=1 python bench.py --steps 400 --warmup 10 --no-cpu-baseline --no-reference-gpu --no-sub-records --profile-host gpurun_out/r02_k_host_profile_skip.txt > gpurun_out/r02_k_a.json 2> gpurun_out/r02_k_a.err
python bench.py --steps 400 --warmup 10 --no-cpu-baseline --no-reference-gpu --no-sub-records --profile-host gpurun_out/r02_k_host_profile_wait.txt > gpurun_out/r02_k_b.json 2> gpurun_out/r02_k_b.err
head -50 gpurun_out/r02_k_host_profile_skip.txt
