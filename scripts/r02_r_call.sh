#!/bin/bash
# call R: compute-sanitizer memcheck over the small 3DGUT parity cases (new binning, hit words, 20-float gradient rows, k-buffer, compact exchange)
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gut_parity_gpu.py tests/test_kbuffer_gpu.py -m gpu -q -x -k "c1 or empty or per_pixel or compact or kbuffer or error or rejects" > gpurun_out/r02_r_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_r_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r02_r_memcheck.log
tail -12 gpurun_out/r02_r_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 10 python -m pytest tests/test_gut_parity_gpu.py -m gpu -q -x -k "c1_forward_and_gradients" > gpurun_out/r02_r_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_r_racecheck.log
tail -8 gpurun_out/r02_r_racecheck.log
