#!/bin/bash
# round 2, first call: the whole GPU suite (no -x: see every failure), once more with the experimental k-buffer, then the default bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA --timeout 900 > gpurun_out/r02_gate_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02_gate_pytest.log
tail -40 gpurun_out/r02_gate_pytest.log | cut -c1-300
GUTB200_EXPERIMENTAL_KBUFFER=1 python -m pytest tests/test_kbuffer_gpu.py -m gpu -q -rA --timeout 600 > gpurun_out/r02_gate_kbuffer.log 2>&1
echo "kbuffer rc=$?" | tee -a gpurun_out/r02_gate_kbuffer.log
tail -30 gpurun_out/r02_gate_kbuffer.log | cut -c1-300
python bench.py > gpurun_out/r02_gate_bench.json 2> gpurun_out/r02_gate_bench.err
echo "bench rc=$?"
cut -c1-1500 gpurun_out/r02_gate_bench.json
