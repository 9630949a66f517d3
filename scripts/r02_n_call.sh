#!/bin/bash
# call N: sub-block width of the backward walk again, now that the per-pair adjoint is 94 instructions shorter; ncu of the default
mkdir -p gpurun_out
for m in 7 23 39; do
  GUTB200_SUBTILE_CULLING=$m python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-reference-gpu --no-sub-records > gpurun_out/r02_n_bench_$m.json 2> gpurun_out/r02_n_bench_$m.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_n_bench_$m.json").read().strip().splitlines()[-1])
print("mode $m:", round(d["value"],1), "frames/s", {k:round(v,3) for k,v in d["stage_ms"].items()})
PY
done
ncu --set full --clock-control none --import-source on -k regex:render_backward -s 12 -c 1 -o gpurun_out/r02_n_bwd python bench.py --steps 4 --warmup 10 --no-cpu-baseline --no-reference-gpu --no-sub-records > /dev/null 2> gpurun_out/r02_n_ncu.err
ls -la gpurun_out/r02_n_bwd.ncu-rep
