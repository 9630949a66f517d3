# 1-GPU reference point, then 2-GPU view-parallel with both gradient exchanges (same box)
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/s1.json
for ex in compact allreduce; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline --exchange $ex 2>&1 | tail -1 > gpurun_out/s2_$ex.json
done
python - <<'PY'
import json
for f in ("s1","s2_compact","s2_allreduce"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4), round(d["e2e"]["value"],1), d["config"].get("exchange"), {k: round(v,3) for k,v in d["stage_ms"].items()})
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/{f}.json").read()[-600:])
PY
