#!/bin/bash
# per-GPU cost of the strong-scaling reading of configs[1] (64 frames split over N GPUs -> 64/N frames per GPU), then
# the default line once more
mkdir -p gpurun_out
for t in 8 32; do
  timeout 200 python bench.py --frames-total $t --no-train --no-cpu-baseline --steps 30 --warmup 3 > gpurun_out/bench_strong$t.json 2> gpurun_out/bench_strong$t.err
done
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<'PY'
import json
for t in ("strong8","strong16","strong32","default"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{t}.json").read().strip().splitlines()[-1])
        print(t, d["scaling"], d["config"]["frames_per_gpu_per_step"], round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d.get("train",{}) and round(d["train"]["value"]))
    except Exception as e:
        print(t, "missing", e)
PY
