#!/bin/bash
# N-GPU bench exactly as the driver launches it.  Usage: gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu8_bench.sh 8 tag'
n=${1:-8}
out=gpurun_out/${2:-gpu8}
mkdir -p $out
NCCL_DEBUG=WARN timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus $n --steps 20 --warmup 5 > $out/bench_${n}gpu.json 2> $out/bench_${n}gpu.err; echo "bench rc=$?"
tail -2 $out/bench_${n}gpu.err | cut -c1-300
python - <<PY
import json
try:
    b = json.loads(open("$out/bench_${n}gpu.json").read().strip().splitlines()[-1])
    print("N=$n value", b["value"], "e2e", b["e2e"]["value"], "fp32 e2e", b["roofline"]["e2e_fp32"]["value"])
    t = b["roofline"].get("train") or {}
    print("train", t.get("value"), t.get("ms_per_step"), "graph", t.get("cuda_graph"), "allreduce bytes", t.get("allreduce_bytes_per_step"), "e2e", (t.get("e2e") or {}).get("value"))
except Exception as e:
    print("bench parse failed", e)
PY
