#!/bin/bash
out=gpurun_out/${1:-gpu2bench}
mkdir -p $out
NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_2gpu.json 2> $out/bench_2gpu.err; echo "bench2 rc=$?"
tail -3 $out/bench_2gpu.err | cut -c1-300
python - <<PY
import json
try:
    b = json.loads(open("$out/bench_2gpu.json").read().strip().splitlines()[-1])
    print("N=2 value", b["value"], "e2e", b["e2e"]["value"], "fp32 e2e", b["roofline"]["e2e_fp32"]["value"])
    t = b["roofline"].get("train") or {}
    print("train", t.get("value"), t.get("ms_per_step"), "graph", t.get("cuda_graph"), "allreduce bytes", t.get("allreduce_bytes_per_step"), "e2e", (t.get("e2e") or {}).get("value"))
    print(b["train"].get("cuda_graph_error"))
except Exception as e:
    print("bench parse failed", e)
PY
