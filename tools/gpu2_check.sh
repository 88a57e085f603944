#!/bin/bash
# 2-GPU run: data-parallel gradient equivalence over NCCL (vub_allreduce_bucket on the package's own communicator)
# and the bench at N=2 (frames sharded, training with the gradient exchange captured in the CUDA graph).
# Usage: gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu2_check.sh [tag]'
tag=${1:-gpu2}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L | tee $out/gpus.txt
timeout 400 python -m pytest tests/test_gpu_ddp.py -x -q -s > $out/pytest_ddp.log 2>&1; echo "pytest ddp rc=$?" | tee $out/summary.txt
tail -5 $out/pytest_ddp.log
NCCL_DEBUG=WARN timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_2gpu.json 2> $out/bench_2gpu.err; echo "bench2 rc=$?" | tee -a $out/summary.txt
tail -3 $out/bench_2gpu.err
python - <<PY
import json
try:
    b = json.loads(open("$out/bench_2gpu.json").read().strip().splitlines()[-1])
    print("N=2 value", b["value"], "e2e", b["e2e"]["value"])
    t = b["roofline"].get("train") or {}
    print("train", t.get("value"), t.get("ms_per_step"), "allreduce bytes", t.get("allreduce_bytes_per_step"), "e2e", (t.get("e2e") or {}).get("value"))
except Exception as e:
    print("bench parse failed", e)
PY
