#!/bin/bash
# smoke + short bench (+ optional pytest selection).  Usage: gpurun --timeout 900 -- 'bash tools/final_check.sh tag [pytest args]'
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
: > $out/summary.txt
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/summary.txt
grep smoke: $out/smoke.log; tail -3 $out/smoke.log | grep -v smoke:
shift
if [ -n "$1" ]; then
  timeout 900 python -m pytest "$@" -m gpu -x -q -s --timeout 600 > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/summary.txt
  grep -v "^$" $out/pytest.log | tail -12
fi
timeout 500 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?" | tee -a $out/summary.txt
tail -3 $out/bench.err
python - <<PY
import json
try:
    b = json.loads(open("$out/bench.json").read().strip().splitlines()[-1])
    print("value", b["value"], "e2e", b["e2e"]["value"], "roofline", {k: b["roofline"].get(k) for k in ("kernel", "frac", "avg_launch_ms")})
    t = b["roofline"].get("train") or {}
    print("train", t.get("value"), t.get("ms_per_step"), t.get("worst_family"))
    print("vunet_org", b["roofline"].get("vunet_org"))
    print("sequence", b["roofline"].get("sequence"))
    print("train_vgg", b["roofline"].get("train_vgg"))
    bs = b["roofline"].get("behavior_sampling") or {}
    print("behavior_sampling", {k: bs.get(k) for k in ("value", "ms_per_batch", "e2e", "error")})
    print("behavior_transfer_e2e", b["roofline"].get("behavior_transfer_e2e"))
    for k in b["kernels"][:12]:
        print(k["kernel"], k["launches"], k["ms"], k["share"], k["tflops"], k["gbs"])
except Exception as e:
    print("bench parse failed", e)
PY
