#!/bin/bash
# One gpurun call: native kernel tests (short timeouts: a hung kernel must not eat the GPU budget), A/B perf of the
# row kernel, the GPU pytest tier, smoke and a short bench.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag] [skip_pytest]'
tag=${1:-check}
out=gpurun_out/$tag
mkdir -p $out
: > $out/summary.txt
cd tests/native
stdbuf -oL timeout 240 ./test_kernels > ../../$out/native_fwd.log 2>&1; rc1=$?; echo "test_kernels rc=$rc1" | tee -a ../../$out/summary.txt
stdbuf -oL timeout 240 ./test_bwd > ../../$out/native_bwd.log 2>&1; rc2=$?; echo "test_bwd rc=$rc2" | tee -a ../../$out/summary.txt
grep -h "FAIL" ../../$out/native_fwd.log ../../$out/native_bwd.log | head -20
if [ $rc1 -ne 0 ] || [ $rc2 -ne 0 ]; then
  tail -3 ../../$out/native_fwd.log
  echo "native tests failed: stopping here" | tee -a ../../$out/summary.txt
  exit 1
fi
stdbuf -oL timeout 200 ./test_kernels perf > ../../$out/perf_fwd.log 2>&1; echo "perf rc=$?" | tee -a ../../$out/summary.txt
VUB_ROWS_NBLK=0 stdbuf -oL timeout 200 ./test_kernels perf > ../../$out/perf_fwd_nblk0.log 2>&1
VUB_ROWS_TMA=0 stdbuf -oL timeout 200 ./test_kernels perf > ../../$out/perf_fwd_tma0.log 2>&1
VUB_ROWS_TST=0 stdbuf -oL timeout 200 ./test_kernels perf > ../../$out/perf_fwd_tst0.log 2>&1
stdbuf -oL timeout 200 ./test_bwd perf > ../../$out/perf_bwd.log 2>&1
cd ../..
grep PERF $out/perf_fwd.log | head -14
echo "--- NBLK=0"
grep PERF $out/perf_fwd_nblk0.log | head -8
echo "--- TMA=0"
grep PERF $out/perf_fwd_tma0.log | head -8
echo "--- TST=0"
grep PERF $out/perf_fwd_tst0.log | head -8
if [ "$2" == "quick" ]; then exit 0; fi
if [ -z "$2" ]; then
  timeout 1100 python -m pytest tests -m gpu -x -q -s --timeout 400 > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/summary.txt
  tail -15 $out/pytest_gpu.log
fi
timeout 200 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/summary.txt
tail -2 $out/smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?" | tee -a $out/summary.txt
tail -3 $out/bench.err
python - <<PY
import json
try:
    b = json.loads(open("$out/bench.json").read().strip().splitlines()[-1])
    print("value", b["value"], "e2e", b["e2e"]["value"], "roofline", {k: b["roofline"].get(k) for k in ("kernel", "frac", "avg_launch_ms")})
    print("e2e_fp32", b["roofline"].get("e2e_fp32"))
    t = b["roofline"].get("train") or {}
    print("train", t.get("value"), t.get("ms_per_step"), t.get("worst_family"))
    for k in b["kernels"][:16]:
        print(k["kernel"], k["launches"], k["ms"], k["share"], k["tflops"], k["gbs"])
except Exception as e:
    print("bench parse failed", e)
PY
