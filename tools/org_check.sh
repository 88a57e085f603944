#!/bin/bash
# One gpurun call for the split-K / VunetOrg work: native small-map cases with split-K on and off (+ timings), then the
# VunetOrg GPU tests.  Usage: gpurun --timeout 600 -- 'bash tools/org_check.sh [tag] [pytest args]'
tag=${1:-org}
out=gpurun_out/$tag
mkdir -p $out
: > $out/summary.txt
cd tests/native
stdbuf -oL timeout 120 ./test_kernels splitk > ../../$out/native_splitk.log 2>&1; rc1=$?; echo "test_kernels splitk rc=$rc1" | tee -a ../../$out/summary.txt
grep -h "FAIL" ../../$out/native_splitk.log | head -20
grep -h "PERF\|---" ../../$out/native_splitk.log
stdbuf -oL timeout 120 ./test_bwd wide > ../../$out/native_bwd_wide.log 2>&1; rc2=$?; echo "test_bwd wide rc=$rc2" | tee -a ../../$out/summary.txt
grep -h "FAIL" ../../$out/native_bwd_wide.log | head -20
cd ../..
if [ $rc2 -eq 124 ]; then echo "native bwd test hung: stopping" | tee -a $out/summary.txt; tail -3 $out/native_bwd_wide.log; exit 1; fi
if [ $rc1 -eq 124 ]; then echo "native test hung: stopping" | tee -a $out/summary.txt; tail -3 $out/native_splitk.log; exit 1; fi
shift
timeout 400 python -m pytest ${@:-tests/test_gpu_org.py} -m gpu -x -q -s --timeout 200 > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/summary.txt
grep -v "^$" $out/pytest.log | tail -40
