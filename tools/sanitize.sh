#!/bin/bash
# compute-sanitizer pass over the native kernel tests (race / memory checking of the mbarrier / TMEM /
# cluster kernels).  Run on the GPU box:  gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Writes gpurun_out/sanitize/<tool>_<binary>.log and a one-line-per-run summary (copied to profiles/ by hand).
out=gpurun_out/sanitize
mkdir -p $out
cd tests/native
CS=/usr/local/cuda/bin/compute-sanitizer
summary=../../$out/summary.txt
: > $summary
for tool in memcheck racecheck synccheck; do
  for bin in test_kernels test_bwd; do
    log=../../$out/${tool}_${bin}.log
    # VUB_SANITIZE=1: the binaries keep only the small cases (every kernel family incl. the CTA-pair kernel)
    VUB_SANITIZE=1 timeout 420 $CS --tool $tool --print-limit 20 ./$bin > $log 2>&1
    rc=$?
    errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $log | tail -1)
    passed=$(grep -c "^\[PASS\]" $log)
    failed=$(grep -c "^\[FAIL\]" $log)
    echo "$tool $bin rc=$rc cases_passed=$passed cases_failed=$failed :: $errs" | tee -a $summary
  done
done
