#!/bin/bash
# correctness and timing of the CTA-pair tile kernel (VUB_PAIR=1) against the single-CTA kernel
cd tests/native
export LD_LIBRARY_PATH=../../behavior-driven-video-synthesis_b200:$LD_LIBRARY_PATH
mkdir -p ../../gpurun_out
VUB_EXP_PIPE=1 timeout 120 ./exp_cta2 | grep pipe
VUB_PAIR=1 timeout 180 ./test_kernels > ../../gpurun_out/pair_kernels.log 2>&1; echo "test_kernels pair rc=$?"
tail -3 ../../gpurun_out/pair_kernels.log
grep -c FAIL ../../gpurun_out/pair_kernels.log
VUB_PAIR=1 timeout 180 ./test_bwd > ../../gpurun_out/pair_bwd.log 2>&1; echo "test_bwd pair rc=$?"
tail -2 ../../gpurun_out/pair_bwd.log
for pair in 0 1; do
  echo "== VUB_PAIR=$pair"
  for shape in "64 64 64 128 1 128 3 1 1 1 0 0" "64 32 32 128 2 128 3 1 1 1 0 0" "64 32 32 128 1 128 3 1 1 1 0 0" \
               "64 32 32 128 1 256 3 1 0 1 1 0" "64 16 16 128 1 128 3 1 1 1 0 0" "64 16 16 128 1 512 3 1 0 1 1 0" "64 64 64 64 1 128 3 1 0 1 1 0"; do
    VUB_ENABLE_SLAB=0 VUB_PAIR=$pair timeout 60 ./test_kernels layer $shape | grep PERF
  done
done
