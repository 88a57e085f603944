#!/bin/bash
# Timing of the behavior net's linear layers (batch 256) on the per-tap kernel under different split-K settings.
cd tests/native
for sk in 1 0 2 4 6 8; do
  echo "--- VUB_SPLITK=$sk"
  for shape in "256 1 1 2048 1 2048 1 1 0 0 0 0" "256 1 1 512 1 2048 1 1 0 0 0 0" "256 1 1 2048 1 512 1 1 0 0 0 0" "256 1 1 1088 1 4096 1 1 0 0 0 0" "256 1 1 1024 1 64 1 1 0 0 0 0"; do
    VUB_SPLITK=$sk timeout 60 ./test_kernels layer $shape | grep PERF | sed "s/layer/$shape/"
  done
done
