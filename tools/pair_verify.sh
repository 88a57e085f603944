#!/bin/bash
mkdir -p gpurun_out
cd tests/native
export LD_LIBRARY_PATH=../../behavior-driven-video-synthesis_b200:$LD_LIBRARY_PATH
timeout 400 ./test_kernels > ../../gpurun_out/pair_kernels.log 2>&1; echo "rc=$?"; grep -E "FAIL|ALL" ../../gpurun_out/pair_kernels.log
timeout 400 ./test_bwd > ../../gpurun_out/pair_bwd.log 2>&1; echo "rc=$?"; grep -E "pair |FAIL|ALL" ../../gpurun_out/pair_bwd.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_pair --csv --log-file /tmp/pl.csv ./test_kernels > /dev/null 2>&1
echo "conv_pair launches in test_kernels: $(grep -c conv_pair_kernel /tmp/pl.csv)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_pair --csv --log-file /tmp/pb.csv ./test_bwd > /dev/null 2>&1
echo "conv_pair launches in test_bwd: $(grep -c conv_pair_kernel /tmp/pb.csv)"
echo "== forward concat [128|128]->128 @32x32 B64: slab (default) vs pair (slab off) vs single tiles"
./test_kernels layer 64 32 32 128 2 128 3 1 1 1 0 0 | grep PERF
VUB_ENABLE_SLAB=0 ./test_kernels layer 64 32 32 128 2 128 3 1 1 1 0 0 | grep PERF
VUB_ENABLE_SLAB=0 VUB_PAIR=0 ./test_kernels layer 64 32 32 128 2 128 3 1 1 1 0 0 | grep PERF
echo "== backward, B=32: 128->128 @64x64 ; [128|128]->128 @32x32 ; 128->128 @32x32 ; D2S 128->256 @32x32: default / slab off / slab off + single"
for shape in "32 64 64 128 1 128 3 1 0" "32 32 32 128 2 128 3 1 0" "32 32 32 128 1 128 3 1 0" "32 32 32 128 1 256 3 1 1"; do
  echo "-- $shape"
  ./test_bwd layer $shape | grep "dgrad "
  VUB_ENABLE_SLAB=0 ./test_bwd layer $shape | grep "dgrad "
  VUB_ENABLE_SLAB=0 VUB_PAIR=0 ./test_bwd layer $shape | grep "dgrad "
done
