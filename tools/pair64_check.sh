#!/bin/bash
mkdir -p gpurun_out
cd tests/native
export LD_LIBRARY_PATH=../../behavior-driven-video-synthesis_b200:$LD_LIBRARY_PATH
timeout 400 ./test_kernels > ../../gpurun_out/pair_kernels.log 2>&1; echo "rc=$?"; grep -E "BN64|FAIL|ALL" ../../gpurun_out/pair_kernels.log
timeout 400 ./test_bwd > ../../gpurun_out/pair_bwd.log 2>&1; echo "rc=$?"; grep -E "BN64|FAIL|ALL" ../../gpurun_out/pair_bwd.log
for p in 0 1; do
echo "== VUB_PAIR=$p  [64|64]->64 @128x128 B64 ; 64->64 @64x64 B64 (fwd) ; dgrad 64<-64 @64x64 B32"
VUB_PAIR=$p ./test_kernels layer 64 128 128 64 2 64 3 1 1 1 0 0 | grep PERF
VUB_PAIR=$p ./test_kernels layer 64 64 64 64 1 64 3 1 1 1 0 0 | grep PERF
VUB_PAIR=$p ./test_bwd layer 32 64 64 64 1 64 3 1 0 | grep "dgrad "
done
cd ../..
for p in 0 1 0 1; do
  VUB_PAIR=$p timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pair_ab_$p.json 2> gpurun_out/bench_pair_ab_$p.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pair_ab_$p.json').read().strip().splitlines()[-1])
print('VUB_PAIR=$p', round(d['value']), round(d['train']['value']), round(d['sequence']['value']), d['clocks']['sm_mhz'])
PY
done
