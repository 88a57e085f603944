#!/bin/bash
mkdir -p gpurun_out
for p in 0 1 0 1; do
  VUB_PAIR=$p timeout 600 python bench.py > gpurun_out/bench_pair_ab_$p.json 2> gpurun_out/bench_pair_ab_$p.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pair_ab_$p.json').read().strip().splitlines()[-1])
print('VUB_PAIR=$p', round(d['value']), round(d['train']['value']), round(d['sequence']['value']), d['clocks']['sm_mhz'])
PY
done
