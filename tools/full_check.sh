#!/bin/bash
# round-end style verification: smoke, GPU tests, native tests, bench (writes logs to gpurun_out/)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
( cd tests/native && export LD_LIBRARY_PATH=../../behavior-driven-video-synthesis_b200:$LD_LIBRARY_PATH && timeout 200 ./test_kernels | tail -1 && timeout 200 ./test_bwd | tail -1 )
timeout 600 python bench.py > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; tail -c 1500 gpurun_out/bench_pair.json
