#!/bin/bash
# Usage: gpurun --timeout 600 -- 'bash tools/py_check.sh tag <pytest args>'
tag=$1; shift
out=gpurun_out/$tag
mkdir -p $out
timeout 500 python -m pytest "$@" -m gpu -x -q -s --timeout 300 > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee $out/summary.txt
grep -v "^$" $out/pytest.log | tail -${TAILN:-40}
