#!/bin/bash
# ncu capture of the row-streaming kernel on the benchmark's dominant layer (32->32 3x3 + residual @256x256, 64 frames,
# lean I/O: raw source, ELU on load, raw output) in its round-1 configuration (VUB_ROWS_NBLK=0 VUB_ROWS_TMA=0) and in
# the current one.  Usage: gpurun --timeout 900 -- 'bash tools/rows_ncu.sh [tag]'
tag=${1:-rows_ncu}
out=gpurun_out/$tag
mkdir -p $out
cd tests/native
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,\
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed,\
l1tex__data_pipe_tc_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,\
l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,\
smsp__issue_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,l1tex__m_xbar2l1tex_read_bytes.sum,\
smsp__inst_executed.sum,launch__registers_per_thread"
run() {  # name, env..., layer args
  name=$1; shift
  env "$@" VUB_TEST_ELU_IN=1 timeout 200 ncu --metrics $M --clock-control none -k regex:conv_rows --launch-skip 3 -c 1 --csv \
    --log-file ../../$out/$name.csv ./test_kernels layer 64 256 256 32 1 32 3 1 1 0 0 0 > /dev/null 2>&1
  echo "$name rc=$?"
}
run r01_config VUB_ROWS_NBLK=0 VUB_ROWS_TMA=0
run nblk_only VUB_ROWS_NBLK=1 VUB_ROWS_TMA=0
run current VUB_ROWS_NBLK=1
# full capture of the current configuration with source-level stall attribution
VUB_TEST_ELU_IN=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_rows --launch-skip 3 -c 1 \
  -o ../../$out/rows_full ./test_kernels layer 64 256 256 32 1 32 3 1 1 0 0 0 > /dev/null 2>&1
echo "full rc=$?"
cd ../..
python - <<PY
import csv, glob
for f in sorted(glob.glob("$out/*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    if not rows: continue
    hdr = rows[0]
    mi, vi = hdr.index("Metric Name"), hdr.index("Metric Value")
    print(f)
    for r in rows[1:]:
        print("   %-72s %s" % (r[mi], r[vi]))
PY
