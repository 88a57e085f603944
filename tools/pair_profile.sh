#!/bin/bash
# ncu evidence for the CTA-pair tile kernel + launch list of the synthesis bench (written to gpurun_out/)
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/behavior-driven-video-synthesis_b200:$LD_LIBRARY_PATH
( cd tests/native && VUB_ENABLE_SLAB=0 timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_pair --launch-skip 3 -c 1 \
  -o /tmp/pair ./test_kernels layer 64 64 64 128 1 128 3 1 1 1 0 0 > /dev/null 2>&1 )
ncu -i /tmp/pair.ncu-rep --page raw --csv > /tmp/pair_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/pair_raw.csv')))
hdr,units,val=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg','sm__cycles_active.avg',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__cluster_size','launch__cluster_max_active',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_tc_wavefronts_mem_shared.sum','sm__inst_executed.sum','lts__t_bytes.sum','l1tex__m_xbar2l1tex_read_bytes.sum']
with open('gpurun_out/pair_ncu_summary.txt','w') as o:
    o.write('# ncu --set full --clock-control none, conv_pair_kernel<128,64>, 3x3 128->128 @64x64 +res dual, B=64 (tests/native/test_kernels layer ...)\n')
    for h,u,v in zip(hdr,units,val):
        if any(h.endswith(w) or h==w for w in want):
            o.write(f'{h:90s} {v} {u}\n')
print(open('gpurun_out/pair_ncu_summary.txt').read())
PY
ncu -i /tmp/pair.ncu-rep --page source --csv --print-source sass > /tmp/pair_src.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/pair_src.csv')))
h0=next(i for i,r in enumerate(rows) if len(r)>5)
rows=rows[h0:]; hdr=rows[0]
i=[k for k,h in enumerate(hdr) if 'Sampling (All' in h][0]
def f(r):
    try: return float(r[i])
    except: return 0.0
body=[r for r in rows[1:] if len(r)==len(hdr)]
tot=sum(f(r) for r in body)
top=sorted(body,key=f,reverse=True)[:25]
with open('gpurun_out/pair_ncu_summary.txt','a') as o:
    o.write(f'\n# top SASS lines by warp-stall samples (total {tot:.0f}): samples, executions, instruction\n')
    for r in top: o.write(f'{r[i]:>7s} {r[5]:>9s}  {r[1][:120]}\n')
PY
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 2 --warmup 1 --no-train --no-cpu-baseline > gpurun_out/launches_v4_bench.log 2>&1
python - <<'PY'
import csv,collections,re
rows=[r for r in csv.reader(open('gpurun_out/launches_v4.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    name=re.sub(r'\(.*$','',r[ki].replace('(int)','').replace('(bool)',''))
    agg[name][0]+=1; agg[name][1]+=v/1000.0
tot=sum(v[1] for v in agg.values())
with open('gpurun_out/launches_v4_summary.csv','w') as o:
    o.write('# ncu --metrics gpu__time_duration.sum --clock-control none -c 600, python bench.py --steps 2 --warmup 1 --no-train --no-cpu-baseline (first 600 launches: eager warm-up + CUDA-graph capture + replays); cold-cache serialised times: shares, not absolutes\n')
    o.write('kernel,launches,total_us,share\n')
    for k,v in sorted(agg.items(),key=lambda kv:-kv[1][1]): o.write(f'"{k}",{v[0]},{v[1]:.1f},{v[1]/tot:.4f}\n')
print(open('gpurun_out/launches_v4_summary.csv').read()[:1500])
PY
