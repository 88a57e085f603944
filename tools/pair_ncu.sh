#!/bin/bash
cd tests/native
export LD_LIBRARY_PATH=../../behavior-driven-video-synthesis_b200:$LD_LIBRARY_PATH
mkdir -p ../../gpurun_out
timeout 120 ./exp_cta2 > ../../gpurun_out/exp_cta2_pipe.log 2>&1
echo "exp rc=$?"; cat ../../gpurun_out/exp_cta2_pipe.log
VUB_ENABLE_SLAB=0 VUB_PAIR=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_pair --launch-skip 3 -c 1 \
  -o /tmp/pair ./test_kernels layer 64 64 64 128 1 128 3 1 1 1 0 0 > /dev/null 2>&1
ncu -i /tmp/pair.ncu-rep --page source --csv --print-source sass > /tmp/pair_src.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/pair_src.csv')))
h0=next(i for i,r in enumerate(rows) if len(r)>5)
rows=rows[h0:]
hdr=rows[0]
idx=[i for i,h in enumerate(hdr) if 'Sampling (All' in h]
print(hdr)
if idx:
    i=idx[0]
    def f(r):
        try: return float(r[i])
        except: return 0.0
    body=[r for r in rows[1:] if len(r)==len(hdr)]
    tot=sum(f(r) for r in body)
    top=sorted(body,key=f,reverse=True)[:50]
    with open('../../gpurun_out/pair_src_top.csv','w') as o:
        w=csv.writer(o); w.writerow(hdr)
        for r in top: w.writerow(r)
    print('total samples',tot)
PY
ls -la ../../gpurun_out/
