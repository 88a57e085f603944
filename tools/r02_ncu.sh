#!/bin/bash
# Round-2 ncu evidence in one gpurun call:
#   1. launch list of the default bench command (every launch with its device time; shares, not absolutes)
#   2. metrics of the split-K cluster kernel on its benchmark shapes
#   3. metrics of the VGG-loss kernels (memory-bound kernels + the 256 / 512-channel per-tap convs)
# Usage: gpurun --timeout 900 -- 'bash tools/r02_ncu.sh [tag]'
tag=${1:-r02_ncu}
out=gpurun_out/$tag
mkdir -p $out
# 1 -- launch list: skip the construction / warm-up launches, then 600 launches (6+ graph replays of ~90 kernels)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 580 -c 400 --csv --log-file $out/launches.csv \
  python bench.py --steps 4 --warmup 3 --no-train --no-cpu-baseline > $out/launches_bench.log 2>&1
echo "launch list rc=$?"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_elapsed.max,\
smsp__inst_executed.sum,launch__registers_per_thread,launch__cluster_dim_z,launch__grid_size"
cd tests/native
for sk in 1 0; do
  for shape in "64 4 4 128 1 128 3 1 1 1 0 0" "64 4 4 128 2 128 3 1 1 1 0 0" "64 4 4 512 2 512 3 1 1 1 0 0" "64 2 2 512 2 512 3 1 1 1 0 0"; do
    name=splitk${sk}_$(echo $shape | tr ' ' '_')
    VUB_SPLITK=$sk timeout 120 ncu --metrics $M --clock-control none -k regex:conv_fwd --launch-skip 3 -c 1 --csv \
      --log-file ../../$out/$name.csv ./test_kernels layer $shape > /dev/null 2>&1
    echo "$name rc=$?"
  done
done
cd ../..
# 3 -- VGG loss, batch 8, one instance of every kernel family after warm-up
timeout 400 ncu --metrics $M --clock-control none -k regex:'l1_|maxpool|vgg_|conv_fwd_kernel|conv_tiles|conv_pair|conv_rows' \
  --launch-skip 300 -c 120 --csv --log-file $out/vgg_kernels.csv python tools/vgg_perf.py 8 > $out/vgg_ncu.log 2>&1
echo "vgg rc=$?"
python - <<PY
import csv, glob, collections
def rows(f):
    r = [x for x in csv.reader(open(f, errors="replace")) if len(x) > 10]
    return r[0], r[1:]
try:
    hdr, rs = rows("$out/launches.csv")
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rs:
        k = r[ki][:90]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open("$out/launches_summary.csv", "w") as f:
        f.write("kernel,launches,total_ns,share\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('"%s",%d,%.0f,%.4f\n' % (k, a[0], a[1], a[1] / tot))
    print(open("$out/launches_summary.csv").read()[:2500])
except Exception as e:
    print("launch list parse failed", e)
for f in sorted(glob.glob("$out/splitk*.csv")) + ["$out/vgg_kernels.csv"]:
    try:
        hdr, rs = rows(f)
        ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
        ii = hdr.index("ID")
        per = collections.OrderedDict()
        for r in rs:
            per.setdefault((r[ii], r[ki][:60]), {})[r[mi]] = r[vi]
        print(f)
        seen = set()
        for (i, k), m in per.items():
            key = (k, m.get("launch__grid_size"))
            if key in seen: continue
            seen.add(key)
            print("  %-60s t=%sns dramR=%s dramW=%s dram%%=%s tensor%%=%s grid=%s clusterZ=%s" % (k, m.get("gpu__time_duration.sum"),
                  m.get("dram__bytes_read.sum"), m.get("dram__bytes_write.sum"), m.get("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                  m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"), m.get("launch__grid_size"), m.get("launch__cluster_dim_z")))
    except Exception as e:
        print(f, "parse failed", e)
PY
