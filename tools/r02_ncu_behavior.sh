#!/bin/bash
# ncu metrics of one eager behavior-net sampling pass (batch 256, 50 frames; 502 launches): per kernel family and grid
# the launch count, mean duration, DRAM bytes and tensor-pipe activity.  Usage: gpurun --timeout 600 -- 'bash tools/r02_ncu_behavior.sh [tag]'
tag=${1:-r02_ncu_behavior}
out=gpurun_out/$tag
mkdir -p $out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__cluster_dim_z,launch__grid_size"
timeout 420 ncu --metrics $M --print-units base --clock-control none -k regex:'conv_fwd_kernel|gather_affine|couple_inv|lstm_|decoder_' \
  --launch-skip 1004 -c 502 --csv --log-file $out/behavior_kernels.csv python tools/behavior_ncu.py > $out/behavior_ncu.log 2>&1
echo "ncu rc=$?"; tail -4 $out/behavior_ncu.log
python - <<PY
import csv, collections
try:
    r = [x for x in csv.reader(open("$out/behavior_kernels.csv", errors="replace")) if len(x) > 10]
    hdr, rs = r[0], r[1:]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for x in rs:
        per.setdefault((x[ii], x[ki][:70]), {})[x[mi]] = float(x[vi].replace(",", ""))
    agg = collections.OrderedDict()
    for (i, k), m in per.items():
        a = agg.setdefault((k, int(m.get("launch__grid_size", 0)), int(m.get("launch__cluster_dim_z", 0))), collections.Counter())
        a["n"] += 1
        for name in m:
            a[name] += m[name]
    tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
    lines = ["kernel,grid,cluster_z,launches,mean_ns,share,mean_dram_read_B,mean_dram_write_B,mean_dram_pct,mean_tensor_pct"]
    for (k, g, cz), a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = a["n"]
        lines.append('"%s",%d,%d,%d,%.0f,%.4f,%.0f,%.0f,%.1f,%.1f' % (
            k, g, cz, n, a["gpu__time_duration.sum"] / n, a["gpu__time_duration.sum"] / tot, a["dram__bytes_read.sum"] / n,
            a["dram__bytes_write.sum"] / n, a["dram__throughput.avg.pct_of_peak_sustained_elapsed"] / n,
            a["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"] / n))
    lines.append("# %d launches, %.3f ms of kernel time (serialised, cold-cache ncu replays: shares, not absolutes)" % (
        sum(a["n"] for a in agg.values()), tot * 1e-6))
    open("$out/behavior_kernels_summary.csv", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
except Exception as e:
    print("parse failed", e)
PY
