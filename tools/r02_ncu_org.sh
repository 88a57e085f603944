#!/bin/bash
# ncu launch list of one eager VunetOrg.transfer step (H256, 64 frames): per kernel family and grid the launch count,
# mean duration and share.  Usage: gpurun --timeout 600 -- 'bash tools/r02_ncu_org.sh [tag]'
tag=${1:-r02_ncu_org}
out=gpurun_out/$tag
mkdir -p $out
timeout 420 ncu --metrics gpu__time_duration.sum,launch__cluster_dim_z,launch__grid_size --print-units base --clock-control none \
  --profile-from-start off --csv --log-file $out/org_launches.csv python tools/org_ncu.py > $out/org_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $out/org_ncu.log
python - <<PY
import csv, collections
try:
    r = [x for x in csv.reader(open("$out/org_launches.csv", errors="replace")) if len(x) > 10]
    hdr, rs = r[0], r[1:]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for x in rs:
        per.setdefault((x[ii], x[ki][:70]), {})[x[mi]] = float(x[vi].replace(",", ""))
    agg = collections.OrderedDict()
    for (i, k), m in per.items():
        a = agg.setdefault((k, int(m.get("launch__cluster_dim_z", 0))), collections.Counter())
        a["n"] += 1
        a["ns"] += m["gpu__time_duration.sum"]
    tot = sum(a["ns"] for a in agg.values())
    lines = ["kernel,cluster_z,launches,mean_ns,total_ns,share"]
    for (k, cz), a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        lines.append('"%s",%d,%d,%.0f,%.0f,%.4f' % (k, cz, a["n"], a["ns"] / a["n"], a["ns"], a["ns"] / tot))
    lines.append("# %d launches, %.3f ms of kernel time (serialised, cold-cache ncu replays: shares, not absolutes)" % (
        sum(a["n"] for a in agg.values()), tot * 1e-6))
    open("$out/org_launches_summary.csv", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
except Exception as e:
    print("parse failed", e)
PY
