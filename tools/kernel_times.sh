#!/usr/bin/env bash
# per-kernel ncu launch list (gpu__time_duration etc.) of one steady-state step of tools/sweep.py; usage: tools/kernel_times.sh <out.csv> "<sweep spec>"
out=$1; shift
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"k_look|k_stitch|k_spec|k_diag_w|k_hint|k_pack|k_fix|k_emit|k_mark|k_link|k_resolve|k_scan|k_results|k_fill|k_make" --launch-skip 150 -c 40 --csv --log-file "$out" python tools/sweep.py "$@" > "${out%.csv}.log" 2>&1
python - "$out" <<PY
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, mi, vi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
out = {}
for r in rows[hdr + 1:]:
    if len(r) > vi:
        out.setdefault((int(r[0]), r[ki].split("(")[0][-28:]), {})[r[mi]] = r[vi]
for k in sorted(out)[-24:]:
    v = out[k]
    print("%3d %-28s %8.1f us  lanes %5s  winst %6.1fM  dram %7.1f MB  warps %s%%" % (k[0], k[1], float(v["gpu__time_duration.sum"]) / 1e3,
          v["smsp__thread_inst_executed_per_inst_executed.ratio"], float(v["smsp__inst_executed.sum"]) / 1e6,
          (float(v["dram__bytes_read.sum"]) + float(v["dram__bytes_write.sum"])) / 1e6, v["sm__warps_active.avg.pct_of_peak_sustained_active"]))
PY
