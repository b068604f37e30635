#!/bin/bash
# per-kernel durations of one cfg3 frame with the fine binning (ncu launch list; shares only, never bench values)
mkdir -p gpurun_out
export RT_SORT_BITS=7
python scripts/profile_frame.py cfg3 gpurun_out/frame_cfg3_bits7.json > /dev/null 2> gpurun_out/l24.err || { tail -5 gpurun_out/l24.err; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg3_bits7.csv python scripts/profile_frame.py cfg3 > gpurun_out/l24.log 2>&1
python - <<'PY'
import csv, re
rows = list(csv.reader(l for l in open("gpurun_out/launches_cfg3_bits7.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki])[:60]
    print(f"{name:60s} {r[vi]:>14s} {r[ui]}")
PY
