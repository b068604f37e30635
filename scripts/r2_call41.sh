#!/bin/bash
# refresh the cfg3 captures after the k_shade / k_combine change, then the default bench of the final build
P=gpurun_out/prof
mkdir -p $P
M=gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
for w in cfg3 cfg4; do
  export RT_PROFILE_EMIT=$([ $w = cfg4 ] && echo 1 || echo "")
  python scripts/profile_frame.py $w $P/frame_$w.json > $P/plain_$w.log 2>&1 && \
  timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $P/ncu_$w.csv python scripts/profile_frame.py $w $P/frame_ncu_$w.json > $P/ncu_$w.log 2>&1
  echo "ncu metrics $w rc=$?"
done
unset RT_PROFILE_EMIT
python scripts/profile_frame.py cfg3 > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_shade" -c 3 -o /tmp/r2_cfg3_knn python scripts/profile_frame.py cfg3 > $P/full_r2_cfg3_knn.log 2>&1
echo "full rc=$?"
python scripts/ncu_export.py /tmp/r2_cfg3_knn.ncu-rep $P/r2_cfg3_knn_full.csv > /dev/null 2>&1
python scripts/sass_hot.py /tmp/r2_cfg3_knn.ncu-rep k_shade 0 > $P/r2_cfg3_knn_seg0_sass.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench41.json 2> gpurun_out/bench41.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench41.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"])
for k, v in d["extra"].items():
    if "value" in v: print(k, v["value"], v["ms_per_step"], v["e2e"]["value"], v["roofline"]["frac"])
    else: print(k, v.get("ms_per_call"), v.get("speedup_vs_reference_program"))
PY
