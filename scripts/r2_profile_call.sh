#!/bin/bash
# profiles of the shipped build: per-class instruction counts (roofline records), full captures of the top kernels
# (exported to CSV / SASS listings ON THE BOX: the reports themselves are too large to travel), launch list of the bench
P=gpurun_out/prof
mkdir -p $P
M=gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
for w in cfg2 cfg3 cfg5 cfg4; do
  export RT_PROFILE_EMIT=$([ $w = cfg4 ] && echo 1 || echo "")
  python scripts/profile_frame.py $w $P/frame_$w.json > $P/plain_$w.log 2>&1 && \
  timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $P/ncu_$w.csv python scripts/profile_frame.py $w $P/frame_ncu_$w.json > $P/ncu_$w.log 2>&1
  echo "ncu metrics $w rc=$?"
done
unset RT_PROFILE_EMIT
full() {  # name workload kernel-regex launches [env]
  python scripts/profile_frame.py $2 > /dev/null 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$3" -c $4 -o /tmp/$1 python scripts/profile_frame.py $2 > $P/full_$1.log 2>&1
  echo "full $1 rc=$?"
  python scripts/ncu_export.py /tmp/$1.ncu-rep $P/$1_full.csv > /dev/null 2>&1
}
full r2_cfg2 cfg2 "k_trace|k_shade|k_combine|k_sort_scatter" 12
python scripts/sass_hot.py /tmp/r2_cfg2.ncu-rep k_trace 1 > $P/r2_cfg2_anyhit_seg0_sass.txt 2>&1
python scripts/sass_hot.py /tmp/r2_cfg2.ncu-rep k_trace 0 > $P/r2_cfg2_nearest_seg0_sass.txt 2>&1
python scripts/sass_hot.py /tmp/r2_cfg2.ncu-rep k_shade 0 > $P/r2_cfg2_shade_seg0_sass.txt 2>&1
full r2_cfg3_knn cfg3 "k_shade" 3
python scripts/sass_hot.py /tmp/r2_cfg3_knn.ncu-rep k_shade 0 > $P/r2_cfg3_knn_seg0_sass.txt 2>&1
export RT_PROFILE_EMIT=1
full r2_cfg4_emit cfg4 "k_emit|k_shade" 2
python scripts/sass_hot.py /tmp/r2_cfg4_emit.ncu-rep k_emit 0 > $P/r2_cfg4_emit_sass.txt 2>&1
unset RT_PROFILE_EMIT
full r2_cfg5_trace cfg5 "k_trace" 6
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $P/bench_plain.json 2> $P/bench_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $P/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $P/bench_ncu.log 2>&1; echo "launch list rc=$?"
du -sh $P
