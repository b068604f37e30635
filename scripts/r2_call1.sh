#!/bin/bash
# first GPU call of round 2: tests, bench, probes, ncu metric captures for the roofline records
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke1.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest1.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest1.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
timeout 600 python scripts/r2_probe.py own batch cfg4 > gpurun_out/probe1.jsonl 2> gpurun_out/probe1.err; echo "probe rc=$?"
M=gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
for w in cfg2 cfg3 cfg5; do
  python scripts/profile_frame.py $w gpurun_out/frame_$w.json > gpurun_out/plain_$w.log 2>&1 && \
  timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/ncu_$w.csv python scripts/profile_frame.py $w gpurun_out/frame_ncu_$w.json > gpurun_out/ncu_$w.log 2>&1
  echo "ncu $w rc=$?"
done
