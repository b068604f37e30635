"""Export a curated, transposed metric table from an .ncu-rep (one column per launch) into profiles/.
usage: ncu_export.py report.ncu-rep out.csv"""
import csv, re, subprocess, sys
rep, out_path = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = re.compile(r"^(Kernel Name|Grid Size|Block Size|gpu__time_duration\.sum|launch__(registers_per_thread|grid_size|block_size|occupancy_limit_\w+|shared_mem_per_block_dynamic|waves_per_multiprocessor)|sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__thread_inst_executed_per_inst_executed\.ratio|smsp__inst_executed\.sum|sm__inst_executed_pipe_(fma|alu|lsu|fp64|xu|fmaheavy|fmalite)\.avg\.pct_of_peak_sustained_active|l1tex__t_sector_hit_rate\.pct|lts__t_sector_hit_rate\.pct|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|lts__t_bytes\.sum\.per_second|l1tex__t_bytes\.sum\.per_second|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__pcsamp_warps_issue_stalled_\w+|smsp__sass_average_branch_targets_threads_uniform\.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|smsp__cycles_active\.avg)$")
out = [["metric", "unit"] + [f"launch{i}" for i in range(len(data))]]
for j, h in enumerate(hdr):
    if want.match(h) and "not_issued" not in h:
        out.append([h, units[j]] + [r[j] for r in data])
csv.writer(open(out_path, "w")).writerows(out)
print(len(out), "metrics x", len(data), "launches ->", out_path)
