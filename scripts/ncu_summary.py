"""Summarise ncu outputs: `launches` CSV -> per-kernel totals; raw CSV of a --set full report -> key metrics."""
import collections, csv, sys

def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            agg.setdefault(r[ki][:70], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':72s} {'n':>4s} {'total ms':>10s} {'share':>6s}  first launches (ms)")
    for k, v in agg.items():
        print(f"{k:72s} {len(v):4d} {sum(v)/1e6:10.3f} {sum(v)/tot:6.3f}  {[round(x/1e6,2) for x in v[:9]]}")

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum.per_second",
        "l1tex__t_bytes.sum.per_second", "smsp__sass_average_branch_targets_threads_uniform.pct"]

def full(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    names = [r[hdr.index("Kernel Name")][:28] for r in rows[2:]] if "Kernel Name" in hdr else []
    print("kernels:", names)
    for k in KEYS + sorted(h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h):
        if k in hdr:
            i = hdr.index(k)
            print(f"{k[-62:]:62s} {rows[1][i]:>12s}", " ".join(f"{r[i][:11]:>11s}" for r in rows[2:]))

if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else full)(sys.argv[2])
