"""Join an ncu --metrics CSV of one profiled frame (scripts/profile_frame.py under `ncu --profile-from-start off
--metrics ... --csv --log-file`) with the frame's logical work into profiles/ncu_<workload>.json -- the file
bench.py reads to compute the issue-slot roofline of the dominant kernel.

usage: ncu_roofline.py launches.csv frame.json out.json [ptxas.log]"""
import collections, csv, json, re, subprocess, sys

METRICS = ["gpu__time_duration.sum", "smsp__thread_inst_executed.sum", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers_per_thread",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg"]


def kernel_class(name):
    if "k_trace" in name:
        m = re.search(r"k_trace\w*<\s*(?:\(bool\))?(\d)", name)
        return "trace_any" if m and m.group(1) == "1" else "trace_nearest"
    for key, cls in (("k_raygen", "raygen"), ("k_sort", "sort"), ("k_shade", "shade"), ("k_combine", "combine"),
                     ("k_resolve", "resolve"), ("k_emit", "emit")):
        if key in name:
            return cls
    return "other"


def main():
    csv_path, frame_path, out_path = sys.argv[1:4]
    rows = list(csv.reader(open(csv_path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    col = {h: hdr.index(h) for h in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        d = launches.setdefault(r[col["ID"]], {"name": r[col["Kernel Name"]]})
        try:
            v = float(r[col["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        unit = r[col["Metric Unit"]]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit) if r[col["Metric Name"]].startswith("gpu__time") else None
        scale = scale or {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)
        d[r[col["Metric Name"]]] = v * scale
    frame = json.load(open(frame_path))
    classes = collections.OrderedDict()
    for d in launches.values():
        c = classes.setdefault(kernel_class(d["name"]), dict(launches=0, ms_under_ncu=0.0, thread_inst=0.0, warp_inst=0.0,
                                                               dram_bytes=0.0, l2_bytes=0.0, issue_active_pct=[],
                                                               warps_active_pct=[], registers=set(), l1_hit_pct=[],
                                                               l2_hit_pct=[], kernels=set()))
        c["launches"] += 1
        c["ms_under_ncu"] += d.get("gpu__time_duration.sum", 0.0)
        c["thread_inst"] += d.get("smsp__thread_inst_executed.sum", 0.0)
        c["warp_inst"] += d.get("smsp__inst_executed.sum", 0.0)
        c["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        c["l2_bytes"] += d.get("lts__t_bytes.sum", 0.0)
        for key, dst in (("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                         ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
                         ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct")):
            if key in d:
                c[dst].append(round(d[key], 1))
        if "launch__registers_per_thread" in d:
            c["registers"].add(int(d["launch__registers_per_thread"]))
        c["kernels"].add(re.sub(r"\(.*", "", d["name"])[:80])
    total_ms = sum(c["ms_under_ncu"] for c in classes.values()) or 1.0
    for c in classes.values():
        c["lanes_per_warp_inst"] = round(c["thread_inst"] / c["warp_inst"], 2) if c["warp_inst"] else None
        c["share_under_ncu"] = round(c["ms_under_ncu"] / total_ms, 4)
        c["registers"], c["kernels"] = sorted(c["registers"]), sorted(c["kernels"])
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out = dict(source=csv_path, frame=frame, classes=classes, captured_at_commit=head,
               note="per-class sums over ONE frame; times under ncu are cold-cache and serialised: compare shares")
    json.dump(out, open(out_path, "w"), indent=1)
    print(f"{out_path}: {sum(c['launches'] for c in classes.values())} launches")
    for k, c in classes.items():
        print(f"  {k:14s} n={c['launches']:3d} share {c['share_under_ncu']:.3f} lanes {c['lanes_per_warp_inst']} "
              f"issue {c['issue_active_pct'][:6]} regs {c['registers']} dram {c['dram_bytes'] / 1e9:.2f} GB")


if __name__ == "__main__":
    main()
