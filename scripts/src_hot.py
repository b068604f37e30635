"""Per-source-line view of one kernel launch of an ncu report with --import-source: warp-instructions executed,
thread-instructions, average lanes and stall samples per CUDA source line (SASS aggregated by its -lineinfo line).
usage: src_hot.py report.ncu-rep kernel-regex launch-skip [top]"""
import collections, csv, subprocess, sys
rep, rx, skip = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
agg = collections.OrderedDict()
path, hdr = "", None
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        path = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        print(r[1][:110])
        continue
    if r[0] == "Line No":
        hdr = r
        iex, ith, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ith or not r[iex].isdigit():
        continue
    key = (path, r[0])
    a = agg.setdefault(key, [0, 0, 0, r[1].strip()[:105], 0])
    a[0] += int(r[iex]); a[1] += int(r[ith]); a[2] += int(r[isamp]) if r[isamp].isdigit() else 0; a[4] += 1
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[2] for a in agg.values()) or 1
print(f"total warp-inst {tot/1e9:.3f} G, thread-inst {sum(a[1] for a in agg.values())/1e9:.2f} G, samples {tots}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/tot:5.2f}% inst {100*a[2]/tots:5.2f}% samp lanes {a[1]/max(a[0],1):5.1f} sass {a[4]:3d}  {f}:{ln:>4s}  {a[3]}")
