"""Per-SASS-instruction view of one kernel launch of an ncu report (samples, executions, lanes).
usage: sass_hot.py report.ncu-rep kernel-regex launch-skip [min_samples]"""
import csv, subprocess, sys
rep, rx, skip = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
hdr = rows[1]
ia, isamp, iex, ith = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
ilsb, iw, iss = hdr.index("stall_long_sb"), hdr.index("stall_wait"), hdr.index("stall_short_sb")
data = [r for r in rows[2:] if len(r) > iw and r[isamp].isdigit()]
tot = sum(int(r[isamp]) for r in data); totex = sum(int(r[iex]) for r in data)
print("total samples", tot, "total warp-inst", totex, "n sass", len(data))
for k, r in enumerate(data):
    print(f"{k:4d} {int(r[isamp]):6d} {100*int(r[isamp])/tot:5.2f}% {int(r[iex])/1e6:8.2f}M {r[ith]:>3s} lsb={r[ilsb]:>5s} ssb={r[iss]:>5s} {r[ia].strip()[:90]}")
