"""Condense a scripts/sass_hot.py listing (one line per SASS instruction of one launch) into its hot regions:
runs of consecutive instructions with the same execution count = one basic-block chain of the kernel's loops.
usage: sass_regions.py listing.txt [min_share_pct] > summary.txt"""
import re, sys
path = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
rows, head = [], []
for l in open(path):
    m = re.match(r'\s*(\d+)\s+(\d+)\s+([\d.]+)%\s+([\d.]+)M\s+(\d+)\s+lsb=\s*(\d+)\s+ssb=\s*(\d+)\s+(.*)', l)
    if m:
        rows.append((int(m.group(1)), int(m.group(2)), float(m.group(4)), int(m.group(5)), int(m.group(6)), int(m.group(7)), m.group(8)))
    elif len(head) < 2:
        head.append(l.rstrip())
# a report may list the function twice (two identical images): keep the first copy
n = len(rows)
if n % 2 == 0 and n > 0 and all(rows[i][2] == rows[i + n // 2][2] and rows[i][6] == rows[i + n // 2][6] for i in range(0, n // 2, max(1, n // 200))):
    rows = rows[: n // 2]
tot = sum(r[2] for r in rows) or 1.0
tsamp = sum(r[1] for r in rows) or 1
regions, cur = [], None
for r in rows:
    if cur and abs(cur["cnt"] - r[2]) <= 0.02 * max(cur["cnt"], 1) + 0.3 and r[0] == cur["end"] + 1:
        cur["end"] = r[0]; cur["n"] += 1; cur["sum"] += r[2]; cur["samp"] += r[1]; cur["lanes"] += r[3]
        cur["lsb"] += r[4]; cur["ssb"] += r[5]; cur["ops"].append(r[6].split()[0] if not r[6].startswith("@") else r[6].split()[1])
    else:
        cur = dict(start=r[0], end=r[0], cnt=r[2], n=1, sum=r[2], samp=r[1], lanes=r[3], lsb=r[4], ssb=r[5], first=r[6][:60],
                   ops=[r[6].split()[0] if not r[6].startswith("@") else r[6].split()[1]])
        regions.append(cur)
for h in head:
    print(h)
print(f"{len(rows)} SASS instructions, {tot / 1e3:.3f} G warp-instructions, {tsamp} stall samples; regions with >= {min_share} % of the instructions:")
print(f"{'sass range':>12s} {'n':>4s} {'executions':>11s} {'warp-inst':>10s} {'share':>6s} {'lanes':>5s} {'samples':>8s} {'long-sb':>7s} {'short-sb':>8s}  opcode mix / first instruction")
for g in sorted(regions, key=lambda g: -g["sum"]):
    if 100 * g["sum"] / tot < min_share:
        continue
    mix = {}
    for o in g["ops"]:
        o = o.split(".")[0]
        mix[o] = mix.get(o, 0) + 1
    top = " ".join(f"{k}x{v}" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:7])
    print(f"{g['start']:5d}-{g['end']:5d} {g['n']:4d} {g['cnt']:10.1f}M {g['sum']:9.1f}M {100 * g['sum'] / tot:5.1f}% {g['lanes'] / g['n']:5.1f} "
          f"{100 * g['samp'] / tsamp:7.1f}% {100 * g['lsb'] / tsamp:6.1f}% {100 * g['ssb'] / tsamp:7.1f}%  {top}  | {g['first']}")
