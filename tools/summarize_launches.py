"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <command>`): per kernel and grid the
launch count, total and average duration and the share of the summed GPU time.   python tools/summarize_launches.py X.csv [title]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[start]
ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
d = collections.OrderedDict()
for r in rows[start + 2:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    key = (r[ki].split("(")[0].replace("void ", "").replace("lic360::", ""), r[gi], r[bi])
    d.setdefault(key, []).append(v)
tot = sum(sum(v) for v in d.values())
print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.  total GPU time %.2f ms in %d launches"
      % (tot / 1e6, sum(len(v) for v in d.values())))
print("%-34s %-16s %-14s %6s %10s %9s %7s" % ("kernel", "grid", "block", "count", "total ms", "avg us", "share"))
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print("%-34s %-16s %-14s %6d %10.3f %9.1f %6.1f%%" % (k[0][:34], k[1], k[2], len(v), sum(v) / 1e6, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
