"""Developer tool: per-kernel count / mean duration from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi or not r[vi].replace(",", "").replace(".", "").isdigit():
        continue
    agg.setdefault(r[ki].split("(")[0][-48:], []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k:50s} n={len(v):4d}  mean {sum(v) / len(v) / 1000:8.2f} us  share {100 * sum(v) / tot:5.1f} %")
