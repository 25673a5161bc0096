"""Developer tool: top stall sites of an ncu source page export (ncu -i x.ncu-rep --page source --csv > x_src.csv)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hdr = rows[1]
data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


tot = sum(f(r, '# Samples') for r in data)
print("total samples", tot, "n instr", len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
top = sorted(range(len(data)), key=lambda i: -f(data[i], '# Samples'))[:ntop]
for i in sorted(top):
    r = data[i]
    st = sorted(((f(r, s), s) for s in stalls), reverse=True)[:3]
    print(f"{i:5d} {r[ix['Address']][-5:]} {r[ix['Source']][:64]:64s} smp={f(r, '# Samples'):7.0f} ({100 * f(r, '# Samples') / tot:4.1f}%) "
          f"ex={f(r, 'Instructions Executed'):9.0f} " + " ".join(f"{s[6:]}={v:.0f}" for v, s in st if v > 0))
print({s[6:]: int(sum(f(r, s) for r in data)) for s in stalls})
