"""Summarise `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` output: stall mix + hottest SASS lines."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        if r and r[0] == "Kernel Name":
            break
        continue
    data.append(r)
g = lambda r, k: int(r[idx[k]] or 0)
tot = sum(g(r, "# Samples") for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(g(r, s) for r in data) for s in stalls}
print("samples", tot, {k: f"{100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > 0})
for r in sorted(data, key=lambda r: -g(r, "# Samples"))[:n]:
    st = {s[6:]: g(r, s) for s in stalls if g(r, s) > 0}
    print(f"{g(r, '# Samples'):6d} {g(r, 'Instructions Executed'):10d}  {r[idx['Source']].strip()[:64]:64s} {dict(sorted(st.items(), key=lambda x: -x[1])[:3])}")
