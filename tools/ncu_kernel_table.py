"""Aggregate an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log per kernel."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    k = r["Kernel Name"][:60]
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    m = r["Metric Name"]
    e = agg.setdefault(k, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
    if m.startswith("gpu__time"):
        e["n"] += 1
        e["us"] += v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
    else:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        e["rd" if "read" in m else "wr"] += v * scale
for k, e in agg.items():
    n = max(e["n"], 1)
    us = e["us"] / n
    tb = (e["rd"] + e["wr"]) / n
    print(f"{us:9.1f} us/launch x{e['n']:3d}  dram rd {e['rd'] / n / 1e6:8.2f} MB wr {e['wr'] / n / 1e6:8.2f} MB  {tb / us / 1e3 if us else 0:7.1f} GB/s dram  {k}")
