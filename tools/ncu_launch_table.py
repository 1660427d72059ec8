"""One line per launch from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log."""
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
cur = {}
for r in csv.DictReader(lines):
    e = cur.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"])[:52], "grid": r["Grid Size"]})
    e[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
for k, e in cur.items():
    t, tu = e.get("gpu__time_duration.sum", (0, "ns"))
    us = t / 1000 if tu in ("ns", "nsecond") else (t * 1000 if tu in ("ms", "msecond") else t)

    def b(m):
        v, u = e.get(m, (0, "byte"))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    print("%5s %9.1f us rd %8.1f MB wr %8.1f MB grid %-18s %s" % (k, us, b("dram__bytes_read.sum") / 1e6, b("dram__bytes_write.sum") / 1e6, e["grid"], e["name"]))
