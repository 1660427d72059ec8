"""Aggregate `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` by CUDA source line:
samples and executed instructions per (file, line).  usage: ncu_lines.py file.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.Counter()
inst = collections.Counter()
src = {}
cur_file = ""
hdr = None
cur_line = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        samp_i = r.index("# Samples")
        inst_i = r.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= samp_i:
        continue
    if r[0].strip():          # a CUDA source line row
        cur_line = (cur_file, int(r[0]))
        src[cur_line] = r[1].strip()[:90]
        continue
    if cur_line is None:      # SASS row under the current CUDA line
        continue
    try:
        agg[cur_line] += int(r[samp_i] or 0)
        inst[cur_line] += int(r[inst_i] or 0)
    except ValueError:
        pass
tot = sum(agg.values())
print("total samples", tot)
for (f, l), n in agg.most_common(top):
    print(f"{n:8d} {100.0 * n / max(tot, 1):5.1f}%  inst {inst[(f, l)]:12d}  {f}:{l}  {src[(f, l)]}")
