"""DRAM traffic per launch out of an `ncu --set full` report -> profiles/r02_dram_traffic.json (read by bench.py).

    ncu -i gpurun_out/<rep>.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_traffic.py /tmp/raw.csv <report name> [--merge]

Every kernel of the report gets an entry keyed by its (short) name: launches seen, mean duration, mean
dram__bytes_read.sum + dram__bytes_write.sum per launch.  Composite keys that bench.py asks for are the sum of
their member kernels' per-launch means (one launch of each per step):
    softmax_fwd_bwd_B65536_d64 = tc_softmax_fwd_kernel + tc_softmax_bwd_fused_kernel (+ finalize, rowdot if captured)
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}

COMPOSITES = {
    "softmax_fwd_bwd_B65536_d64": ["tc_softmax_fwd_kernel", "tc_softmax_bwd_fused_kernel", "tc_softmax_bwd2_kernel",
                                   "softmax_bwd_finalize_kernel", "rowdot_bf16_kernel"],
}


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"^(tt::)?(tc::)?", "", name)
    return re.split(r"[<(]", name)[0]


def main():
    path, rep = sys.argv[1], sys.argv[2]
    merge = "--merge" in sys.argv
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    need = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"]
    for n in need:
        if n not in col:
            raise SystemExit("column missing from the raw page: " + n)

    def val(r, name, table):
        return float(r[col[name]].replace(",", "")) * table.get(units[col[name]], 1.0)

    agg = {}
    for r in body:
        k = short(r[col["Kernel Name"]])
        e = agg.setdefault(k, {"launches": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "grid": r[col["Grid Size"]] if "Grid Size" in col else None})
        e["launches"] += 1
        e["us"] += val(r, "gpu__time_duration.sum", TIME)
        e["rd"] += val(r, "dram__bytes_read.sum", UNIT)
        e["wr"] += val(r, "dram__bytes_write.sum", UNIT)
    out = {}
    for k, e in agg.items():
        n = e["launches"]
        out[k] = {"dram_bytes": round((e["rd"] + e["wr"]) / n), "dram_read": round(e["rd"] / n), "dram_write": round(e["wr"] / n),
                  "us": round(e["us"] / n, 2), "launches": n, "grid": e["grid"], "report": rep}
    for key, members in COMPOSITES.items():
        have = [m for m in members if m in out]
        if have:
            out[key] = {"dram_bytes": sum(out[m]["dram_bytes"] for m in have), "members": have, "report": rep,
                        "us": round(sum(out[m]["us"] for m in have), 2)}
    dst = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
    if merge and os.path.exists(dst):
        old = json.load(open(dst))
        old.update(out)
        out = old
    with open(dst, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for k, e in sorted(out.items()):
        print("%-40s %10.2f MB  %9.2f us" % (k, e["dram_bytes"] / 1e6, e.get("us", 0.0)))


if __name__ == "__main__":
    main()
