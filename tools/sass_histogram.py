"""SASS opcode histogram of the tensor-core / TMA kernels in the built library (no GPU needed).

    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt

For every kernel of libtt_b200.so: total instructions and the counts of the opcodes that prove the Blackwell path --
UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTCCP (tcgen05.cp), UTMALDG/UTMASTG/UTMAREDG (TMA load/store/reduce),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), MUFU.EX2, FFMA2/FADD2/FMNMX3 (packed math), RED/ATOM -- and the arch of the cubin.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "two_tower_recommender_model_b200", "lib", "libtt_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTCCP", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTCBAR",
         "UTCATOMSWS", "SYNCS", "MUFU.EX2", "MUFU", "FFMA2", "FADD2", "FMUL2", "FMNMX3", "HMMA", "RED", "ATOM", "LDG", "STG", "LDS", "STS",
         "BAR", "USETMAXREG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    arch = None
    fn = None
    stats = collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch = m.group(1)
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("tt::", "")
            fn = stats.setdefault(name, {"arch": arch, "n": 0, "ops": collections.Counter()})
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and fn is not None:
            op = m.group(1)
            fn["n"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                    fn["ops"][w] += 1
    print("# cuobjdump -sass two_tower_recommender_model_b200/lib/libtt_b200.so | opcode histogram per kernel (tools/sass_histogram.py)")
    tot = collections.Counter()
    for name, s in stats.items():
        ops = " ".join("%s=%d" % (k, v) for k, v in s["ops"].items() if k not in ("LDG", "STG", "LDS", "STS", "BAR", "MUFU") or v)
        print("%-8s %6d instr  %-58s %s" % (s["arch"], s["n"], name[:58], ops))
        tot.update(s["ops"])
    print("# total over %d kernels: %s" % (len(stats), " ".join("%s=%d" % kv for kv in tot.items())))


if __name__ == "__main__":
    sys.exit(main())
