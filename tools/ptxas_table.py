"""Per-kernel resource usage of the shipped library as ptxas reports it (registers, stack, spills, static shared memory):
compiles every csrc/*.cu with the build's own flags plus ``-Xptxas -v`` into a scratch directory and prints one row per
kernel.  No GPU needed.

    python tools/ptxas_table.py > profiles/r02_ptxas_resource_usage.txt
"""
import os
import re
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from two_tower_recommender_model_b200 import build  # noqa: E402


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        return out if len(out) == len(names) else names
    except OSError:
        return names


def short(sig: str) -> str:
    """`void ns::kernel<args>(params)` -> `kernel<args>`."""
    sig = re.sub(r"^void ", "", sig)
    depth, cut = 0, len(sig)
    for i, ch in enumerate(sig):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            cut = i
            break
    return re.sub(r"(\w+::)+", "", sig[:cut])


def field(pattern: str, blob: str) -> int:
    m = re.search(pattern, blob)
    return int(m.group(1)) if m else 0


def main() -> None:
    nvcc = build._nvcc()
    with tempfile.TemporaryDirectory() as tmp:
        def one(src):
            r = subprocess.run([nvcc] + build.NVCC_FLAGS + ["-Xptxas", "-v", "-c", src, "-o", os.path.join(tmp, os.path.basename(src) + ".o")],
                               capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(r.stderr)
            return os.path.basename(src), r.stderr
        with ThreadPoolExecutor(max_workers=8) as ex:
            logs = list(ex.map(one, build.sources()))
    rows = []
    for src, log in logs:
        lines = log.splitlines()
        for i, ln in enumerate(lines):
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", ln)
            if not m:
                continue
            blob = " ".join(lines[i + 1:i + 4])
            rows.append((src, m.group(1), field(r"Used (\d+) registers", blob), field(r"(\d+) bytes stack frame", blob),
                         field(r"(\d+) bytes spill stores", blob), field(r"(\d+) bytes spill loads", blob),
                         field(r"(\d+) bytes smem", blob), field(r"used (\d+) barriers", blob)))
    names = [short(n) for n in demangle([r[1] for r in rows])]
    print("# ptxas -v, sm_100a, flags of two_tower_recommender_model_b200/build.py; dynamic shared memory is not listed here")
    print("# (the tcgen05 kernels take theirs at launch); stack / spills: bytes per thread")
    print(f"{'file':<20} {'kernel':<84} {'regs':>4} {'stack':>5} {'spill st':>8} {'spill ld':>8} {'smem':>6} {'bar':>3}")
    for (src, _m, regs, stack, st, ld, smem, bar), name in sorted(zip(rows, names), key=lambda x: (x[0][0], x[1])):
        print(f"{src:<20} {name[:84]:<84} {regs:>4} {stack:>5} {st:>8} {ld:>8} {smem:>6} {bar:>3}")
    spilled = sum(1 for r in rows if r[4] or r[5])
    print(f"\n# {len(rows)} kernels, {spilled} with spills")


if __name__ == "__main__":
    main()
