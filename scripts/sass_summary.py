#!/usr/bin/env python
"""Per-kernel SASS evidence of the Blackwell-specific instructions -> profiles/sass_summary_<tag>.txt
UTCHMMA/UTCQMMA = tcgen05.mma (f16/bf16 kind, fp8 kind), UTCBAR = tcgen05.commit, LDTM = tcgen05.ld,
UTMALDG = TMA tensor load, SYNCS = mbarrier ops, POPC = the bit-serial XNOR path, RED/ATOM = global atomics.
usage: python scripts/sass_summary.py [tag]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bdbnn_b200", "libbdbnn_b200.so")
COLS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "SYNCS", "POPC", "RED", "ATOM"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per[name] = collections.Counter()
            continue
        if name:
            m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                op = m.group(1)
                for c in COLS:
                    if op.startswith(c):
                        per[name][c] += 1
    names = list(per)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines() \
        if names else []
    out = [f"# cuobjdump -sass bdbnn_b200/libbdbnn_b200.so — instruction counts per kernel (static SASS)",
           "# " + " ".join(f"{c:>7}" for c in COLS) + "  kernel"]
    tot = collections.Counter()
    for n, d in zip(names, dem):
        c = per[n]
        tot.update(c)
        if sum(c[k] for k in COLS if k not in ("SYNCS", "RED", "ATOM")) == 0:
            continue
        short = re.sub(r"\(.*", "", d)[:90]
        out.append("  " + " ".join(f"{c[k]:7d}" for k in COLS) + "  " + short)
    out.append("# totals: " + ", ".join(f"{k}={tot[k]}" for k in COLS))
    elf = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout.split()
    out.append("# cubins: " + " ".join(e for e in elf if e.endswith(".cubin")))
    path = os.path.join(ROOT, "profiles", f"sass_summary_{tag}.txt")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
