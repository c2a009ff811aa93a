#!/usr/bin/env python
"""Trim an `ncu --page source --csv` dump to the lines that matter:  python scripts/ncu_source_hot.py <in.csv> <out.csv> "<title>"
Keeps SASS lines with >= 0.3 % of the stall samples plus every TMA / mbarrier / tcgen05 instruction."""
import csv, sys
src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[si]) for r in data if r[si].isdigit())
keys = ("UTMALDG", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "SYNCS.PHASECHK", "SYNCS.ARRIVE", "BAR.SYNC", "UTCATOMSWS")
with open(dst, "w") as fh:
    fh.write(f"# {title}; total stall samples {tot}\n")
    fh.write("line,sass,stall_samples,pct,instructions_executed\n")
    for i, r in enumerate(data):
        n = int(r[si]) if r[si].isdigit() else 0
        if n >= 0.003 * tot or any(k in r[1] for k in keys):
            fh.write(f'{i},"{r[1].strip()}",{n},{100.0 * n / max(tot, 1):.2f},{r[ie]}\n')
print("wrote", dst)
