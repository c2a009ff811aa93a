#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export (gpurun_out/<tag>_raw.csv) into tracked files:
     profiles/<round>_ncu_full.csv       key metrics per launch
     profiles/<round>_dram_traffic.json  per kernel: launches, DRAM bytes (read+write) and time per launch
   python scripts/ncu_raw_summary.py <tag> <round-name>"""
import collections, csv, json, os, re, sys

tag, rnd = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", f"{tag}_raw.csv"))))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max"]
idx = [hdr.index(w) for w in want if w in hdr]
P = os.path.join(ROOT, "profiles")
with open(os.path.join(P, f"{rnd}_ncu_full.csv"), "w") as fh:
    w = csv.writer(fh)
    w.writerow([f"{hdr[i]} [{units[i]}]" for i in idx])
    for r in rows[2:]:
        w.writerow([r[i][:90] for i in idx])


def num(r, name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    return v * scale.get(u, 1.0)


agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").replace("bdbnn::", "").strip()
    a = agg[name]
    a[0] += 1
    a[1] += num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
    a[2] += num(r, "gpu__time_duration.sum")
out = {k: {"launches": n, "dram_bytes_per_launch": b / n, "time_us_per_launch": t / n}
       for k, (n, b, t) in sorted(agg.items(), key=lambda kv: -kv[1][2])}
json.dump(out, open(os.path.join(P, f"{rnd}_dram_traffic.json"), "w"), indent=1)
for k, v in out.items():
    print(f"{k:36s} n={v['launches']:3d} dram={v['dram_bytes_per_launch'] / 1e6:8.1f} MB  t={v['time_us_per_launch']:7.1f} us  "
          f"-> {v['dram_bytes_per_launch'] / v['time_us_per_launch'] / 1e3:7.1f} GB/s")
