#!/usr/bin/env python
"""Summarise gpurun_out artefacts into profiles/ (tracked):  python scripts/ncu_summary.py <tag> <round-name>
  <tag>_launches.csv   (ncu gpu__time_duration per launch)  -> profiles/<round>_launches_by_kernel.csv
  <tag>_prof.ncu-rep   (ncu --set full)                     -> profiles/<round>_ncu_full.csv (key metrics/launch)
  <tag>_bench.json                                           -> profiles/<round>_bench.json
"""
import collections, csv, os, re, shutil, subprocess, sys

tag, rnd = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

lp = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(lp):
    lines = [l for l in open(lp) if not l.startswith("==")]
    agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        ns = v * 1e3 if row["Metric Unit"] == "us" else (v * 1e6 if row["Metric Unit"] == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).strip()[:120]
        agg[name][0] += 1; agg[name][1] += ns; tot += ns
    with open(os.path.join(P, f"{rnd}_launches_by_kernel.csv"), "w") as fh:
        fh.write("# one timed train step (bench.py --profile-mode under ncu --profile-from-start off); "
                 "cold-cache serialised launch times: compare SHARES\n")
        fh.write("kernel,launches,total_us,share\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"\"{k}\",{n},{t / 1e3:.1f},{t / tot:.4f}\n")
    print("launch list:", len(agg), "kernels, total", round(tot / 1e6, 3), "ms")

rp = os.path.join(G, f"{tag}_prof.ncu-rep")
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(os.path.join(P, f"{rnd}_ncu_full.csv"), "w") as fh:
        w = csv.writer(fh)
        w.writerow([f"{hdr[i]} [{units[i]}]" for i in idx])
        for r in rows[2:]:
            w.writerow([r[i][:100] for i in idx])
    print("ncu full:", len(rows) - 2, "launches")

for suffix in ("bench.json", "kernels.json"):
    bp = os.path.join(G, f"{tag}_{suffix}")
    if os.path.exists(bp):
        shutil.copy(bp, os.path.join(P, f"{rnd}_{suffix}"))
