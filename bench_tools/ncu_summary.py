"""Transpose `ncu -i rep --page raw --csv` into one row per metric, one column per launch (the format of
profiles/*_ncu_full.csv).  usage: python bench_tools/ncu_summary.py rep.ncu-rep kernel-substring out.csv [metric ...]"""
import csv
import subprocess
import sys

rep, want, out = sys.argv[1], sys.argv[2], sys.argv[3]
metrics = sys.argv[4:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
launches = [r for r in rows[2:] if want in r[h.index("Kernel Name")]]
if not metrics:
    metrics = [m for m in h if "__" in m]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(launches))])
    for m in metrics:
        if m in h:
            j = h.index(m)
            w.writerow([m, units[j]] + [r[j] for r in launches])
print("%d launches, %d metrics -> %s" % (len(launches), len(metrics), out))
