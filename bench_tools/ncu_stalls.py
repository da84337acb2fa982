"""Summarise an `ncu --page source --csv` dump: stall reasons overall and per code region.

usage: ncu -i prof.ncu-rep --page source --csv > src.csv ; python bench_tools/ncu_stalls.py src.csv [n_regions]
Regions are delimited by BAR.SYNC instructions (the stage boundaries of the resident kernel).
"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = Counter()
regions = []
cur = {"first": None, "n": 0, "samples": 0, "stalls": Counter(), "ops": Counter(), "wave": 0, "wave_ideal": 0, "exec": 0}
for r in rows[hdr_i + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break  # next launch in the same report
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]]
    op = src.split()[0] if src.split() else "?"
    if op.startswith("@"):
        op = src.split()[1]
    try:
        ns = int(r[col["# Samples"]] or 0)
    except ValueError:
        ns = 0
    cur["n"] += 1
    cur["samples"] += ns
    cur["ops"][op.split(".")[0]] += 1
    cur["exec"] += int(r[col["Instructions Executed"]] or 0)
    try:
        cur["wave"] += int(r[col["L1 Wavefronts Shared"]] or 0)
        cur["wave_ideal"] += int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
    except ValueError:
        pass
    for s in stalls:
        try:
            v = int(r[col[s]] or 0)
        except ValueError:
            v = 0
        cur["stalls"][s] += v
        tot[s] += v
    if cur["first"] is None:
        cur["first"] = r[col["Address"]]
    if op.startswith("BAR"):
        regions.append(cur)
        cur = {"first": None, "n": 0, "samples": 0, "stalls": Counter(), "ops": Counter(), "wave": 0, "wave_ideal": 0, "exec": 0}
regions.append(cur)
allS = sum(tot.values())
print("total samples", allS)
for s, v in tot.most_common(8):
    print("  %-22s %6.1f%%" % (s, 100.0 * v / allS))
print()
for i, g in enumerate(regions):
    if g["samples"] == 0:
        continue
    top = ", ".join("%s %.0f%%" % (s.replace("stall_", ""), 100.0 * v / max(1, sum(g["stalls"].values())))
                    for s, v in g["stalls"].most_common(4))
    ops = ", ".join("%s:%d" % kv for kv in g["ops"].most_common(5))
    print("region %2d  instr=%5d exec=%9d samples=%6d (%4.1f%%)  smem wavefronts=%9d (ideal %9d)  [%s]  {%s}"
          % (i, g["n"], g["exec"], g["samples"], 100.0 * g["samples"] / allS, g["wave"], g["wave_ideal"], top, ops))
