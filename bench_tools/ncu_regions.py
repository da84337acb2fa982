"""Split an `ncu --page source --csv --print-source sass` dump at barrier instructions and report, per region, the
executed warp instructions, the share of stall samples and the opcode mix.
usage: ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv ; python bench_tools/ncu_regions.py src.csv"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
i = idx[0]
hdr = rows[i]
col = {h: j for j, h in enumerate(hdr)}
end = idx[1] if len(idx) > 1 else len(rows)
body = [r for r in rows[i + 1:end] if len(r) >= len(hdr)]
regions, cur = [], [0, 0, Counter()]
for r in body:
    ex = int(r[col["Instructions Executed"]] or 0)
    smp = int(r[col["# Samples"]] or 0)
    s = r[col["Source"]].split()
    op = (s[1] if s[0].startswith("@") else s[0]) if s else "?"
    cur[0] += ex
    cur[1] += smp
    cur[2][op.split(".")[0]] += ex
    if op.startswith("BAR") or "CGABAR_WAIT" in op:
        regions.append(cur)
        cur = [0, 0, Counter()]
regions.append(cur)
tot = sum(r[0] for r in regions) or 1
ts = sum(r[1] for r in regions) or 1
for k, (ex, smp, c) in enumerate(regions):
    if ex:
        print("%2d exec=%7d (%4.1f%%) samples=%4.1f%%  %s" % (k, ex, 100 * ex / tot, 100 * smp / ts, dict(c.most_common(7))))
