"""Print an `ncu --csv --metrics ...` log as one line per launch.  usage: python bench_tools/ncu_metrics_table.py log.csv [n]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]
d = OrderedDict()
for r in rows[1:]:
    rec = dict(zip(h, r))
    d.setdefault((rec["ID"], rec["Kernel Name"]), OrderedDict())[rec["Metric Name"]] = rec["Metric Value"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(d)
for (i, name), m in list(d.items())[:n]:
    short = name.split("::")[-1].split("(")[0][:26]
    print("%3s %-26s " % (i, short) + "  ".join("%s=%s" % (k.split("__")[-1].split(".")[0][:18], v) for k, v in m.items()))
