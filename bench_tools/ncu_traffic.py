"""DRAM bytes per launch of the headline kernel from an `ncu --set full` report, stamped with the sha256 of the kernel
sources it was taken from (bench.py refuses the number when the sources have changed since).

usage: python bench_tools/ncu_traffic.py report.ncu-rep kernel-substring profiles/r2_resident_traffic.json"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep, want, out = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
launches = [r for r in rows[2:] if want in r[h.index("Kernel Name")]]
jr, jw, jt = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
rd = [float(r[jr]) * scale[units[jr]] for r in launches]
wr = [float(r[jw]) * scale[units[jw]] for r in launches]
rec = {"kernel": launches[0][h.index("Kernel Name")].split("(")[0], "launches": len(launches),
       "dram_bytes_read_per_launch": sum(rd) / len(rd), "dram_bytes_write_per_launch": sum(wr) / len(wr),
       "dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd),
       "gpu_time_us_per_launch": sum(float(r[jt]) for r in launches) / len(launches) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units[jt], 1.0),
       "kernel_source_sha256": bench.kernel_source_sha(), "report": os.path.basename(rep),
       "how": "ncu --set full --clock-control none (caches flushed per launch); 4096 networks of 21x21x36 float32 per launch"}
json.dump(rec, open(out, "w"), indent=1)
print(rec)
