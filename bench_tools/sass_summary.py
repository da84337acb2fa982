"""Per-kernel counts of the SASS mnemonics that show what the Blackwell kernels are made of (profiles/r2_sass_summary.txt).

usage: python bench_tools/sass_summary.py [libpyratslam_b200.so] > profiles/r2_sass_summary.txt
Columns: static instruction count, FFMA2 / FFMA / DFMA (packed, scalar fp32 and fp64 FMA), LOP3 / POPC (bit-sliced
compare), UBLKCP (cp.async.bulk: the bulk-copy engine), UTMALDG (cp.async.bulk.tensor: tensor-map TMA), LDGSTS (cp.async),
SYNCS (mbarrier), ELECT, REDUX, UCGABAR / MEMBAR (cluster barriers), UTCxMMA / LDTM (tcgen05: none expected -- the path
has no GEMM-shaped stage).
"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pyratslam_b200", "_lib", "libpyratslam_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["FFMA2", "FFMA", "DFMA", "LOP3", "POPC", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "ELECT", "REDUX", "UCGABAR", "UTCMMA", "LDTM"]
rows, name, ops, arch = [], None, Counter(), set()


def demangle(n):
    try:
        d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception:
        d = n
    d = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", d)
    d = re.sub(r"^void ", "", d)
    return d.split("(")[0]


def flush():
    if name and ops:
        rows.append((demangle(name), sum(ops.values()), [sum(v for k, v in ops.items() if k.startswith(c) and (c != "FFMA" or k == "FFMA")) for c in cols]))


for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, ops = m.group(1), Counter()
        continue
    m = re.search(r"EF_CUDA_SM(\d+)", line) or re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        ops[m.group(1)] += 1
flush()
print("# %s: %d kernels, arch %s" % (os.path.basename(lib), len(rows), ",".join(sorted(arch)) or "sm_100a"))
print("%-58s %7s " % ("kernel", "instr") + " ".join("%7s" % c for c in cols))
tot = [0] * len(cols)
for n, k, v in sorted(rows, key=lambda r: -r[1]):
    print("%-58s %7d " % (n[:58], k) + " ".join("%7d" % x for x in v))
    tot = [a + b for a, b in zip(tot, v)]
print("%-58s %7d " % ("total", sum(r[1] for r in rows)) + " ".join("%7d" % x for x in tot))
