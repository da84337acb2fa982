"""Static opcode mix of each kernel in an object file (cuobjdump -sass): a quick issue-slot budget without a GPU.

usage: python bench_tools/sass_mix.py file.o [name-substring]
Counts are per static instruction (unrolled straight-line code ~ executed counts); branches are not followed.
"""
import re
import subprocess
import sys
from collections import Counter

out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
want = sys.argv[2] if len(sys.argv) > 2 else ""
name, ops = None, Counter()


def flush():
    if name and want in name and ops:
        n = sum(ops.values())
        print("%s  (%d instructions)" % (name[:100], n))
        print("    " + ", ".join("%s:%d" % kv for kv in ops.most_common(16)))


for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, ops = m.group(1), Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        ops[m.group(1)] += 1
flush()
