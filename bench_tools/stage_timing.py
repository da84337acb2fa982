"""Per-stage cycles of the resident kernel measured inside the kernel (clock64 between barriers, CTA 0).

Builds a profiling copy of the library with -DPRS_RESIDENT_TIMING into bench_tools/_timing/ and runs the
ensemble workload on it.  usage (on the GPU box):  python bench_tools/stage_timing.py
"""
import ctypes
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "bench_tools", "_timing")


def build():
    os.makedirs(OUT, exist_ok=True)
    from pyratslam_b200 import build as b
    objs = []
    for src in b.SOURCES:
        obj = os.path.join(OUT, src.replace(".cu", ".o"))
        subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + ["-DPRS_RESIDENT_TIMING", "-c", os.path.join(b.CSRC, src), "-o", obj])
        objs.append(obj)
    lib = os.path.join(OUT, "libpyratslam_b200.so")
    subprocess.check_call([b._nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
    return lib


if __name__ == "__main__":
    lib = build() if "--no-build" not in sys.argv else os.path.join(OUT, "libpyratslam_b200.so")
    from pyratslam_b200 import _native as nat
    nat.LIB_PATH = lib
    import numpy as np
    import torch
    import bench
    from pyratslam_b200 import PoseCellEnsemble
    B = 4096
    gis, odom = bench.ensemble_inputs(B, 64, 3)
    ens = PoseCellEnsemble(bench.SHAPE, B, global_inhibition=gis)
    ens.inject(1.0, (10, 10, 18))
    od = torch.from_numpy(odom).cuda()
    for t in range(5):
        ens.update_async(od[t])
    buf = (ctypes.c_ulonglong * 8)()
    L = nat.lib()
    L.prs_debug_stage_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.prs_debug_stage_cycles(buf, 1)
    steps = 20
    for t in range(steps):
        ens.update_async(od[t % 64])
    L.prs_debug_stage_cycles(buf, 0)
    nets = steps * ((B + 147) // 148)
    names = ["1 theta+scatter", "2 y pass", "3 x pass+sum(a)", "3 store A2+sum(b)", "4 7x7", "5 theta+argmax", "loop tail/plan"]
    tot = sum(buf[i] for i in range(7))
    for i, n in enumerate(names):
        print("%-20s %8.0f clk/network  %5.1f%%" % (n, buf[i] / nets, 100.0 * buf[i] / tot))
    print("%-20s %8.0f clk/network" % ("total", tot / nets))
