"""Cycles between the barriers of k_pc_active (clock64, CTA 0), per network: where the latency of one CTA's update goes.
Builds a profiling copy of the library with -DPRS_ACTIVE_TIMING into bench_tools/_timing/.
usage (on the GPU box):  python bench_tools/active_stage_timing.py [21x21x36 4096]"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "bench_tools", "_timing")


def build():
    os.makedirs(OUT, exist_ok=True)
    from pyratslam_b200 import build as b
    objs, procs = [], []
    for src in b.SOURCES:
        obj = os.path.join(OUT, src.replace(".cu", ".o"))
        procs.append(subprocess.Popen([b._nvcc()] + b.NVCC_FLAGS + ["-DPRS_ACTIVE_TIMING", "-c", os.path.join(b.CSRC, src), "-o", obj]))
        objs.append(obj)
    for pr in procs:
        assert pr.wait() == 0
    lib = os.path.join(OUT, "libpyratslam_b200.so")
    subprocess.check_call([b._nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
    return lib


if __name__ == "__main__":
    lib = build()
    from pyratslam_b200 import _native as nat
    nat.LIB_PATH = lib
    import numpy as np
    import torch
    from pyratslam_b200 import PoseCellEnsemble
    shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "21x21x36").split("x"))
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    rng = np.random.default_rng(5)
    od = torch.from_numpy(np.stack([rng.uniform(0, 0.3, (16, B)), rng.uniform(-0.1, 0.1, (16, B))], axis=-1)).cuda()
    ens = PoseCellEnsemble(shape, B, global_inhibition=np.linspace(0.05, 0.25, B), active_set=2)
    ens.inject(1.0, tuple(s // 2 for s in shape))
    for t in range(6):
        ens.update_async(od[t])
    buf = (ctypes.c_ulonglong * 32)()
    L = nat.lib()
    L.prs_debug_active_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.prs_debug_active_cycles(buf, 1)
    steps = 20
    for t in range(steps):
        ens.update_async(od[(6 + t) % 16])
    L.prs_debug_active_cycles(buf, 0)
    names = ["0 plan, masks zeroed, F copied", "1 occupancy of the axes", "2 dilated masks", "3 sets S, G", "4 spos, Xc zeroed",
             "5 scatter into Xc", "6 theta pass", "7 y pass", "8 x pass + inhibition", "9 total, masks of the result",
             "10 sets SA, DA, D_th", "11 tables", "12 masks B_x, B_y", "13 sets B_x, B_y", "14 zero old cells, 7x7 stage",
             "15 theta stage, state written, arg-max", "16 (dead networks only)", "17 finalise"]
    tot = sum(buf[i] for i in range(18))
    for i, n in enumerate(names):
        print("%-42s %8.0f clk/network  %5.1f%%" % (n, buf[i] / steps, 100.0 * buf[i] / tot))
    print("%-42s %8.0f clk/network (network 0 of %d, %s)" % ("total", tot / steps, B, "x".join(map(str, shape))))
