"""Circular-mode sweep of the bit-sliced library with and without lock-step warps (prs_vt_tune knob 4)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyratslam_b200 import _native as nat  # noqa: E402

L = nat.lib()
g = torch.Generator(device="cuda").manual_seed(4)
key = torch.zeros(1, dtype=torch.int64, device="cuda")
scratch = torch.zeros(4096, dtype=torch.uint8, device="cuda")
n = (1 << 20) - 3
lib = torch.randint(0, 256, (n, 32, 32), dtype=torch.uint8, device="cuda", generator=g)
qs = torch.randint(0, 256, (4, 32, 32), dtype=torch.uint8, device="cuda", generator=g)
packed = torch.zeros(int(L.prs_vt_packed_bytes(n)), dtype=torch.uint8, device="cuda")
nat.check(L.prs_vt_pack_u8(lib.data_ptr(), n, packed.data_ptr(), 0, nat.stream_ptr()))
del lib


def sweep(t, sc=None):
    nat.check(L.prs_vt_sweep_packed_u8(packed.data_ptr(), n, qs[t % 4].data_ptr(), 1, 0, key.data_ptr(),
                                       sc.data_ptr() if sc is not None else None, scratch.data_ptr(), nat.stream_ptr()))


want = None
for lock in (0, 1, 0, 1):
    nat.check(L.prs_vt_tune(4, lock))
    sc = torch.zeros(n, dtype=torch.int32, device="cuda")
    sweep(0, sc)
    torch.cuda.synchronize()
    if want is None:
        want = sc.clone()
    ok = bool(torch.equal(sc, want))
    for t in range(3):
        sweep(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(10):
        sweep(t)
    e1.record()
    torch.cuda.synchronize()
    print("circular lock=%d: %.4f ms/query  %s" % (lock, e0.elapsed_time(e1) / 10, "OK" if ok else "MISMATCH"), flush=True)
nat.check(L.prs_vt_tune(4, 1))
