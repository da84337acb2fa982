"""Determinism / race check of the ring sweeps: many back-to-back sweeps with changing queries, every result compared
with the register-prefetch kernels (knob 0 / knob 2 = 0).  usage: python bench_tools/vt_stress.py [rounds]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyratslam_b200 import _native as nat  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = nat.lib()
g = torch.Generator(device="cuda").manual_seed(9)
key = torch.zeros(1, dtype=torch.int64, device="cuda")
scratch = torch.zeros(4096, dtype=torch.uint8, device="cuda")
DEF = (34, 5, 13, 2)
bad = 0
# ---- bit-sliced uint8
n = (1 << 19) + 77
lib = torch.randint(0, 256, (n, 32, 32), dtype=torch.uint8, device="cuda", generator=g)
qs = torch.randint(0, 256, (6, 32, 32), dtype=torch.uint8, device="cuda", generator=g)
packed = torch.zeros(int(L.prs_vt_packed_bytes(n)), dtype=torch.uint8, device="cuda")
nat.check(L.prs_vt_pack_u8(lib.data_ptr(), n, packed.data_ptr(), 0, nat.stream_ptr()))
del lib


def sweep_u8(q, sc):
    nat.check(L.prs_vt_sweep_packed_u8(packed.data_ptr(), n, q.data_ptr(), 0, 0, key.data_ptr(), sc.data_ptr(),
                                       scratch.data_ptr(), nat.stream_ptr()))


nat.check(L.prs_vt_tune(0, 0))
want = []
for q in qs:
    sc = torch.zeros(n, dtype=torch.int32, device="cuda")
    sweep_u8(q, sc)
    want.append(sc)
for depth, ctas in ((34, 5), (34, 6), (44, 1), (4, 5), (8, 3), (2, 5)):
    nat.check(L.prs_vt_tune(0, depth))
    nat.check(L.prs_vt_tune(1, ctas))
    for r in range(rounds):
        outs = [torch.zeros(n, dtype=torch.int32, device="cuda") for _ in qs]
        for q, sc in zip(qs, outs):
            sweep_u8(q, sc)
        torch.cuda.synchronize()
        b = sum(int((a != w).sum().item()) for a, w in zip(outs, want))
        bad += b
        if b:
            print("u8 depth=%d ctas=%d round %d: %d wrong scores" % (depth, ctas, r, b), flush=True)
del packed, want
# ---- float32
nf = (1 << 17) + 5
libf = torch.rand((nf, 32, 32), dtype=torch.float32, device="cuda", generator=g) * 255
qf = torch.rand((6, 32, 32), dtype=torch.float32, device="cuda", generator=g) * 255


def sweep_f(q, sc):
    nat.check(L.prs_vt_sweep_f32(libf.data_ptr(), nf, q.data_ptr(), 0, 0, key.data_ptr(), sc.data_ptr(), nat.stream_ptr()))


nat.check(L.prs_vt_tune(2, 0))
wantf = []
for q in qf:
    sc = torch.zeros(nf, dtype=torch.float32, device="cuda")
    sweep_f(q, sc)
    wantf.append(sc)
for depth, ctas in ((13, 2), (12, 3), (11, 6), (2, 3), (1, 8)):
    nat.check(L.prs_vt_tune(2, depth))
    nat.check(L.prs_vt_tune(3, ctas))
    first = None
    for r in range(rounds):
        outs = [torch.zeros(nf, dtype=torch.float32, device="cuda") for _ in qf]
        for q, sc in zip(qf, outs):
            sweep_f(q, sc)
        torch.cuda.synchronize()
        if depth < 10:
            b = sum(int((a != w).sum().item()) for a, w in zip(outs, wantf))
        else:  # another summation order: close to the register kernel, and identical from run to run
            b = sum(int(((a - w).abs() > 1e-5 * w.abs()).sum().item()) for a, w in zip(outs, wantf))
            if first is None:
                first = outs
            else:
                b += sum(int((a != f).sum().item()) for a, f in zip(outs, first))
        bad += b
        if b:
            print("f32 depth=%d ctas=%d round %d: %d wrong scores" % (depth, ctas, r, b), flush=True)
for knob, v in enumerate(DEF):
    L.prs_vt_tune(knob, v)
print("vt_stress: %d wrong scores" % bad)
sys.exit(1 if bad else 0)
