"""The two product sweeps with their default knobs, a few launches each (the target of the ncu captures):
2^20 bit-sliced uint8 templates and 2^18 float32 templates, reference mode.  usage: python bench_tools/vt_profile.py [launches]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyratslam_b200 import _native as nat  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 5
L = nat.lib()
g = torch.Generator(device="cuda").manual_seed(4)
key = torch.zeros(1, dtype=torch.int64, device="cuda")
scratch = torch.zeros(4096, dtype=torch.uint8, device="cuda")
n = 1 << 20
lib = torch.randint(0, 256, (n, 32, 32), dtype=torch.uint8, device="cuda", generator=g)
q = torch.randint(0, 256, (32, 32), dtype=torch.uint8, device="cuda", generator=g)
packed = torch.zeros(int(L.prs_vt_packed_bytes(n)), dtype=torch.uint8, device="cuda")
nat.check(L.prs_vt_pack_u8(lib.data_ptr(), n, packed.data_ptr(), 0, nat.stream_ptr()))
del lib
for _ in range(k):
    nat.check(L.prs_vt_sweep_packed_u8(packed.data_ptr(), n, q.data_ptr(), 0, 0, key.data_ptr(), None, scratch.data_ptr(),
                                       nat.stream_ptr()))
torch.cuda.synchronize()
del packed
nf = 1 << 18
libf = torch.rand((nf, 32, 32), dtype=torch.float32, device="cuda", generator=g) * 255
qf = torch.rand((32, 32), dtype=torch.float32, device="cuda", generator=g) * 255
for _ in range(k):
    nat.check(L.prs_vt_sweep_f32(libf.data_ptr(), nf, qf.data_ptr(), 0, 0, key.data_ptr(), None, nat.stream_ptr()))
torch.cuda.synchronize()
print("ok", hex(int(key.item()) & (2**64 - 1)))
