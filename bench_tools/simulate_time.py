"""Wall time of the simulate.py scenario (50x50x10 grid, 40 updates, construction included), repeated."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyratslam_b200 import simulate  # noqa: E402

simulate.main(steps=40, verbose=False)
ts = []
for i in range(20):
    t0 = time.perf_counter()
    simulate.main(steps=40, verbose=False)
    ts.append(time.perf_counter() - t0)
ts.sort()
print("simulate.py scenario: best %.3f ms, median %.3f ms per 40-step run (%.1f us per step)"
      % (ts[0] * 1e3, ts[10] * 1e3, ts[10] / 40 * 1e6))
