"""Seeded synthetic inputs shared by the fixture generator and the tests (no reference needed)."""
from __future__ import annotations

import numpy as np


def synth_frames(rng, T):
    """Frames with revisits the uint8 wrap rule accepts (darker copy, even vertical shift)."""
    frames = np.zeros((T, 256, 256), dtype=np.uint8)
    for t in range(T):
        if t % 4 == 3 and t > 3:
            src = frames[rng.integers(0, t)]
            dark = rng.integers(0, 4, src.shape, dtype=np.int16)
            f = np.clip(src.astype(np.int16) - dark, 0, 255).astype(np.uint8)
            frames[t] = np.roll(f, 2 * int(rng.integers(-7, 8)), axis=0)
        else:
            frames[t] = rng.integers(0, 256, (256, 256), dtype=np.uint8)
    return frames
