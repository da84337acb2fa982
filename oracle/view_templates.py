"""CPU oracle for the view-template matcher.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates ``ratslam/view_templates.py``.  The arithmetic that matters:

* ``ViewTemplate.match`` (``view_templates.py:16-28``) compares stored rows
  ``8+o .. 23+o`` with query rows ``8 .. 23`` for the 15 offsets ``o = -7..7``
  and keeps the smallest sum.  With uint8 images (the ROS path,
  ``ros_simulate.py:100-101``) ``a - b`` wraps modulo 256 and ``abs`` is the
  identity, so the score is ``sum((a-b) mod 256)`` accumulated by numpy in
  uint64 -- an integer, reproduced exactly here.  With float images it is an
  ordinary sum of absolute differences.
* ``ViewTemplates.match`` (``view_templates.py:63-75``) creates a new template
  when the library is empty or the best score is strictly above the threshold,
  else returns the **first** best template (``numpy.argmin``).
* ``ViewTemplates.__init__`` (``view_templates.py:42-57``) uses Python-2 integer
  division for the template shape and for the row index ``base/im_x``.

``circular_scores`` is an extension with no reference counterpart (BASELINE
config 5's "full circular-shift search"): all 32 cyclic row rotations of the
whole template.
"""
from __future__ import annotations

import numpy as np


def match_score(stored, new, max_offset=8):
    """``ViewTemplate.match``: min over 15 windowed row offsets.  ``view_templates.py:16-28``."""
    best = np.inf
    for off in range(-max_offset + 1, max_offset):
        a = stored[max_offset + off: -max_offset + off, :]
        b = new[max_offset:-max_offset, :]
        diff = np.sum(np.abs(a - b))  # uint8: wraps, abs is a no-op, uint64 accumulate
        if diff < best:
            best = diff
    return best


def match_scores_all_offsets(stored, new, max_offset=8):
    """The 15 per-offset scores (for tests that look below the min)."""
    out = []
    for off in range(-max_offset + 1, max_offset):
        a = stored[max_offset + off: -max_offset + off, :]
        b = new[max_offset:-max_offset, :]
        out.append(np.sum(np.abs(a - b)))
    return np.array(out)


def circular_score(stored, new):
    """Extension: min over all cyclic row shifts ``s`` of ``sum |roll(stored, -s) - new|``."""
    best = np.inf
    n = stored.shape[0]
    for s in range(n):
        diff = np.sum(np.abs(np.roll(stored, -s, axis=0) - new))
        if diff < best:
            best = diff
    return best


def library_scores(library, query, mode="ref"):
    """Vectorised scores of ``query`` against ``library[n, R, C]`` (same arithmetic as above).

    Used where the Python loop over templates would take minutes; checked
    against ``match_score`` in the CPU tests.
    """
    lib = np.asarray(library)
    q = np.asarray(query)
    n, R, _ = lib.shape
    is_int = np.issubdtype(lib.dtype, np.integer)
    acc = np.uint64 if is_int else lib.dtype
    if mode == "ref":
        mo = 8
        best = None
        for off in range(-mo + 1, mo):
            d = np.abs(lib[:, mo + off: R - mo + off, :] - q[None, mo: R - mo, :])
            s = d.reshape(n, -1).sum(axis=1, dtype=acc)
            best = s if best is None else np.minimum(best, s)
        return best
    elif mode == "circular":
        best = None
        for sft in range(R):
            d = np.abs(np.roll(lib, -sft, axis=1) - q[None])
            s = d.reshape(n, -1).sum(axis=1, dtype=acc)
            best = s if best is None else np.minimum(best, s)
        return best
    raise ValueError(mode)


class ViewTemplate:
    """``view_templates.py:4-37``."""

    def __init__(self, pc_x, pc_y, pc_th, index, template):
        self.pc_x, self.pc_y, self.pc_th = pc_x, pc_y, pc_th
        self.template = template
        self.index = index
        self.max_offset = 8

    def match(self, new_template):
        return match_score(self.template, new_template, self.max_offset)

    def location(self):
        return (self.pc_x, self.pc_y, self.pc_th)

    def get_index(self):
        return self.index


class ViewTemplates:
    """``view_templates.py:40-75``."""

    def __init__(self, x_range, y_range, x_step, y_step, im_x, im_y, match_threshold):
        self.templates = []
        self.shape = ((x_range[1] - x_range[0]) // x_step, (y_range[1] - y_range[0]) // y_step)
        self.match_threshold = match_threshold
        base = np.arange(im_x * im_y)
        row, col = base // im_x, base % im_x  # py2 '/' on ints floors
        m = ((row > y_range[0]) & (row < y_range[1]) & (col > x_range[0]) & (col < x_range[1])
             & ((row - y_range[0]) % y_step != 0) & ((col - x_range[0]) % x_step != 0))
        self.mask = m.reshape((im_x, im_y))

    def match(self, input, pc_x, pc_y, pc_th):
        template = input[self.mask].reshape(self.shape)
        vals = [t.match(template) for t in self.templates]
        if len(vals) == 0 or min(vals) > self.match_threshold:
            t = ViewTemplate(pc_x, pc_y, pc_th, len(self.templates), template)
            self.templates.append(t)
            return t
        return self.templates[int(np.argmin(vals))]
