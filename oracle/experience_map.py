"""CPU oracle for the experience map.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates ``ratslam/experience_map.py``: dead-reckoned pose accumulation and one
appended experience per update (``experience_map.py:52-60``); no links, no
relaxation (the reference's TODOs at ``:49,59``).
"""
from __future__ import annotations

import math


def clip_rad_180(angle):
    """``experience_map.py:6-11``."""
    if angle > math.pi:
        angle -= math.ceil(angle / (2 * math.pi)) * 2 * math.pi
    elif angle <= -math.pi:
        angle += math.ceil(abs(angle) / (2 * math.pi)) * 2 * math.pi
    return angle


class Experience:
    def __init__(self, pc_loc, em_loc, vt):
        self.pc_x, self.pc_y, self.pc_th = pc_loc
        self.vt = vt
        self.m_x, self.m_y = em_loc

    def get_point(self):
        return (self.m_x, self.m_y)


class ExperienceMap:
    def __init__(self):
        self.accum_delta_x = 0
        self.accum_delta_y = 0
        self.accum_delta_th = 0
        self.experiences = []
        self.current_exp = None

    def create(self, pc_loc, vt=None):
        e = Experience(pc_loc, (self.accum_delta_x, self.accum_delta_y), vt)
        self.experiences.append(e)
        self.current_exp = e

    def update(self, vtrans, vrot, pc_loc, vt=None):
        self.accum_delta_th = clip_rad_180(self.accum_delta_th + vrot)
        self.accum_delta_x += vtrans * math.cos(self.accum_delta_th)
        self.accum_delta_y += vtrans * math.sin(self.accum_delta_th)
        self.create(pc_loc)

    def get_points(self):
        return [e.get_point() for e in self.experiences]

    def get_current_point(self):
        return self.current_exp.get_point()
