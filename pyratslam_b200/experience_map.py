"""Experience map -- host bookkeeping, as in the reference (``ratslam/experience_map.py``).

O(1) per step and not part of the GPU path (BASELINE north_star: "ExperienceMap ... stay on the
host as non-hot code").  It dead-reckons a 2-D pose from the odometry and records one experience
per update, tagged with the arg-max pose cell; the reference has no links or relaxation yet
(``experience_map.py:49,59``), so neither does this.
"""
from __future__ import annotations

import math


def clip_rad_180(angle):
    """Wrap an angle into (-pi, pi] the way ``experience_map.py:6-11`` does."""
    two_pi = 2 * math.pi
    if angle > math.pi:
        return angle - math.ceil(angle / two_pi) * two_pi
    if angle <= -math.pi:
        return angle + math.ceil(abs(angle) / two_pi) * two_pi
    return angle


class Experience:
    def __init__(self, pc_loc, em_loc, vt):
        self.pc_x, self.pc_y, self.pc_th = pc_loc[0], pc_loc[1], pc_loc[2]
        self.vt = vt
        self.m_x, self.m_y = em_loc[0], em_loc[1]

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
        self.current_exp = Experience(pc_loc, (self.accum_delta_x, self.accum_delta_y), vt)
        self.experiences.append(self.current_exp)

    def update(self, vtrans, vrot, pc_loc, vt=None):
        th = clip_rad_180(self.accum_delta_th + vrot)
        self.accum_delta_th = th
        self.accum_delta_x += vtrans * math.cos(th)
        self.accum_delta_y += vtrans * math.sin(th)
        self.create(pc_loc)

    def get_points(self):
        return [e.get_point() for e in self.experiences]

    def get_current_point(self):
        return self.current_exp.get_point()
