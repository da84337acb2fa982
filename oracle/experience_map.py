"""CPU oracle for the experience map.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates ``ratslam/experience_map.py``: dead-reckoned pose accumulation and one
appended experience per update (``experience_map.py:52-60``); no links, no
relaxation (the reference's TODOs at ``:49,59``).  ``LinkedExperienceMap`` below is the
SPECIFICATION of the links / loop-closure extension, which the reference lacks.
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


# ----------------------------------------------------------------------------------------------------------
# Links and loop-closure relaxation -- SURVEY 8(f) row 3.  The reference stops at two TODOs
# (``experience_map.py:49`` "linking will go here", ``:59`` "leaving vt stuff out"), so there is nothing to restate:
# THIS FILE IS THE SPECIFICATION the product (pyratslam_b200/experience_map.py, ``linked=True``) is tested against.
# It follows the published RatSLAM experience-map algorithm (Milford & Wyeth; openRatSLAM ``experience_map.cpp``:
# on_create_experience / on_create_link / on_set_experience / iterate), written as plain loops over Python objects.
#
# * Odometry accumulates RELATIVE to the current experience: rel_th = clip(rel_th + vrot), rel_x += vtrans cos(rel_th),
#   rel_y += vtrans sin(rel_th) (the reference's absolute accumulators keep running next to them, unchanged).
# * ``update(vtrans, vrot, pc_loc, vt)``: among the experiences created with the same view template ``vt`` (not None)
#   take those whose pose cell lies within ``delta_pc`` of ``pc_loc`` (Euclidean over the three wrapped cell deltas,
#   wrap lengths ``pc_dims``); the nearest (ties: lowest id) that is not the current experience closes a loop: a link
#   current -> match is recorded (unless that pair is already linked) and the match becomes the current experience.
#   With no candidate -- or no current experience yet -- a new experience is created at the current experience's map
#   pose plus the relative motion and linked from the current one.  A match with the current experience itself does
#   nothing.  After a create or a loop closure the relative accumulators restart (rel_x = rel_y = 0, rel_th = the
#   current experience's heading).
# * A link stores d = hypot(rel_x, rel_y), heading_rad = signed_delta(from.th, atan2(rel_y, rel_x)) and
#   facing_rad = signed_delta(from.th, rel_th).
# * ``iterate(loops)``: Gauss-Seidel relaxation in experience order, link order = creation order, gain ``correction``
#   (0.5): both ends of a link move half of the position error towards each other; headings likewise.


def signed_delta_rad(a, b):
    """Smallest signed rotation that takes heading ``a`` to heading ``b`` (openRatSLAM get_signed_delta_rad)."""
    d = clip_rad_180(b) - clip_rad_180(a)
    if d > math.pi:
        d -= 2 * math.pi
    elif d <= -math.pi:
        d += 2 * math.pi
    return d


class LinkedExperience(Experience):
    def __init__(self, pc_loc, em_loc, th, vt, index):
        Experience.__init__(self, pc_loc, em_loc, vt)
        self.th = th
        self.index = index
        self.links_from = []    # indices into ExperienceMap.links of the links that START here
        self.links_to = []      # ... that END here


class Link:
    def __init__(self, exp_from, exp_to, d, heading_rad, facing_rad):
        self.exp_from, self.exp_to = exp_from, exp_to
        self.d, self.heading_rad, self.facing_rad = d, heading_rad, facing_rad


class LinkedExperienceMap(ExperienceMap):
    def __init__(self, pc_dims=(21, 21, 36), delta_pc=1.0, correction=0.5):
        ExperienceMap.__init__(self)
        self.pc_dims = tuple(pc_dims)
        self.delta_pc = delta_pc
        self.correction = correction
        self.links = []
        self.rel_x = self.rel_y = self.rel_th = 0.0
        self.n_loop_closures = 0

    def _pc_delta(self, e, pc_loc):
        s = 0.0
        for a, b, n in zip((e.pc_x, e.pc_y, e.pc_th), pc_loc, self.pc_dims):
            d = abs(float(a) - float(b)) % n
            d = min(d, n - d)
            s += d * d
        return math.sqrt(s)

    def _link(self, a, b):
        for li in a.links_from:
            if self.links[li].exp_to == b.index:
                return False
        d = math.hypot(self.rel_x, self.rel_y)
        heading = signed_delta_rad(a.th, math.atan2(self.rel_y, self.rel_x))
        facing = signed_delta_rad(a.th, self.rel_th)
        self.links.append(Link(a.index, b.index, d, heading, facing))
        a.links_from.append(len(self.links) - 1)
        b.links_to.append(len(self.links) - 1)
        return True

    def create(self, pc_loc, vt=None):
        cur = self.current_exp
        if cur is None:
            e = LinkedExperience(pc_loc, (self.rel_x, self.rel_y), clip_rad_180(self.rel_th), vt, 0)
        else:
            e = LinkedExperience(pc_loc, (cur.m_x + self.rel_x, cur.m_y + self.rel_y), clip_rad_180(self.rel_th), vt,
                                 len(self.experiences))
        self.experiences.append(e)
        if cur is not None:
            self._link(cur, e)
        self._set_current(e)

    def _set_current(self, e):
        self.current_exp = e
        self.rel_x = self.rel_y = 0.0
        self.rel_th = e.th

    def update(self, vtrans, vrot, pc_loc, vt=None):
        self.accum_delta_th = clip_rad_180(self.accum_delta_th + vrot)          # experience_map.py:55-57, unchanged
        self.accum_delta_x += vtrans * math.cos(self.accum_delta_th)
        self.accum_delta_y += vtrans * math.sin(self.accum_delta_th)
        self.rel_th = clip_rad_180(self.rel_th + vrot)
        self.rel_x += vtrans * math.cos(self.rel_th)
        self.rel_y += vtrans * math.sin(self.rel_th)
        best, best_d = None, None
        if vt is not None and self.current_exp is not None:
            for e in self.experiences:
                if e.vt is None or e.vt != vt:
                    continue
                d = self._pc_delta(e, pc_loc)
                if d <= self.delta_pc and (best is None or d < best_d):
                    best, best_d = e, d
        if best is None:
            self.create(pc_loc, vt)
        elif best is not self.current_exp:
            self._link(self.current_exp, best)
            self.n_loop_closures += 1
            self._set_current(best)

    def iterate(self, loops=1):
        c = self.correction
        for _ in range(loops):
            for a in self.experiences:
                for li in a.links_from:
                    ln = self.links[li]
                    b = self.experiences[ln.exp_to]
                    lx = a.m_x + ln.d * math.cos(a.th + ln.heading_rad)
                    ly = a.m_y + ln.d * math.sin(a.th + ln.heading_rad)
                    ex, ey = (b.m_x - lx) * c, (b.m_y - ly) * c
                    a.m_x += ex
                    a.m_y += ey
                    b.m_x -= ex
                    b.m_y -= ey
                    df = signed_delta_rad(a.th + ln.facing_rad, b.th)
                    a.th = clip_rad_180(a.th + df * c)
                    b.th = clip_rad_180(b.th - df * c)

    def get_poses(self):
        return [(e.m_x, e.m_y, e.th) for e in self.experiences]
