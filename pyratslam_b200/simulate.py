"""Offline driver: the loop of ``ratslam/simulate.py`` on the CUDA pose-cell network.

``RatSLAM`` / ``main`` keep the reference's names and data (``simulate.py:13-40,56-58``): a
(50, 50, 10) grid, unit energy injected at the centre, 40 odometry steps with vtrans = 3 and a
pi/4 turn on steps 4..8.  The matplotlib scatter of ``simulate.py:47-67`` is viewer-side and
optional here (``plot=True`` needs matplotlib, which the build image does not have).
"""
from __future__ import annotations

import math

import numpy as np

from .posecell_network import PoseCellNetwork

POSE_SIZE = (50, 50, 10)


class RatSLAM(object):
    def __init__(self, data=None, shape=POSE_SIZE, **pcn_kwargs):
        self.cur_step = 0
        self._pcn_kwargs = pcn_kwargs
        self.pcn = self.init_pcn(shape)
        self.data = np.zeros((20, 2)) if data is None else data

    def init_pcn(self, shape):
        pcn = PoseCellNetwork(shape, **self._pcn_kwargs)
        midpoint = (math.floor(shape[0] / 2), math.floor(shape[1] / 2), math.floor(shape[2] / 2))
        pcn.inject(1, midpoint)
        self.current_pose_cell = midpoint
        return pcn

    def step(self):
        self.current_pose_cell = self.pcn.update(self.data[self.cur_step, :])
        self.cur_step += 1


def default_data(steps=40):
    data = np.zeros((steps, 2))
    data[:, 0] = 3
    data[4:9, 1] = math.pi / 4
    return data


def main(steps=40, plot=False, verbose=True, **pcn_kwargs):
    data = default_data(40)
    sim = RatSLAM(data=data, shape=POSE_SIZE, **pcn_kwargs)
    ax = None
    if plot:
        import matplotlib.pyplot as plt  # viewer side; not needed for the computation
        fig = plt.figure()
        ax = fig.add_subplot(111, projection="3d")
        plt.ion()
        plt.show()
    trace = []
    for _ in range(steps):
        sim.step()
        trace.append(sim.current_pose_cell)
        if ax is not None:
            pc = sim.pcn.posecells
            idx = np.nonzero(pc > .002)
            ax.clear()
            ax.scatter(idx[0], idx[1], idx[2], s=pc[idx] * 100)
            ax.set_xlim3d([0, POSE_SIZE[0]])
            ax.set_ylim3d([0, POSE_SIZE[1]])
            ax.set_zlim3d([0, POSE_SIZE[2]])
            plt.pause(.01)
    if verbose:
        print(trace)
    return trace


if __name__ == "__main__":
    main()
