import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from synth import synth_frames
from pyratslam_b200 import ros_simulate
T = 80
frames = synth_frames(np.random.default_rng(1), T)
rng = np.random.default_rng(1)
odom = np.stack([rng.uniform(0, 3, T), rng.uniform(-1, 1, T)], axis=1)
node = ros_simulate.RatslamRos()
node.replay_native(frames, odom, n_plans=2)
print("ok")
