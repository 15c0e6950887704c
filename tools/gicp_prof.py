"""DEV / profiling target: index build, covariances and two aligns of the bench-size GICP scene (20k-pt scan, 1M-pt map)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_slam_b200 import api, synth  # noqa: E402

world = synth.make_world(synth.SEED, beams=True)
mp = synth.sample_map(1_000_000, synth.SEED, world=world)
p_true = np.array([3.0, -2.0, 1.2, 0.0, 0.0, 0.6])
T = synth.pose_vec_to_matrix(p_true)
scan = np.ascontiguousarray(synth.raycast(T[:3, 3], T[:3, :3], synth.livox_dirs(24000, synth.SEED), world, seed=synth.SEED)[:20000])
guess = synth.pose_vec_to_matrix(p_true + np.array([0.15, -0.1, 0.05, 0.01, -0.01, 0.03]))
g = api.GeneralizedIterativeClosestPoint()
g.setInputTarget(mp)
g.setInputSource(scan)
for _ in range(2):
    rc = g.align(guess)
print("rc", rc, "passes", g.result.iterations, "gpu_ms", g.result.gpu_ms, "launches", api.kernel_launches())
