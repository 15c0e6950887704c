"""Dev helper for ncu: config-1 map, then the standalone k=5 search at 20k and 1M queries (one warm-up each)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pointcloud_slam_b200 import api, synth
params = sys.argv[1] if len(sys.argv) > 1 else "livox"
res, nearby = (0.2, 26) if params == "livox" else (0.5, 18)
c = synth.config1()
g = api.IVox(resolution=res, nearby=nearby)
g.AddPoints(c["map"])
ol, Rl = synth.lidar_pose(c["x_prop"])
qw = (c["scan"].astype(np.float64) @ Rl.T + ol).astype(np.float32)
rng = np.random.default_rng(1)
qbig = np.ascontiguousarray(np.concatenate([qw + rng.normal(0, 0.05, qw.shape).astype(np.float32) for _ in range(50)], 0))
for q in (qw, qbig, qw, qbig):
    api.flush_l2(0)
    g.GetClosestPoint(q)
    print(len(q), g.last_knn_ms())
