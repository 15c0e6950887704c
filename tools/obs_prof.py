"""Dev helper for ncu: config-1 map, a few IEKF updates (plain launches, no graph)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pointcloud_slam_b200 import api, synth
params = sys.argv[1] if len(sys.argv) > 1 else "livox"
res, nearby, ext = (0.2, 26, False) if params == "livox" else (0.5, 18, True)
c = synth.config1()
g = api.IVox(resolution=res, nearby=nearby)
g.AddPoints(c["map"])
kf = api.Esekf(g, extrinsic_est_en=ext)
api.lib().b200_iekf_set_graph(kf.h, 0)
for r in range(4):
    kf.change_x(c["x_prop"]); kf.change_P(c["P"])
    kf.update_iterated_dyn_share_modified(c["scan"])
    print(kf.stats.gpu_ms)
