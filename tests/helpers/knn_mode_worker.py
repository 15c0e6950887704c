"""Worker of tests/test_gpu_knn_modes.py: one process per candidate-walk variant (B200_KNN_MODE is read once per process).
Prints a digest of the k = 5 search results and of one IEKF posterior on a sparse and on a dense map, plus the oracle check."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointcloud_slam_b200 import api, synth  # noqa: E402
from oracle import binding as ob  # noqa: E402

cfg = synth.config1(n_map=60_000, n_scan=2_500)
ol, Rl = synth.lidar_pose(cfg["x_true"])
q = (cfg["scan"].astype(np.float64) @ Rl.T + ol).astype(np.float32)
out = {}
# sparse: 0.2 m voxels, NEARBY26; dense: 2.5 m voxels (a hundred candidates per query: beyond the flat list capacity), NEARBY18
for name, res, nearby in (("sparse", 0.2, 26), ("dense", 2.5, 18)):
    g = api.IVox(resolution=res, nearby=nearby)
    g.AddPoints(cfg["map"])
    i1, d1, c1 = g.GetClosestPoint(q)
    o = ob.OracleLio(resolution=res, nearby=nearby)
    o.insert(cfg["map"])
    i0, d0, c0 = o.knn5(q)
    kf = api.Esekf(g)
    kf.change_x(cfg["x_prop"]); kf.change_P(cfg["P"])
    rc = kf.update_iterated_dyn_share_modified(cfg["scan"])
    out[name] = dict(knn=hashlib.md5(i1.tobytes() + d1.tobytes() + c1.tobytes()).hexdigest(),
                     oracle_equal=bool(np.array_equal(i0, i1) and np.array_equal(d0, d1) and np.array_equal(c0, c1)),
                     x=hashlib.md5(kf.get_x().tobytes()).hexdigest(), rc=int(rc), mean_candidates=float(g.stencil_points(q)[0]) / len(q))
out["tma_timeouts"] = api.knn_tma_timeouts()
print(json.dumps(out))
