"""Dev helper: A/B of the k-NN candidate-walk variants (B200_KNN_MODE=0|1|5|6 in the environment, one process per
variant).  Prints the search time at 20k and 1M queries (cold L2), a digest of the results (must be equal across
variants: the neighbour sets are bit-exact by contract) and the IEKF update time with its per-kernel split."""
import ctypes as C
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pointcloud_slam_b200 import api, synth  # noqa: E402

params = sys.argv[1] if len(sys.argv) > 1 else "livox"
res, nearby, ext = (0.2, 26, False) if params == "livox" else (0.5, 18, True)
c = synth.config1()
g = api.IVox(resolution=res, nearby=nearby)
g.AddPoints(c["map"])
ol, Rl = synth.lidar_pose(c["x_prop"])
qw = (c["scan"].astype(np.float64) @ Rl.T + ol).astype(np.float32)
rng = np.random.default_rng(1)
qbig = np.ascontiguousarray(np.concatenate([qw + rng.normal(0, 0.05, qw.shape).astype(np.float32) for _ in range(50)], 0))
out = {}
for name, q in (("20k", qw), ("1M", qbig)):
    g.GetClosestPoint(q)
    ms = []
    for _ in range(5):
        api.flush_l2(0)
        i, d, n = g.GetClosestPoint(q)
        ms.append(g.last_knn_ms())
    h = hashlib.md5(i.tobytes() + d.tobytes() + n.tobytes()).hexdigest()[:12]
    out[name] = (float(np.mean(ms)), float(np.min(ms)), h)
kf = api.Esekf(g, extrinsic_est_en=ext)
up = []
for r in range(8):
    api.flush_l2(0)
    kf.change_x(c["x_prop"]); kf.change_P(c["P"])
    kf.update_iterated_dyn_share_modified(c["scan"])
    up.append(kf.stats.gpu_ms)
xh = hashlib.md5(kf.get_x().tobytes()).hexdigest()[:12]
warm = []
for r in range(8):
    kf.change_x(c["x_prop"]); kf.change_P(c["P"])
    kf.update_iterated_dyn_share_modified(c["scan"])
    warm.append(kf.stats.gpu_ms)
api.lib().b200_iekf_set_profiling(kf.h, 1)
kt = None
for r in range(3):
    api.flush_l2(0)
    kf.change_x(c["x_prop"]); kf.change_P(c["P"])
    kf.update_iterated_dyn_share_modified(c["scan"])
    ms = (C.c_float * 17)()
    k = api.lib().b200_iekf_kernel_times(kf.h, ms, 17)
    kt = ["%.1f" % (ms[i] * 1e3) for i in range(k)]
print("MODE", os.environ.get("B200_KNN_MODE", "auto"), params, "| knn 20k: %.4f ms (min %.4f) %s | 1M: %.4f ms (min %.4f) %s" % (out["20k"] + out["1M"]),
      "| update cold %.4f warm %.4f ms x=%s | kernels us %s" % (float(np.median(up[3:])), float(np.median(warm[3:])), xh, kt))
