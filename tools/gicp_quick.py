"""DEV: one GICP align at bench size on the GPU with timings (python tools/gicp_quick.py [n_map])."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_slam_b200 import api, synth  # noqa: E402

n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
world = synth.make_world(synth.SEED, beams=True)
mp = synth.sample_map(n_map, synth.SEED, world=world)
p_true = np.array([3.0, -2.0, 1.2, 0.0, 0.0, 0.6])
T = synth.pose_vec_to_matrix(p_true)
scan = np.ascontiguousarray(synth.raycast(T[:3, 3], T[:3, :3], synth.livox_dirs(24000, synth.SEED), world, seed=synth.SEED)[:20000])
guess = synth.pose_vec_to_matrix(p_true + np.array([0.15, -0.1, 0.05, 0.01, -0.01, 0.03]))
g = api.GeneralizedIterativeClosestPoint()
t0 = time.perf_counter(); g.setInputTarget(mp); g.setInputSource(scan); g._handle(); t1 = time.perf_counter()
g.covariances("target"); t2 = time.perf_counter()
out = {"set_ms": (t1 - t0) * 1e3, "target_cov_ms": (t2 - t1) * 1e3, "index": g.index_info("target"), "index_src": g.index_info("source")}
for k in range(4):
    t0 = time.perf_counter(); rc = g.align(guess); w = (time.perf_counter() - t0) * 1e3
    r = g.result
    out[f"align{k}"] = dict(rc=rc, wall_ms=w, gpu_ms=r.gpu_ms, it=r.iterations, inner=r.inner_total, calls=[r.n_f, r.n_df, r.n_fdf], m=r.last_m, status=r.last_status)
fin = g.getFinalTransformation()
out["err"] = [float(np.abs(fin[:3, 3] - T[:3, 3]).max()), float(np.abs(fin[:3, :3] - T[:3, :3]).max())]
out["launches"] = api.kernel_launches()
print(json.dumps(out))
