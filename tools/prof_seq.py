"""Dev: the k-NN search in the sliding-map regime of configs[2] (dense voxels after many MapIncremental calls)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pointcloud_slam_b200 import api, synth
import bench

n_scans = int(os.environ.get("SEQ_SCANS", 300))
world = synth.make_world(synth.SEED)
per_lap = 290
def true_state(k):
    a = 2 * np.pi * k / per_lap
    return synth.make_state(np.array([30.0 * np.cos(a), 15.0 * np.sin(a), 1.2]), [0.0, 0.0, np.arctan2(15.0 * np.cos(a), -30.0 * np.sin(a))])
def scan_of(k):
    o, Rl = synth.lidar_pose(true_state(k))
    return np.ascontiguousarray(synth.raycast(o, Rl, synth.livox_dirs(25_000, seed=synth.SEED + k), world, seed=synth.SEED + 7 * k)[:20000])
ivox = api.IVox(resolution=0.5, nearby=18)
kf = api.Esekf(ivox, filter_size_map=0.5)
P0 = synth.init_cov() * 0.01
x = true_state(0)
ol, Rl = synth.lidar_pose(x)
ivox.AddPoints((scan_of(0).astype(np.float64) @ Rl.T + ol).astype(np.float32))
for k in range(1, n_scans):
    scan = scan_of(k)
    kf.change_x(synth.perturb_state(true_state(k), seed=k, dpos=0.02, drot_deg=0.2)); kf.change_P(P0)
    kf.update_iterated_dyn_share_modified(scan)
    kf.MapIncremental(kf.get_x(), True)
print("map voxels", ivox.NumValidGrids(), "points", ivox.NumPoints(), "last update ms", kf.stats.gpu_ms, flush=True)
scan = scan_of(n_scans)
xt = true_state(n_scans)
ol, Rl = synth.lidar_pose(xt)
qw = (scan.astype(np.float64) @ Rl.T + ol).astype(np.float32)
pts, cells = ivox.stencil_points(qw)
print(f"candidates/query {pts/len(qw):.1f} occupied cells/query {cells/len(qw):.2f}")
idx, d2, cnt = ivox.GetClosestPoint(qw)
print("standalone knn ms", ivox.last_knn_ms())
kf.set_profiling(True)
for r in range(3):
    kf.change_x(synth.perturb_state(xt, seed=7, dpos=0.02, drot_deg=0.2)); kf.change_P(P0)
    kf.update_iterated_dyn_share_modified(scan)
    print("profiled", f"{kf.stats.gpu_ms:.3f} ms", "kernels(us)", [f"{1e3*t:.1f}" for t in kf.kernel_times_ms()], "knn", list(kf.stats.knn)[:4])
kf.set_profiling(False)
for r in range(3):
    kf.change_x(synth.perturb_state(xt, seed=7, dpos=0.02, drot_deg=0.2)); kf.change_P(P0)
    kf.update_iterated_dyn_share_modified(scan)
    print("graph", f"{kf.stats.gpu_ms:.3f} ms")
