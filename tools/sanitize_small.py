"""A small pass over every kernel of the library for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pointcloud_slam_b200 import api, synth

cfg = synth.config1(n_map=40_000, n_scan=1_500)
g = api.IVox(resolution=0.5, nearby=18, capacity=3000, max_points=200_000)   # small capacity: the LRU path runs
for k in range(4):
    g.AddPoints(cfg["map"][k * 10_000:(k + 1) * 10_000])
o_l, Rl = synth.lidar_pose(cfg["x_true"])
q = (cfg["scan"].astype(np.float64) @ Rl.T + o_l).astype(np.float32)
g.GetClosestPoint(q)
g2 = api.IVox(resolution=0.2, nearby=26, max_points=200_000)
g2.AddPoints(cfg["map"])
kf = api.Esekf(g2)
for k in range(3):
    kf.change_x(cfg["x_prop"]); kf.change_P(cfg["P"])
    kf.update_iterated_dyn_share_modified(cfg["scan"])
    kf.MapIncremental(kf.get_x(), True)
print("iekf", kf.stats.passes, list(kf.stats.n_eff)[:4], "evicted", g.evicted())
n = api.NormalDistributionsTransform()
n.setTransformationEpsilon(0.01)
n.setInputTarget(cfg["map"]); n.setInputSource(q)
p6 = np.array([0.05, -0.03, 0.02, 0.001, -0.002, 0.004])
n.computeDerivatives(p6); n.computeHessian(p6)
n.align(synth.pose_vec_to_matrix(p6).astype(np.float32))
poses = synth.hypothesis_grid(np.zeros(6), 3, 3, 2, 1.0)
print("reloc", api.relocalize(n, poses)[:2], "fitness", n.getFitnessScore(), "batch", n.alignBatch(poses[:3])[1][0].iters)
vg = api.VoxelGrid(); vg.setLeafSize(0.5); vg.setInputCloud(cfg["scan"]); c, cnt = vg.filter()
pts = np.zeros((1500, 12), np.float32); pts[:, :3] = cfg["scan"]; pts[:, 9] = np.linspace(0, 99, 1500)
poses22 = np.zeros((5, 22)); poses22[:, 0] = np.arange(5) * 0.025; poses22[:, 13:22] = np.eye(3).reshape(9); poses22[:, 4:7] = 0.1
vg.undistort(pts, 9, 8, poses22, cfg["x_true"]); vg.filter_staged()
b = api.FullMapBuilder(leaf=0.2, capacity_voxels=200_000)
b.add_keyframe(np.concatenate([cfg["scan"], np.ones((1500, 1), np.float32)], 1), [0, 0, 0, 1, 0, 0, 0])
print("voxels", len(c), b.num_voxels(), len(b.extract()[0]))
