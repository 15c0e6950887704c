"""Short program for the ncu pass over the kernels added after the first profiles: map builder, VoxelGrid, undistortion, LOAM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pointcloud_slam_b200 import api, synth

world = synth.make_world(synth.SEED, beams=True)
frames, poses = [], []
for k in range(6):
    a = 2 * np.pi * k / 6
    pos = np.array([30.0 * np.cos(a), 15.0 * np.sin(a), 1.2])
    q = synth.quat_from_rotvec([0.0, 0.0, a + np.pi / 2])
    pts = synth.raycast(pos, synth.quat_to_R(q), synth.avia_dirs(115_000, seed=900 + k), world, seed=950 + k)[:100_000]
    frames.append(np.ascontiguousarray(np.concatenate([pts, np.ones((len(pts), 1), np.float32)], 1)))
    poses.append(np.array([pos[0], pos[1], pos[2], q[3], q[0], q[1], q[2]]))
d = [torch.from_numpy(f).cuda() for f in frames]
b = api.FullMapBuilder(leaf=0.1, capacity_voxels=4_000_000)
for rep in range(3):
    for k in range(6):
        b.add_keyframe_device(d[k].data_ptr(), len(frames[k]), poses[k])
print("fullmap voxels", b.num_voxels())
cfg = synth.config1(n_map=400_000, n_scan=20_000)
vg = api.VoxelGrid(); vg.setLeafSize(0.2); vg.setInputCloud(cfg["scan"])
for _ in range(2):
    c, n = vg.filter()
print("voxelgrid", len(c), vg.last_ms())
pts = np.zeros((20000, 12), np.float32); pts[:, :3] = cfg["scan"]; pts[:, 9] = np.linspace(0, 99.9, 20000)
p22 = np.zeros((21, 22)); p22[:, 0] = np.arange(21) * 0.005; p22[:, 13:22] = np.eye(3).reshape(9); p22[:, 4:7] = [0.1, -0.2, 0.3]; p22[:, 7:10] = [1, 0, 0]
for _ in range(2):
    vg.undistort(pts, 9, 8, p22, cfg["x_true"], want_host=False)
print("undistort", vg.last_ms())
sc = synth.loam_scene(n_surf_map=400_000, surf_stride=4, corner_stride=2)
g = api.ScanToMap(max_map_points=1_000_000)
g.setInputCloud(sc["corner_map"], sc["surf_map"])
guess = sc["t_true"] + np.array([0.01, -0.01, 0.02, 0.15, -0.1, 0.05], np.float32)
for _ in range(2):
    t, rc = g.scan2MapOptimization(sc["corner"], sc["surf"], guess)
print("loam", rc, g.stats.iters, g.stats.gpu_ms)
