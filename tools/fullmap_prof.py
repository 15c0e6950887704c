"""Dev helper for ncu: the map builder's batched accumulation at bench density (32 keyframes, replayed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pointcloud_slam_b200 import api, synth
world = synth.make_world(synth.SEED, beams=True)
frames, poses = [], []
K = 24
for k in range(K):
    a = 2 * np.pi * k / 32
    pos = np.array([30.0 * np.cos(a), 15.0 * np.sin(a), 1.2])
    q = synth.quat_from_rotvec([0.0, 0.0, a + np.pi / 2])
    pts = synth.raycast(pos, synth.quat_to_R(q), synth.avia_dirs(115_000, seed=900 + k), world, seed=950 + k)[:100_000]
    frames.append(np.ascontiguousarray(np.concatenate([pts, np.ones((len(pts), 1), np.float32)], 1)))
    poses.append(np.array([pos[0], pos[1], pos[2], q[3], q[0], q[1], q[2]]))
d = [torch.from_numpy(f).cuda() for f in frames]
b = api.FullMapBuilder(leaf=0.1, capacity_voxels=4_000_000)
for rep in range(4):
    b.add_keyframes_device([t.data_ptr() for t in d], [len(f) for f in frames], np.stack(poses))
print("fullmap voxels", b.num_voxels())
