"""Profiling target without torch (fast start): two launches of the map builder's batched accumulation, 24 keyframes x 100k points per
launch at bench density - the first creates the voxels, the second touches them again.  Device buffers come from cudart through ctypes."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_slam_b200 import api, synth  # noqa: E402

rt = C.CDLL("libcudart.so")
_cache = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_frames_cache.npz")
CACHE = np.load(_cache) if os.path.exists(_cache) else None
world = None if CACHE is not None else synth.make_world(synth.SEED, beams=True)
K, POOL = 24, 4
frames, poses = [], []
for k in range(K):
    a = 2 * np.pi * k / 32
    pos = np.array([30.0 * np.cos(a), 15.0 * np.sin(a), 1.2])
    q = synth.quat_from_rotvec([0.0, 0.0, a + np.pi / 2])
    if k < POOL:
        if CACHE is not None:   # ray casting costs a second per keyframe: a cache made on the CPU box rides along with the snapshot
            frames.append(np.ascontiguousarray(CACHE[f"arr_{k}"]))
        else:
            pts = synth.raycast(pos, synth.quat_to_R(q), synth.avia_dirs(115_000, seed=900 + k), world, seed=950 + k)[:100_000]
            frames.append(np.ascontiguousarray(np.concatenate([pts, np.ones((len(pts), 1), np.float32)], 1)))
    poses.append(np.array([pos[0], pos[1], pos[2], q[3], q[0], q[1], q[2]]))
dev = []
for f in frames:
    p = C.c_void_p()
    assert rt.cudaMalloc(C.byref(p), C.c_size_t(f.nbytes)) == 0
    assert rt.cudaMemcpy(p, f.ctypes.data_as(C.c_void_p), C.c_size_t(f.nbytes), 1) == 0
    dev.append(p.value)
b = api.FullMapBuilder(leaf=0.1, capacity_voxels=2_000_000)
for rep in range(2):
    b.add_keyframes_device([dev[k % POOL] for k in range(K)], [len(frames[k % POOL]) for k in range(K)], np.stack(poses))
print("fullmap voxels", b.num_voxels(), "points per launch", sum(len(frames[k % POOL]) for k in range(K)))
