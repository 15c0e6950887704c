"""GPU voxel-grid paths against the oracle: scan downsample (bit-exact) and the keyframe-merge map builder."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("leaf", [0.2, 0.5])
def test_scan_downsample_bit_exact(oracle, api, synth, small_cfg, leaf):
    rng = np.random.default_rng(3)
    scan = np.concatenate([small_cfg["scan"], rng.uniform(0, 255, (len(small_cfg["scan"]), 1)).astype(np.float32)], 1)
    vg = api.VoxelGrid()
    vg.setLeafSize(leaf, leaf, leaf)
    vg.setInputCloud(scan)
    c1, n1 = vg.filter()
    c0, n0 = oracle.voxel_grid(scan, leaf)
    np.testing.assert_array_equal(n1, n0)
    np.testing.assert_array_equal(c1, c0)     # same fp32 sums in the same order
    assert n1.sum() == len(scan)


def test_downsample_edge_cases(oracle, api):
    vg = api.VoxelGrid()
    vg.setLeafSize(1.0)
    p = np.array([[0.1, 0.1, 0.1, 1], [0.2, 0.2, 0.2, 3], [np.nan, 0, 0, 5], [5.0, 5.0, 5.0, 7], [-3, -3, -3, 9]], np.float32)
    vg.setInputCloud(p)
    c1, n1 = vg.filter()
    c0, n0 = oracle.voxel_grid(p, 1.0)
    np.testing.assert_array_equal(c1, c0)
    np.testing.assert_array_equal(n1, n0)
    vg.setMinimumPointsNumberPerVoxel(2)
    c1, n1 = vg.filter()
    assert list(n1) == [2]
    vg.setMinimumPointsNumberPerVoxel(0)
    vg.setInputCloud(p[:, :3])                 # 12-byte records: intensity reads as 0
    c1, n1 = vg.filter()
    assert (c1[:, 3] == 0).all() and len(c1) == 3
    vg.setInputCloud(np.full((4, 4), np.nan, np.float32))
    c1, n1 = vg.filter()
    assert len(c1) == 0


def test_downsampled_scan_feeds_the_update_on_device(oracle, api, synth, small_cfg):
    """scan -> VoxelGrid -> IEKF update without leaving the device equals the host round trip."""
    g = api.IVox(resolution=0.5, nearby=18)
    g.AddPoints(small_cfg["map"])
    vg = api.VoxelGrid()
    vg.setLeafSize(0.5)
    vg.setInputCloud(small_cfg["scan"])
    c, n = vg.filter()
    kf = api.Esekf(g)
    kf.change_x(small_cfg["x_prop"]); kf.change_P(small_cfg["P"])
    kf.update_iterated_dyn_share_modified(c[:, :3])
    x_host = kf.get_x().copy()
    kf2 = api.Esekf(g)
    kf2.change_x(small_cfg["x_prop"]); kf2.change_P(small_cfg["P"])
    ptr, m = vg.device_points()
    assert m == len(c)
    kf2.update_device(ptr, m)
    np.testing.assert_array_equal(kf2.get_x(), x_host)


def make_frames(synth, k=12, n=6000):
    world = synth.make_world(synth.SEED, beams=True)
    frames, poses = [], []
    for i in range(k):
        pos = np.array([-10.0 + 2.0 * i, 3.0 * np.sin(0.5 * i), 1.2])
        q = synth.quat_from_rotvec([0.0, 0.0, 0.3 * i])
        R = synth.quat_to_R(q)
        pts = synth.raycast(pos, R, synth.avia_dirs(int(n * 1.2), seed=50 + i), world, seed=70 + i)[:n]
        inten = np.linspace(0, 100, len(pts), dtype=np.float32)[:, None]
        frames.append(np.ascontiguousarray(np.concatenate([pts, inten], 1)))
        poses.append([pos[0], pos[1], pos[2], q[3], q[0], q[1], q[2]])
    return frames, np.array(poses)


def test_full_map_builder_matches_oracle(oracle, api, synth):
    frames, poses = make_frames(synth)
    c0, n0 = oracle.full_map(frames, poses, 0.1)
    b = api.FullMapBuilder(leaf=0.1, capacity_voxels=500_000)
    for f, p in zip(frames, poses):
        b.add_keyframe(f, p)
    assert b.num_voxels() == len(c0)
    c1, n1 = b.extract()
    np.testing.assert_array_equal(n1, n0)                 # same voxels in the same (z, y, x) order with the same counts
    np.testing.assert_allclose(c1, c0, rtol=0, atol=2e-5)  # fp64 sums here vs fp32 running sums in pcl::CentroidPoint
    assert n1.sum() == sum(len(f) for f in frames)
    # keyframe order does not matter
    b2 = api.FullMapBuilder(leaf=0.1, capacity_voxels=500_000)
    for f, p in list(zip(frames, poses))[::-1]:
        b2.add_keyframe(f, p)
    c2, n2 = b2.extract()
    np.testing.assert_array_equal(n2, n1)
    # sums are fp32 atomics: offsets from the voxel corner (< one leaf) for x, y, z, the plain value for the intensity (0..100 here)
    np.testing.assert_allclose(c2[:, :3], c1[:, :3], rtol=0, atol=4e-6)   # one float32 ulp of a 20 m coordinate is 1.9e-6
    np.testing.assert_allclose(c2[:, 3], c1[:, 3], rtol=0, atol=1e-3)


def test_full_map_batched_device_keyframes_match_one_by_one(oracle, api, synth):
    """b200_mapbuild_add_keyframes_device (several keyframes per launch, ragged sizes, more keyframes than one launch
    batch) builds the same map as the one-by-one host path."""
    import torch
    frames, poses = make_frames(synth)
    frames = [f[: len(f) - 37 * i] for i, f in enumerate(frames)] * 9        # ragged, and > 24 keyframes
    poses = list(poses) * 9
    b = api.FullMapBuilder(leaf=0.1, capacity_voxels=500_000)
    for f, p in zip(frames, poses):
        b.add_keyframe(f, p)
    c0, n0 = b.extract()
    dev = []
    for f in frames:
        a = np.zeros((len(f), 4), np.float32)
        a[:, : f.shape[1]] = f
        dev.append(torch.from_numpy(a).cuda())
    b2 = api.FullMapBuilder(leaf=0.1, capacity_voxels=500_000)
    b2.add_keyframes_device([d.data_ptr() for d in dev], [len(f) for f in frames], np.stack(poses))
    assert b2.num_voxels() == len(c0)
    c1, n1 = b2.extract()
    np.testing.assert_array_equal(n1, n0)
    np.testing.assert_allclose(c1[:, :3], c0[:, :3], rtol=0, atol=4e-6)   # the centroid's float32 rounding may flip by one ulp (1.9e-6 at 20 m)
    np.testing.assert_allclose(c1[:, 3], c0[:, 3], rtol=0, atol=1e-3)   # intensity: fp32 sum, order of the atomics


def test_full_map_capacity_error(api, synth):
    frames, poses = make_frames(synth, k=2)
    b = api.FullMapBuilder(leaf=0.1, capacity_voxels=100)
    b.add_keyframe(frames[0], poses[0])
    with pytest.raises(api.B200Error):
        b.num_voxels()


def imu_scene(synth, n=8000, first_time_ms=0.0, seed=4):
    """A scan with per-point time offsets (ms, in curvature position 9 of a 12-float PointXYZINormal record) and the IMU
    poses of a smooth motion: 21 poses, 5 ms apart."""
    rng = np.random.default_rng(seed)
    K = 21
    w = np.array([0.3, -0.2, 0.8])          # rad/s
    acc = np.array([0.5, -0.3, 0.1])
    v0 = np.array([1.5, 0.2, -0.1])
    poses = np.zeros((K, 22))
    for k in range(K):
        t = 0.005 * k
        q = synth.quat_from_rotvec(w * t)
        poses[k, 0] = t
        poses[k, 1:4] = acc + rng.normal(0, 0.01, 3)
        poses[k, 4:7] = w + rng.normal(0, 0.005, 3)
        poses[k, 7:10] = v0 + acc * t
        poses[k, 10:13] = v0 * t + 0.5 * acc * t * t
        poses[k, 13:22] = synth.quat_to_R(q).reshape(9)
    t_end = 0.1
    x_end = synth.make_state(v0 * t_end + 0.5 * acc * t_end ** 2, w * t_end)
    pts = np.zeros((n, 12), np.float32)
    pts[:, :3] = rng.uniform(-30, 30, (n, 3))
    pts[:, 8] = rng.uniform(0, 255, n)
    pts[:, 9] = np.round(rng.uniform(first_time_ms, 100.0, n), 3)
    pts[::97, 9] = pts[3::97, 9][: len(pts[::97])]     # ties in time
    if first_time_ms == 0.0:
        pts[5, 9] = 0.0                                 # a point at t = 0 is not compensated
    return pts, poses, x_end


@pytest.mark.parametrize("first_ms", [0.0, 12.5])
def test_undistort_matches_oracle(oracle, api, synth, first_ms):
    pts, poses, x_end = imu_scene(synth, first_time_ms=first_ms)
    o_xyzi, o_order = oracle.undistort(pts, 9, 8, poses, x_end)
    vg = api.VoxelGrid()
    g_xyzi, g_order = vg.undistort(pts, 9, 8, poses, x_end)
    np.testing.assert_array_equal(g_order, o_order)                 # same time order, ties in input order
    np.testing.assert_array_equal(g_xyzi[:, 3], o_xyzi[:, 3])
    # fp64 chain narrowed to fp32: device sin/cos may differ from libm in the last bit of the double
    np.testing.assert_allclose(g_xyzi[:, :3], o_xyzi[:, :3], rtol=0, atol=4e-6)
    assert (g_xyzi[:, :3] == o_xyzi[:, :3]).mean() > 0.999
    moved = np.abs(o_xyzi[:, :3] - pts[o_order, :3]).max(1)
    assert moved.max() > 0.05                                       # the motion is not a no-op
    if first_ms == 0.0:
        assert moved[pts[o_order, 9] == 0.0].max() == 0.0           # t = 0: untouched (imu_processing.hpp:262)


def test_raw_scan_to_update_stays_on_device(oracle, api, synth, small_cfg):
    """raw scan -> undistort -> VoxelGrid -> IEKF update, chained on the device, equals the same chain through the host."""
    pts, poses, x_end = imu_scene(synth, n=len(small_cfg["scan"]))
    pts[:, :3] = small_cfg["scan"]
    vg = api.VoxelGrid()
    vg.setLeafSize(0.5)
    und, order = vg.undistort(pts, 9, 8, poses, x_end)
    m = vg.filter_staged()
    ptr, m2 = vg.device_points()
    assert m == m2 > 100
    o_und, _ = oracle.undistort(pts, 9, 8, poses, x_end)
    c0, n0 = oracle.voxel_grid(und, 0.5)        # the device's own undistorted points through the oracle's VoxelGrid
    assert len(c0) == m
    g = api.IVox(resolution=0.5, nearby=18)
    g.AddPoints(small_cfg["map"])
    kf = api.Esekf(g)
    kf.change_x(small_cfg["x_prop"]); kf.change_P(small_cfg["P"])
    kf.update_device(ptr, m)
    x_dev = kf.get_x().copy()
    kf2 = api.Esekf(g)
    kf2.change_x(small_cfg["x_prop"]); kf2.change_P(small_cfg["P"])
    kf2.update_iterated_dyn_share_modified(c0[:, :3])
    np.testing.assert_array_equal(x_dev, kf2.get_x())
