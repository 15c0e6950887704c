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
    np.testing.assert_allclose(c2, c1, rtol=0, atol=1e-6)


def test_full_map_capacity_error(api, synth):
    frames, poses = make_frames(synth, k=2)
    b = api.FullMapBuilder(leaf=0.1, capacity_voxels=100)
    b.add_keyframe(frames[0], poses[0])
    with pytest.raises(api.B200Error):
        b.num_voxels()
