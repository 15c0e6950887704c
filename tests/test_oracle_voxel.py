"""Self-checks of the VoxelGrid / map-builder oracle (pcl::VoxelGrid is third-party; parity unpinned)."""
import numpy as np


def numpy_voxel_grid(p, leaf):
    inv = np.float32(1) / np.float32(leaf)
    ijk = np.floor(p[:, :3] * inv).astype(np.int64)
    ijk -= ijk.min(0)
    dv = ijk.max(0) + 1
    idx = ijk[:, 0] + ijk[:, 1] * dv[0] + ijk[:, 2] * dv[0] * dv[1]
    u, inv_i, cn = np.unique(idx, return_inverse=True, return_counts=True)
    ref = np.zeros((len(u), p.shape[1]))
    np.add.at(ref, inv_i, p.astype(np.float64))
    return ref / cn[:, None], cn


def test_voxel_grid_matches_numpy(oracle):
    rng = np.random.default_rng(0)
    p = rng.uniform(-7, 9, (30000, 4)).astype(np.float32)
    for leaf in (0.2, 0.5, 1.3):
        c, n = oracle.voxel_grid(p, leaf)
        ref, cn = numpy_voxel_grid(p, leaf)
        assert len(c) == len(ref) and (n == cn).all()
        np.testing.assert_allclose(c, ref, atol=5e-6)
        assert n.sum() == len(p)


def test_voxel_grid_edge_cases(oracle):
    p = np.array([[0.1, 0.1, 0.1, 1], [0.2, 0.2, 0.2, 3], [np.nan, 0, 0, 5], [5.0, 5.0, 5.0, 7]], np.float32)
    c, n = oracle.voxel_grid(p, 1.0)
    assert list(n) == [2, 1]
    np.testing.assert_allclose(c[0], [0.15, 0.15, 0.15, 2.0], atol=1e-6)
    c, n = oracle.voxel_grid(p, 1.0, min_points=2)
    assert list(n) == [2]
    c, n = oracle.voxel_grid(p[:, :3], 1.0)           # no intensity column
    assert c.shape == (2, 4) and (c[:, 3] == 0).all()
    one = np.array([[1.0, 2.0, 3.0, 4.0]], np.float32)
    c, n = oracle.voxel_grid(one, 0.1)
    np.testing.assert_array_equal(c, one)


def test_full_map_equals_voxel_grid_of_the_moved_union(oracle, synth):
    rng = np.random.default_rng(1)
    frames = [rng.uniform(-5, 5, (2000 + 100 * k, 4)).astype(np.float32) for k in range(4)]
    poses = []
    moved = []
    for k, f in enumerate(frames):
        q = synth.quat_from_rotvec([0.0, 0.02 * k, 0.3 * k])          # x y z w
        t = np.array([1.0 * k, -0.5 * k, 0.1])
        poses.append([t[0], t[1], t[2], q[3], q[0], q[1], q[2]])   # x y z qw qx qy qz
        R = synth.quat_to_R(q)
        moved.append(np.concatenate([(f[:, :3].astype(np.float64) @ R.T + t), f[:, 3:4]], 1))
    c, n = oracle.full_map(frames, np.array(poses), 0.25)
    allp = np.concatenate(moved, 0)
    assert n.sum() == len(allp)
    ref, cn = numpy_voxel_grid(allp.astype(np.float32), 0.25)
    # float vs double transforms move a few points across voxel borders: compare the bulk statistics
    assert abs(len(c) - len(ref)) <= 0.002 * len(ref)
    np.testing.assert_allclose((c[:, :3] * n[:, None]).sum(0) / n.sum(), allp[:, :3].mean(0), atol=1e-4)


def test_undistort_oracle_properties(oracle, synth):
    """No motion -> points unchanged (only sorted); pure translation at constant velocity -> each point shifted by -v * (T - t)
    expressed in the end frame."""
    rng = np.random.default_rng(2)
    n, K = 500, 11
    pts = np.zeros((n, 12), np.float32)
    pts[:, :3] = rng.uniform(-10, 10, (n, 3))
    pts[:, 9] = rng.uniform(0.001, 100.0, n)
    poses = np.zeros((K, 22))
    for k in range(K):
        poses[k, 0] = 0.01 * k
        poses[k, 13:22] = np.eye(3).reshape(9)
    x_end = synth.make_state([0, 0, 0], [0, 0, 0], ext_t=np.zeros(3))
    out, order = oracle.undistort(pts, 9, -1, poses, x_end)
    assert (np.diff(pts[order, 9]) >= 0).all()
    np.testing.assert_array_equal(out[:, :3], pts[order, :3])
    v = np.array([2.0, -1.0, 0.5])
    for k in range(K):
        poses[k, 7:10] = v
        poses[k, 10:13] = v * poses[k, 0]
    x_end = synth.make_state(v * 0.1, [0, 0, 0], ext_t=np.zeros(3))
    out, order = oracle.undistort(pts, 9, -1, poses, x_end)
    t = pts[order, 9].astype(np.float64) / 1000.0
    expect = pts[order, :3].astype(np.float64) - v * (0.1 - t)[:, None]
    np.testing.assert_allclose(out[:, :3], expect, atol=2e-6)
