"""Self-checks that pin the LIO oracle (the reference ships no vectors for this path: parity unpinned).

1. stencil kNN == numpy brute force over the same stencil cells;
2. esti_plane == numpy least squares within fp32 tolerance, and its accept/reject rule;
3. manifold boxplus/boxminus are inverse to each other; inverse() agrees with numpy;
4. Jacobian rows == central differences of the residual w.r.t. boxplus perturbations;
5. the IEKF recovers the seeded pose on noise-free data;
6. LRU voxel eviction at capacity.
"""
import numpy as np
import pytest

from conftest import world_scan

STENCIL26 = np.array([[0, 0, 0], [-1, 0, 0], [1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, -1], [0, 0, 1], [1, 1, 0], [-1, 1, 0],
                      [1, -1, 0], [-1, -1, 0], [1, 0, 1], [-1, 0, 1], [1, 0, -1], [-1, 0, -1], [0, 1, 1], [0, -1, 1],
                      [0, 1, -1], [0, -1, -1], [1, 1, 1], [-1, 1, 1], [1, -1, 1], [1, 1, -1], [-1, -1, 1], [-1, 1, -1],
                      [1, -1, -1], [-1, -1, -1]])


def cell_of(p, res):
    inv = np.float32(1.0 / np.float64(np.float32(res)))
    v = p.astype(np.float32) * inv
    return (np.sign(v) * np.floor(np.abs(v) + np.float32(0.5))).astype(np.int64)  # std::round: half away from zero


def brute_force_knn(mp, q, res, nstencil):
    cells = cell_of(mp, res)
    lut = {}
    for i, c in enumerate(map(tuple, cells)):
        lut.setdefault(c, []).append(i)
    out = []
    for p in q:
        c = cell_of(p[None], res)[0]
        cand = []
        for s in STENCIL26[:nstencil]:
            cand += lut.get(tuple(c + s), [])
        if not cand:
            out.append([])
            continue
        cand = np.array(cand)
        d = mp[cand].astype(np.float32) - p.astype(np.float32)
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        keep = d2 < 25.0
        order = np.lexsort((np.arange(len(cand))[keep], d2[keep]))[:5]
        out.append(list(cand[keep][order]))
    return out


@pytest.mark.parametrize("res,nearby,nst", [(0.5, 18, 19), (0.2, 26, 27), (0.5, 6, 7), (1.0, 0, 1)])
def test_knn_equals_brute_force(oracle, synth, small_cfg, res, nearby, nst):
    mp = small_cfg["map"][:60000]
    q = world_scan(synth, small_cfg)[:400]
    o = oracle.OracleLio(resolution=res, nearby=nearby)
    o.insert(mp)
    idx, d2, cnt = o.knn5(q)
    ref = brute_force_knn(mp, q, res, nst)
    for i in range(len(q)):
        assert cnt[i] == len(ref[i])
        assert list(idx[i, :cnt[i]]) == ref[i]
        assert np.all(np.diff(d2[i, :cnt[i]]) >= 0)


def test_map_counts_and_ordinals(oracle, small_cfg):
    mp = small_cfg["map"][:20000]
    o = oracle.OracleLio(resolution=0.5, nearby=18)
    assert o.insert(mp[:12000]) == 12000
    assert o.insert(mp[12000:]) == 20000
    assert o.num_points == 20000
    assert o.num_voxels == len({tuple(c) for c in cell_of(mp, 0.5)})
    idx, _, _ = o.knn5(mp[15000:15010])  # a map point is its own nearest neighbour
    assert list(idx[:, 0]) == list(range(15000, 15010))


def test_lru_eviction_at_capacity(oracle):
    # capacity 4: the map holds at most 3 voxels after any insertion (ivox3d.h:271-274)
    o = oracle.OracleLio(resolution=1.0, nearby=0, capacity=4)
    pts = np.array([[0, 0, 0], [10, 0, 0], [20, 0, 0], [30, 0, 0]], np.float32)
    o.insert(pts[:3])
    assert o.num_voxels == 3
    o.insert(pts[:1])          # touch voxel 0 -> most recent
    o.insert(pts[3:])          # new voxel -> size hits capacity -> evict LRU tail (voxel at x=10)
    assert o.num_voxels == 3
    _, _, cnt = o.knn5(pts)
    assert list(cnt) == [2, 0, 1, 1]


def test_esti_plane_matches_least_squares(oracle):
    rng = np.random.default_rng(3)
    for trial in range(200):
        n = rng.integers(3, 6)
        nrm = rng.normal(size=3)
        nrm /= np.linalg.norm(nrm)
        c = rng.uniform(-40, 40, 3)
        basis = np.linalg.svd(nrm[None])[2][1:]
        pts = c + rng.uniform(-0.3, 0.3, (n, 2)) @ basis + rng.normal(0, 0.005, (n, 1)) * nrm
        pts = pts.astype(np.float32)
        ok, plane = oracle.esti_plane(pts, 0.1)
        x = np.linalg.lstsq(pts.astype(np.float64), -np.ones(n), rcond=None)[0]
        ref = np.append(x / np.linalg.norm(x), 1 / np.linalg.norm(x))
        # absolute coordinates make A ill-conditioned in fp32: compare point-to-plane distances
        d_or = pts @ plane[:3] + plane[3]
        d_ref = pts @ ref[:3] + ref[3]
        if n == 5:
            assert np.max(np.abs(d_or - d_ref)) < 2e-2
        else:
            assert np.max(np.abs(d_or - d_ref)) < 1e-3
        assert ok == bool(np.all(np.abs(d_or) <= 0.1 + 1e-6)) or abs(np.max(np.abs(d_or)) - 0.1) < 1e-4
    ok, _ = oracle.esti_plane(np.zeros((2, 3), np.float32), 0.1)
    assert not ok


def test_esti_plane_rejects_non_planar(oracle):
    pts = np.array([[10, 0, 0], [10.5, 0, 0.4], [10, 0.5, -0.4], [10.5, 0.5, 0.5], [10.2, 0.2, -0.5]], np.float32)
    ok, _ = oracle.esti_plane(pts, 0.1)
    assert not ok


def test_manifold_roundtrip(oracle, synth):
    rng = np.random.default_rng(5)
    x = synth.make_state([1, 2, 3], [0.3, -0.2, 0.9])
    x[7:11] = synth.quat_from_rotvec([0.01, 0.02, -0.03])
    for _ in range(50):
        dx = rng.normal(0, 0.05, 23)
        y = oracle.boxplus(x, dx)
        assert abs(np.linalg.norm(y[23:26]) - 9.809) < 1e-12
        assert abs(np.linalg.norm(y[3:7]) - 1) < 1e-12
        np.testing.assert_allclose(oracle.boxminus(y, x), dx, atol=1e-9)
    np.testing.assert_allclose(oracle.boxminus(x, x), 0, atol=1e-15)


def test_inverse_matches_numpy(oracle, synth):
    P = synth.init_cov() / 0.001
    np.testing.assert_allclose(oracle.inverse(P), np.linalg.inv(P), rtol=1e-8, atol=1e-12)


def _residuals(oracle_mod, lio, scan, x):
    lio.obs_model(scan, x, converge=False)
    ps = lio.point_state(len(scan))
    return ps["residual"].astype(np.float64), ps["selected"].astype(bool)


@pytest.mark.parametrize("ext", [False, True])
def test_jacobian_rows_match_finite_differences(oracle, small_cfg, ext):
    scan = small_cfg["scan"][:1500]
    x0 = small_cfg["x_prop"].copy()
    lio = oracle.OracleLio(resolution=0.5, nearby=18, extrinsic_est_en=ext)
    lio.insert(small_cfg["map"])
    rc, HtH, Hth, ne = lio.obs_model(scan, x0, converge=True)  # fixes neighbours and planes
    hx, hv, n_rows = lio.last_rows(len(scan))
    ps = lio.point_state(len(scan))
    sel = ps["selected"].astype(bool)
    r0 = ps["residual"].astype(np.float64)
    assert n_rows == sel.sum() == ne
    np.testing.assert_allclose(hv, -r0[sel], atol=0)
    np.testing.assert_allclose(HtH, hx.T @ hx, rtol=1e-12)
    np.testing.assert_allclose(Hth, hx.T @ hv, rtol=1e-10, atol=1e-9)
    # tangent order: pos(0-2) rot(3-5) offR(6-8) offT(9-11); row = d residual / d tangent
    ncol = 12 if ext else 6
    eps = 1e-3  # residuals are fp32 (~1e-6 abs noise at tens of metres)
    for j in range(ncol):
        d = np.zeros(23)
        d[j] = eps
        rp, sp = _residuals(oracle, lio, scan, oracle.boxplus(x0, d))
        rm, sm = _residuals(oracle, lio, scan, oracle.boxplus(x0, -d))
        ok = sel & sp & sm
        fd = (rp - rm)[ok] / (2 * eps)
        ana = hx[ok[sel], j]
        assert np.median(np.abs(fd - ana)) < 5e-3
        assert np.percentile(np.abs(fd - ana), 99) < 5e-2 * max(1.0, np.abs(ana).max())
    if not ext:
        assert np.all(hx[:, 6:] == 0)


def test_iekf_recovers_pose_on_noise_free_data(oracle, synth):
    world = synth.make_world()
    mp = synth.sample_map(300_000, sigma=0.0, world=world)
    x_true = synth.make_state([3.0, -2.0, 1.2], [0.01, -0.02, 0.6])
    o, Rl = synth.lidar_pose(x_true)
    scan = synth.raycast(o, Rl, synth.livox_dirs(6000), world, sigma=0.0)[:5000]
    lio = oracle.OracleLio(resolution=0.5, nearby=18, max_iter=4)
    lio.insert(mp)
    x = synth.perturb_state(x_true)
    P = synth.init_cov()
    for _ in range(3):  # a few scans' worth of updates
        rc, x, P, st = lio.update(scan, x, P)
        assert rc == 0
    d = oracle.boxminus(x, x_true)
    e0 = oracle.boxminus(synth.perturb_state(x_true), x_true)
    # The fixed point of k=5 point-to-plane matching is a few mm off the true pose (planes fitted across
    # wall/floor corners bias it), so the gate is "an order of magnitude better than the 3 cm / 0.5 deg start".
    assert np.linalg.norm(d[0:3]) < 6e-3 < 0.2 * np.linalg.norm(e0[0:3]), d[:3]
    assert np.linalg.norm(d[3:6]) < 2e-4 < 0.05 * np.linalg.norm(e0[3:6]), d[3:6]


def test_update_bookkeeping(oracle, small_cfg):
    lio = oracle.OracleLio(resolution=0.5, nearby=18)
    lio.insert(small_cfg["map"])
    rc, x, P, st = lio.update(small_cfg["scan"], small_cfg["x_prop"], small_cfg["P"])
    assert rc == 0 and 2 <= st.passes <= 4 and st.knn[0] == 1
    assert np.allclose(P, P.T, atol=1e-9)
    assert np.all(np.linalg.eigvalsh(0.5 * (P + P.T)) > 0)
    assert np.all(np.diag(P)[:6] < np.diag(small_cfg["P"])[:6])  # the scan is informative about the pose
    # no map -> no effective points -> state untouched, status 1
    empty = oracle.OracleLio(resolution=0.5, nearby=18)
    rc, x2, P2, st2 = empty.update(small_cfg["scan"], small_cfg["x_prop"], small_cfg["P"])
    assert rc == 1 and st2.passes == 4
    np.testing.assert_array_equal(x2, small_cfg["x_prop"])
    np.testing.assert_array_equal(P2, small_cfg["P"])


def imu_steps(K=20, seed=4):
    """K IMU intervals the way UndistortPcl feeds predict (imu_processing.hpp:190-241): dt ~ 5 ms, gravity-compensated motion."""
    rng = np.random.default_rng(seed)
    steps = np.zeros((K, 8))
    steps[:, 0] = rng.uniform(0.004, 0.006, K)
    steps[:, 1] = np.cumsum(steps[:, 0])
    steps[:, 2:5] = np.array([0.3, -0.2, 9.81]) + rng.normal(0, 0.2, (K, 3))
    steps[:, 5:8] = np.array([0.05, -0.1, 0.4]) + rng.normal(0, 0.05, (K, 3))
    return steps


Q12 = np.array([0.1] * 3 + [0.1] * 3 + [1e-4] * 3 + [1e-4] * 3)   # gyr_cov, acc_cov, b_gyr_cov, b_acc_cov (config/livox.yaml:13-16)


def test_predict_restatement(oracle, synth):
    """esekf::predict (esekfom.hpp:269-374): the mean follows the strap-down kinematics of get_f, the covariance stays symmetric
    positive definite and grows by the process noise, and one step equals the closed form F P F^T + G Q G^T built here in numpy."""
    x0 = synth.make_state([1.0, -2.0, 0.5], [0.02, -0.01, 0.7])
    x0[14:17] = [0.8, -0.3, 0.05]            # vel
    x0[17:20] = [0.001, -0.002, 0.0005]      # bg
    x0[20:23] = [0.01, 0.02, -0.01]          # ba
    P0 = synth.init_cov()
    steps = imu_steps(1)
    x1, P1, poses = oracle.predict(steps, Q12, x0, P0)
    dt, acc, gyr = steps[0, 0], steps[0, 2:5], steps[0, 5:8]
    R = synth.quat_to_R(x0[3:7])
    a_in = R @ (acc - x0[20:23]) + x0[23:26]
    np.testing.assert_allclose(x1[0:3], x0[0:3] + dt * x0[14:17], atol=1e-14)
    np.testing.assert_allclose(x1[14:17], x0[14:17] + dt * a_in, atol=1e-13)
    dq = synth.quat_from_rotvec((gyr - x0[17:20]) * dt)
    np.testing.assert_allclose(x1[3:7], synth.quat_mul(x0[3:7], dq), atol=1e-14)
    for sl in (slice(7, 14), slice(17, 26)):                      # extrinsics, biases, gravity do not move
        np.testing.assert_array_equal(x1[sl], x0[sl])
    np.testing.assert_allclose(P1, P1.T, atol=1e-15)
    assert np.linalg.eigvalsh(0.5 * (P1 + P1.T)).min() > 0
    # the velocity rows of F against central differences of the mean propagation (tangent perturbations of the prior)
    def vel_after(dx):
        xp = oracle.boxplus(x0, dx)
        return oracle.predict(steps, Q12, xp, P0)[0][14:17]
    for col in (3, 4, 5, 12, 13, 18, 19, 20):                     # d vel' / d (rot, vel, ba)
        e = np.zeros(23)
        e[col] = 1e-6
        fd = (vel_after(e) - vel_after(-e)) / 2e-6
        F_col = (np.eye(3)[:, col - 12] if 12 <= col < 15 else 0) + 0
        if 3 <= col < 6:
            am = acc - x0[20:23]
            hat = np.array([[0, -am[2], am[1]], [am[2], 0, -am[0]], [-am[1], am[0], 0]])
            F_col = (-R @ hat * dt)[:, col - 3]
        elif 18 <= col < 21:
            F_col = (-R * dt)[:, col - 18]
        np.testing.assert_allclose(fd, F_col, atol=2e-8)
    # IMUpose_ entry: offset time, world acceleration incl. gravity, bias-free angular rate, velocity, position, rotation
    assert poses.shape == (1, 22) and poses[0, 0] == steps[0, 1]
    np.testing.assert_allclose(poses[0, 7:10], x1[14:17])
    np.testing.assert_allclose(poses[0, 10:13], x1[0:3])
    np.testing.assert_allclose(poses[0, 13:22].reshape(3, 3), synth.quat_to_R(x1[3:7]), atol=1e-14)
