"""Parity of the CUDA path (through the C ABI) against the CPU oracle — local map + IEKF update.

Bars (BASELINE.json north_star): neighbour index sets bit-exact; per-point planes / residuals / flags
bit-exact (fp32, same op sequence); H^T H and H^T h within 1e-6 relative; converged pose within
1e-4 m / 1e-4 rad (asserted far tighter).
"""
import numpy as np
import pytest

from conftest import world_scan

pytestmark = pytest.mark.gpu

PARAMS = {"livox": dict(resolution=0.2, nearby=26, ext=False), "horizon": dict(resolution=0.5, nearby=18, ext=True)}
H_TOL = 1e-6      # north_star: H/b within 1e-6 relative
POSE_TOL = 1e-4   # north_star: 1e-4 m / 1e-4 rad


def pair(oracle, api, pname, mp, **kw):
    p = PARAMS[pname]
    o = oracle.OracleLio(resolution=p["resolution"], nearby=p["nearby"], extrinsic_est_en=p["ext"], **kw)
    g = api.IVox(resolution=p["resolution"], nearby=p["nearby"])
    if mp is not None:
        o.insert(mp)
        g.AddPoints(mp)
    kf = api.Esekf(g, extrinsic_est_en=p["ext"], **{k: v for k, v in kw.items() if k in ("max_iter", "limit", "filter_size_map")})
    return o, g, kf


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("pname", ["livox", "horizon"])
def test_map_build_and_knn_bit_exact(oracle, api, synth, small_cfg, pname):
    o, g, _ = pair(oracle, api, pname, small_cfg["map"])
    assert g.NumValidGrids() == o.num_voxels
    assert g.NumPoints() == o.num_points
    q = world_scan(synth, small_cfg)
    i0, d0, c0 = o.knn5(q)
    i1, d1, c1 = g.GetClosestPoint(q)
    np.testing.assert_array_equal(c1, c0)
    np.testing.assert_array_equal(i1, i0)   # same ordinals in the same (distance, rank) order
    np.testing.assert_array_equal(d1, d0)   # and bit-identical fp32 distances


@pytest.mark.parametrize("pname", ["livox", "horizon"])
def test_incremental_insert_keeps_parity(oracle, api, synth, small_cfg, pname):
    mp = small_cfg["map"]
    o, g, _ = pair(oracle, api, pname, None)
    q = world_scan(synth, small_cfg)[:2000]
    rng = np.random.default_rng(11)
    start = 0
    for k, n in enumerate([1, 7, 50_000, 3, 20_000, 129_939]):  # ragged batches, voxels grow and relocate
        batch = mp[start:start + n]
        start += n
        o.insert(batch)
        g.AddPoints(batch)
        assert g.NumValidGrids() == o.num_voxels and g.NumPoints() == o.num_points
        i0, d0, c0 = o.knn5(q)
        i1, d1, c1 = g.GetClosestPoint(q)
        np.testing.assert_array_equal(i1, i0)
        np.testing.assert_array_equal(d1, d0)
        np.testing.assert_array_equal(c1, c0)
    # duplicates of existing points: ties at equal distance must resolve by enumeration rank
    dup = mp[rng.integers(0, start, 5000)]
    o.insert(dup)
    g.AddPoints(dup)
    i0, d0, c0 = o.knn5(dup[:1000])
    i1, d1, c1 = g.GetClosestPoint(dup[:1000])
    np.testing.assert_array_equal(i1, i0)
    np.testing.assert_array_equal(d1, d0)


def test_knn_edge_cases(oracle, api, small_cfg):
    o, g, _ = pair(oracle, api, "horizon", small_cfg["map"][:50_000])
    far = np.array([[500, 500, 500], [-1e4, 3, 2], [0, 0, 100]], np.float32)       # empty stencils
    i1, d1, c1 = g.GetClosestPoint(far)
    assert list(c1) == [0, 0, 0] and np.all(i1 == -1)
    i0, d0, c0 = o.knn5(far)
    np.testing.assert_array_equal(i1, i0)
    # strided input (48-byte PointXYZINormal records)
    rec = np.zeros((1000, 12), np.float32)
    rec[:, :3] = small_cfg["map"][:1000]
    rec[:, 3:] = 7.0
    i2, d2, c2 = g.GetClosestPoint(rec)
    i3, d3, c3 = o.knn5(rec)
    np.testing.assert_array_equal(i2, i3)
    assert np.all(i2[:, 0] == np.arange(1000)) and np.all(d2[:, 0] == 0)
    # zero queries / zero points are no-ops
    g.AddPoints(np.zeros((0, 3), np.float32))
    assert g.GetClosestPoint(np.zeros((0, 3), np.float32))[0].shape == (0, 5)


def test_bad_points_are_dropped_not_fatal(oracle, api, small_cfg):
    """One NaN / Inf / out-of-range lidar return must not poison the map or fail every later insert: such points
    are skipped (they still take an ordinal), reported through b200_map_dropped, and the rest of the batch goes in."""
    mp = small_cfg["map"][:30_000].copy()
    bad_at = [5, 77, 20_000]
    dirty = mp.copy()
    dirty[bad_at[0]] = [np.nan, 0, 0]
    dirty[bad_at[1]] = [1e9, 0, 0]
    dirty[bad_at[2]] = [0, np.inf, 1]
    g = api.IVox(resolution=0.5, nearby=18)
    g.AddPoints(dirty)                       # no exception
    assert g.dropped() == (3, 3)
    g.AddPoints(small_cfg["map"][30_000:40_000])   # a clean batch after a dirty one is clean
    assert g.dropped() == (3, 0)
    # oracle: the same cloud with the bad points moved far away into voxels no query reaches is equivalent for every query
    o = oracle.OracleLio(resolution=0.5, nearby=18)
    clean = mp.copy()
    clean[bad_at] = [[9e3, 9e3, 9e3], [9.1e3, 9e3, 9e3], [9.2e3, 9e3, 9e3]]
    o.insert(clean)
    o.insert(small_cfg["map"][30_000:40_000])
    assert g.NumPoints() == o.num_points - 3 and g.NumValidGrids() == o.num_voxels - 3
    q = small_cfg["map"][:3000] + np.float32(0.01)
    i0, d0, c0 = o.knn5(q)
    i1, d1, c1 = g.GetClosestPoint(q)
    np.testing.assert_array_equal(i1, i0)
    np.testing.assert_array_equal(d1, d0)
    np.testing.assert_array_equal(c1, c0)


@pytest.mark.parametrize("pname", ["livox", "horizon"])
@pytest.mark.parametrize("converge", [True, False])
def test_obs_model_parity(oracle, api, small_cfg, pname, converge):
    o, g, kf = pair(oracle, api, pname, small_cfg["map"])
    scan, x = small_cfg["scan"], small_cfg["x_prop"]
    if not converge:  # a no-search pass reuses the neighbours/planes of an earlier search pass
        o.obs_model(scan, small_cfg["x_true"], True)
        kf.ObsModel(scan, small_cfg["x_true"], True)
    rc0, H0, h0, n0 = o.obs_model(scan, x, converge)
    rc1, H1, h1, n1 = kf.ObsModel(scan, x, converge)
    assert rc0 == rc1 and n0 == n1 and n0 > 1000
    assert relerr(H1, H0) < H_TOL and relerr(h1, h0) < H_TOL
    s0, s1 = o.point_state(len(scan)), kf.point_state()
    for key in ("plane", "residual", "selected", "nn_idx", "nn_count"):
        np.testing.assert_array_equal(s1[key], s0[key], err_msg=key)


@pytest.mark.parametrize("pname", ["livox", "horizon"])
def test_iekf_update_parity(oracle, api, small_cfg, pname):
    o, g, kf = pair(oracle, api, pname, small_cfg["map"])
    rc0, x0, P0, st0 = o.update(small_cfg["scan"], small_cfg["x_prop"], small_cfg["P"])
    kf.change_x(small_cfg["x_prop"])
    kf.change_P(small_cfg["P"])
    rc1 = kf.update_iterated_dyn_share_modified(small_cfg["scan"])
    st1 = kf.stats
    assert rc0 == rc1 == 0
    assert (st1.passes, st1.knn_passes, st1.converged) == (st0.passes, st0.knn_passes, st0.converged)
    assert list(st1.n_eff)[:st1.passes] == list(st0.n_eff)[:st0.passes]
    assert list(st1.knn)[:st1.passes] == list(st0.knn)[:st0.passes]
    d = oracle.boxminus(kf.get_x(), x0)
    assert np.abs(d[0:3]).max() < POSE_TOL * 1e-3 and np.abs(d[3:6]).max() < POSE_TOL * 1e-3   # far inside 1e-4
    assert np.abs(kf.get_x() - x0).max() < 1e-7
    assert relerr(kf.get_P(), P0) < 1e-6
    for p in range(st1.passes):
        H, h, xin = kf.last_HtH(p)
        assert np.abs(xin - np.array(st0.x_in[p])).max() < 1e-8
        assert relerr(H, np.array(st0.HtH[p]).reshape(12, 12)) < H_TOL
        # H^T h cancels to ~0 near convergence: gate it against the scale of |H|^T |h| instead of its own size
        assert np.abs(h - np.array(st0.Hth[p])).max() < H_TOL * np.sqrt(np.trace(np.array(st0.HtH[p]).reshape(12, 12)) * st0.n_eff[p]) * 1e-2
    s0, s1 = o.point_state(len(small_cfg["scan"])), kf.point_state()
    for key in ("selected", "nn_idx", "nn_count"):
        np.testing.assert_array_equal(s1[key], s0[key], err_msg=key)


def test_pinned_scan_takes_the_unpacked_path_with_identical_results(oracle, api, small_cfg):
    """A scan in page-locked memory (b200_host_alloc) is unpacked on the device: same posterior, bit for bit, for packed
    12-byte records and for 48-byte PointXYZINormal-strided records."""
    scan = small_cfg["scan"]

    def run(cloud):  # fresh filter every time: the per-point arrays persist across scans by design
        o, g, kf = pair(oracle, api, "livox", small_cfg["map"])
        kf.change_x(small_cfg["x_prop"]); kf.change_P(small_cfg["P"])
        assert kf.update_iterated_dyn_share_modified(cloud) == 0
        return kf.get_x().copy(), kf.get_P().copy()

    x_ref, P_ref = run(scan)
    for cols in (3, 12):
        pin = api.PinnedCloud(len(scan), cols)
        pin.array[:] = 7.0
        pin.array[:, :3] = scan
        x, P = run(pin.array)
        assert np.array_equal(x, x_ref) and np.array_equal(P, P_ref)
        pin.close()


def test_update_without_map_reports_no_effective_points(oracle, api, small_cfg):
    o, g, kf = pair(oracle, api, "horizon", None)
    kf.change_x(small_cfg["x_prop"])
    kf.change_P(small_cfg["P"])
    rc = kf.update_iterated_dyn_share_modified(small_cfg["scan"][:777])
    assert rc == 1 and kf.stats.status == 1 and kf.stats.passes == 4     # B200_NO_EFFECTIVE_POINTS, every pass skipped
    np.testing.assert_array_equal(kf.get_x(), small_cfg["x_prop"])
    np.testing.assert_array_equal(kf.get_P(), small_cfg["P"])


@pytest.mark.parametrize("pname", ["livox", "horizon"])
def test_scan_sequence_with_map_incremental(oracle, api, synth, pname):
    """config 3 in miniature: a short trajectory, update + MapIncremental per scan, states carried over.
    Exercises the persistent per-point arrays (stale residuals, shrinking/growing scans) and map growth."""
    world = synth.make_world()
    mp = synth.sample_map(150_000, world=world)
    o, g, kf = pair(oracle, api, pname, mp)
    x = synth.make_state([0.0, 0.0, 1.0], [0.0, 0.0, 0.1])
    xo, Po = x.copy(), synth.init_cov()
    kf.change_x(x)
    kf.change_P(Po)
    rng = np.random.default_rng(2)
    for k in range(8):
        x_true = synth.make_state([0.4 * k, 0.1 * k, 1.0], [0.0, 0.0, 0.1 + 0.02 * k])
        ol, Rl = synth.lidar_pose(x_true)
        nrays = int(rng.integers(2500, 4000))      # scan sizes go up and down
        scan = synth.raycast(ol, Rl, synth.livox_dirs(nrays, seed=100 + k), world, seed=200 + k)
        # prediction step stand-in: move both filters to the same perturbed prior
        prior = synth.perturb_state(x_true, seed=300 + k, dpos=0.03, drot_deg=0.3)
        Pprior = synth.init_cov(seed=400 + k) * 0.01
        rc0, xo, Po, st0 = o.update(scan, prior, Pprior)
        kf.change_x(prior)
        kf.change_P(Pprior)
        rc1 = kf.update_iterated_dyn_share_modified(scan)
        assert rc0 == rc1 == 0
        assert list(kf.stats.n_eff)[:4] == list(st0.n_eff)[:4], f"scan {k}"
        d = oracle.boxminus(kf.get_x(), xo)
        assert np.abs(d[:6]).max() < POSE_TOL * 1e-3, f"scan {k}"
        # both maps are grown from the ORACLE's posterior so the inserted points are bit-identical
        tot, na0, nd0 = o.map_incremental(scan, xo, True)
        na1, nd1 = kf.MapIncremental(xo, True)
        assert (na1, nd1) == (na0, nd0), f"scan {k}"
        assert g.NumValidGrids() == o.num_voxels and g.NumPoints() == o.num_points
    q = mp[:3000] + np.float32(0.01)
    i0, d0, c0 = o.knn5(q)
    i1, d1, c1 = g.GetClosestPoint(q)
    np.testing.assert_array_equal(i1, i0)
    np.testing.assert_array_equal(d1, d0)


@pytest.mark.parametrize("pname", ["livox", "horizon"])
def test_full_size_properties(api, synth, pname):
    """BASELINE.json configs[0] at full size (2M-point map, 20k scan): size-independent properties —
    sorted distances, self-match of map points, counts, and agreement with a numpy brute-force kNN on a
    random sample of queries (exact kNN restricted to d < 5 m equals the stencil kNN whenever the 5th
    distance is below the stencil's guaranteed radius)."""
    c = synth.config1()
    p = PARAMS[pname]
    g = api.IVox(resolution=p["resolution"], nearby=p["nearby"])
    g.AddPoints(c["map"])
    assert g.NumPoints() == len(c["map"])
    o_l, Rl = synth.lidar_pose(c["x_true"])
    q = (c["scan"].astype(np.float64) @ Rl.T + o_l).astype(np.float32)
    idx, d2, cnt = g.GetClosestPoint(q)
    assert np.all(cnt >= 0) and np.all(cnt <= 5) and np.mean(cnt == 5) > 0.9
    for k in range(4):
        ok = cnt > k + 1
        assert np.all(d2[ok, k] <= d2[ok, k + 1])
    full = cnt == 5
    d_chk = ((c["map"][idx[full]] - q[full, None, :]) ** 2).sum(-1)
    np.testing.assert_allclose(d_chk, d2[full], rtol=1e-4, atol=1e-7)
    # guaranteed radius: the query lies within res/2 of its cell centre on every axis, so the full
    # 27-cell stencil contains every map point within `res` of it; the 19-cell stencil (no corner
    # cells) has no such radius, so the brute-force cross-check runs for NEARBY26 only.
    if p["nearby"] == 26:
        rng = np.random.default_rng(0)
        r_safe = p["resolution"] ** 2
        checked = 0
        for i in rng.permutation(np.nonzero(full)[0])[:60]:
            if d2[i, 4] >= r_safe:
                continue
            dd = ((c["map"] - q[i]) ** 2).sum(1)
            bf = np.argsort(dd, kind="stable")[:5]
            assert set(bf) == set(idx[i])
            checked += 1
        assert checked > 10
    # map points find themselves
    i_self, d_self, _ = g.GetClosestPoint(c["map"][::4001])
    assert np.all(i_self[:, 0] == np.arange(0, len(c["map"]), 4001)) and np.all(d_self[:, 0] == 0)
    # one full update runs, converges and lands near the truth
    kf = api.Esekf(g, extrinsic_est_en=p["ext"])
    kf.change_x(c["x_prop"])
    kf.change_P(c["P"])
    assert kf.update_iterated_dyn_share_modified(c["scan"]) == 0
    assert kf.stats.n_eff[0] > 0.9 * len(c["scan"])
    err = kf.get_x()[:3] - c["x_true"][:3]
    assert np.linalg.norm(err) < 0.02


@pytest.mark.parametrize("capacity,batch", [(3000, 2500), (1500, 4000), (800, 700)])
def test_lru_eviction_matches_reference_semantics(oracle, api, synth, small_cfg, capacity, batch):
    """IVox::AddPoints evicts the least recently touched voxel whenever a new voxel brings the map to capacity_
    (ivox3d.h:268-275), point by point.  The GPU map decides the same victims per batch, including voxels that are
    evicted and re-created inside one batch."""
    mp = small_cfg["map"]
    o = oracle.OracleLio(resolution=0.5, nearby=18, capacity=capacity)
    g = api.IVox(resolution=0.5, nearby=18, capacity=capacity)
    q = world_scan(synth, small_cfg)[:1500]
    rng = np.random.default_rng(5)
    start = 0
    for k in range(12):
        # alternate fresh regions with revisits of old points so that recency matters
        if k % 3 == 2:
            sel = rng.integers(0, max(start, 1), batch // 2)
            pts = np.ascontiguousarray(mp[sel])
        else:
            pts = mp[start:start + batch]
            start += batch
        o.insert(pts)
        g.AddPoints(pts)
        assert g.NumValidGrids() == o.num_voxels, k
        assert g.NumPoints() == o.num_points, k
        i0, d0, c0 = o.knn5(q)
        i1, d1, c1 = g.GetClosestPoint(q)
        np.testing.assert_array_equal(c1, c0)
        np.testing.assert_array_equal(i1, i0)
    assert o.num_voxels == capacity - 1   # the map sits at the capacity: evictions did happen


def test_predict_parity(oracle, api, synth, small_cfg):
    """esekf::predict over the IMU intervals of a scan, on the device in one launch, against the oracle: state, covariance and
    the IMUpose_ list the undistortion pass consumes."""
    from test_oracle_lio import imu_steps, Q12
    x0 = synth.make_state([1.0, -2.0, 0.5], [0.02, -0.01, 0.7])
    x0[14:17] = [0.8, -0.3, 0.05]
    x0[17:20] = [0.001, -0.002, 0.0005]
    x0[20:23] = [0.01, 0.02, -0.01]
    P0 = synth.init_cov()
    g = api.IVox(resolution=0.5, nearby=18)
    kf = api.Esekf(g)
    for K in (1, 20, 57):
        steps = imu_steps(K, seed=K)
        x_o, P_o, poses_o = oracle.predict(steps, Q12, x0, P0)
        kf.change_x(x0)
        kf.change_P(P0)
        poses_g = kf.predict(steps, Q12)
        assert np.abs(oracle.boxminus(kf.get_x(), x_o)).max() < 1e-12
        assert relerr(kf.get_P(), P_o) < 1e-12
        assert relerr(poses_g, poses_o) < 1e-12


def test_world_scan_is_point_body_to_world(oracle, api, synth, small_cfg):
    """/cloud_registered: the scan moved by PointBodyToWorld (laser_mapping.cc:855-864, fp64 quaternion arithmetic narrowed on
    store) at the posterior - against the same transform in numpy fp64."""
    o, g, kf = pair(oracle, api, "horizon", small_cfg["map"])
    kf.change_x(small_cfg["x_prop"])
    kf.change_P(small_cfg["P"])
    kf.update_iterated_dyn_share_modified(small_cfg["scan"])
    x = kf.get_x()
    w = kf.world_scan(cols=12)            # 48-byte PointXYZINormal records
    assert w.shape == (len(small_cfg["scan"]), 12) and np.all(w[:, 3:] == 0)
    R, Ro = synth.quat_to_R(x[3:7]), synth.quat_to_R(x[7:11])
    want = (R @ (Ro @ small_cfg["scan"].astype(np.float64).T + x[11:14, None]) + x[0:3, None]).T
    assert np.abs(w[:, :3] - want).max() <= 2e-6 * max(1.0, np.abs(want).max())
    assert (w[:, :3] == want.astype(np.float32)).mean() > 0.98       # identical after narrowing for all but rounding-boundary cases
