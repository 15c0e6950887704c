"""Self-checks that pin the NDT oracle (the reference ships no vectors for pclomp NDT: parity unpinned).

1. voxel leaves == a numpy restatement of VoxelGridCovariance::applyFilter (membership, mean, the covariance formula
   with the Leaf() identity start and the (n-1)/n factor, eigenvalue inflation, inverse);
2. Matrix3f::eulerAngles(0,1,2) round-trips through Translation * Rx * Ry * Rz;
3. the score gradient / Hessian == central differences of the score / gradient (on voxel-interior points);
4. the fp64 computeHessian path agrees with the float path away from the float table's +sy quirk;
5. align() recovers the seeded pose on synthetic data, and takes >= 2 Newton iterations (A.6);
6. calculateScore ranks the true pose first in a hypothesis grid.
"""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def ndt_small(synth):
    world = synth.make_world(synth.SEED, beams=True)
    mp = synth.sample_map(300_000, synth.SEED, world=world)
    p_true = np.array([3.0, -2.0, 1.2, 0.0, 0.0, 0.6])
    T = synth.pose_vec_to_matrix(p_true)
    dirs = synth.livox_dirs(5000, synth.SEED)
    scan = np.ascontiguousarray(synth.raycast(T[:3, 3], T[:3, :3], dirs, world, seed=synth.SEED)[:4000])
    return dict(map=mp, scan=scan, p_true=p_true, T_true=T)


def numpy_leaves(mp, res, min_pts=6, eig_ratio=0.01):
    inv = np.float32(1.0) / np.float32(res)
    mn, mx = mp.min(0), mp.max(0)
    min_b = np.floor(mn * inv).astype(np.int64)
    max_b = np.floor(mx * inv).astype(np.int64)
    div = max_b - min_b + 1
    ijk = (np.floor(mp * inv) - min_b.astype(np.float32)).astype(np.int64)
    ids = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(ids, kind="stable")
    out = {}
    uniq, start, cnt = np.unique(ids[order], return_index=True, return_counts=True)
    P = mp.astype(np.float64)[order]
    for u, s, c in zip(uniq, start, cnt):
        if c < min_pts:
            continue
        pts = P[s:s + c]
        mu = pts.mean(0)
        cov = (np.eye(3) + pts.T @ pts) / c - np.outer(mu, mu)
        cov *= (c - 1.0) / c
        w, V = np.linalg.eigh(cov)
        if w[0] < 0 or w[1] < 0 or w[2] <= 0:
            continue
        if w[0] < eig_ratio * w[2]:
            w[0] = eig_ratio * w[2]
            w[1] = max(w[1], eig_ratio * w[2])
            cov = V @ np.diag(w) @ np.linalg.inv(V)
        out[int(u)] = (c, mu, cov, np.linalg.inv(cov))
    return out, min_b, div


def test_leaves_match_numpy(oracle, ndt_small):
    mp = ndt_small["map"][:120_000]
    o = oracle.OracleNdt(resolution=1.0)
    n_valid = o.set_target(mp)
    L = o.leaves()
    ref, min_b, div = numpy_leaves(mp, 1.0)
    gmn, gdv = o.grid()
    assert list(gmn) == list(min_b) and list(gdv) == list(div)
    assert n_valid == len(L["ids"]) == len(ref)
    assert list(L["ids"]) == sorted(ref)
    for k, i in enumerate(L["ids"][::37]):
        k = k * 37
        c, mu, cov, icov = ref[int(i)]
        assert L["npts"][k] == c
        np.testing.assert_allclose(L["mean"][k], mu, rtol=0, atol=1e-9)
        np.testing.assert_allclose(L["cov"][k], cov, rtol=0, atol=1e-7 * max(1.0, np.abs(cov).max()))
        # the inverse is conditioned by the inflated eigenvalue ratio (<= 100): compare through cov * icov = I
        np.testing.assert_allclose(L["cov"][k] @ L["icov"][k], np.eye(3), atol=1e-8)
        w = np.linalg.eigvalsh(0.5 * (L["cov"][k] + L["cov"][k].T))
        assert w[0] >= 0.01 * w[2] * (1 - 1e-6)


def test_min_points_and_resolution(oracle, ndt_small):
    mp = ndt_small["map"][:50_000]
    for res in (0.5, 1.0, 2.0):
        o = oracle.OracleNdt(resolution=res)
        o.set_target(mp)
        L = o.leaves()
        ref, _, _ = numpy_leaves(mp, res)
        assert list(L["ids"]) == sorted(ref)
        assert (L["npts"] >= 6).all()


def test_euler_round_trip(oracle, synth):
    rng = np.random.default_rng(3)
    for _ in range(200):
        p = np.concatenate([rng.uniform(-5, 5, 3), rng.uniform(-np.pi, np.pi, 3)])
        M = oracle.matrix_from_pose(p)
        np.testing.assert_allclose(M.astype(np.float64), synth.pose_vec_to_matrix(p), atol=3e-7 * max(1.0, np.abs(p[:3]).max()))
        rpy = oracle.euler_from_matrix(M)
        assert 0.0 <= rpy[0] <= np.pi + 1e-6  # Eigen: first angle in [0, pi]
        M2 = oracle.matrix_from_pose(np.concatenate([p[:3], rpy.astype(np.float64)]))
        np.testing.assert_allclose(M2[:3, :3], M[:3, :3], atol=2e-6)


def interior_source(o, radius=15.0, n=3000, seed=5):
    """World-frame points scattered within +-0.2 m of the centres of valid voxels near the origin: small pose
    perturbations never move them across a voxel boundary, so the score is differentiable there (in general it is
    only piecewise smooth: the neighbourhood of a point changes when it crosses a boundary)."""
    L = o.leaves()
    mn, dv = o.grid()
    ids = L["ids"]
    ijk = np.stack([ids % dv[0], (ids // dv[0]) % dv[1], ids // (dv[0] * dv[1])], 1) + mn
    centre = (ijk + 0.5) * 1.0
    centre = centre[np.linalg.norm(centre, axis=1) < radius]
    rng = np.random.default_rng(seed)
    pick = centre[rng.integers(0, len(centre), n)]
    return np.ascontiguousarray((pick + rng.uniform(-0.2, 0.2, (n, 3))).astype(np.float32))


def test_gradient_and_hessian_are_derivatives_of_score(oracle, ndt_small):
    o = oracle.OracleNdt(resolution=1.0)
    o.set_target(ndt_small["map"])
    o.set_source(interior_source(o))
    p = np.array([0.05, -0.03, 0.02, 0.001, 0.0, 0.003])  # pitch = 0: sy = 0, the float table's +sy quirk is silent
    s0, g, H = o.derivatives(p)
    assert np.isfinite(g).all() and np.isfinite(H).all() and np.abs(g).max() > 1.0
    Hnum = np.zeros((6, 6))
    for k, h in [(0, 1e-3), (1, 1e-3), (2, 1e-3), (3, 2e-4), (4, 2e-4), (5, 2e-4)]:
        dp = np.zeros(6)
        dp[k] = h
        sp, gp, _ = o.derivatives(p + dp, compute_hessian=False)
        sm, gm, _ = o.derivatives(p - dp, compute_hessian=False)
        num = (sp - sm) / (2 * h)
        assert abs(num - g[k]) <= 2e-3 * abs(g[k]) + 2e-3 * np.abs(g).max() * (1.0 if k >= 3 else 0.05), (k, num, g[k])
        Hnum[:, k] = (gp - gm) / (2 * h)
    scale = np.sqrt(np.outer(np.abs(np.diag(H)), np.abs(np.diag(H))))
    assert (np.abs(Hnum - H) <= 0.02 * scale + 1e-6 * np.abs(H).max()).all(), (Hnum - H) / scale
    np.testing.assert_allclose(H, H.T, rtol=0, atol=1e-5 * np.abs(H).max())


def test_double_hessian_agrees_with_float_path(oracle, ndt_small):
    o = oracle.OracleNdt(resolution=1.0)
    o.set_target(ndt_small["map"])
    o.set_source(ndt_small["scan"])
    p = ndt_small["p_true"] + np.array([0.05, 0.02, -0.01, 0.0, 0.0, 0.02])  # roll = pitch = 0: sy = 0, no table quirk
    _, _, Hf = o.derivatives(p)
    Hd = o.hessian(p)
    np.testing.assert_allclose(Hf, Hd, rtol=0, atol=2e-4 * np.abs(Hd).max())


def test_align_recovers_pose(oracle, synth, ndt_small):
    """Two different starts (6 cm / 0.3 deg and 40 cm / 2 deg off) converge to the same optimum, which lies within
    the voxel-Gaussian bias (< 10 cm for 1 m voxels) of the seeded pose."""
    o = oracle.OracleNdt(resolution=1.0, trans_eps=0.01, step_size=0.1, max_iter=35)
    o.set_target(ndt_small["map"])
    o.set_source(ndt_small["scan"])
    finals = []
    for d in ([0.05, -0.04, 0.02, 0.0, 0.0, np.deg2rad(0.3)], [0.3, -0.25, 0.0, 0.0, 0.0, np.deg2rad(2.0)]):
        p0 = ndt_small["p_true"] + np.array(d)
        rc, T, r = o.align(synth.pose_vec_to_matrix(p0).astype(np.float32))
        assert rc == 0 and r.converged == 1
        assert r.iters >= 2  # at least two Newton steps are always taken
        assert r.evals >= r.iters + 1
        err = np.linalg.inv(ndt_small["T_true"]) @ T.astype(np.float64)
        assert np.linalg.norm(err[:3, 3]) < 0.10
        assert np.arccos(np.clip((np.trace(err[:3, :3]) - 1) / 2, -1, 1)) < np.deg2rad(0.3)
        assert r.trans_probability > 0
        np.testing.assert_allclose(T.astype(np.float64), synth.pose_vec_to_matrix(np.array(r.p_final)), atol=2e-5)
        finals.append(np.array(r.p_final))
    assert np.linalg.norm(finals[0][:3] - finals[1][:3]) < 0.03
    assert r.iters >= 4  # the 0.1 step clamp: a 0.4 m error needs several Newton steps


def test_identity_guess_is_not_pre_applied(oracle, ndt_small):
    """align(out, I): the source is used as is (ndt_omp_impl.hpp:83-88) and p starts at zero."""
    o = oracle.OracleNdt(resolution=1.0)
    o.set_target(ndt_small["map"])
    world_scan = (ndt_small["scan"].astype(np.float64) @ ndt_small["T_true"][:3, :3].T + ndt_small["T_true"][:3, 3]).astype(np.float32)
    o.set_source(world_scan)
    rc, T, r = o.align(np.eye(4, dtype=np.float32))
    assert rc == 0
    assert np.linalg.norm(T[:3, 3]) < 0.10  # same bias as above


def test_score_batch_ranks_true_pose_first(oracle, synth, ndt_small):
    o = oracle.OracleNdt(resolution=1.0)
    o.set_target(ndt_small["map"])
    o.set_source(ndt_small["scan"])
    poses = synth.hypothesis_grid(ndt_small["p_true"], nx=6, ny=6, nyaw=4, pitch=1.0)
    s = o.score_batch(poses)
    best = int(np.argmax(s))
    # the grid is centred so that (ix, iy, k) = (3, 3, 0) is the true pose
    assert best == (3 * 6 + 3) * 4 + 0
    assert np.isfinite(s).all()
