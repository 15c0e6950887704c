"""Parity of the CUDA NDT path (through the C ABI) against the CPU oracle.

Bars: voxel membership / point counts exact, leaf mean / covariance / inverse covariance bit-exact (the device
accumulates each leaf in input order with the same fp64 op sequence); score / gradient / Hessian within 1e-6
relative (north_star "H/b within 1e-6"); align() poses within 1e-4 m / 1e-4 rad with the same iteration and
evaluation counts; calculateScore within 1e-12 relative and the same argmax.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

H_TOL = 1e-6
POSE_TOL = 1e-4


@pytest.fixture(scope="module")
def ndt_small(synth):
    world = synth.make_world(synth.SEED, beams=True)
    mp = synth.sample_map(300_000, synth.SEED, world=world)
    p_true = np.array([3.0, -2.0, 1.2, 0.0, 0.0, 0.6])
    T = synth.pose_vec_to_matrix(p_true)
    dirs = synth.livox_dirs(5000, synth.SEED)
    scan = np.ascontiguousarray(synth.raycast(T[:3, 3], T[:3, :3], dirs, world, seed=synth.SEED)[:4000])
    return dict(map=mp, scan=scan, p_true=p_true, T_true=T)


def pair(oracle, api, cfg, **kw):
    o = oracle.OracleNdt(**kw)
    o.set_target(cfg["map"])
    o.set_source(cfg["scan"])
    g = api.NormalDistributionsTransform()
    g.setResolution(kw.get("resolution", 1.0))
    g.setTransformationEpsilon(kw.get("trans_eps", 0.01))
    g.setStepSize(kw.get("step_size", 0.1))
    g.setMaximumIterations(kw.get("max_iter", 35))
    g.setNeighborhoodSearchMethod(kw.get("search", 7))
    g.setInputTarget(cfg["map"])
    g.setInputSource(cfg["scan"])
    return o, g


def relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("res", [0.5, 1.0, 2.0])
def test_voxel_leaves_bit_exact(oracle, api, ndt_small, res):
    o, g = pair(oracle, api, ndt_small, resolution=res)
    Lo, Lg = o.leaves(), g.leaves()
    assert g.numVoxels() == len(Lo["ids"])
    gm, gd = g.grid()
    om, od = o.grid()
    assert list(gm) == list(om) and list(gd) == list(od)
    np.testing.assert_array_equal(Lg["ids"], Lo["ids"])
    np.testing.assert_array_equal(Lg["npts"], Lo["npts"])
    np.testing.assert_array_equal(Lg["mean"], Lo["mean"])
    np.testing.assert_array_equal(Lg["cov"], Lo["cov"])
    np.testing.assert_array_equal(Lg["icov"], Lo["icov"])


def test_target_with_nonfinite_and_sparse_points(oracle, api, ndt_small):
    mp = ndt_small["map"][:40_000].copy()
    mp[5] = [np.nan, 0, 0]
    mp[77] = [0, np.inf, 0]
    cfg = dict(map=mp, scan=ndt_small["scan"])
    o, g = pair(oracle, api, cfg, resolution=1.0)   # 40k points: most voxels hold fewer than 6 points
    Lo, Lg = o.leaves(), g.leaves()
    np.testing.assert_array_equal(Lg["ids"], Lo["ids"])
    np.testing.assert_array_equal(Lg["icov"], Lo["icov"])


@pytest.mark.parametrize("search", [1, 7, 27])
def test_derivatives_parity(oracle, api, ndt_small, search):
    o, g = pair(oracle, api, ndt_small, search=search)
    for d in ([0, 0, 0, 0, 0, 0], [0.12, -0.08, 0.03, 0.004, -0.006, 0.01], [0.4, 0.3, -0.1, 0.02, 0.03, -0.05]):
        p = ndt_small["p_true"] + np.array(d)
        s0, g0, H0 = o.derivatives(p)
        s1, g1, H1 = g.computeDerivatives(p)
        assert abs(s1 - s0) <= H_TOL * abs(s0)
        assert relerr(g1, g0) <= H_TOL
        assert relerr(H1, H0) <= H_TOL
        assert g.nbhd_total(p) == o.nbhd_total(p)


def test_double_hessian_parity(oracle, api, ndt_small):
    o, g = pair(oracle, api, ndt_small)
    p = ndt_small["p_true"] + np.array([0.05, 0.02, -0.01, 0.003, -0.002, 0.02])
    H0 = o.hessian(p)
    H1 = g.computeHessian(p)
    assert relerr(H1, H0) <= 1e-10


@pytest.mark.parametrize("start", [[0.05, -0.04, 0.02, 0.0, 0.0, 0.005], [0.3, -0.25, 0.0, 0.0, 0.0, 0.035], [0, 0, 0, 0, 0, 0]])
def test_align_parity(oracle, api, synth, ndt_small, start):
    o, g = pair(oracle, api, ndt_small)
    guess = synth.pose_vec_to_matrix(ndt_small["p_true"] + np.array(start)).astype(np.float32)
    rc0, T0, r0 = o.align(guess)
    rc1 = g.align(guess)
    r1 = g.result
    assert rc1 == rc0 and r1.converged == r0.converged
    assert (r1.iters, r1.evals, r1.hess_evals) == (r0.iters, r0.evals, r0.hess_evals)
    dp = np.array(r1.p_final) - np.array(r0.p_final)
    assert np.abs(dp[:3]).max() < POSE_TOL and np.abs(dp[3:]).max() < POSE_TOL
    np.testing.assert_allclose(g.getFinalTransformation(), T0, atol=POSE_TOL)
    assert abs(r1.trans_probability - r0.trans_probability) <= 1e-6 * abs(r0.trans_probability)
    assert relerr(np.array(r1.hessian), np.array(r0.hessian)) <= 1e-5
    assert g.hasConverged() and g.getFinalNumIteration() == r0.iters


def test_align_identity_guess(oracle, api, ndt_small):
    T = ndt_small["T_true"]
    ws = (ndt_small["scan"].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    cfg = dict(map=ndt_small["map"], scan=ws)
    o, g = pair(oracle, api, cfg)
    rc0, T0, r0 = o.align(np.eye(4, dtype=np.float32))
    rc1 = g.align(None)
    assert rc1 == rc0 and g.result.iters == r0.iters
    np.testing.assert_allclose(g.getFinalTransformation(), T0, atol=POSE_TOL)


def test_align_batch_equals_single(api, synth, ndt_small):
    g = api.NormalDistributionsTransform()
    g.setTransformationEpsilon(0.01)
    g.setInputTarget(ndt_small["map"])
    g.setInputSource(ndt_small["scan"])
    starts = [[0.05, -0.04, 0.02, 0, 0, 0.005], [0.3, -0.25, 0, 0, 0, 0.035], [-0.2, 0.1, 0.05, 0, 0, -0.02], [0, 0, 0, 0, 0, 0]]
    guesses = [synth.pose_vec_to_matrix(ndt_small["p_true"] + np.array(s)).astype(np.float32) for s in starts]
    finals, res = g.alignBatch(np.stack([m.T.reshape(16) for m in guesses]))
    for k, m in enumerate(guesses):
        g.align(m)
        assert (res[k].iters, res[k].evals, res[k].converged) == (g.result.iters, g.result.evals, g.result.converged)
        np.testing.assert_allclose(finals[k], g.getFinalTransformation(), atol=1e-6)


def test_score_batch_and_argmax(oracle, api, synth, ndt_small):
    o, g = pair(oracle, api, ndt_small)
    poses = synth.hypothesis_grid(ndt_small["p_true"], nx=6, ny=6, nyaw=4, pitch=1.0)
    s0 = o.score_batch(poses)
    s1 = g.calculateScore(poses)
    np.testing.assert_allclose(s1, s0, rtol=1e-12, atol=1e-14)
    assert int(np.argmax(s1)) == int(np.argmax(s0)) == (3 * 6 + 3) * 4
    best, score, ms = api.relocalize(g, poses)
    assert best == int(np.argmax(s0))
    assert score == s1[best]
    # sharded the way the multi-GPU path does it: the winner of the slices is the global winner
    wins = []
    for r in range(3):
        b, e = api.shard_range(len(poses), 3, r)
        wins.append(api.relocalize(g, poses[b:e], h_begin=b)[:2])
    assert max(wins, key=lambda w: (w[1], -w[0]))[0] == best
    wins = [api.relocalize(g, poses[r::3], h_begin=r, h_stride=3)[:2] for r in range(3)]   # interleaved slices
    assert max(wins, key=lambda w: (w[1], -w[0]))[0] == best


def test_ndt_errors(api, ndt_small):
    g = api.NormalDistributionsTransform()
    g.setInputSource(ndt_small["scan"])
    with pytest.raises(api.B200Error):
        g.computeDerivatives(np.zeros(6))      # no target yet
    big = np.array([[0, 0, 0], [1e9, 1e9, 1e9]], np.float32)
    g2 = api.NormalDistributionsTransform()
    g2.setResolution(0.01)
    g2.setInputTarget(big)
    with pytest.raises(api.B200Error):
        g2.numVoxels()                          # leaf indices would overflow (applyFilter's guard)


def test_fitness_score_exact_nn(oracle, api, synth, ndt_small):
    """getFitnessScore = mean squared exact nearest-neighbour distance (PCL kd-tree semantics), also for points that fall
    outside the map and for a max_range cut."""
    cfg = dict(map=ndt_small["map"][:80_000], scan=ndt_small["scan"][:1500])
    o, g = pair(oracle, api, cfg)
    for d in ([0, 0, 0, 0, 0, 0], [0.4, -0.3, 0.1, 0.01, 0.02, 0.05], [30.0, 5.0, 2.0, 0, 0, 1.0], [150.0, 100.0, 20.0, 0, 0, 0]):
        T = synth.pose_vec_to_matrix(ndt_small["p_true"] + np.array(d)).astype(np.float32)
        s0, n0 = o.fitness(T)
        s1 = g.getFitnessScore(T=T)
        assert g.fitness_in_range == n0
        assert abs(s1 - s0) <= 1e-12 * abs(s0)
        s0, n0 = o.fitness(T, max_range=0.01)
        s1 = g.getFitnessScore(max_range=0.01, T=T)
        assert g.fitness_in_range == n0 and (abs(s1 - s0) <= 1e-12 * abs(s0) or (n0 == 0 and s1 == s0))
    rc0, T0, r0 = o.align(synth.pose_vec_to_matrix(ndt_small["p_true"]).astype(np.float32))
    g.align(synth.pose_vec_to_matrix(ndt_small["p_true"]).astype(np.float32))
    assert abs(g.getFitnessScore() - o.fitness(T0)[0]) <= 1e-6 * o.fitness(T0)[0]


def test_full_size_properties(oracle, api, synth):
    """BASELINE.json configs[1] / [3] at full size (10M-point prior map, 20k-point scan): size-independent properties.
    The voxel build is deterministic and accounts for every point; align() from the perturbed guess ends at the same
    optimum as the oracle; the hypothesis grid's winner is the true pose and survives any sharding."""
    cfg = synth.config2()
    g = api.NormalDistributionsTransform()
    g.setTransformationEpsilon(0.01)
    g.setInputTarget(cfg["map"])
    g.setInputSource(cfg["scan"])
    L1 = g.leaves()
    g.setInputTarget(cfg["map"])                    # rebuild: bit-identical (sorted, input-order sums)
    L2 = g.leaves()
    for k in ("ids", "npts", "mean", "cov", "icov"):
        np.testing.assert_array_equal(L1[k], L2[k])
    assert (np.diff(L1["ids"]) > 0).all() and L1["npts"].min() >= 6 and L1["npts"].sum() <= len(cfg["map"])
    # every valid leaf: icov * cov = I, eigenvalue ratio floor 0.01
    sel = slice(None, None, 997)
    prod = np.einsum("nij,njk->nik", L1["cov"][sel], L1["icov"][sel])
    np.testing.assert_allclose(prod, np.broadcast_to(np.eye(3), prod.shape), atol=1e-6)
    w = np.linalg.eigvalsh(0.5 * (L1["cov"][sel] + np.transpose(L1["cov"][sel], (0, 2, 1))))
    assert (w[:, 0] >= 0.01 * w[:, 2] * (1 - 1e-6)).all()
    # the mean of a leaf lies inside its voxel
    mn, dv = g.grid()
    ids = L1["ids"][sel]
    ijk = np.stack([ids % dv[0], (ids // dv[0]) % dv[1], ids // (dv[0] * dv[1])], 1) + mn
    assert ((L1["mean"][sel] >= ijk - 1e-4) & (L1["mean"][sel] <= ijk + 1 + 1e-4)).all()
    # align: oracle parity at full size
    o = oracle.OracleNdt(resolution=1.0, trans_eps=0.01)
    o.set_target(cfg["map"])
    o.set_source(cfg["scan"])
    rc0, T0, r0 = o.align(cfg["guess"])
    rc1 = g.align(cfg["guess"])
    assert rc1 == rc0 and (g.result.iters, g.result.evals) == (r0.iters, r0.evals)
    assert np.abs(np.array(g.result.p_final) - np.array(r0.p_final)).max() < POSE_TOL
    # relocalization: the true pose wins, however the grid is cut
    poses = synth.hypothesis_grid(cfg["p_true"], 16, 16, 4, 1.0)
    best, score, _ = api.relocalize(g, poses)
    assert best == (8 * 16 + 8) * 4
    wins = [api.relocalize(g, poses[r::5], h_begin=r, h_stride=5)[:2] for r in range(5)]
    assert max(wins, key=lambda w: (w[1], -w[0])) == (best, score)
    fit = g.getFitnessScore()
    assert 0 < fit < 0.05 and g.fitness_in_range == len(cfg["scan"])


def dense_corner_scene(seed=9, n=120_000):
    """Two walls and a floor sampled so densely that voxels hold > 1000 points: Leaf() starts cov_ from the identity, so only
    such voxels reach the eigenvalue-inflation branch (vgc_impl:344-357), the one consumer of SelfAdjointEigenSolver's vectors."""
    rng = np.random.default_rng(seed)
    floor = np.c_[rng.uniform(-4, 4, n), rng.uniform(-4, 4, n), rng.normal(0, 0.01, n)]
    wall = np.c_[rng.uniform(-4, 4, n), 4 + rng.normal(0, 0.01, n), rng.uniform(0, 3, n)]
    wall2 = np.c_[-4 + rng.normal(0, 0.01, n // 2), rng.uniform(-4, 4, n // 2), rng.uniform(0, 3, n // 2)]
    mp = np.concatenate([floor, wall, wall2]).astype(np.float32)
    src = mp[rng.choice(len(mp), 4000, replace=False)] + rng.normal(0, 0.01, (4000, 3)).astype(np.float32)
    return dict(map=mp, scan=np.ascontiguousarray(src))


def test_inflated_leaves_bit_exact(oracle, api):
    """The device runs Eigen's tridiagonalisation + implicit-QR eigen solver op for op (ndt.cuh eigen_selfadjoint3), so the
    rebuilt covariances V L V^-1 of inflated leaves are bit-identical to the oracle's."""
    cfg = dense_corner_scene()
    o, g = pair(oracle, api, cfg, resolution=1.0)
    Lo, Lg = o.leaves(), g.leaves()
    w = np.linalg.eigvalsh(Lo["cov"])
    assert (np.abs(w[:, 0] / w[:, 2] - 0.01) < 1e-9).sum() >= 20
    np.testing.assert_array_equal(Lg["ids"], Lo["ids"])
    np.testing.assert_array_equal(Lg["cov"], Lo["cov"])
    np.testing.assert_array_equal(Lg["icov"], Lo["icov"])


def test_newton_direction_parity(oracle, api):
    """JacobiSVD(H).solve(-g) (ndt_omp_impl.hpp:112-114): the device's elimination shortcut for well-conditioned Hessians and
    its literal two-sided Jacobi SVD (taken for rank-deficient / non-finite ones) against the oracle's restatement of Eigen."""
    g = api.NormalDistributionsTransform()
    g._handle()
    rng = np.random.default_rng(12)
    for t in range(40):
        B = rng.normal(size=(6, 6))
        H = -(B @ B.T) * rng.uniform(1e2, 1e6) - np.eye(6)
        rhs = rng.normal(size=6) * 1e3
        x0, sv = oracle.jacobi_svd_solve6(H, rhs)
        x1, path = g.newtonDirection(H, rhs)
        assert path == 0
        assert np.abs(x1 - x0).max() <= 1e-9 * (sv[0] / sv[-1]) * np.abs(x0).max()
        x2, path = g.newtonDirection(H, rhs, force_svd=True)
        assert path == 1
        np.testing.assert_array_equal(x2, x0)          # same algorithm, same op sequence: bit-identical
    # rank deficient (a floor-only scene leaves x, y, yaw unobservable): singular values below 6 eps sigma_max are dropped
    U = np.linalg.qr(rng.normal(size=(6, 6)))[0]
    for s in ([5e6, 3e5, 1e4, 0.0, 0.0, 0.0], [1.0, 0.5, 0.25, 1e-17, 1e-18, 0.0], [0.0] * 6):
        H = -(U @ np.diag(s) @ U.T)
        rhs = rng.normal(size=6)
        x0, sv = oracle.jacobi_svd_solve6(H, rhs)
        x1, path = g.newtonDirection(H, rhs)
        assert path == 1
        np.testing.assert_array_equal(x1, x0)
    H = np.eye(6)
    H[1, 2] = np.nan
    x1, path = g.newtonDirection(H, np.ones(6))
    assert path == 1 and np.all(np.isnan(x1))           # NaN step: computeTransformation returns not converged (:119-123)


def test_align_with_no_overlap_stops_like_the_reference(oracle, api, synth, ndt_small):
    """No source point lands in a valid voxel: score, gradient and Hessian are zero, JacobiSVD.solve returns a zero step and
    computeTransformation leaves at once with converged_ = true (ndt_omp_impl.hpp:119-123)."""
    cfg = dict(map=ndt_small["map"], scan=ndt_small["scan"] + np.float32(500.0))
    o, g = pair(oracle, api, cfg)
    guess = np.eye(4, dtype=np.float32)
    rc0, T0, r0 = o.align(guess)
    rc1 = g.align(guess)
    assert rc1 == rc0 and (g.result.iters, g.result.evals, g.result.converged) == (r0.iters, r0.evals, r0.converged)
    np.testing.assert_allclose(g.getFinalTransformation(), T0, atol=1e-6)


def test_kdtree_neighbourhood_mode(oracle, api, synth, ndt_small):
    """KDTREE search method (ndt_omp_impl.hpp:216-219): radiusSearch(point, resolution) over the fp32 leaf centroids.  Same
    neighbourhood sizes as the oracle, derivatives within 1e-6, the hits are a subset of the DIRECT26 block, and align agrees."""
    o, g = pair(oracle, api, ndt_small, search=0)
    o26, g26 = pair(oracle, api, ndt_small, search=27)
    for d in ([0, 0, 0, 0, 0, 0], [0.12, -0.08, 0.03, 0.004, -0.006, 0.01]):
        p = ndt_small["p_true"] + np.array(d)
        s0, g0, H0 = o.derivatives(p)
        s1, g1, H1 = g.computeDerivatives(p)
        assert abs(s1 - s0) <= H_TOL * abs(s0) and relerr(g1, g0) <= H_TOL and relerr(H1, H0) <= H_TOL
        n_kd, n_26 = g.nbhd_total(p), g26.nbhd_total(p)
        assert n_kd == o.nbhd_total(p) and 0 < n_kd < n_26
    guess = synth.pose_vec_to_matrix(ndt_small["p_true"] + np.array([0.05, -0.04, 0.02, 0.0, 0.0, 0.005])).astype(np.float32)
    rc0, T0, r0 = o.align(guess)
    rc1 = g.align(guess)
    assert rc1 == rc0 and (g.result.iters, g.result.evals) == (r0.iters, r0.evals)
    assert np.abs(np.array(g.result.p_final) - np.array(r0.p_final)).max() < POSE_TOL
    poses = synth.hypothesis_grid(ndt_small["p_true"], 3, 3, 2, 1.0)
    np.testing.assert_allclose(g.calculateScore(poses), o.score_batch(poses), rtol=1e-12)


def test_get_max_eigen(oracle, api, synth, ndt_small):
    """getMaxEigen (ndt_omp.h:209-223): max eigenvalue of the final Hessian / 100000 (LAPACK's general eigen solver as the check)."""
    o, g = pair(oracle, api, ndt_small)
    guess = synth.pose_vec_to_matrix(ndt_small["p_true"] + np.array([0.05, -0.04, 0.02, 0.0, 0.0, 0.005])).astype(np.float32)
    g.align(guess)
    H = np.array(g.result.hessian).reshape(6, 6)
    want = np.linalg.eigvals(H).real.max() / 100000.0
    assert abs(g.getMaxEigen() - want) <= 1e-9 * abs(want)
