"""GICP on the device (b200_gicp_*, csrc/gicp.cu) against the oracle restatement of pclomp::GeneralizedIterativeClosestPoint
(oracle/gicp_oracle.cpp), through the C ABI.

Bit-exact: neighbour index lists (ascending distance, ties to the lower index), correspondences, Mahalanobis matrices.
<= 1e-12: regularised covariances (same fp64 operation order on both sides).
<= 1e-9 relative: the cost functor's value and gradient (the device adds the per-point terms in a tree, the oracle serially).
<= 1e-4 m / 1e-4: the converged transformation (BASELINE.json's pose tolerance), with the same number of outer passes.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene(synth):
    world = synth.make_world(synth.SEED, beams=True)
    mp = synth.sample_map(150_000, synth.SEED, world=world)
    p_true = np.array([3.0, -2.0, 1.2, 0.0, 0.0, 0.6])
    T = synth.pose_vec_to_matrix(p_true)
    scan = np.ascontiguousarray(synth.raycast(T[:3, 3], T[:3, :3], synth.livox_dirs(9000, synth.SEED), world, seed=synth.SEED)[:7000])
    guess = synth.pose_vec_to_matrix(p_true + np.array([0.08, -0.05, 0.03, 0.004, -0.003, 0.01]))
    return dict(map=mp, scan=scan, p_true=p_true, T_true=T, guess=guess)


@pytest.fixture(scope="module")
def pair(oracle, api, scene):
    o = oracle.OracleGicp()
    o.set_target(scene["map"])
    o.set_source(scene["scan"])
    g = api.GeneralizedIterativeClosestPoint()
    g.setInputTarget(scene["map"])
    g.setInputSource(scene["scan"])
    yield o, g
    g.close()


def test_neighbours_and_covariances_match_oracle(oracle, pair, scene):
    o, g = pair
    for which, cloud in (("source", scene["scan"]), ("target", scene["map"])):
        cov, knn = g.covariances(which, with_neighbours=True)
        sel = np.arange(0, len(cloud), 7 if which == "target" else 1)
        idx, _ = oracle.exact_knn(cloud, cloud[sel], 20)
        np.testing.assert_array_equal(knn[sel], idx)                 # same 20 neighbours in the same order
        np.testing.assert_allclose(cov, o.covariances(which), rtol=0, atol=1e-12)
        info = g.index_info(which)
        assert len(cloud) / info["occupied"] > 1.0 and info["cells"] <= 2 ** 27


def test_small_and_degenerate_clouds(oracle, api):
    rng = np.random.default_rng(9)
    pts = rng.uniform(-1, 1, (64, 3)).astype(np.float32)
    pts[5] = pts[4]                                                   # duplicates: distance ties go to the lower index
    pts[17] = pts[4]
    g = api.GeneralizedIterativeClosestPoint()
    g.setCorrespondenceRandomness(32)
    g.setInputTarget(pts)
    g.setInputSource(pts[:40])
    cov, knn = g.covariances("target", with_neighbours=True)
    idx, _ = oracle.exact_knn(pts, pts, 32)
    np.testing.assert_array_equal(knn, idx)
    o = oracle.OracleGicp(k_correspondences=32)
    o.set_target(pts)
    o.set_source(pts[:40])
    np.testing.assert_allclose(cov, o.covariances("target"), rtol=0, atol=1e-12)
    # a collinear cloud: every covariance is rank one before the regularisation
    line = np.zeros((50, 3), np.float32)
    line[:, 0] = np.linspace(0, 5, 50)
    g.setCorrespondenceRandomness(20)
    g.setInputTarget(line)
    g.setInputSource(line)
    o = oracle.OracleGicp()
    o.set_target(line)
    o.set_source(line)
    np.testing.assert_allclose(g.covariances("target"), o.covariances("target"), rtol=0, atol=1e-12)
    # fewer points than k_correspondences: the reference refuses (gicp_omp_impl.hpp:54-58)
    g.setInputTarget(pts[:10])
    g.setInputSource(pts[:10])
    with pytest.raises(api.B200Error):
        g.align()
    with pytest.raises(api.B200Error):
        h = api.GeneralizedIterativeClosestPoint()
        h.setCorrespondenceRandomness(33)                              # the neighbour list lives in one warp: k <= 32
        h.setInputTarget(pts)
        h.setInputSource(pts)
        h.align()
    g.close()


def test_correspondences_and_cost_match_oracle(pair, scene):
    o, g = pair
    for trans in (np.eye(4), None):
        if trans is None:   # a second pass: some transformation_ on top of the guess
            trans = np.eye(4)
            trans[:3, 3] = [-0.05, 0.03, -0.02]
            c, s = np.cos(0.004), np.sin(0.004)
            trans[:2, :2] = [[c, -s], [s, c]]
        m0, idx0, maha0, d0 = o.correspondences(trans, scene["guess"])
        m1, idx1, maha1, d1 = g.correspondences(trans, scene["guess"])
        assert m1 == m0 > 0.9 * len(scene["scan"])
        np.testing.assert_array_equal(idx1, idx0)
        hit = idx0 >= 0
        np.testing.assert_array_equal(d1[hit], d0[hit])
        np.testing.assert_array_equal(maha1[hit], maha0[hit])
        for x in (np.zeros(6), np.array([0.01, -0.02, 0.005, 0.002, -0.001, 0.003])):
            f_op0, f_fdf0, g_df0, g_fdf0 = o.cost(x)
            f_op1, f_fdf1, g1, m = g.cost(x)
            assert m == m0
            assert abs(f_op1 - f_op0) <= 1e-9 * abs(f_op0) and abs(f_fdf1 - f_fdf0) <= 1e-9 * abs(f_fdf0)
            np.testing.assert_allclose(g1, g_fdf0, rtol=0, atol=1e-9 * np.abs(g_fdf0).max())


def test_distance_gate(oracle, api, scene):
    """Matches beyond setMaxCorrespondenceDistance are dropped (gicp_omp_impl.hpp:436)."""
    o = oracle.OracleGicp(corr_dist_threshold=0.3)
    o.set_target(scene["map"])
    o.set_source(scene["scan"])
    g = api.GeneralizedIterativeClosestPoint()
    g.setMaxCorrespondenceDistance(0.3)
    g.setInputTarget(scene["map"])
    g.setInputSource(scene["scan"])
    far = scene["guess"].copy()
    far[:3, 3] += [0.25, 0.2, 0.0]
    m0, idx0, _, _ = o.correspondences(np.eye(4), far)
    m1, idx1, _, _ = g.correspondences(np.eye(4), far)
    assert 0 < m0 < len(scene["scan"]) and m1 == m0
    np.testing.assert_array_equal(idx1, idx0)
    g.close()


def test_align_matches_oracle(pair, scene):
    o, g = pair
    rc0, fin0, r0 = o.align(scene["guess"])
    rc1 = g.align(scene["guess"])
    fin1, r1 = g.getFinalTransformation(), g.result
    assert rc1 == rc0 == 0 and g.hasConverged() and r1.converged == r0.converged == 1
    assert r1.iterations == r0.iterations and r1.last_m == r0.last_m
    assert np.abs(fin1 - fin0).max() <= 1e-4                           # north_star tolerance: 1e-4 m / 1e-4 rad
    assert np.abs(fin1[:3, 3] - scene["T_true"][:3, 3]).max() < 0.03
    # the optimiser ran on the device: functor calls were counted there
    assert r1.n_fdf >= r1.iterations and r1.n_f > 0 and r1.inner_total >= r1.iterations
    # getFitnessScore of the aligned scan == mean squared exact-NN distance
    s1 = g.getFitnessScore()
    assert s1 > 0 and g.fitness_in_range == len(scene["scan"])
    # started at the truth it stays there
    assert g.align(scene["T_true"]) == 0
    assert np.abs(g.getFinalTransformation()[:3, 3] - scene["T_true"][:3, 3]).max() < 5e-3


def test_fitness_score_matches_exact_search(oracle, pair, scene):
    o, g = pair
    T = scene["T_true"].astype(np.float32)
    s = g.getFitnessScore(T=T)
    src = scene["scan"]
    moved = np.stack([((T[r, 0] * src[:, 0] + T[r, 1] * src[:, 1]) + T[r, 2] * src[:, 2]) + T[r, 3] for r in range(3)], 1).astype(np.float32)
    _, d2 = oracle.exact_knn(scene["map"], moved, 1)
    ref = d2[:, 0].astype(np.float64)
    assert abs(s - ref.mean()) <= 1e-12 * ref.mean()
    s_r = g.getFitnessScore(max_range=0.01, T=T)
    keep = ref <= 0.01
    assert g.fitness_in_range == int(keep.sum()) and abs(s_r - ref[keep].mean()) <= 1e-12 * ref[keep].mean()


def test_too_few_correspondences_break_the_loop(pair, scene):
    """The optimiser's NotEnoughPointsException ends the loop with converged_ == false and final = guess (gicp_omp_impl.hpp:206-211,496-500,513)."""
    o, g = pair
    far = scene["T_true"].copy()
    far[0, 3] += 500.0
    rc0, fin0, r0 = o.align(far)
    rc1 = g.align(far)
    assert rc1 == rc0 == 2 and not g.hasConverged() and g.result.iterations == r0.iterations == 0
    np.testing.assert_array_equal(g.getFinalTransformation(), fin0)


def test_full_size_pose_recovery(api, synth):
    """20k-point scan against a 1M-point map (the density the estimator is meant for): the seeded pose comes back to millimetres,
    a size-independent property (the oracle's exact search needs seconds per cloud at this size)."""
    world = synth.make_world(synth.SEED, beams=True)
    mp = synth.sample_map(1_000_000, synth.SEED, world=world)
    p_true = np.array([3.0, -2.0, 1.2, 0.0, 0.0, 0.6])
    T = synth.pose_vec_to_matrix(p_true)
    scan = np.ascontiguousarray(synth.raycast(T[:3, 3], T[:3, :3], synth.livox_dirs(24000, synth.SEED), world, seed=synth.SEED)[:20000])
    g = api.GeneralizedIterativeClosestPoint()
    g.setInputTarget(mp)
    g.setInputSource(scan)
    assert g.align(synth.pose_vec_to_matrix(p_true + np.array([0.15, -0.1, 0.05, 0.01, -0.01, 0.03]))) == 0
    fin = g.getFinalTransformation()
    assert np.abs(fin[:3, 3] - T[:3, 3]).max() < 0.01 and np.abs(fin[:3, :3] - T[:3, :3]).max() < 1e-3
    assert g.result.last_m == len(scan)
    g.close()


def test_cuda_gicp_reproduces_golden(api):
    """tests/golden/r02_gicp_small.npz (oracle outputs frozen by tools/make_golden.py gicp): every stage of the device path."""
    import os
    from conftest import ROOT
    gd = np.load(os.path.join(ROOT, "tests", "golden", "r02_gicp_small.npz"))
    g = api.GeneralizedIterativeClosestPoint()
    g.setInputTarget(gd["map"])
    g.setInputSource(gd["scan"])
    cov, knn = g.covariances("source", with_neighbours=True)
    np.testing.assert_array_equal(knn, gd["knn_src"])
    np.testing.assert_allclose(cov, gd["cov_src"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(g.covariances("target")[::16], gd["cov_tgt_every_16"], rtol=0, atol=1e-12)
    m, idx, maha, d2 = g.correspondences(np.eye(4), gd["guess"])
    assert m == int(gd["corr_m"])
    np.testing.assert_array_equal(idx, gd["corr_idx"])
    hit = idx >= 0
    np.testing.assert_array_equal(maha[hit], gd["corr_maha"][hit])
    np.testing.assert_array_equal(d2[hit], gd["corr_d2"][hit])
    f_op, f_fdf, gr, mm = g.cost(gd["cost_x"])
    assert mm == m and abs(f_op - float(gd["cost_f_op"])) <= 1e-9 * abs(float(gd["cost_f_op"])) and abs(f_fdf - float(gd["cost_f_fdf"])) <= 1e-9 * abs(float(gd["cost_f_fdf"]))
    np.testing.assert_allclose(gr, gd["cost_g"], rtol=0, atol=1e-9 * np.abs(gd["cost_g"]).max())
    rc = g.align(gd["guess"])
    assert rc == int(gd["align_rc"]) and g.result.iterations == int(gd["align_iterations"]) and g.result.last_m == int(gd["align_last_m"])
    assert np.abs(g.getFinalTransformation() - gd["align_final"]).max() <= 1e-4
    g.close()
