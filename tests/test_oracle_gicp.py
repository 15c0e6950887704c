"""Self-checks that pin the GICP oracle (the reference ships no vectors for pclomp GICP: parity unpinned), and CPU checks of
the PRODUCT's GICP arithmetic: pointcloud-slam_b200/csrc/gicp_math.cuh is plain __host__ __device__ code, so
tests/helpers/gicp_host_harness.cpp compiles it with g++ and every piece is compared with the oracle here, without a GPU.

oracle:
1. the grid-accelerated exact k-NN == brute force (duplicates, queries outside the cloud, k > points in reach);
2. computeCovariances == a numpy restatement (brute-force neighbours, covariance, SVD, (1, 1, eps) spectrum);
3. applyState == Rz Ry Rx; the functor's gradient == central differences of its value; operator() == fdf's value up to
   float rounding; df == fdf;
4. BFGS reaches the minimum of a smooth test function; align() recovers the seeded pose on a dense scene and stays put
   when started at the truth.
product arithmetic (harness) vs oracle:
5. apply_state, the 3x3 JacobiSVD regularisation, the Mahalanobis matrix, the cost functor, the BFGS minimiser (same
   iterates, same functor-call counts) and the rigid estimate of one pass.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


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
def gicp(oracle, scene):
    g = oracle.OracleGicp()
    g.set_target(scene["map"])
    g.set_source(scene["scan"])
    return g


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    """The product's gicp_math.cuh compiled for the host (same flags as the oracle: no FMA contraction)."""
    so = str(tmp_path_factory.mktemp("hh") / "libgicp_harness.so")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-x", "c++",
           "-I", os.path.join(ROOT, "pointcloud-slam_b200", "csrc"), os.path.join(ROOT, "tests", "helpers", "gicp_host_harness.cpp"), "-o", so]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    L = C.CDLL(so)
    vp, i32, dbl = C.c_void_p, C.c_int32, C.c_double
    L.hh_apply_state.argtypes = [vp, vp]
    L.hh_state_from_transform.argtypes = [vp, vp]
    L.hh_cov_regularize.argtypes = [vp, vp, i32, dbl, vp]
    L.hh_svd3_u.argtypes = [vp, vp, vp]
    L.hh_mahalanobis.argtypes = [vp, vp, vp, vp, vp]
    L.hh_cost.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    L.hh_estimate.restype = i32
    L.hh_estimate.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp]
    L.hh_bfgs_test.restype = i32
    L.hh_bfgs_test.argtypes = [vp, i32, vp, vp]
    L.hh_transform_delta.restype = dbl
    L.hh_transform_delta.argtypes = [vp, vp, dbl, dbl]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def rows12(T44):
    return np.ascontiguousarray(np.asarray(T44, dtype=np.float32)[:3, :].reshape(-1))


# ------------------------------------------------------------------ oracle self-checks
def test_exact_knn_equals_brute_force(oracle):
    rng = np.random.default_rng(3)
    pts = np.concatenate([rng.uniform(-20, 20, (6000, 2)), rng.normal(0, 0.02, (6000, 1))], 1).astype(np.float32)   # a sheet
    pts = np.concatenate([pts, rng.uniform(-5, 5, (3000, 3)).astype(np.float32), pts[:200]])                        # a blob and duplicates
    q = np.concatenate([pts[::37], rng.uniform(-30, 30, (300, 3)).astype(np.float32), np.array([[100.0, 0, 0], [-70.0, 55.0, 9.0]], np.float32)])
    for k in (1, 5, 20):
        i0, d0 = oracle.exact_knn(pts, q, k, brute=True)
        i1, d1 = oracle.exact_knn(pts, q, k, brute=False)
        np.testing.assert_array_equal(i1, i0)
        np.testing.assert_array_equal(d1, d0)
    # more neighbours asked for than the cloud holds
    i0, d0 = oracle.exact_knn(pts[:7], q[:5], 12, brute=False)
    assert (i0[:, 7:] == -1).all() and (np.sort(i0[:, :7], 1) == np.arange(7)).all()


def test_covariances_match_numpy(oracle):
    rng = np.random.default_rng(5)
    pts = np.concatenate([rng.uniform(-3, 3, (1500, 2)), rng.normal(0, 0.01, (1500, 1))], 1).astype(np.float32)
    g = oracle.OracleGicp(k_correspondences=20, gicp_epsilon=0.001)
    g.set_target(pts)
    g.set_source(pts[:40])
    cov = g.covariances("target")
    P = pts.astype(np.float64)
    for i in range(0, 1500, 97):
        d2 = ((pts - pts[i]) ** 2).sum(1)
        nn = np.lexsort((np.arange(len(pts)), d2))[:20]
        c = np.cov(P[nn].T, bias=True)
        U, s, _ = np.linalg.svd(c)
        ref = U @ np.diag([1.0, 1.0, 0.001]) @ U.T
        np.testing.assert_allclose(cov[i], ref, atol=1e-6)
        w = np.linalg.eigvalsh(cov[i])
        np.testing.assert_allclose(w, [0.001, 1.0, 1.0], atol=1e-9)
        # the small direction is the sheet's normal
        n = np.linalg.eigh(cov[i])[1][:, 0]
        assert abs(n[2]) > 0.99
    with pytest.raises(ValueError):
        h = oracle.OracleGicp(k_correspondences=20)
        h.set_target(pts[:10])
        h.set_source(pts[:10])
        h.covariances("target")


def test_apply_state_and_gradient(oracle, gicp, scene):
    x = np.array([0.3, -0.2, 0.1, 0.05, -0.07, 0.4])
    T = oracle.gicp_apply_state(x)
    cr, sr, cp, sp, cy, sy = np.cos(x[3]), np.sin(x[3]), np.cos(x[4]), np.sin(x[4]), np.cos(x[5]), np.sin(x[5])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    np.testing.assert_allclose(T[:3, :3], Rz @ Ry @ Rx, atol=2e-7)
    np.testing.assert_allclose(T[:3, 3], x[:3], atol=1e-7)
    m, idx, maha, d2 = gicp.correspondences(np.eye(4), scene["guess"])
    assert m > 0.9 * len(scene["scan"]) and (idx >= 0).sum() == m
    assert np.abs(maha - maha.transpose(0, 2, 1)).max() < 1e-3 * np.abs(maha).max()
    x = np.array([0.01, -0.02, 0.005, 0.002, -0.001, 0.003])
    f_op, f_fdf, g_df, g_fdf = gicp.cost(x)
    assert abs(f_op - f_fdf) < 1e-5 * abs(f_fdf)            # operator() multiplies in float, fdf in double
    np.testing.assert_allclose(g_df, g_fdf, rtol=1e-12, atol=1e-12)
    num = np.zeros(6)
    for i in range(6):
        h = 1e-4
        xp, xm = x.copy(), x.copy()
        xp[i] += h
        xm[i] -= h
        num[i] = (gicp.cost(xp)[1] - gicp.cost(xm)[1]) / (2 * h)
    np.testing.assert_allclose(g_fdf, num, rtol=2e-3, atol=2e-3 * np.abs(num).max())


def test_bfgs_reaches_the_minimum(oracle):
    st, x, inner, calls = oracle.bfgs_test(np.zeros(6), max_inner=50)
    assert st == 0 and 2 <= inner < 50
    np.testing.assert_allclose(x, [0.3, -0.2, 0.5, 0.05, -0.04, 0.08], atol=5e-3)   # gradient tolerance 1e-2
    st2, x2, inner2, _ = oracle.bfgs_test(np.zeros(6), max_inner=2)
    assert inner2 == 2 and st2 == -1                                               # Running: stopped by the iteration cap


def test_align_recovers_the_seeded_pose(oracle, gicp, scene):
    rc, fin, r = gicp.align(scene["guess"])
    assert rc == 0 and r.converged == 1 and 1 <= r.iterations < 50 and r.last_m > 0.9 * len(scene["scan"])
    # GICP's own stopping rule is loose (a pass that moves < 0.5 mm / 2e-3 ends the loop, BFGS stops at |g| < 1e-2) and the sparse
    # test scene has few returns on the walls that fix x: centimetres here, 2 mm on the bench's 20k-point scan / 1M-point map
    assert np.abs(fin[:3, 3] - scene["T_true"][:3, 3]).max() < 0.03 < 0.08
    assert np.abs(fin[:3, :3] - scene["T_true"][:3, :3]).max() < 2e-3
    rc, fin2, r2 = gicp.align(scene["T_true"])
    assert rc == 0 and np.abs(fin2[:3, 3] - scene["T_true"][:3, 3]).max() < 5e-3
    # too few correspondences: the optimiser throws, the loop breaks, converged_ stays false
    far = scene["T_true"].copy()
    far[0, 3] += 500.0
    rc, fin3, r3 = gicp.align(far)
    assert rc == 2 and r3.converged == 0 and r3.iterations == 0
    np.testing.assert_allclose(fin3, far.astype(np.float32), atol=1e-6)          # final = Identity * guess


# ------------------------------------------------------------------ product arithmetic (gicp_math.cuh on the host) vs oracle
def test_product_apply_state_matches_oracle(oracle, harness):
    rng = np.random.default_rng(11)
    for _ in range(200):
        x = np.concatenate([rng.uniform(-50, 50, 3), rng.uniform(-np.pi, np.pi, 3)])
        T12 = np.zeros(12, np.float32)
        harness.hh_apply_state(_p(x), _p(T12))
        np.testing.assert_array_equal(T12.reshape(3, 4), oracle.gicp_apply_state(x)[:3])
        x6 = np.zeros(6)
        harness.hh_state_from_transform(_p(T12), _p(x6))
        if abs(x[4]) < 1.5:   # away from the gimbal lock the ZYX angles come back
            np.testing.assert_allclose(np.cos(x6[3:] - x[3:]), 1.0, atol=1e-5)


def test_product_covariance_regularisation_matches_oracle(oracle, harness):
    rng = np.random.default_rng(7)
    pts = np.concatenate([rng.uniform(-3, 3, (900, 2)), rng.normal(0, 0.01, (900, 1))], 1).astype(np.float32)
    pts[:300] = rng.uniform(-2, 2, (300, 3)).astype(np.float32)   # some volume-like neighbourhoods too
    g = oracle.OracleGicp()
    g.set_target(pts)
    g.set_source(pts[:30])
    cov = g.covariances("target")
    idx, _ = oracle.exact_knn(pts, pts, 20)
    out = np.zeros(9)
    for i in range(0, 900, 13):
        nb = pts[idx[i]]
        mean = np.zeros(3)
        c6 = np.zeros(6)
        for p in nb:      # the reference's order: float products, fp64 accumulation
            mean += p.astype(np.float64)
            c6 += np.array([p[0] * p[0], p[1] * p[0], p[1] * p[1], p[2] * p[0], p[2] * p[1], p[2] * p[2]], np.float32).astype(np.float64)
        harness.hh_cov_regularize(_p(mean), _p(c6), 20, 0.001, _p(out))
        np.testing.assert_allclose(out.reshape(3, 3), cov[i], rtol=0, atol=1e-13)


def test_product_cost_mahalanobis_and_estimate_match_oracle(oracle, harness, gicp, scene, synth):
    guess = scene["guess"]
    m, idx, maha, d2 = gicp.correspondences(np.eye(4), guess)
    G = guess.astype(np.float32)
    src = scene["scan"]
    out = np.stack([((G[r, 0] * src[:, 0] + G[r, 1] * src[:, 1]) + G[r, 2] * src[:, 2]) + G[r, 3] for r in range(3)], 1).astype(np.float32)
    tgt = np.ascontiguousarray(scene["map"][np.maximum(idx, 0)])
    out = np.ascontiguousarray(out)
    M = np.ascontiguousarray(maha.reshape(-1, 9))
    # Mahalanobis matrices from the oracle's covariances
    C1, C2 = gicp.covariances("source"), gicp.covariances("target")
    T12, G12 = rows12(np.eye(4)), rows12(guess)
    M9 = np.zeros(9, np.float32)
    for i in np.flatnonzero(idx >= 0)[::211]:
        harness.hh_mahalanobis(_p(T12), _p(G12), _p(np.ascontiguousarray(C1[i])), _p(np.ascontiguousarray(C2[idx[i]])), _p(M9))
        np.testing.assert_array_equal(M9, M[i])
    # the functor: same serial order -> same bits
    for x in (np.zeros(6), np.array([0.01, -0.02, 0.005, 0.002, -0.001, 0.003])):
        f0, f1 = C.c_double(), C.c_double()
        g0, g1 = np.zeros(6), np.zeros(6)
        harness.hh_cost(_p(out), _p(tgt), _p(M), _p(idx), len(out), _p(x), C.byref(f0), C.byref(f1), _p(g0), _p(g1))
        o0, o1, og0, og1 = gicp.cost(x)
        assert f0.value == o0 and f1.value == o1
        np.testing.assert_array_equal(g0, og0)
        np.testing.assert_array_equal(g1, og1)
    # one rigid estimate (the BFGS loop of a pass)
    T = rows12(np.eye(4)).copy()
    inner, calls = C.c_int32(), np.zeros(3, np.int32)
    st = harness.hh_estimate(_p(out), _p(tgt), _p(M), _p(idx), len(out), 20, _p(T), C.byref(inner), _p(calls))
    rc, To, inner_o, st_o, calls_o = gicp.estimate(np.eye(4))
    assert rc == 0 and st == st_o and inner.value == inner_o
    np.testing.assert_array_equal(calls, calls_o)
    np.testing.assert_array_equal(T.reshape(3, 4), To[:3])
    d = harness.hh_transform_delta(_p(rows12(np.eye(4))), _p(T), 2e-3, 5e-4)
    To64 = (To - np.eye(4, dtype=np.float32)).astype(np.float64)   # the reference subtracts in float, then scales in double
    ref = max(np.abs(To64[:3, :3]).max() / 2e-3, np.abs(To64[:3, 3]).max() / 5e-4)
    assert abs(d - ref) <= 1e-12 * ref


def test_product_bfgs_matches_oracle(oracle, harness):
    for x0 in (np.zeros(6), np.array([1.0, -1.0, 0.5, 0.3, 0.2, -0.4])):
        for cap in (3, 20, 60):
            st_o, x_o, inner_o, calls_o = oracle.bfgs_test(x0, max_inner=cap)
            x = x0.copy()
            inner, calls = C.c_int32(), np.zeros(3, np.int32)
            st = harness.hh_bfgs_test(_p(x), cap, C.byref(inner), _p(calls))
            assert st == st_o and inner.value == inner_o
            np.testing.assert_array_equal(calls, calls_o)
            np.testing.assert_allclose(x, x_o, rtol=0, atol=1e-15)


def test_product_svd3_against_numpy(harness):
    """jacobi_svd3_u (the JacobiSVD<Matrix3d>(cov, ComputeFullU) of computeCovariances): U orthonormal, singular values descending
    and equal to numpy's, A == U diag(sv) U^T for symmetric positive semi-definite input - full rank, rank 2 (a plane), rank 1 (a line)
    and the zero matrix."""
    rng = np.random.default_rng(21)
    U, sv = np.zeros(9), np.zeros(3)
    for case in range(300):
        rank = (3, 3, 2, 1)[case % 4]
        B = rng.normal(0, 1, (3, rank)) * rng.uniform(1e-3, 10.0)
        A = np.ascontiguousarray(B @ B.T)
        harness.hh_svd3_u(_p(A), _p(U), _p(sv))
        Um = U.reshape(3, 3)
        scale = max(np.abs(A).max(), 1e-300)
        np.testing.assert_allclose(Um.T @ Um, np.eye(3), atol=1e-12)
        assert sv[0] >= sv[1] >= sv[2] >= 0
        np.testing.assert_allclose(sv, np.linalg.svd(A, compute_uv=False), atol=1e-12 * scale)
        np.testing.assert_allclose(Um @ np.diag(sv) @ Um.T, A, atol=1e-12 * scale)
    Z = np.zeros(9)
    harness.hh_svd3_u(_p(Z), _p(U), _p(sv))
    np.testing.assert_array_equal(U.reshape(3, 3), np.eye(3))
    np.testing.assert_array_equal(sv, 0.0)
