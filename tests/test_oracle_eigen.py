"""The oracle's restatements of the Eigen decompositions the NDT path calls (oracle/smallmat.h), which follow the Eigen
sources vendored in the reference (fast_gicp/thirdparty/Eigen/Eigen/src/{Eigenvalues,SVD,Jacobi,misc}) line by line:
checked against LAPACK (numpy) and against the round-1 substitutes (cyclic Jacobi) they replace."""
import numpy as np
import pytest


def test_selfadjoint3_matches_lapack(oracle):
    rng = np.random.default_rng(5)
    cases = []
    for t in range(3000):
        B = rng.normal(size=(3, 3)) * rng.uniform(1e-3, 1e2)
        cases.append(B @ B.T)
    cases += [np.diag([3.0, 1.0, 2.0]), np.diag([1.0, 1.0, 1.0]), np.zeros((3, 3)), np.diag([0.0, 0.0, 5.0]),
              np.array([[2.0, 1.0, 0.0], [1.0, 2.0, 0.0], [0.0, 0.0, 1.0]]),            # mat(2,0) == 0: tridiagonal already
              np.array([[1e-300, 0, 0], [0, 1e-300, 0], [0, 0, 1e-300]])]
    flat = np.array([0.2, 0.7, 0.1])
    for k in range(200):  # near-planar covariances, the case voxel Gaussians of a surface map produce
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        cases.append(R @ np.diag([1e-4 * rng.uniform(0.5, 2), 0.1 * rng.uniform(0.5, 2), 0.3]) @ R.T)
    for A in cases:
        A = 0.5 * (A + A.T)
        ok, w, V = oracle.eigen_selfadjoint3(A)
        assert ok
        w0 = np.linalg.eigvalsh(A)
        scale = max(np.abs(w0).max(), 1e-300)
        assert np.all(np.diff(w) >= 0)                                    # ascending, like Eigen
        np.testing.assert_allclose(w, w0, rtol=0, atol=4e-15 * scale)
        np.testing.assert_allclose(V @ np.diag(w) @ V.T, A, rtol=0, atol=8e-15 * scale)
        np.testing.assert_allclose(V.T @ V, np.eye(3), rtol=0, atol=8e-15)


def test_selfadjoint3_reads_only_the_lower_triangle(oracle):
    A = np.array([[2.0, 99.0, -99.0], [0.5, 3.0, 99.0], [0.25, -0.75, 1.0]])
    L = np.tril(A) + np.tril(A, -1).T
    _, w, _ = oracle.eigen_selfadjoint3(A)
    np.testing.assert_allclose(w, np.linalg.eigvalsh(L), atol=1e-14)


def test_jacobi_svd_solve6_matches_lapack(oracle):
    rng = np.random.default_rng(6)
    eps = np.finfo(float).eps
    for t in range(1500):
        B = rng.normal(size=(6, 6))
        H = B @ B.T * rng.uniform(1e-3, 1e3) + np.eye(6) * 1e-3
        if t % 4 == 0:
            H = -H                                                            # NDT Hessians are negative definite near the optimum
        if t % 5 == 0:
            H = H + 1e-13 * rng.normal(size=(6, 6))                           # slightly unsymmetric, as accumulated Hessians are
        rhs = rng.normal(size=6)
        x, sv = oracle.jacobi_svd_solve6(H, rhs)
        sv0 = np.linalg.svd(H, compute_uv=False)
        np.testing.assert_allclose(sv, sv0, rtol=0, atol=1e-13 * sv0.max())
        np.testing.assert_allclose(x, np.linalg.solve(H, rhs), rtol=0, atol=1e-11 * (sv0.max() / sv0.min()) * np.abs(x).max())
    # rank deficient: singular values below 6 eps sigma_max are dropped (SVDBase::rank), the rest is a pseudo-inverse
    U = np.linalg.qr(rng.normal(size=(6, 6)))[0]
    s = np.array([5.0, 3.0, 1.0, 0.5, 1e-17, 0.0])
    H = U @ np.diag(s) @ U.T
    rhs = rng.normal(size=6)
    x, sv = oracle.jacobi_svd_solve6(H, rhs)
    x0 = U[:, :4] @ ((U[:, :4].T @ rhs) / s[:4])
    np.testing.assert_allclose(x, x0, atol=1e-12)
    assert np.all(np.diff(sv) <= 0) and sv[4] < 6 * eps * sv[0]
    # non-finite input -> InvalidInput -> NaN step (computeTransformation then returns not converged, ndt_omp_impl.hpp:119-123)
    H[2, 3] = np.nan
    x, _ = oracle.jacobi_svd_solve6(H, rhs)
    assert np.all(np.isnan(x))


def test_leaves_and_align_agree_with_the_round1_substitutes(oracle, synth):
    """Where the restated Eigen path replaced a substitute algorithm (cyclic Jacobi for SelfAdjointEigenSolver and for
    JacobiSVD.solve) the results must agree to rounding: leaves to 1e-12, the alignment to the same iterations.
    Leaf() starts cov_ from the identity (voxel_grid_covariance_omp.h:98-116), which adds (n-1)/n^2 to every eigenvalue, so
    the eigenvalue-inflation branch (vgc_impl:344-357) - the only consumer of the eigenvectors - needs voxels with more than
    ~1000 points: a dense two-wall corner scene."""
    rng = np.random.default_rng(9)
    n = 120_000
    floor = np.c_[rng.uniform(-4, 4, n), rng.uniform(-4, 4, n), rng.normal(0, 0.01, n)]
    wall = np.c_[rng.uniform(-4, 4, n), 4 + rng.normal(0, 0.01, n), rng.uniform(0, 3, n)]
    wall2 = np.c_[-4 + rng.normal(0, 0.01, n // 2), rng.uniform(-4, 4, n // 2), rng.uniform(0, 3, n // 2)]
    mp = np.concatenate([floor, wall, wall2]).astype(np.float32)
    src = mp[rng.choice(len(mp), 4000, replace=False)] + rng.normal(0, 0.01, (4000, 3)).astype(np.float32)
    out = {}
    try:
        for legacy in (1, 0):
            oracle.set_legacy_eigen(legacy)
            o = oracle.OracleNdt(resolution=1.0, trans_eps=0.01)
            o.set_target(mp)
            o.set_source(src)
            p6 = np.array([0.2, -0.15, 0.05, 0.004, -0.003, 0.02])
            rc, T, r = o.align(synth.pose_vec_to_matrix(p6).astype(np.float32))
            out[legacy] = (o.leaves(), r.iters, r.evals, np.array(r.p_final))
    finally:
        oracle.set_legacy_eigen(0)
    (La, ia, ea, pa), (Lb, ib, eb, pb) = out[1], out[0]
    np.testing.assert_array_equal(La["ids"], Lb["ids"])
    w = np.linalg.eigvalsh(Lb["cov"])
    inflated = np.abs(w[:, 0] / w[:, 2] - 0.01) < 1e-9
    assert inflated.sum() >= 20                                              # a good number of leaves took the inflation branch
    assert np.any(La["icov"] != Lb["icov"])                                  # and the two eigen paths do differ in the last bits
    scale = np.abs(La["icov"]).reshape(len(La["ids"]), -1).max(1)[:, None, None]
    assert (np.abs(La["icov"] - Lb["icov"]) / scale).max() <= 1e-12
    assert (ia, ea) == (ib, eb) and ia >= 2
    np.testing.assert_allclose(pa, pb, rtol=0, atol=1e-7)   # float per-point math amplifies the last-bit leaf differences; bar is 1e-4
    assert np.abs(pb[:3]).max() < 0.02 and np.abs(pb[3:]).max() < 0.005      # and the alignment did recover the identity pose
