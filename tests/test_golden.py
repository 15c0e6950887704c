"""tests/golden/r01_small.npz (made by tools/make_golden.py): the oracle reproduces it on CPU, the CUDA path on GPU.
These vectors freeze our own restatement against drift; they are not reference pins (parity unpinned)."""
import os

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "r01_small.npz"))


def test_oracle_reproduces_golden(oracle, gold):
    lio = oracle.OracleLio(resolution=0.5, nearby=18)
    lio.insert(gold["map"])
    idx, d2, cnt = lio.knn5(gold["knn_query"])
    np.testing.assert_array_equal(idx, gold["knn_idx"])
    np.testing.assert_array_equal(d2, gold["knn_d2"])
    rc, x, P, st = lio.update(gold["scan"], gold["x_prop"], gold["P"])
    assert rc == int(gold["iekf_rc"]) and st.passes == int(gold["passes"])
    np.testing.assert_allclose(x, gold["x_post"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.array(st.HtH[0]).reshape(12, 12), gold["HtH0"], rtol=1e-12)
    ndt = oracle.OracleNdt(resolution=2.0, trans_eps=0.01)
    ndt.set_target(gold["map"])
    ndt.set_source(gold["scan"])
    L = ndt.leaves()
    np.testing.assert_array_equal(L["ids"], gold["ndt_ids"])
    np.testing.assert_array_equal(L["icov"], gold["ndt_icov"])
    s, g, H = ndt.derivatives(gold["ndt_p6"])
    np.testing.assert_allclose([s], [gold["ndt_score"]], rtol=1e-12)
    np.testing.assert_allclose(H, gold["ndt_H"], rtol=1e-9, atol=1e-9 * np.abs(gold["ndt_H"]).max())
    rc, T, r = ndt.align(gold["ndt_guess"])
    assert (r.iters, r.evals) == (int(gold["ndt_iters"]), int(gold["ndt_evals"]))
    np.testing.assert_allclose(T, gold["ndt_final"], atol=1e-6)
    np.testing.assert_allclose(ndt.score_batch(gold["reloc_poses"]), gold["reloc_scores"], rtol=1e-12)
    c, n = oracle.voxel_grid(gold["scan"], 0.5)
    np.testing.assert_array_equal(c, gold["vg_centroids"])
    np.testing.assert_array_equal(n, gold["vg_counts"])


@pytest.mark.gpu
def test_cuda_path_reproduces_golden(api, oracle, gold):
    g = api.IVox(resolution=0.5, nearby=18)
    g.AddPoints(gold["map"])
    idx, d2, cnt = g.GetClosestPoint(gold["knn_query"])
    np.testing.assert_array_equal(idx, gold["knn_idx"])
    np.testing.assert_array_equal(d2, gold["knn_d2"])
    np.testing.assert_array_equal(cnt, gold["knn_cnt"])
    kf = api.Esekf(g)
    kf.change_x(gold["x_prop"])
    kf.change_P(gold["P"])
    assert kf.update_iterated_dyn_share_modified(gold["scan"]) == int(gold["iekf_rc"])
    assert kf.stats.passes == int(gold["passes"]) and list(kf.stats.n_eff) == list(gold["n_eff"])
    d = oracle.boxminus(kf.get_x(), gold["x_post"])
    assert np.abs(d[:6]).max() < 1e-7                      # north_star: 1e-4 m / rad
    HtH, Hth, _ = kf.last_HtH(0)
    assert np.abs(HtH - gold["HtH0"]).max() <= 1e-6 * np.abs(gold["HtH0"]).max()
    ndt = api.NormalDistributionsTransform()
    ndt.setResolution(2.0)
    ndt.setTransformationEpsilon(0.01)
    ndt.setInputTarget(gold["map"])
    ndt.setInputSource(gold["scan"])
    L = ndt.leaves()
    np.testing.assert_array_equal(L["ids"], gold["ndt_ids"])
    np.testing.assert_array_equal(L["npts"], gold["ndt_npts"])
    np.testing.assert_array_equal(L["mean"], gold["ndt_mean"])
    np.testing.assert_array_equal(L["icov"], gold["ndt_icov"])
    s, gr, H = ndt.computeDerivatives(gold["ndt_p6"])
    assert abs(s - gold["ndt_score"]) <= 1e-6 * abs(gold["ndt_score"])
    assert np.abs(H - gold["ndt_H"]).max() <= 1e-6 * np.abs(gold["ndt_H"]).max()
    assert np.abs(gr - gold["ndt_g"]).max() <= 1e-6 * np.abs(gold["ndt_g"]).max()
    ndt.align(gold["ndt_guess"])
    assert (ndt.result.iters, ndt.result.evals) == (int(gold["ndt_iters"]), int(gold["ndt_evals"]))
    assert np.abs(np.array(ndt.result.p_final) - gold["ndt_p_final"]).max() < 1e-4
    np.testing.assert_allclose(ndt.calculateScore(gold["reloc_poses"]), gold["reloc_scores"], rtol=1e-12)
    fit = ndt.getFitnessScore(T=gold["ndt_final"])
    assert abs(fit - gold["fitness"][0]) <= 1e-12 * gold["fitness"][0] and ndt.fitness_in_range == int(gold["fitness"][1])
    vg = api.VoxelGrid()
    vg.setLeafSize(0.5)
    vg.setInputCloud(gold["scan"])
    c, n = vg.filter()
    np.testing.assert_array_equal(c, gold["vg_centroids"])
    np.testing.assert_array_equal(n, gold["vg_counts"])


# ------------------------------------------------------------------ GICP (tests/golden/r02_gicp_small.npz, `python tools/make_golden.py gicp`)
@pytest.fixture(scope="module")
def gold_gicp():
    return np.load(os.path.join(ROOT, "tests", "golden", "r02_gicp_small.npz"))


def test_oracle_reproduces_gicp_golden(oracle, gold_gicp):
    gd = gold_gicp
    g = oracle.OracleGicp()
    g.set_target(gd["map"])
    g.set_source(gd["scan"])
    np.testing.assert_array_equal(g.covariances("source"), gd["cov_src"])
    np.testing.assert_array_equal(g.covariances("target")[::16], gd["cov_tgt_every_16"])
    knn, _ = oracle.exact_knn(gd["scan"], gd["scan"], 20)
    np.testing.assert_array_equal(knn, gd["knn_src"])
    m, idx, maha, d2 = g.correspondences(np.eye(4), gd["guess"])
    assert m == int(gd["corr_m"])
    np.testing.assert_array_equal(idx, gd["corr_idx"])
    np.testing.assert_array_equal(maha, gd["corr_maha"])
    f_op, f_fdf, g_df, g_fdf = g.cost(gd["cost_x"])
    assert f_op == float(gd["cost_f_op"]) and f_fdf == float(gd["cost_f_fdf"])
    np.testing.assert_array_equal(g_fdf, gd["cost_g"])
    rc, fin, r = g.align(gd["guess"])
    assert rc == int(gd["align_rc"]) and r.iterations == int(gd["align_iterations"]) > 1 and r.inner_total == int(gd["align_inner_total"])
    np.testing.assert_array_equal([r.n_f, r.n_df, r.n_fdf], gd["align_calls"])
    np.testing.assert_allclose(fin, gd["align_final"], rtol=0, atol=1e-7)
