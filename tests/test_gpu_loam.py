"""GPU LOAM scan-to-map (b200_loam_*) against the oracle: feature selection and coefficients bit-exact, the LM loop with the
same iteration count and the same transform."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene(synth):
    return synth.loam_scene()


def pair(oracle, api, sc):
    o = oracle.OracleLoam()
    o.set_map(sc["corner_map"], sc["surf_map"])
    g = api.ScanToMap(max_map_points=200_000)
    g.setInputCloud(sc["corner_map"], sc["surf_map"])
    return o, g


def test_features_bit_exact(oracle, api, scene):
    o, g = pair(oracle, api, scene)
    for d in ([0, 0, 0, 0, 0, 0], [0.01, -0.01, 0.02, 0.15, -0.1, 0.05], [0.0, 0.0, 0.3, 2.0, 1.0, 0.0]):
        t6 = scene["t_true"] + np.array(d, np.float32)
        n0, f0, c0 = o.features(scene["corner"], scene["surf"], t6)
        n1, f1, c1 = g.features(scene["corner"], scene["surf"], t6)
        assert n1 == n0
        np.testing.assert_array_equal(f1, f0)
        np.testing.assert_array_equal(c1, c0)     # same fp32 op sequence, same neighbours


def test_optimize_parity(oracle, api, scene):
    o, g = pair(oracle, api, scene)
    for d in ([0.01, -0.01, 0.02, 0.15, -0.1, 0.05], [-0.02, 0.01, -0.03, -0.2, 0.15, -0.05]):
        guess = scene["t_true"] + np.array(d, np.float32)
        t0, s0 = o.optimize(scene["corner"], scene["surf"], guess)
        t1, rc = g.scan2MapOptimization(scene["corner"], scene["surf"], guess)
        assert rc == 0 and g.stats.converged == int(s0["converged"])
        assert g.stats.iters == s0["iters"] and g.stats.n_sel == s0["n_sel"] and g.stats.degenerate == int(s0["degenerate"])
        np.testing.assert_allclose(np.array(g.stats.AtA_first).reshape(6, 6), s0["AtA"], rtol=1e-6)
        np.testing.assert_allclose(t1, t0, rtol=0, atol=1e-6)      # north_star: 1e-4 m / 1e-4 rad
        assert np.abs(t1[3:] - scene["t_true"][3:]).max() < 0.01


def test_map_replacement_and_few_features(oracle, api, scene):
    o, g = pair(oracle, api, scene)
    guess = scene["t_true"] + np.array([0, 0, 0, 0.1, 0, 0], np.float32)
    t1, rc = g.scan2MapOptimization(scene["corner"][:10], scene["surf"][:20], guess, iter_num=5)
    assert rc == 2 and g.stats.iters == 5 and not g.stats.converged
    np.testing.assert_array_equal(t1, guess)
    # setInputCloud replaces the maps (the kd-trees are rebuilt for every scan)
    g.setInputCloud(scene["corner_map"][::2], scene["surf_map"][::2])
    o.set_map(scene["corner_map"][::2], scene["surf_map"][::2])
    n0, f0, c0 = o.features(scene["corner"], scene["surf"], scene["t_true"])
    n1, f1, c1 = g.features(scene["corner"], scene["surf"], scene["t_true"])
    np.testing.assert_array_equal(f1, f0)
    np.testing.assert_array_equal(c1, c0)
