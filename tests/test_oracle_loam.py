"""Self-checks of the LOAM scan-to-map oracle (jueying_slam mapOptmization.cpp:1255-1590; parity unpinned)."""
import numpy as np


def test_features_select_planes_and_edges(oracle, synth):
    sc = synth.loam_scene()
    o = oracle.OracleLoam()
    o.set_map(sc["corner_map"], sc["surf_map"])
    n, flags, coeff = o.features(sc["corner"], sc["surf"], sc["t_true"])
    nc = len(sc["corner"])
    assert flags[:nc].mean() > 0.8 and flags[nc:].mean() > 0.7 and n == flags.sum()
    # at the true pose the residuals (coeff.intensity = s * distance) are at the noise level, the directions are unit-ish
    sel = flags.astype(bool)
    assert np.abs(coeff[sel, 3]).mean() < 0.03
    norms = np.linalg.norm(coeff[sel, :3], axis=1)
    assert norms.max() <= 1.0 + 1e-5 and norms.min() > 0.1


def test_optimize_recovers_pose(oracle, synth):
    sc = synth.loam_scene()
    o = oracle.OracleLoam()
    o.set_map(sc["corner_map"], sc["surf_map"])
    guess = sc["t_true"] + np.array([0.01, -0.01, 0.02, 0.15, -0.1, 0.05], np.float32)
    t, st = o.optimize(sc["corner"], sc["surf"], guess)
    assert st["converged"] and 2 <= st["iters"] <= 30 and not st["degenerate"]
    assert np.abs(t[3:] - sc["t_true"][3:]).max() < 0.01 and np.abs(t[:3] - sc["t_true"][:3]).max() < 2e-3
    assert np.allclose(st["AtA"], st["AtA"].T) and np.linalg.eigvalsh(st["AtA"]).min() > 100   # not degenerate


def test_too_few_features_leave_the_transform(oracle, synth):
    sc = synth.loam_scene()
    o = oracle.OracleLoam()
    o.set_map(sc["corner_map"], sc["surf_map"])
    guess = sc["t_true"] + np.array([0, 0, 0, 0.1, 0, 0], np.float32)
    t, st = o.optimize(sc["corner"][:10], sc["surf"][:20], guess, iter_num=5)   # < 50 selected: LMOptimization returns false
    assert not st["converged"] and st["iters"] == 5
    np.testing.assert_array_equal(t, guess)
