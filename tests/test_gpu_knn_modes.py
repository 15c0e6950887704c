"""Every candidate-walk variant of the k = 5 stencil search (lane-owned, cooperative, per-query hybrid, flattened list with
its cooperative fallback) must return the same neighbours, bit for bit, as the oracle and as each other - the variant is
picked per launch from the map's density, so a map may see several of them over its lifetime.  One process per variant:
B200_KNN_MODE is read once per process."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_mode(mode):
    env = dict(os.environ)
    if mode is None:
        env.pop("B200_KNN_MODE", None)
    else:
        env["B200_KNN_MODE"] = str(mode)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "helpers", "knn_mode_worker.py")], env=env, capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


def test_all_walk_variants_agree_bit_for_bit():
    ref = run_mode(None)
    for regime in ("sparse", "dense"):
        assert ref[regime]["oracle_equal"], regime
        assert ref[regime]["rc"] == 0
    assert ref["sparse"]["mean_candidates"] < 40 < 64 < ref["dense"]["mean_candidates"]   # both sides of the flat-list capacity
    for mode in (9, 8, 7, 0, 1, 4, 5):   # 9: candidate runs staged in shared memory by 1-D TMA bulk copies (the dense regime exceeds the
        got = run_mode(mode)             #    staging area on most queries and takes the gather path)
        assert got["tma_timeouts"] == 0, mode
        for regime in ("sparse", "dense"):
            assert got[regime]["oracle_equal"], (mode, regime)
            assert got[regime]["knn"] == ref[regime]["knn"], (mode, regime)
            assert got[regime]["x"] == ref[regime]["x"], (mode, regime)
