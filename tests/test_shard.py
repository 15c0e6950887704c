"""The sharded (N > 1) relocalization path: slice arithmetic, and the allreduce-argmin protocol over 2 gloo ranks on CPU."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_range_partitions_exactly(api):
    for total in (0, 1, 7, 75, 4096):
        for n in (1, 2, 3, 4, 8):
            cuts = [api.shard_range(total, n, r) for r in range(n)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(n - 1))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_score_key_is_order_preserving(api):
    rng = np.random.default_rng(1)
    v = np.concatenate([rng.normal(0, 1, 500), [0.0, -0.0, 1e-300, -1e-300, 1e300, -1e300, np.inf, -np.inf]])
    k = np.array([api.score_key(x) for x in v], dtype=np.uint64)
    order_v = np.argsort(v, kind="stable")
    assert (np.diff(k[order_v].astype(np.float64)) >= 0).all()
    assert api.score_key(float("nan")) == 0
    for x in v:
        assert api.score_from_key(api.score_key(x)) == x or (x == 0 and api.score_from_key(api.score_key(x)) == 0)


def test_argmin_protocol_single_rank(api):
    s = np.array([0.1, 0.7, 0.7, -0.2])
    best, score = api.argmin_protocol_host(s, 10, lambda t: t[None])
    assert (best, score) == (11, 0.7)       # ties go to the lower index
    best, score = api.argmin_protocol_host(np.zeros(0), 0, lambda t: t[None])
    assert best == -1
    import torch
    two = lambda t: torch.stack([t, torch.tensor([api.score_key(0.7) - (1 << 63), 3], dtype=torch.int64)])
    assert api.argmin_protocol_host(s, 10, two) == (3, 0.7)   # equal scores on two ranks: the lower global index wins


@pytest.mark.timeout(300)
def test_two_rank_gloo_argmin(tmp_path):
    out = tmp_path / "res.json"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "helpers", "reloc_worker.py"), "gloo", str(out)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=280)
    assert p.returncode == 0, p.stderr[-2000:]
    r = json.load(open(out))
    assert r["world"] == 2 and r["slice"] == [0, 38]
    assert r["best"] == r["expect"]
    assert r["score"] == r["expect_score"]
    # the map builder's merge protocol over the same two ranks: disjoint ownership by the key hash, union == the oracle's map
    assert r["fullmap_disjoint"] and r["fullmap_owner_rule"] and r["fullmap_voxels"] == r["oracle_voxels"] > 1000
    assert r["fullmap_counts_equal"] and r["fullmap_max_centroid_diff"] <= 2e-5
