"""Multi-GPU relocalization through the product path (NCCL): needs >= 2 visible GPUs, skipped otherwise."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def gpu_count():
    try:
        import ctypes
        n = ctypes.c_int(0)
        return n.value if ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        import torch
        return torch.cuda.device_count()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2])
def test_sharded_reloc_matches_single_gpu(tmp_path, world):
    if gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = tmp_path / "res.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(ROOT, "tests", "helpers", "reloc_worker.py"), "nccl", str(out)]
    p = subprocess.run(cmd, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True, text=True, timeout=580)
    assert p.returncode == 0, p.stderr[-3000:]
    r = json.load(open(out))
    assert r["same_on_all_ranks"]
    assert r["best"] == r["expect"] and r["score"] == r["expect_score"]
    # sharded construct_full_map: same voxels, counts and centroids as one GPU, each voxel on exactly one rank
    assert r["fullmap_voxels"] == r["single_voxels"] and r["fullmap_points"] == r["single_points"]
    assert r["fullmap_match"]
