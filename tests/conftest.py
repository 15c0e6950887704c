import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
        n = ctypes.c_int(0)
        return cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:
            return False


HAVE_GPU = None


def have_gpu():
    global HAVE_GPU
    if HAVE_GPU is None:
        HAVE_GPU = _have_gpu()
    return HAVE_GPU


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def api():
    """The product: ctypes view of libb200reg.so.  GPU tests fail loudly when it is missing."""
    from pointcloud_slam_b200 import api as _api
    assert os.path.exists(_api.lib_path()), "libb200reg.so missing: run __graft_entry__.build()"
    return _api


@pytest.fixture(scope="session")
def synth():
    from pointcloud_slam_b200 import synth as _s
    return _s


@pytest.fixture(scope="session")
def small_cfg(synth):
    """200k-point map, 5k-point scan: the oracle finishes in well under a second."""
    return synth.config1(n_map=200_000, n_scan=5_000)


def world_scan(synth, cfg, x=None):
    x = cfg["x_true"] if x is None else x
    o, Rl = synth.lidar_pose(x)
    return (cfg["scan"].astype(np.float64) @ Rl.T + o).astype(np.float32)
