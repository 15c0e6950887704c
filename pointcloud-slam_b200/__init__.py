"""pointcloud-slam_b200 — B200-native scan-to-map registration engine (hot path of matiable/pointcloud-slam).

csrc/   hand-written sm_100a CUDA kernels + the C ABI (include/b200reg.h) -> libb200reg.so
host/   header-only C++ adaptors that re-create the reference's interfaces on top of the C ABI
api.py  ctypes mirror of the same interfaces for tests/ and bench.py
synth.py seeded synthetic worlds / maps / scans
"""
from . import api, synth  # noqa: F401
