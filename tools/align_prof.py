"""Dev helper for ncu: configs[1] NDT - a few derivative evaluations and aligns."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pointcloud_slam_b200 import api, synth
n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
cfg = synth.config2(n_map, 20_000)
g = api.NormalDistributionsTransform()
g.setTransformationEpsilon(0.01)
g.setInputTarget(cfg["map"])
g.setInputSource(cfg["scan"])
for _ in range(3):
    g.computeDerivatives(cfg["p_guess"]); print("deriv", g.last_ms())
for _ in range(3):
    g.align(cfg["guess"]); print("align", g.result.gpu_ms, g.result.iters, g.result.evals, g.last_launches())
import ctypes as C
a, b = C.c_int64(0), C.c_int64(0)
api.lib().b200_ndt_debug_cycles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
api.lib().b200_ndt_debug_cycles(g._handle(), C.byref(a), C.byref(b))
print("advance() cycles over the align: %d (%.1f us at 1.92 GHz), of which Newton solves %d; per evaluation %.1f us" % (a.value, a.value / 1920.0, b.value, a.value / 1920.0 / max(g.result.evals, 1)))
