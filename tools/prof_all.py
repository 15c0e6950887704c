"""Short program for the ncu passes: a few invocations of every hot kernel at bench sizes (map sizes reduced where the
kernel's cost per unit does not depend on them).  Run plain first, then under ncu (see profiles/README.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pointcloud_slam_b200 import api, synth

n_map = int(os.environ.get("PROF_N_MAP", 2_000_000))
n_prior = int(os.environ.get("PROF_N_PRIOR", 4_000_000))
n_hyp = int(os.environ.get("PROF_N_HYP", 512))

data = synth.config1(n_map, 20_000)
ivox = api.IVox(resolution=0.2, nearby=26)
ivox.AddPoints(data["map"])
kf = api.Esekf(ivox)
kf.lib_graph = api.lib().b200_iekf_set_graph(kf.h, 0)   # plain launches so that ncu sees each kernel
for k in range(3):
    kf.change_x(data["x_prop"]); kf.change_P(data["P"])
    kf.update_iterated_dyn_share_modified(data["scan"])
print("iekf", kf.stats.passes, kf.stats.knn_passes, list(kf.stats.n_eff)[:4], f"{kf.stats.gpu_ms:.3f} ms")

cfg = synth.config2(n_prior, 20_000)
g = api.NormalDistributionsTransform()
g.setTransformationEpsilon(0.01)
g.setInputTarget(cfg["map"]); g.setInputSource(cfg["scan"]); g._handle()
g.setInputTarget(cfg["map"])
print("ndt voxels", g.numVoxels(), f"build {g.last_ms():.3f} ms")
for k in range(3):
    s, gr, H = g.computeDerivatives(cfg["p_guess"])
print("deriv", s, f"{g.last_ms():.4f} ms")
for k in range(2):
    rc = g.align(cfg["guess"])
print("align", rc, g.result.iters, g.result.evals, f"{g.result.gpu_ms:.3f} ms")
poses = synth.hypothesis_grid(cfg["p_true"], 32, 32, 4, 1.0)[:: max(1, 4096 // n_hyp)][:n_hyp]
for k in range(2):
    best, score, ms = api.relocalize(g, poses)
print("reloc", len(poses), best, f"{ms:.3f} ms")
