"""Round-2 ncu program: one cold IEKF update (plain launches), the standalone 1M-query search, NDT derivative evaluations and
a relocalization batch.  Run plain first, then under ncu (profiles/README.md); tools/ncu_traffic.py turns the report into
profiles/traffic.json (dram bytes per launch, read by bench.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pointcloud_slam_b200 import api, synth

n_prior = int(os.environ.get("PROF_N_PRIOR", 4_000_000))
n_hyp = int(os.environ.get("PROF_N_HYP", 512))
data = synth.config1(2_000_000, 20_000)
ivox = api.IVox(resolution=0.2, nearby=26)
ivox.AddPoints(data["map"])
kf = api.Esekf(ivox)
api.lib().b200_iekf_set_graph(kf.h, 0)   # plain launches so that ncu sees each kernel
for k in range(2):
    api.flush_l2(0)                      # cold, like the timed steps of bench.py
    kf.change_x(data["x_prop"]); kf.change_P(data["P"])
    kf.update_iterated_dyn_share_modified(data["scan"])
print("iekf", kf.stats.passes, kf.stats.knn_passes, list(kf.stats.n_eff)[:4], f"{kf.stats.gpu_ms:.3f} ms")
ol, Rl = synth.lidar_pose(data["x_prop"])
qw = (data["scan"].astype(np.float64) @ Rl.T + ol).astype(np.float32)
rng = np.random.default_rng(1)
qbig = np.ascontiguousarray(np.concatenate([qw + rng.normal(0, 0.05, qw.shape).astype(np.float32) for _ in range(50)], 0))
for q in (qw, qbig):
    api.flush_l2(0)
    ivox.GetClosestPoint(q)
    print("knn", len(q), f"{ivox.last_knn_ms():.4f} ms")
cfg = synth.config2(n_prior, 20_000)
g = api.NormalDistributionsTransform()
g.setTransformationEpsilon(0.01)
g.setInputTarget(cfg["map"]); g.setInputSource(cfg["scan"]); g._handle()
for k in range(2):
    api.flush_l2(0)
    s, gr, H = g.computeDerivatives(cfg["p_guess"])
print("deriv", s, f"{g.last_ms():.4f} ms")
poses = synth.hypothesis_grid(cfg["p_true"], 32, 32, 4, 1.0)[:: max(1, 4096 // n_hyp)][:n_hyp]
for k in range(2):
    api.flush_l2(0)
    best, score, ms = api.relocalize(g, poses)
print("reloc", len(poses), best, f"{ms:.3f} ms")
