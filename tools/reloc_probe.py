"""Dev helper: configs[3] relocalization scoring (4096 hypotheses, 20k-pt scan, 10M-pt map) timing + score digest."""
import hashlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pointcloud_slam_b200 import api, synth
n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
cfg = synth.config2(n_map, 20_000)
g = api.NormalDistributionsTransform()
g.setTransformationEpsilon(0.01)
g.setInputTarget(cfg["map"])
g.setInputSource(cfg["scan"])
poses = synth.hypothesis_grid(cfg["p_true"], 32, 32, 4, 1.0)
for _ in range(2):
    best, score, ms = api.relocalize(g, poses, None)
t = []
for _ in range(5):
    api.flush_l2(0)
    best, score, ms = api.relocalize(g, poses, None)
    t.append(ms)
s = g.calculateScore(poses)
print("reloc ms %.4f (min %.4f) hyp/s %.3e best %d score %.12f digest %s voxels %d" % (np.mean(t), np.min(t), len(poses) / (np.mean(t) * 1e-3), best, score,
      hashlib.md5(np.round(s, 9).tobytes()).hexdigest()[:10], g.numVoxels()))
d = []
for _ in range(5):
    api.flush_l2(0)
    g.computeDerivatives(cfg["p_guess"]); d.append(g.last_ms())
a = []
for _ in range(5):
    api.flush_l2(0)
    g.align(cfg["guess"]); a.append(g.result.gpu_ms)
print("derivatives ms %.4f align ms %.4f iters %d evals %d" % (np.mean(d), np.mean(a), g.result.iters, g.result.evals))
