"""Dev timing of the NDT path at BASELINE.json configs[1] / configs[3] sizes (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pointcloud_slam_b200 import api, synth

n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
t0 = time.perf_counter()
cfg = synth.config2(n_map, 20_000)
print(f"synth {time.perf_counter()-t0:.1f}s map {cfg['map'].shape} scan {cfg['scan'].shape}", flush=True)
g = api.NormalDistributionsTransform()
g.setTransformationEpsilon(0.01)
for k in range(3):
    t0 = time.perf_counter()
    g.setInputTarget(cfg["map"])
    if k == 0:
        g.setInputSource(cfg["scan"])
        g._handle()
    t1 = time.perf_counter()
    print(f"set_target wall {1e3*(t1-t0):.1f} ms  device {g.last_ms():.3f} ms  voxels {g.numVoxels()}", flush=True)
for k in range(3):
    t0 = time.perf_counter()
    s, gr, H = g.computeDerivatives(cfg["p_guess"])
    t1 = time.perf_counter()
    print(f"derivatives wall {1e3*(t1-t0):.3f} ms device {g.last_ms():.4f} ms score {s:.3f} pairs {g.nbhd_total(cfg['p_guess'])}")
for k in range(3):
    t0 = time.perf_counter()
    rc = g.align(cfg["guess"])
    t1 = time.perf_counter()
    r = g.result
    print(f"align rc {rc} wall {1e3*(t1-t0):.3f} ms device {r.gpu_ms:.4f} ms iters {r.iters} evals {r.evals} hess {r.hess_evals} launches {g.last_launches()} "
          f"dp {np.round(np.array(r.p_final)-cfg['p_true'],4)}")
poses = synth.hypothesis_grid(cfg["p_true"], 32, 32, 4, 1.0)
for k in range(3):
    t0 = time.perf_counter()
    best, score, ms = api.relocalize(g, poses)
    t1 = time.perf_counter()
    print(f"reloc 4096 hyps wall {1e3*(t1-t0):.3f} ms device {ms:.4f} ms best {best} (true {(16*32+16)*4}) score {score:.5f} -> {4096/ms*1e3:.0f} hyp/s")
t0 = time.perf_counter()
finals, res = g.alignBatch(poses[:512])
t1 = time.perf_counter()
it = [r.iters for r in res]
print(f"align_batch 512 wall {1e3*(t1-t0):.2f} ms device {g.last_ms():.3f} ms iters mean {np.mean(it):.1f} max {max(it)} launches {g.last_launches()}")
