"""Dev: time the candidate-walk variants of the k-NN kernel (B200_KNN_MODE) in the sparse (livox, 0.2 m) and dense
(sliding-map, 0.5 m voxels with ~25 points each) regimes, 20k and 1M queries, cold and warm L2."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    import numpy as np
    from pointcloud_slam_b200 import api, synth
    out = {}
    data = synth.config1(2_000_000, 20_000)
    o_l, Rl = synth.lidar_pose(data["x_prop"])
    qw = (data["scan"].astype(np.float64) @ Rl.T + o_l).astype(np.float32)
    rng = np.random.default_rng(1)
    qbig = np.ascontiguousarray(np.concatenate([qw + rng.normal(0, 0.05, qw.shape).astype(np.float32) for _ in range(50)], 0))
    for name, res, nearby, mp in (("sparse", 0.2, 26, data["map"]), ("dense", 0.5, 18, data["map"][:1_600_000])):
        ivox = api.IVox(resolution=res, nearby=nearby)
        ivox.AddPoints(mp)
        for qn, q in (("20k", qw), ("1M", qbig)):
            ivox.GetClosestPoint(q)
            cold, warm = [], []
            for _ in range(5):
                api.flush_l2(0)
                ivox.GetClosestPoint(q); cold.append(ivox.last_knn_ms())
                ivox.GetClosestPoint(q); warm.append(ivox.last_knn_ms())
            pts, cells = ivox.stencil_points(q)
            out[f"{name}_{qn}"] = dict(cold_us=1e3 * float(np.median(cold)), warm_us=1e3 * float(np.median(warm)), cand=pts / len(q))
        idx, d2, cnt = ivox.GetClosestPoint(qw)
        out[f"{name}_checksum"] = int(idx.astype(np.int64).sum())
        ivox.close()
    print(json.dumps(out))
else:
    for mode in (7, 0, 1, 5):   # 7 = warp per query, 0 = lane-owned walk, 1 = cooperative walk, 5 = flattened walk (8 lanes per query)
        r = subprocess.run([sys.executable, __file__, "worker"], env=dict(os.environ, B200_KNN_MODE=str(mode)), capture_output=True, text=True)
        print("mode", mode, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:], flush=True)
