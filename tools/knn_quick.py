"""Dev: standalone k=5 search timing for the modes given on the command line (default 7 0 5), sparse livox map, 20k / 1M queries."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    import numpy as np
    from pointcloud_slam_b200 import api, synth
    data = synth.config1(2_000_000, 20_000)
    o_l, Rl = synth.lidar_pose(data["x_prop"])
    qw = (data["scan"].astype(np.float64) @ Rl.T + o_l).astype(np.float32)
    rng = np.random.default_rng(1)
    qbig = np.ascontiguousarray(np.concatenate([qw + rng.normal(0, 0.05, qw.shape).astype(np.float32) for _ in range(50)], 0))
    out = {}
    ivox = api.IVox(resolution=0.2, nearby=26)
    ivox.AddPoints(data["map"])
    for qn, q in (("20k", qw), ("1M", qbig)):
        ivox.GetClosestPoint(q)
        cold, warm = [], []
        for _ in range(7):
            api.flush_l2(0)
            ivox.GetClosestPoint(q); cold.append(ivox.last_knn_ms())
            ivox.GetClosestPoint(q); warm.append(ivox.last_knn_ms())
        out[qn] = dict(cold_us=round(1e3 * float(np.median(cold)), 1), warm_us=round(1e3 * float(np.median(warm)), 1))
    idx, d2, cnt = ivox.GetClosestPoint(qw)
    out["checksum"] = int(idx.astype(np.int64).sum())
    print(json.dumps(out))
else:
    modes = sys.argv[1:] or ["7", "0", "5"]
    for mode in modes:
        r = subprocess.run([sys.executable, __file__, "worker"], env=dict(os.environ, B200_KNN_MODE=str(mode)), capture_output=True, text=True)
        print("mode", mode, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:], flush=True)
