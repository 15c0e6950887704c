"""Dev: the configs[0] IEKF leg of bench.py alone for the search variants given on the command line (B200_KNN_MODE)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    import argparse, torch
    import bench
    from pointcloud_slam_b200 import api, synth
    a = argparse.Namespace(params=sys.argv[2] if len(sys.argv) > 2 else "livox", steps=30, warmup=5, no_cpu=True)
    r = bench.iekf_leg(a, 0, 0, 1, api, synth, torch, full=True)
    print(json.dumps({"ms": round(r["ms_per_update"], 4), "warm": round(r["ms_per_update_warm_l2"], 4), "e2e": round(r["e2e"]["ms_per_update"], 4),
                      "k_search_us": round(1e3 * r["kernels"]["k_search_ms"], 2), "k_obs_us": round(1e3 * r["kernels"]["k_obs_ms"], 2),
                      "frac": round(r["roofline"]["frac"], 4), "frac_1M": round(r["roofline"]["batched"]["frac"], 4),
                      "ms_1M": round(r["roofline"]["batched"]["kernel_ms"], 4), "n_eff": r["n_eff"]}))
else:
    params = "livox"
    modes = [m for m in sys.argv[1:] if m.isdigit()] or ["7", "8", "0"]
    for m in sys.argv[1:]:
        if not m.isdigit(): params = m
    for mode in modes:
        r = subprocess.run([sys.executable, __file__, "worker", params], env=dict(os.environ, B200_KNN_MODE=str(mode)), capture_output=True, text=True)
        print("mode", mode, params, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-800:], flush=True)
