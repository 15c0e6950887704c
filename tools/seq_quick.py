"""Dev: the configs[2] sliding-map sequence leg of bench.py alone.  usage: seq_quick.py [scans] [parity_scans]"""
import os, sys, json, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pointcloud_slam_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
par = int(sys.argv[2]) if len(sys.argv) > 2 else 40
a = argparse.Namespace(no_cpu=(par == 0))
r = bench.sequence_leg(a, 0, api, synth, n, par)
print(json.dumps({k: r[k] for k in ("scans", "ms_update_device", "ms_per_scan_e2e", "launch_modes", "map_points", "parity_vs_oracle", "final_position_error_m")}))
if os.environ.get("B200_SEQ_TRACE"):
    import numpy as np
    t = np.array(r["trace_ms_map_incremental"])
    print("incr: median %.4f  top:" % np.median(t), [(int(i), float(t[i])) for i in np.argsort(-t)[:12]])
    u = np.array(r["trace_ms_update_device"])
    print("update: median %.4f  top:" % np.median(u), [(int(i), float(u[i])) for i in np.argsort(-u)[:12]])
