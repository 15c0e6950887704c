"""Dev helper: per-scan device time of the configs[2] sequence leg (B200_KNN_MODE in the environment picks the walk)."""
import json, os, sys, types
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ["B200_SEQ_TRACE"] = "1"
import bench
from pointcloud_slam_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
args = types.SimpleNamespace(no_cpu=True)
r = bench.sequence_leg(args, 0, api, synth, n, n_parity=0)
t = r.pop("trace_ms_update_device")
print("MODE", os.environ.get("B200_KNN_MODE", "auto"), "median %.3f mean %.3f" % (r["ms_update_device_median"], r["ms_update_device"]["mean"]),
      "e2e mean %.3f" % r["ms_per_scan_e2e"]["mean"], "voxels", r["map_voxels"][-1], "points", r["map_points"][-1])
print(" ".join("%.2f" % v for v in t[::4]))
