"""Dev: the configs[4] construct_full_map leg of bench.py alone (single GPU).  usage: fullmap_quick.py [frames]"""
import os, sys, json, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from pointcloud_slam_b200 import api, synth
a = argparse.Namespace(fullmap_frames=int(sys.argv[1]) if len(sys.argv) > 1 else 10000, fullmap_pool=32, fullmap_reuse=5, fullmap_host_frames=400, no_cpu=False)
r = bench.fullmap_leg(a, 0, 0, 1, api, synth, torch, None)
print(json.dumps({k: r[k] for k in ("value", "seconds", "map_voxels", "e2e", "parity")}))
