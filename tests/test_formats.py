"""File and wire formats of the map pipeline (host/pcd_io.hpp): area-list CSV of the tiled prior maps
(jueying_slam/include/dynamic_map.h:14-156), TUM trajectory (jueying_lio/src/laser_mapping.cc:825-841), the PointCloud2 body
of /cloud_registered (laser_mapping.cc:747-773).  Plain host C++, no GPU."""
import os
import subprocess

from conftest import ROOT


def test_formats_cpp(tmp_path):
    exe = str(tmp_path / "formats_test")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "pointcloud-slam_b200", "host"),
           os.path.join(ROOT, "tests", "helpers", "formats_test.cpp"), "-o", exe]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    p = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 0 and "formats ok" in p.stdout, p.stdout + p.stderr
