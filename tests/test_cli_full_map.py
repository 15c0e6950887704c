"""The construct_full_map command-line tool (pointcloud-slam_b200/apps) against the oracle, through real files:
frames/<i>.pcd (binary and ascii PointXYZI), poses.txt (x y z qw qx qy qz), merged map .pcd."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, have_gpu


def write_pcd(path, pts, ascii_=False, extra_field=False):
    n = len(pts)
    fields = "x y z intensity" + (" ring" if extra_field else "")
    cols = 5 if extra_field else 4
    hdr = (f"# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS {fields}\nSIZE {' '.join(['4'] * cols)}\n"
           f"TYPE {' '.join(['F'] * 4 + (['U'] if extra_field else []))}\nCOUNT {' '.join(['1'] * cols)}\nWIDTH {n}\nHEIGHT 1\n"
           f"VIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {'ascii' if ascii_ else 'binary'}\n")
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if ascii_:
            for p in pts:
                f.write((" ".join(repr(float(v)) for v in p) + (" 7" if extra_field else "") + "\n").encode())
        elif extra_field:
            rec = np.zeros(n, dtype=[("p", np.float32, 4), ("ring", np.uint32)])
            rec["p"] = pts
            rec["ring"] = 7
            f.write(rec.tobytes())
        else:
            f.write(np.ascontiguousarray(pts, np.float32).tobytes())


def read_pcd(path):
    raw = open(path, "rb").read()
    i = raw.index(b"DATA binary\n") + len(b"DATA binary\n")
    n = int([l for l in raw[:i].decode().splitlines() if l.startswith("POINTS")][0].split()[1])
    return np.frombuffer(raw[i:], np.float32).reshape(n, 4)


def build_tool(tmp_path, api):
    exe = str(tmp_path / "construct_full_map")
    libdir = os.path.dirname(api.lib_path())
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(libdir, "host"),
           os.path.join(libdir, "apps", "construct_full_map.cpp"), "-o", exe, "-L", libdir, "-lb200reg", f"-Wl,-rpath,{libdir}",
           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    return exe


def make_dataset(tmp_path, synth):
    from test_gpu_voxel import make_frames
    frames, poses = make_frames(synth, k=6, n=3000)
    d = tmp_path / "frames"
    d.mkdir()
    for i, f in enumerate(frames):   # ids 0, 1, 2, 10, 11, 12: numeric, not lexicographic, order
        write_pcd(str(d / f"{[0, 1, 2, 10, 11, 12][i]}.pcd"), f, ascii_=(i == 1), extra_field=(i == 2))
    (d / "notes.txt").write_text("ignored")
    np.savetxt(str(tmp_path / "poses.txt"), poses, fmt="%.17g")
    return frames, poses


@pytest.mark.skipif(have_gpu(), reason="only meaningful on a box without a GPU")
def test_cli_builds_and_fails_loudly_without_gpu(tmp_path, api, synth):
    exe = build_tool(tmp_path, api)
    make_dataset(tmp_path, synth)
    p = subprocess.run([exe, str(tmp_path / "poses.txt"), str(tmp_path / "frames"), str(tmp_path / "out.pcd"), "0.1"], capture_output=True, text=True)
    assert p.returncode == 1 and "no CUDA device" in p.stderr
    assert subprocess.run([exe], capture_output=True).returncode == 64


@pytest.mark.gpu
def test_cli_full_map_matches_oracle(tmp_path, api, oracle, synth):
    exe = build_tool(tmp_path, api)
    frames, poses = make_dataset(tmp_path, synth)
    out = tmp_path / "map.pcd"
    p = subprocess.run([exe, str(tmp_path / "poses.txt"), str(tmp_path / "frames"), str(out), "0.1"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    got = read_pcd(str(out))
    c0, n0 = oracle.full_map(frames, poses, 0.1)
    assert len(got) == len(c0)
    np.testing.assert_allclose(got, c0, rtol=0, atol=2e-5)
