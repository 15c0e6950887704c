"""The header-only C++ adaptors (pointcloud-slam_b200/host/*.hpp) compile against include/b200reg.h, link
libb200reg.so and behave: without a GPU construction fails loudly (no CPU fallback), with one the IVox / Esekf /
NormalDistributionsTransform mirrors register a small scene."""
import os
import subprocess

import pytest

from conftest import ROOT, have_gpu


def build(tmp_path, api):
    exe = str(tmp_path / "host_smoke")
    libdir = os.path.dirname(api.lib_path())
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(libdir, "host"),
           os.path.join(ROOT, "tests", "helpers", "host_smoke.cpp"), "-o", exe, "-L", libdir, "-lb200reg", f"-Wl,-rpath,{libdir}",
           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    return exe


@pytest.mark.skipif(have_gpu(), reason="only meaningful on a box without a GPU")
def test_host_adaptors_compile_and_fail_loudly_without_gpu(tmp_path, api):
    p = subprocess.run([build(tmp_path, api)], capture_output=True, text=True)
    assert p.returncode == 10 and "FAILED LOUDLY" in p.stdout and "no CUDA device" in p.stdout, p.stdout + p.stderr


@pytest.mark.gpu
def test_host_adaptors_run_on_gpu(tmp_path, api):
    p = subprocess.run([build(tmp_path, api)], capture_output=True, text=True)
    assert p.returncode == 0 and "host adaptors ok" in p.stdout, p.stdout + p.stderr
