"""The C-ABI library loads and exports every symbol include/b200reg.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, have_gpu


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "b200reg.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for s in ["b200_map_create", "b200_map_insert", "b200_map_knn5", "b200_iekf_update", "b200_iekf_map_incremental",
              "b200_ndt_set_target", "b200_ndt_align", "b200_ndt_derivatives", "b200_ndt_score_batch", "b200_reloc_argmin"]:
        assert s in syms


def test_library_exports_every_declared_symbol(api):
    lib = ctypes.CDLL(api.lib_path())
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"libb200reg.so does not export: {missing}"


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/b200reg.h must compile as C99 (no C++ in the signatures), which is what a cgo / JNI / ctypes
    binding generator would consume."""
    import subprocess
    src = tmp_path / "c_abi_check.c"
    src.write_text('#include "b200reg.h"\nint main(void) { b200_gicp_params g; b200_ndt_params n; (void)g; (void)n; return b200_version() == 0; }\n')
    p = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-fsyntax-only", str(src)],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr


def test_version_string(api):
    assert b"sm_100a" in api.lib().b200_version()


@pytest.mark.skipif(have_gpu(), reason="only meaningful on a box without a GPU")
def test_no_gpu_fails_loudly(api):
    """There is no CPU fallback: without a device create() must return an error, not a working handle."""
    with pytest.raises(api.B200Error):
        api.IVox(resolution=0.5, nearby=18)


@pytest.mark.skipif(have_gpu(), reason="only meaningful on a box without a GPU")
def test_no_gpu_fails_loudly_for_every_handle_type(api):
    """NDT, GICP, the map builder and the LOAM estimator refuse to exist without a device as well."""
    import numpy as np
    pts = np.zeros((64, 3), np.float32)
    for make in (lambda: api.NormalDistributionsTransform().setInputTarget(pts) or api.NormalDistributionsTransform()._handle(),
                 lambda: api.GeneralizedIterativeClosestPoint()._handle(),
                 lambda: api.FullMapBuilder(leaf=0.1, capacity_voxels=1000),
                 lambda: api.ScanToMap(max_map_points=1000)):
        with pytest.raises(api.B200Error):
            make()


def test_bad_arguments_are_rejected(api):
    lib = api.lib()
    assert lib.b200_map_create(None, 0, None) == -1  # B200_ERR_ARG
    assert lib.b200_map_insert(None, None, 0, 12) == -1
    assert b"null" in lib.b200_last_error()
    assert lib.b200_gicp_create(None, 0, None) == -1
    assert lib.b200_gicp_set_target(None, None, 0, 12) == -1
    assert lib.b200_gicp_align(None, None, None, None) == -1
