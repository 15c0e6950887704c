"""TEST INFRASTRUCTURE ONLY — ctypes binding of the CPU oracle (oracle/liboracle.so).

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
PARITY UNPINNED: see oracle/oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ORC_MAX_PASSES = 8


class LioParams(C.Structure):
    _fields_ = [("resolution", C.c_float), ("nearby", C.c_int32), ("capacity_voxels", C.c_uint64),
                ("max_iter", C.c_int32), ("plane_thr", C.c_float), ("extrinsic_est_en", C.c_int32),
                ("R", C.c_double), ("limit", C.c_double * 23), ("filter_size_map", C.c_double),
                ("num_threads", C.c_int32)]


class IekfStats(C.Structure):
    _fields_ = [("status", C.c_int32), ("passes", C.c_int32), ("knn_passes", C.c_int32), ("converged", C.c_int32),
                ("n_eff", C.c_int32 * ORC_MAX_PASSES), ("knn", C.c_int32 * ORC_MAX_PASSES),
                ("x_in", (C.c_double * 26) * ORC_MAX_PASSES), ("HtH", (C.c_double * 144) * ORC_MAX_PASSES),
                ("Hth", (C.c_double * 12) * ORC_MAX_PASSES),
                ("ms_match", C.c_double), ("ms_jacobian", C.c_double), ("ms_solve", C.c_double)]


class NdtParams(C.Structure):
    _fields_ = [("resolution", C.c_float), ("step_size", C.c_double), ("outlier_ratio", C.c_double),
                ("trans_eps", C.c_double), ("max_iter", C.c_int32), ("search", C.c_int32), ("min_pts", C.c_int32),
                ("eig_ratio", C.c_double), ("num_threads", C.c_int32)]


class GicpParams(C.Structure):
    _fields_ = [("k_correspondences", C.c_int32), ("gicp_epsilon", C.c_double), ("rotation_epsilon", C.c_double),
                ("transformation_epsilon", C.c_double), ("corr_dist_threshold", C.c_double), ("max_iterations", C.c_int32),
                ("max_inner_iterations", C.c_int32), ("num_threads", C.c_int32)]


class GicpResult(C.Structure):
    _fields_ = [("converged", C.c_int32), ("iterations", C.c_int32), ("last_m", C.c_int32), ("last_inner", C.c_int32),
                ("last_status", C.c_int32), ("inner_total", C.c_int32), ("n_f", C.c_int32), ("n_df", C.c_int32), ("n_fdf", C.c_int32)]


class NdtResult(C.Structure):
    _fields_ = [("converged", C.c_int32), ("iters", C.c_int32), ("evals", C.c_int32), ("hess_evals", C.c_int32),
                ("trans_probability", C.c_double), ("hessian", C.c_double * 36), ("score", C.c_double),
                ("p_final", C.c_double * 6)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("lio_oracle.cpp", "ndt_oracle.cpp", "voxelgrid_oracle.cpp", "undistort_oracle.cpp", "loam_oracle.cpp", "gicp_oracle.cpp", "smallmat.h", "oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
        L.orc_lio_create.restype = vp
        L.orc_lio_create.argtypes = [C.POINTER(LioParams)]
        L.orc_lio_destroy.argtypes = [vp]
        L.orc_map_insert.restype = i64
        L.orc_map_insert.argtypes = [vp, vp, i64, i64]
        L.orc_map_num_voxels.restype = i64
        L.orc_map_num_voxels.argtypes = [vp]
        L.orc_map_num_points.restype = i64
        L.orc_map_num_points.argtypes = [vp]
        L.orc_map_knn5.argtypes = [vp, vp, i64, i64, vp, vp, vp]
        L.orc_map_knn_candidates.restype = i64
        L.orc_map_knn_candidates.argtypes = [vp, vp, i64, i64]
        L.orc_iekf_update.restype = i32
        L.orc_iekf_update.argtypes = [vp, vp, i64, i64, vp, vp, C.POINTER(IekfStats)]
        L.orc_obs_model.restype = i32
        L.orc_obs_model.argtypes = [vp, vp, i64, i64, vp, i32, vp, vp, vp]
        L.orc_point_state.argtypes = [vp, i64, vp, vp, vp, vp, vp]
        L.orc_last_rows.restype = i32
        L.orc_last_rows.argtypes = [vp, vp, vp, i32]
        L.orc_map_incremental.restype = i64
        L.orc_map_incremental.argtypes = [vp, vp, i64, i64, vp, i32, vp, vp]
        L.orc_esti_plane.restype = i32
        L.orc_esti_plane.argtypes = [vp, i32, C.c_float, vp]
        L.orc_state_boxplus.argtypes = [vp, vp]
        L.orc_state_boxminus.argtypes = [vp, vp, vp]
        L.orc_inverse.argtypes = [vp, i32, vp]
        L.orc_predict.argtypes = [vp, i32, vp, vp, vp, vp]
        L.orc_ndt_create.restype = vp
        L.orc_ndt_create.argtypes = [C.POINTER(NdtParams)]
        L.orc_ndt_destroy.argtypes = [vp]
        L.orc_ndt_set_target.restype = i64
        L.orc_ndt_set_target.argtypes = [vp, vp, i64, i64]
        L.orc_ndt_set_source.argtypes = [vp, vp, i64, i64]
        L.orc_ndt_num_leaves.restype = i64
        L.orc_ndt_num_leaves.argtypes = [vp]
        L.orc_ndt_leaves.restype = i64
        L.orc_ndt_leaves.argtypes = [vp, i64, vp, vp, vp, vp, vp]
        L.orc_ndt_grid.argtypes = [vp, vp, vp]
        L.orc_ndt_derivatives.restype = C.c_double
        L.orc_ndt_derivatives.argtypes = [vp, vp, vp, vp, i32]
        L.orc_ndt_hessian.argtypes = [vp, vp, vp]
        L.orc_ndt_align.restype = i32
        L.orc_ndt_align.argtypes = [vp, vp, vp, C.POINTER(NdtResult)]
        L.orc_ndt_score_batch.argtypes = [vp, vp, i64, vp]
        L.orc_ndt_fitness.restype = C.c_double
        L.orc_ndt_fitness.argtypes = [vp, vp, C.c_double, vp]
        L.orc_ndt_nbhd_total.restype = i64
        L.orc_ndt_nbhd_total.argtypes = [vp, vp]
        L.orc_voxel_grid.restype = i64
        L.orc_voxel_grid.argtypes = [vp, i64, i64, C.c_float, i32, vp, vp, i64]
        L.orc_full_map.restype = i64
        L.orc_full_map.argtypes = [vp, vp, i64, vp, C.c_float, vp, vp, i64]
        L.orc_undistort.argtypes = [vp, i64, i64, i32, i32, vp, i32, vp, vp, vp]
        L.orc_loam_create.restype = vp
        L.orc_loam_create.argtypes = [i32]
        L.orc_loam_destroy.argtypes = [vp]
        L.orc_loam_set_map.argtypes = [vp, vp, i64, i64, vp, i64, i64]
        L.orc_loam_features.restype = i32
        L.orc_loam_features.argtypes = [vp, vp, i64, i64, vp, i64, i64, vp, vp, vp]
        L.orc_loam_optimize.restype = i32
        L.orc_loam_optimize.argtypes = [vp, vp, i64, i64, vp, i64, i64, vp, i32, vp, vp, vp, vp]
        L.orc_euler_from_matrix.argtypes = [vp, vp]
        L.orc_set_legacy_eigen.argtypes = [i32]
        L.orc_eigen_selfadjoint3.restype = i32
        L.orc_eigen_selfadjoint3.argtypes = [vp, vp, vp]
        L.orc_jacobi_svd_solve6.argtypes = [vp, vp, vp, vp]
        L.orc_matrix_from_pose.argtypes = [vp, vp]
        L.orc_gicp_create.restype = vp
        L.orc_gicp_create.argtypes = [C.POINTER(GicpParams)]
        L.orc_gicp_destroy.argtypes = [vp]
        L.orc_gicp_set_target.argtypes = [vp, vp, i64, i64]
        L.orc_gicp_set_source.argtypes = [vp, vp, i64, i64]
        L.orc_gicp_covariances.restype = i32
        L.orc_gicp_covariances.argtypes = [vp, i32, vp]
        L.orc_gicp_align.restype = i32
        L.orc_gicp_align.argtypes = [vp, vp, vp, C.POINTER(GicpResult)]
        L.orc_gicp_correspondences.restype = i64
        L.orc_gicp_correspondences.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_gicp_cost.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_gicp_apply_state.argtypes = [vp, vp]
        L.orc_gicp_estimate.restype = i32
        L.orc_gicp_estimate.argtypes = [vp, vp, vp, vp, vp]
        L.orc_gicp_knn.argtypes = [vp, i64, i64, vp, i64, i64, i32, i32, vp, vp]
        L.orc_bfgs_test.restype = i32
        L.orc_bfgs_test.argtypes = [vp, i32, vp, vp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] >= 3
    return a


def lio_params(resolution=0.5, nearby=18, capacity=1_000_000, max_iter=3, plane_thr=0.1, extrinsic_est_en=False,
               R=0.001, limit=0.001, filter_size_map=0.5, num_threads=0) -> LioParams:
    p = LioParams()
    p.resolution, p.nearby, p.capacity_voxels = resolution, nearby, capacity
    p.max_iter, p.plane_thr, p.extrinsic_est_en = max_iter, plane_thr, int(extrinsic_est_en)
    p.R, p.filter_size_map, p.num_threads = R, filter_size_map, num_threads
    for i in range(23):
        p.limit[i] = limit
    return p


class OracleLio:
    """IVox + ObsModel + IEKF update, CPU oracle."""

    def __init__(self, **kw):
        self.params = lio_params(**kw)
        self.h = lib().orc_lio_create(C.byref(self.params))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_lio_destroy(self.h)
            self.h = None

    def insert(self, pts):
        pts = _f32(pts)
        return lib().orc_map_insert(self.h, _p(pts), pts.shape[0], pts.strides[0])

    @property
    def num_voxels(self):
        return lib().orc_map_num_voxels(self.h)

    @property
    def num_points(self):
        return lib().orc_map_num_points(self.h)

    def knn5(self, q):
        q = _f32(q)
        n = q.shape[0]
        idx = np.empty((n, 5), np.int32)
        d = np.empty((n, 5), np.float32)
        cnt = np.empty(n, np.int32)
        lib().orc_map_knn5(self.h, _p(q), n, q.strides[0], _p(idx), _p(d), _p(cnt))
        return idx, d, cnt

    def knn_candidates(self, q):
        q = _f32(q)
        return lib().orc_map_knn_candidates(self.h, _p(q), q.shape[0], q.strides[0])

    def update(self, scan, x, P):
        scan = _f32(scan)
        x = np.array(x, dtype=np.float64)
        P = np.array(P, dtype=np.float64)
        st = IekfStats()
        rc = lib().orc_iekf_update(self.h, _p(scan), scan.shape[0], scan.strides[0], _p(x), _p(P), C.byref(st))
        return rc, x, P, st

    def obs_model(self, scan, x, converge=True):
        scan = _f32(scan)
        x = np.ascontiguousarray(x, dtype=np.float64)
        HtH = np.zeros((12, 12))
        Hth = np.zeros(12)
        ne = C.c_int32(0)
        rc = lib().orc_obs_model(self.h, _p(scan), scan.shape[0], scan.strides[0], _p(x), int(converge), _p(HtH), _p(Hth),
                                 C.byref(ne))
        return rc, HtH, Hth, ne.value

    def point_state(self, n):
        plane = np.empty((n, 4), np.float32)
        res = np.empty(n, np.float32)
        sel = np.empty(n, np.uint8)
        nn = np.empty((n, 5), np.int32)
        cnt = np.empty(n, np.int32)
        lib().orc_point_state(self.h, n, _p(plane), _p(res), _p(sel), _p(nn), _p(cnt))
        return dict(plane=plane, residual=res, selected=sel, nn_idx=nn, nn_count=cnt)

    def last_rows(self, max_rows):
        hx = np.zeros((max_rows, 12))
        hv = np.zeros(max_rows)
        n = lib().orc_last_rows(self.h, _p(hx), _p(hv), max_rows)
        return hx[:min(n, max_rows)], hv[:min(n, max_rows)], n

    def map_incremental(self, scan, x, ekf_inited=True):
        scan = _f32(scan)
        x = np.ascontiguousarray(x, dtype=np.float64)
        na, nd = C.c_int32(0), C.c_int32(0)
        tot = lib().orc_map_incremental(self.h, _p(scan), scan.shape[0], scan.strides[0], _p(x), int(ekf_inited),
                                        C.byref(na), C.byref(nd))
        return tot, na.value, nd.value


def set_legacy_eigen(on: bool):
    """Test switch: the round-1 cyclic-Jacobi substitutes instead of the restated Eigen SelfAdjointEigenSolver / JacobiSVD."""
    lib().orc_set_legacy_eigen(int(on))


def eigen_selfadjoint3(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    w, V = np.zeros(3), np.zeros((3, 3))
    ok = lib().orc_eigen_selfadjoint3(_p(A), _p(w), _p(V))
    return bool(ok), w, V


def jacobi_svd_solve6(H, rhs):
    H = np.ascontiguousarray(H, dtype=np.float64)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    x, sv = np.zeros(6), np.zeros(6)
    lib().orc_jacobi_svd_solve6(_p(H), _p(rhs), _p(x), _p(sv))
    return x, sv


def esti_plane(pts, thr=0.1):
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    plane = np.zeros(4, np.float32)
    ok = lib().orc_esti_plane(_p(pts), pts.shape[0], thr, _p(plane))
    return bool(ok), plane


def boxplus(x, dx):
    x = np.array(x, dtype=np.float64)
    dx = np.ascontiguousarray(dx, dtype=np.float64)
    lib().orc_state_boxplus(_p(x), _p(dx))
    return x


def boxminus(x, y):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    d = np.zeros(23)
    lib().orc_state_boxminus(_p(x), _p(y), _p(d))
    return d


def predict(steps, Q12, x, P):
    """esekf::predict over K IMU intervals (steps [K,8] = dt, offs_t, acc_avr, angvel_avr); returns x, P, IMUpose_ [K,22]."""
    steps = np.ascontiguousarray(steps, dtype=np.float64).reshape(-1, 8)
    Q12 = np.ascontiguousarray(Q12, dtype=np.float64)
    x = np.array(x, dtype=np.float64)
    P = np.array(P, dtype=np.float64)
    poses = np.zeros((len(steps), 22))
    lib().orc_predict(_p(steps), len(steps), _p(Q12), _p(x), _p(P), _p(poses))
    return x, P, poses


def inverse(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    out = np.empty_like(A)
    lib().orc_inverse(_p(A), A.shape[0], _p(out))
    return out


def ndt_params(resolution=1.0, step_size=0.1, outlier_ratio=0.55, trans_eps=0.01, max_iter=35, search=7, min_pts=6,
               eig_ratio=0.01, num_threads=0) -> NdtParams:
    p = NdtParams()
    p.resolution, p.step_size, p.outlier_ratio, p.trans_eps = resolution, step_size, outlier_ratio, trans_eps
    p.max_iter, p.search, p.min_pts, p.eig_ratio, p.num_threads = max_iter, search, min_pts, eig_ratio, num_threads
    return p


class OracleNdt:
    def __init__(self, **kw):
        self.params = ndt_params(**kw)
        self.h = lib().orc_ndt_create(C.byref(self.params))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ndt_destroy(self.h)
            self.h = None

    def set_target(self, pts):
        pts = _f32(pts)
        return lib().orc_ndt_set_target(self.h, _p(pts), pts.shape[0], pts.strides[0])

    def set_source(self, pts):
        pts = _f32(pts)
        lib().orc_ndt_set_source(self.h, _p(pts), pts.shape[0], pts.strides[0])

    def leaves(self):
        n = lib().orc_ndt_leaves(self.h, 0, None, None, None, None, None)
        ids = np.empty(n, np.int64)
        npts = np.empty(n, np.int32)
        mean = np.empty((n, 3))
        cov = np.empty((n, 3, 3))
        icov = np.empty((n, 3, 3))
        lib().orc_ndt_leaves(self.h, n, _p(ids), _p(npts), _p(mean), _p(cov), _p(icov))
        return dict(ids=ids, npts=npts, mean=mean, cov=cov, icov=icov)

    def grid(self):
        mn = np.zeros(3, np.int32)
        dv = np.zeros(3, np.int32)
        lib().orc_ndt_grid(self.h, _p(mn), _p(dv))
        return mn, dv

    def derivatives(self, p6, compute_hessian=True):
        p6 = np.ascontiguousarray(p6, dtype=np.float64)
        g = np.zeros(6)
        H = np.zeros((6, 6))
        s = lib().orc_ndt_derivatives(self.h, _p(p6), _p(g), _p(H), int(compute_hessian))
        return s, g, H

    def hessian(self, p6):
        p6 = np.ascontiguousarray(p6, dtype=np.float64)
        H = np.zeros((6, 6))
        lib().orc_ndt_hessian(self.h, _p(p6), _p(H))
        return H

    def align(self, guess):
        g = np.ascontiguousarray(np.asarray(guess, dtype=np.float32).T)  # column-major bytes
        out = np.zeros((4, 4), np.float32)
        r = NdtResult()
        rc = lib().orc_ndt_align(self.h, _p(g), _p(out), C.byref(r))
        return rc, out.T.copy(), r

    def score_batch(self, poses_cm16):
        poses = np.ascontiguousarray(poses_cm16, dtype=np.float32)
        s = np.zeros(poses.shape[0])
        lib().orc_ndt_score_batch(self.h, _p(poses), poses.shape[0], _p(s))
        return s

    def nbhd_total(self, p6):
        p6 = np.ascontiguousarray(p6, dtype=np.float64)
        return lib().orc_ndt_nbhd_total(self.h, _p(p6))

    def fitness(self, T, max_range=1.7976931348623157e308):
        t = np.ascontiguousarray(np.asarray(T, dtype=np.float32).T)
        nr = C.c_int64(0)
        s = lib().orc_ndt_fitness(self.h, _p(t), max_range, C.byref(nr))
        return s, nr.value


def euler_from_matrix(M):
    m = np.ascontiguousarray(np.asarray(M, dtype=np.float32).T)
    r = np.zeros(3, np.float32)
    lib().orc_euler_from_matrix(_p(m), _p(r))
    return r


def matrix_from_pose(p6):
    p6 = np.ascontiguousarray(p6, dtype=np.float64)
    m = np.zeros((4, 4), np.float32)
    lib().orc_matrix_from_pose(_p(p6), _p(m))
    return m.T.copy()


def voxel_grid(xyzi, leaf, min_points=0):
    """pcl::VoxelGrid::filter on [n,3] or [n,4] (x,y,z,intensity) float32: returns (centroids [m,4], counts [m])."""
    a = np.ascontiguousarray(xyzi, dtype=np.float32)
    n = a.shape[0]
    out = np.empty((max(n, 1), 4), np.float32)
    cnt = np.empty(max(n, 1), np.int32)
    m = lib().orc_voxel_grid(_p(a), n, a.strides[0], leaf, min_points, _p(out), _p(cnt), n)
    return out[:m].copy(), cnt[:m].copy()


def full_map(frames, poses7, leaf):
    """construct_full_map oracle: frames = list of [n_k,4] float32 clouds, poses7 [k,7] (x y z qw qx qy qz)."""
    offs = np.zeros(len(frames) + 1, np.int64)
    offs[1:] = np.cumsum([len(f) for f in frames])
    allp = np.ascontiguousarray(np.concatenate(frames, 0), dtype=np.float32)
    poses = np.ascontiguousarray(poses7, dtype=np.float64)
    out = np.empty((len(allp), 4), np.float32)
    cnt = np.empty(len(allp), np.int32)
    m = lib().orc_full_map(_p(allp), _p(offs), len(frames), _p(poses), leaf, _p(out), _p(cnt), len(allp))
    return out[:m].copy(), cnt[:m].copy()


def undistort(points, time_index, intensity_index, poses22, x_end26):
    """ImuProcess::UndistortPcl, backward half: returns (xyzi [n,4] in time order, order [n])."""
    a = np.ascontiguousarray(points, dtype=np.float32)
    poses = np.ascontiguousarray(poses22, dtype=np.float64).reshape(-1, 22)
    x = np.ascontiguousarray(x_end26, dtype=np.float64)
    out = np.zeros((a.shape[0], 4), np.float32)
    order = np.zeros(a.shape[0], np.int32)
    lib().orc_undistort(_p(a), a.shape[0], a.strides[0], time_index, intensity_index, _p(poses), poses.shape[0], _p(x), _p(out), _p(order))
    return out, order


class OracleLoam:
    """jueying_slam scan2MapOptimization (corner + surf features, 6x6 LM); transforms are (roll, pitch, yaw, x, y, z) float32."""

    def __init__(self, num_threads=0):
        self.h = lib().orc_loam_create(num_threads)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_loam_destroy(self.h)
            self.h = None

    def set_map(self, corner, surf):
        c, s = _f32(corner), _f32(surf)
        lib().orc_loam_set_map(self.h, _p(c), c.shape[0], c.strides[0], _p(s), s.shape[0], s.strides[0])

    def features(self, corner, surf, t6):
        c, s = _f32(corner), _f32(surf)
        t = np.ascontiguousarray(t6, dtype=np.float32)
        flags = np.zeros(len(c) + len(s), np.uint8)
        coeff = np.zeros((len(c) + len(s), 4), np.float32)
        n = lib().orc_loam_features(self.h, _p(c), c.shape[0], c.strides[0], _p(s), s.shape[0], s.strides[0], _p(t), _p(flags), _p(coeff))
        return n, flags, coeff

    def optimize(self, corner, surf, t6, iter_num=30):
        c, s = _f32(corner), _f32(surf)
        t = np.array(t6, dtype=np.float32)
        nsel, conv, deg = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        AtA = np.zeros((6, 6))
        it = lib().orc_loam_optimize(self.h, _p(c), c.shape[0], c.strides[0], _p(s), s.shape[0], s.strides[0], _p(t), iter_num,
                                     C.byref(nsel), C.byref(conv), C.byref(deg), _p(AtA))
        return t, dict(iters=it, n_sel=nsel.value, converged=bool(conv.value), degenerate=bool(deg.value), AtA=AtA)


# ------------------------------------------------------------------ pclomp::GeneralizedIterativeClosestPoint
def gicp_params(k_correspondences=20, gicp_epsilon=0.001, rotation_epsilon=2e-3, transformation_epsilon=5e-4, corr_dist_threshold=5.0,
                max_iterations=200, max_inner_iterations=20, num_threads=0) -> GicpParams:
    """Defaults of the constructor (gicp_omp.h:115-135)."""
    p = GicpParams()
    p.k_correspondences, p.gicp_epsilon, p.rotation_epsilon = k_correspondences, gicp_epsilon, rotation_epsilon
    p.transformation_epsilon, p.corr_dist_threshold = transformation_epsilon, corr_dist_threshold
    p.max_iterations, p.max_inner_iterations, p.num_threads = max_iterations, max_inner_iterations, num_threads
    return p


class OracleGicp:
    def __init__(self, **kw):
        self.params = gicp_params(**kw)
        self.h = lib().orc_gicp_create(C.byref(self.params))
        self.n_src = self.n_tgt = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_gicp_destroy(self.h)
            self.h = None

    def set_target(self, pts):
        pts = _f32(pts)
        self.n_tgt = pts.shape[0]
        lib().orc_gicp_set_target(self.h, _p(pts), pts.shape[0], pts.strides[0])

    def set_source(self, pts):
        pts = _f32(pts)
        self.n_src = pts.shape[0]
        lib().orc_gicp_set_source(self.h, _p(pts), pts.shape[0], pts.strides[0])

    def covariances(self, which):
        """which: 'source' | 'target' -> [n, 3, 3]"""
        n = self.n_tgt if which == "target" else self.n_src
        out = np.zeros((n, 3, 3))
        rc = lib().orc_gicp_covariances(self.h, 1 if which == "target" else 0, _p(out))
        if rc != 0:
            raise ValueError("cloud smaller than k_correspondences")
        return out

    def align(self, guess=None):
        g = np.ascontiguousarray((np.eye(4) if guess is None else np.asarray(guess)).T, dtype=np.float32)  # column-major
        fin = np.zeros(16, np.float32)
        r = GicpResult()
        rc = lib().orc_gicp_align(self.h, _p(g), _p(fin), C.byref(r))
        return rc, fin.reshape(4, 4).T.copy(), r

    def correspondences(self, trans=None, guess=None):
        t = np.ascontiguousarray((np.eye(4) if trans is None else np.asarray(trans)).T, dtype=np.float32)
        g = np.ascontiguousarray((np.eye(4) if guess is None else np.asarray(guess)).T, dtype=np.float32)
        idx = np.zeros(self.n_src, np.int32)
        maha = np.zeros((self.n_src, 3, 3), np.float32)
        d2 = np.zeros(self.n_src, np.float32)
        m = lib().orc_gicp_correspondences(self.h, _p(t), _p(g), _p(idx), _p(maha), _p(d2))
        return int(m), idx, maha, d2

    def estimate(self, trans=None):
        """estimateRigidTransformationBFGS on the correspondences of the last correspondences() call, from `trans`."""
        t = np.ascontiguousarray((np.eye(4) if trans is None else np.asarray(trans)).T, dtype=np.float32).reshape(-1)
        inner, status = C.c_int32(), C.c_int32()
        calls = np.zeros(3, np.int32)
        rc = lib().orc_gicp_estimate(self.h, _p(t), C.byref(inner), C.byref(status), _p(calls))
        return rc, t.reshape(4, 4).T.copy(), inner.value, status.value, calls

    def cost(self, x6):
        """operator()(x), fdf's f, df's gradient, fdf's gradient on the correspondences of the last correspondences() call"""
        x = np.ascontiguousarray(x6, dtype=np.float64)
        f0, f1 = C.c_double(), C.c_double()
        g0, g1 = np.zeros(6), np.zeros(6)
        lib().orc_gicp_cost(self.h, _p(x), C.byref(f0), C.byref(f1), _p(g0), _p(g1))
        return f0.value, f1.value, g0, g1


def gicp_apply_state(x6):
    x = np.ascontiguousarray(x6, dtype=np.float64)
    t = np.zeros(16, np.float32)
    lib().orc_gicp_apply_state(_p(x), _p(t))
    return t.reshape(4, 4).T.copy()


def exact_knn(pts, queries, k, brute=False):
    pts, queries = _f32(pts), _f32(queries)
    idx = np.zeros((len(queries), k), np.int32)
    d2 = np.zeros((len(queries), k), np.float32)
    lib().orc_gicp_knn(_p(pts), pts.shape[0], pts.strides[0], _p(queries), queries.shape[0], queries.strides[0], k, int(brute), _p(idx), _p(d2))
    return idx, d2


def bfgs_test(x0, max_inner=20):
    x = np.array(x0, dtype=np.float64)
    inner = C.c_int32()
    calls = np.zeros(3, np.int32)
    st = lib().orc_bfgs_test(_p(x), max_inner, C.byref(inner), _p(calls))
    return int(st), x, inner.value, calls
