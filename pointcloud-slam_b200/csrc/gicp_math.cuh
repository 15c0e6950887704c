// b200reg — GICP per-point arithmetic and the BFGS minimiser (a-14).  Replaces the bodies of
// pclomp::GeneralizedIterativeClosestPoint (pointcloud_match/ndt_omp/include/pclomp/gicp_omp_impl.hpp):
//   cov_regularize      computeCovariances, the part after the k-NN sums                   :97-123
//   apply_state         applyState (ZYX Euler angles through three float quaternions)        :518-529
//   r_derivative        computeRDerivative + matricesInnerProd                               :127-184, gicp_omp.h:312-322
//   mahalanobis3        M = (R C1 R^T + C2)^-1 of the correspondence loop                    :438-450
//   point_terms         one correspondence of operator() / df / fdf                          :245-368
//   Bfgs<Eval>          pcl::BFGS<FunctorType> (PCL registration/bfgs.h = GSL vector_bfgs2 + linear_minimize; PCL is not in
//                       the reference tree, restated from its published source) driven as estimateRigidTransformationBFGS
//                       drives it                                                            :188-242
// Everything here is plain scalar code marked __host__ __device__: on the GPU every thread of the optimiser block runs the
// SAME scalar control flow on the same reduced sums (Eval is a block-wide collective), so there is no broadcast of
// decisions and no divergence; tests/helpers/gicp_host_harness.cpp compiles this very header with g++ and checks it
// against the oracle on the CPU (the kernels in gicp.cu only add the indexing around it).
#pragma once
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define G_HD __host__ __device__ __forceinline__
#else
#define G_HD inline
#endif

namespace b200 {
namespace gicp {

constexpr int kAcc = 14;  // f of operator(), f of fdf, g[3], R[9]

// ---- JacobiSVD<Matrix3d>(A, ComputeFullU): U (row-major) with singular values sorted descending.
// Two-sided Jacobi sweeps as in Eigen/src/SVD/JacobiSVD.h:666-792 (2x2 step: misc/RealSvd2x2.h:18-51, makeJacobi
// Jacobi/Jacobi.h:83-113); a square matrix takes no QR preconditioner.  V is not accumulated (ComputeFullU only).
G_HD void jacobi_svd3_u(const double* A, double* U, double* sv) {
    const double eps = DBL_EPSILON, dmin = DBL_MIN, precision = 2.0 * DBL_EPSILON;
    double scale = 0.0;
    for (int i = 0; i < 9; ++i) {
        const double a = fabs(A[i]);
        if (a > scale) scale = a;
    }
    if (!(scale <= DBL_MAX)) {  // NaN or Inf: Eigen leaves the decomposition unset; propagate NaN
        for (int i = 0; i < 9; ++i) U[i] = NAN;
        sv[0] = sv[1] = sv[2] = NAN;
        return;
    }
    if (scale == 0.0) scale = 1.0;
    double W[9];
    for (int i = 0; i < 9; ++i) {
        W[i] = A[i] / scale;
        U[i] = (i % 4 == 0) ? 1.0 : 0.0;
    }
    double max_diag = fmax(fabs(W[0]), fmax(fabs(W[4]), fabs(W[8])));
    (void)eps;
    bool finished = false;
    for (int sweep = 0; sweep < 64 && !finished; ++sweep) {  // Eigen loops until no rotation is needed; 64 sweeps is a watchdog
        finished = true;
        for (int p = 1; p < 3; ++p)
            for (int q = 0; q < p; ++q) {
                const double thr = fmax(dmin, precision * max_diag);
                if (!(fabs(W[p * 3 + q]) > thr || fabs(W[q * 3 + p]) > thr)) continue;
                finished = false;
                double m00 = W[p * 3 + p], m01 = W[p * 3 + q], m10 = W[q * 3 + p], m11 = W[q * 3 + q];
                double c1, s1;
                const double t = m00 + m11, d = m10 - m01;
                if (fabs(d) < dmin) {
                    s1 = 0.0;
                    c1 = 1.0;
                } else {
                    const double u = t / d;
                    const double tmp = sqrt(1.0 + u * u);
                    s1 = 1.0 / tmp;
                    c1 = u / tmp;
                }
                if (!(c1 == 1.0 && s1 == 0.0)) {
                    const double a0 = m00, a1 = m01, b0 = m10, b1 = m11;
                    m00 = c1 * a0 + s1 * b0;
                    m01 = c1 * a1 + s1 * b1;
                    m10 = -s1 * a0 + c1 * b0;
                    m11 = -s1 * a1 + c1 * b1;
                }
                double cr, sr;
                const double deno = 2.0 * fabs(m01);
                if (deno < dmin) {
                    cr = 1.0;
                    sr = 0.0;
                } else {
                    const double tau = (m00 - m11) / deno;
                    const double w = sqrt(tau * tau + 1.0);
                    const double tt = tau > 0.0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
                    const double sign_t = tt > 0.0 ? 1.0 : -1.0;
                    const double nn = 1.0 / sqrt(tt * tt + 1.0);
                    sr = -sign_t * (m01 / fabs(m01)) * fabs(tt) * nn;
                    cr = nn;
                }
                const double c2 = cr, s2 = -sr;  // j_left = rot1 * j_right^T
                const double cl = c1 * c2 - s1 * s2, sl = c1 * s2 + s1 * c2;
                if (!(cl == 1.0 && sl == 0.0)) {
                    for (int k = 0; k < 3; ++k) {
                        const double xi = W[p * 3 + k], yi = W[q * 3 + k];
                        W[p * 3 + k] = cl * xi + sl * yi;
                        W[q * 3 + k] = -sl * xi + cl * yi;
                    }
                    for (int k = 0; k < 3; ++k) {
                        const double xi = U[k * 3 + p], yi = U[k * 3 + q];
                        U[k * 3 + p] = cl * xi + sl * yi;
                        U[k * 3 + q] = -sl * xi + cl * yi;
                    }
                }
                if (!(cr == 1.0 && -sr == 0.0)) {
                    for (int k = 0; k < 3; ++k) {
                        const double xi = W[k * 3 + p], yi = W[k * 3 + q];
                        W[k * 3 + p] = cr * xi - sr * yi;
                        W[k * 3 + q] = sr * xi + cr * yi;
                    }
                }
                max_diag = fmax(max_diag, fmax(fabs(W[p * 3 + p]), fabs(W[q * 3 + q])));
            }
    }
    for (int i = 0; i < 3; ++i) {
        const double a = W[i * 3 + i];
        sv[i] = fabs(a);
        if (a < 0.0)
            for (int k = 0; k < 3; ++k) U[k * 3 + i] = -U[k * 3 + i];
    }
    for (int i = 0; i < 3; ++i) sv[i] *= scale;
    for (int i = 0; i < 3; ++i) {  // selection sort, descending, columns of U follow
        int pos = 0;
        double mx = sv[i];
        for (int j = 1; j < 3 - i; ++j)
            if (sv[i + j] > mx) {
                mx = sv[i + j];
                pos = j;
            }
        if (mx == 0.0) break;
        if (pos) {
            pos += i;
            const double ts = sv[i];
            sv[i] = sv[pos];
            sv[pos] = ts;
            for (int k = 0; k < 3; ++k) {
                const double tu = U[k * 3 + pos];
                U[k * 3 + pos] = U[k * 3 + i];
                U[k * 3 + i] = tu;
            }
        }
    }
}

// sums: mean[3] = sum x, y, z; c = sums of xx, yx, yy, zx, zy, zz (lower triangle, row by row) over the k neighbours, added in
// the neighbours' order.  out = the regularised covariance (row-major 3x3): singular values replaced by (1, 1, gicp_epsilon).
G_HD void cov_regularize(const double* mean_sum, const double* c, int k, double gicp_epsilon, double* out) {
    const double kd = (double)k;
    double mean[3] = {mean_sum[0] / kd, mean_sum[1] / kd, mean_sum[2] / kd};
    double cov[9];
    int t = 0;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b <= a; ++b) {
            double v = c[t++] / kd;
            v -= mean[a] * mean[b];
            cov[a * 3 + b] = v;
            cov[b * 3 + a] = v;
        }
    double U[9], sv[3];
    jacobi_svd3_u(cov, U, sv);
    for (int i = 0; i < 9; ++i) out[i] = 0.0;
    for (int kk = 0; kk < 3; ++kk) {
        const double v = kk == 2 ? gicp_epsilon : 1.0;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) out[i * 3 + j] += (v * U[i * 3 + kk]) * U[j * 3 + kk];
    }
}

// float sine / cosine as the correctly rounded float of the fp64 value: what differs least between libm versions and the
// device (same convention as the NDT angle tables).
G_HD float sin_f(float a) { return (float)sin((double)a); }
G_HD float cos_f(float a) { return (float)cos((double)a); }

struct Quatf {
    float x, y, z, w;
};
// Eigen's SSE float quaternion product (Eigen/src/Geometry/arch/Geometry_SIMD.h:33-45), lane by lane
G_HD Quatf quat_mul(const Quatf& a, const Quatf& b) {
    Quatf r;
    r.x = (a.x * b.w - a.z * b.y) + (a.y * b.z + a.w * b.x);
    r.y = (a.y * b.w - a.x * b.z) + (a.z * b.x + a.w * b.y);
    r.z = (a.z * b.w - a.y * b.x) + (a.x * b.y + a.w * b.z);
    r.w = (a.w * b.w - a.x * b.x) - (a.z * b.z + a.y * b.y);
    return r;
}

// transformation_matrix = Identity, applyState(transformation_matrix, x): T = 3x4 row-major {R | t}, float.
// base_transformation_ is the identity for the whole life of the object (gicp_omp_impl.hpp:396), so R * I = R exactly.
G_HD void apply_state(const double* x, float* T) {
    const float hz = 0.5f * (float)x[5], hy = 0.5f * (float)x[4], hx = 0.5f * (float)x[3];
    // Quaternion = AngleAxis: w = cos(ha), vec = sin(ha) * axis (Eigen/src/Geometry/Quaternion.h:561-569)
    const float sz = sin_f(hz), sy = sin_f(hy), sx = sin_f(hx);
    const Quatf qz = {sz * 0.0f, sz * 0.0f, sz * 1.0f, cos_f(hz)};
    const Quatf qy = {sy * 0.0f, sy * 1.0f, sy * 0.0f, cos_f(hy)};
    const Quatf qx = {sx * 1.0f, sx * 0.0f, sx * 0.0f, cos_f(hx)};
    const Quatf q = quat_mul(quat_mul(qz, qy), qx);
    // toRotationMatrix (Quaternion.h:600-621)
    const float tx = 2.0f * q.x, ty = 2.0f * q.y, tz = 2.0f * q.z;
    const float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const float txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const float tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    T[0] = 1.0f - (tyy + tzz);
    T[1] = txy - twz;
    T[2] = txz + twy;
    T[4] = txy + twz;
    T[5] = 1.0f - (txx + tzz);
    T[6] = tyz - twx;
    T[8] = txz - twy;
    T[9] = tyz + twx;
    T[10] = 1.0f - (txx + tyy);
    T[3] = 0.0f + (float)x[0];
    T[7] = 0.0f + (float)x[1];
    T[11] = 0.0f + (float)x[2];
}

// x from a transformation (estimateRigidTransformationBFGS, gicp_omp_impl.hpp:203-209): float atan2 / asin of float entries
G_HD void state_from_transform(const float* T /*3x4 row-major*/, double* x) {
    x[0] = T[3];
    x[1] = T[7];
    x[2] = T[11];
    x[3] = (double)(float)atan2((double)T[9], (double)T[10]);
    x[4] = (double)(float)asin((double)(-T[8]));
    x[5] = (double)(float)atan2((double)T[4], (double)T[0]);
}

// g[3..5] = <dR/dphi, R>, <dR/dtheta, R>, <dR/dpsi, R>; R row-major 3x3 (the accumulated p_src temp^T, times 2/m)
G_HD void r_derivative(const double* x, const double* R, double* g) {
    const double phi = x[3], theta = x[4], psi = x[5];
    const double cphi = cos(phi), sphi = sin(phi), ctheta = cos(theta), stheta = sin(theta), cpsi = cos(psi), spsi = sin(psi);
    double dphi[9], dth[9], dpsi[9];
    dphi[0] = 0.0; dphi[3] = 0.0; dphi[6] = 0.0;
    dphi[1] = sphi * spsi + cphi * cpsi * stheta;
    dphi[4] = -cpsi * sphi + cphi * spsi * stheta;
    dphi[7] = cphi * ctheta;
    dphi[2] = cphi * spsi - cpsi * sphi * stheta;
    dphi[5] = -cphi * cpsi - sphi * spsi * stheta;
    dphi[8] = -ctheta * sphi;
    dth[0] = -cpsi * stheta;
    dth[3] = -spsi * stheta;
    dth[6] = -ctheta;
    dth[1] = cpsi * ctheta * sphi;
    dth[4] = ctheta * sphi * spsi;
    dth[7] = -sphi * stheta;
    dth[2] = cphi * cpsi * ctheta;
    dth[5] = cphi * ctheta * spsi;
    dth[8] = -cphi * stheta;
    dpsi[0] = -ctheta * spsi;
    dpsi[3] = cpsi * ctheta;
    dpsi[6] = 0.0;
    dpsi[1] = -cphi * cpsi - sphi * spsi * stheta;
    dpsi[4] = -cphi * spsi + cpsi * sphi * stheta;
    dpsi[7] = 0.0;
    dpsi[2] = cpsi * sphi - cphi * spsi * stheta;
    dpsi[5] = sphi * spsi + cphi * cpsi * stheta;
    dpsi[8] = 0.0;
    // matricesInnerProd(mat1, mat2): r += mat1(j, i) * mat2(i, j), i outer, j inner
    double r0 = 0.0, r1 = 0.0, r2 = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            r0 += dphi[j * 3 + i] * R[i * 3 + j];
            r1 += dth[j * 3 + i] * R[i * 3 + j];
            r2 += dpsi[j * 3 + i] * R[i * 3 + j];
        }
    g[3] = r0;
    g[4] = r1;
    g[5] = r2;
}

G_HD void inverse3d(const double* m, double* inv) {  // Eigen compute_inverse_size3 (cofactors / determinant)
    const double c00 = m[4] * m[8] - m[5] * m[7];
    const double c10 = m[5] * m[6] - m[3] * m[8];
    const double c20 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 + m[1] * c10 + m[2] * c20;
    const double id = 1.0 / det;
    inv[0] = c00 * id;
    inv[1] = (m[2] * m[7] - m[1] * m[8]) * id;
    inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    inv[3] = c10 * id;
    inv[4] = (m[0] * m[8] - m[2] * m[6]) * id;
    inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    inv[6] = c20 * id;
    inv[7] = (m[1] * m[6] - m[0] * m[7]) * id;
    inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// transform_R = transformation_ * guess in double, rotation block (gicp_omp_impl.hpp:414-420): k = 0..3 accumulated from zero.
// Tt, G: 3x4 row-major float (bottom rows 0 0 0 1).
G_HD void rotation_of_product(const float* Tt, const float* G, double* R) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += (double)Tt[i * 4 + k] * (double)G[k * 4 + j];
            s += (double)Tt[i * 4 + 3] * 0.0;  // k = 3: guess(3, j) = 0
            R[i * 3 + j] = s;
        }
}

// mahalanobis_[i].block<3,3> = (R C1 R^T + C2)^-1 cast to float (gicp_omp_impl.hpp:438-450); products are left-to-right sums
G_HD void mahalanobis3(const double* R, const double* C1, const double* C2, float* M) {
    double A[9], B[9], inv[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i * 3 + j] = (R[i * 3] * C1[j] + R[i * 3 + 1] * C1[3 + j]) + R[i * 3 + 2] * C1[6 + j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) B[i * 3 + j] = ((A[i * 3] * R[j * 3] + A[i * 3 + 1] * R[j * 3 + 1]) + A[i * 3 + 2] * R[j * 3 + 2]) + C2[i * 3 + j];
    inverse3d(B, inv);
    for (int i = 0; i < 9; ++i) M[i] = (float)inv[i];
}

// One correspondence of the cost functor.  T: 3x4 row-major float; ps: source point as computeTransformation holds it
// (moved by the guess); pt: its target point; M: float 3x3 (the rest of the 4x4 Mahalanobis matrix is zero).
//   acc[0]    operator():  res = T ps - pt (float), ret = res . (maha res) in float, summed in double       :264-272
//   acc[1]    fdf:         res in double from the float difference, temp = M res, f += res . temp            :350-356
//   acc[2..4] df / fdf:    g.head<3>() += temp                                                               :312-324, 359
//   acc[5..13]             R += p_src3 temp^T with p_src3 = base_transformation_ ps = ps                      :320-325, 360-363
G_HD void point_terms(const float* T, float sx, float sy, float sz, float tx, float ty, float tz, const float* M, double* acc) {
    const float px = ((T[0] * sx + T[1] * sy) + T[2] * sz) + T[3];
    const float py = ((T[4] * sx + T[5] * sy) + T[6] * sz) + T[7];
    const float pz = ((T[8] * sx + T[9] * sy) + T[10] * sz) + T[11];
    const float rx = px - tx, ry = py - ty, rz = pz - tz;
    {
        const float mx = (M[0] * rx + M[1] * ry) + M[2] * rz;
        const float my = (M[3] * rx + M[4] * ry) + M[5] * rz;
        const float mz = (M[6] * rx + M[7] * ry) + M[8] * rz;
        const float ret = (rx * mx + ry * my) + rz * mz;
        acc[0] += (double)ret;
    }
    const double dx = (double)rx, dy = (double)ry, dz = (double)rz;
    const double t0 = ((double)M[0] * dx + (double)M[1] * dy) + (double)M[2] * dz;
    const double t1 = ((double)M[3] * dx + (double)M[4] * dy) + (double)M[5] * dz;
    const double t2 = ((double)M[6] * dx + (double)M[7] * dy) + (double)M[8] * dz;
    acc[1] += (dx * t0 + dy * t1) + dz * t2;
    acc[2] += t0;
    acc[3] += t1;
    acc[4] += t2;
    const double ps[3] = {(double)sx, (double)sy, (double)sz};
    const double tt[3] = {t0, t1, t2};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) acc[5 + i * 3 + j] += ps[i] * tt[j];
}

// what the functor returns from the reduced sums over m correspondences
G_HD double cost_f(const double* acc, int m) { return acc[0] / (double)m; }                 // operator()   :273-274
G_HD void cost_gradient(const double* acc, int m, const double* x, double* g) {             // df / fdf     :330-339, 365-368
    const double s = 2.0 / (double)m;
    double R[9];
    for (int i = 0; i < 3; ++i) g[i] = acc[2 + i] * s;
    for (int i = 0; i < 9; ++i) R[i] = acc[5 + i] * s;
    r_derivative(x, R, g);
}
G_HD double cost_f_fdf(const double* acc, int m) { return acc[1] / (double)m; }             // fdf          :364

// ---- pcl::BFGS<Functor> (PCL registration/bfgs.h; memoryless BFGS direction of GSL's vector_bfgs2 with Fletcher's line
// search).  Eval provides   double f(const double* x);   void df(const double* x, double* g);
//                            void fdf(const double* x, double& f, double* g);
enum BfgsStatus { kNegativeGradientEpsilon = -3, kNotStarted = -2, kRunning = -1, kSuccess = 0, kNoProgress = 1 };

template <class Eval, int N = 6>
struct Bfgs {
    Eval& ev;
    // parameters as estimateRigidTransformationBFGS sets them (gicp_omp_impl.hpp:221-226); the others are BFGS::Parameters' defaults
    double sigma = 0.01, rho = 0.01, tau1 = 9.0, tau2 = 0.05, tau3 = 0.5, step_size = 1.0;
    int order = 3, bracket_iters = 100, section_iters = 100;
    int n_f = 0, n_df = 0, n_fdf = 0;  // functor calls (statistics)
    // state
    double f, delta_f, fp0, g0norm, pnorm;
    double x0[N], g0[N], p[N], gradient[N], dx0[N], dg0[N];
    // wrapper (GSL "wrapper_t"): cached values along the line x0 + alpha p
    double f_alpha, df_alpha, x_alpha[N], g_alpha[N];
    double f_cache_key, df_cache_key, x_cache_key, g_cache_key;

    G_HD explicit Bfgs(Eval& e) : ev(e) {}

    G_HD static double dot(const double* a, const double* b) {
        double s = 0.0;
        for (int i = 0; i < N; ++i) s += a[i] * b[i];
        return s;
    }
    G_HD static double norm(const double* a) { return sqrt(dot(a, a)); }

    G_HD void move_to(double alpha) {
        if (alpha == x_cache_key) return;
        for (int i = 0; i < N; ++i) x_alpha[i] = x0[i] + alpha * p[i];
        x_cache_key = alpha;
    }
    G_HD double slope() const { return dot(g_alpha, p); }
    G_HD double apply_f(double alpha) {
        if (alpha == f_cache_key) return f_alpha;
        move_to(alpha);
        f_alpha = ev.f(x_alpha);
        ++n_f;
        f_cache_key = alpha;
        return f_alpha;
    }
    G_HD double apply_df(double alpha) {
        if (alpha == df_cache_key) return df_alpha;
        move_to(alpha);
        if (alpha != g_cache_key) {
            ev.df(x_alpha, g_alpha);
            ++n_df;
            g_cache_key = alpha;
        }
        df_alpha = slope();
        df_cache_key = alpha;
        return df_alpha;
    }
    G_HD void apply_fdf(double alpha, double& fo, double& dfo) {
        if (alpha == f_cache_key && alpha == df_cache_key) {
            fo = f_alpha;
            dfo = df_alpha;
            return;
        }
        if (alpha == f_cache_key || alpha == df_cache_key) {
            fo = apply_f(alpha);
            dfo = apply_df(alpha);
            return;
        }
        move_to(alpha);
        ev.fdf(x_alpha, f_alpha, g_alpha);
        ++n_fdf;
        f_cache_key = alpha;
        g_cache_key = alpha;
        df_alpha = slope();
        df_cache_key = alpha;
        fo = f_alpha;
        dfo = df_alpha;
    }
    G_HD void update_position(double alpha, double* x, double& fo, double* g) {
        double fa, dfa;
        apply_fdf(alpha, fa, dfa);
        fo = fa;
        for (int i = 0; i < N; ++i) {
            x[i] = x_alpha[i];
            g[i] = g_alpha[i];
        }
    }
    G_HD void change_direction() {
        for (int i = 0; i < N; ++i) {
            x_alpha[i] = x0[i];
            g_alpha[i] = g0[i];
        }
        x_cache_key = 0.0;
        f_cache_key = 0.0;
        g_cache_key = 0.0;
        df_alpha = slope();
        df_cache_key = 0.0;
    }

    G_HD int minimize_init(const double* x) {
        delta_f = 0.0;
        for (int i = 0; i < N; ++i) dx0[i] = 0.0;
        ev.fdf(x, f, gradient);
        ++n_fdf;
        for (int i = 0; i < N; ++i) {
            x0[i] = x[i];
            g0[i] = gradient[i];
        }
        g0norm = norm(g0);
        for (int i = 0; i < N; ++i) p[i] = gradient[i] * (-1.0 / g0norm);
        pnorm = norm(p);
        fp0 = -g0norm;
        for (int i = 0; i < N; ++i) {
            x_alpha[i] = x0[i];
            g_alpha[i] = g0[i];
        }
        x_cache_key = 0.0;
        f_alpha = f;
        f_cache_key = 0.0;
        g_cache_key = 0.0;
        df_alpha = slope();
        df_cache_key = 0.0;
        return kNotStarted;
    }

    // PCL's interpolate(): the cubic branch is guarded by `order > 2 && !(fpb != fpa) && fpb != inf`, which holds only when
    // the two slopes are EQUAL - in practice every call takes the quadratic branch, whose curvature test reads `c > a`
    // (not `c > 0`).  Both are kept as published.
    G_HD static double interpolate(double a, double fa, double fpa, double b, double fb, double fpb, double xmin, double xmax, int order) {
        double y, ymin = (xmin - a) / (b - a), ymax = (xmax - a) / (b - a);
        if (ymin > ymax) {
            const double t = ymin;
            ymin = ymax;
            ymax = t;
        }
        if (order > 2 && !(fpb != fpa) && fpb != INFINITY) {
            fpa = fpa * (b - a);
            fpb = fpb * (b - a);
            const double eta = 3.0 * (fb - fa) - 2.0 * fpa - fpb;
            const double xi = fpa + fpb - 2.0 * (fb - fa);
            const double c0 = fa, c1 = fpa, c2 = eta, c3 = xi;
            y = ymin;
            double fmin = c0 + ymin * (c1 + ymin * (c2 + ymin * c3));  // poly_eval: Horner
            check_extremum(c0, c1, c2, c3, ymax, y, fmin);
            // roots of c1 + 2 c2 z + 3 c3 z^2 (PolynomialSolver<Scalar, 2>)
            const double q0 = c1, q1 = 2.0 * c2, q2 = 3.0 * c3;
            const double a2 = 2.0 * q2;
            const double disc = q1 * q1 - 4.0 * q0 * q2;
            if (0.0 < disc) {
                const double dr = sqrt(disc);
                double y0 = (-q1 - dr) / a2, y1 = (-q1 + dr) / a2;
                if (y0 > y1) {
                    const double t = y0;
                    y0 = y1;
                    y1 = t;
                }
                if (y0 > ymin && y0 < ymax) check_extremum(c0, c1, c2, c3, y0, y, fmin);
                if (y1 > ymin && y1 < ymax) check_extremum(c0, c1, c2, c3, y1, y, fmin);
            } else if (0.0 == disc) {
                const double y0 = -q1 / a2;
                if (y0 > ymin && y0 < ymax) check_extremum(c0, c1, c2, c3, y0, y, fmin);
            }
        } else {
            fpa = fpa * (b - a);
            const double fl = fa + ymin * (fpa + ymin * (fb - fa - fpa));
            const double fh = fa + ymax * (fpa + ymax * (fb - fa - fpa));
            const double c = 2.0 * (fb - fa - fpa);
            y = ymin;
            double fmin = fl;
            if (fh < fmin) {
                y = ymax;
                fmin = fh;
            }
            if (c > a) {
                const double z = -fpa / c;
                if (z > ymin && z < ymax) {
                    const double fz = fa + z * (fpa + z * (fb - fa - fpa));
                    if (fz < fmin) {
                        y = z;
                        fmin = fz;
                    }
                }
            }
        }
        return a + y * (b - a);
    }
    G_HD static void check_extremum(double c0, double c1, double c2, double c3, double z, double& zmin, double& fmin) {
        const double yv = c0 + z * (c1 + z * (c2 + z * c3));
        if (yv < fmin) {
            zmin = z;
            fmin = yv;
        }
    }

    G_HD int line_search(double alpha1, double& alpha_new) {
        double f0v, fp0v, falpha, falpha_prev, fpalpha = 0.0, fpalpha_prev, delta, alpha_next;
        double alpha = alpha1, alpha_prev = 0.0;
        double a, b, fa, fb, fpa, fpb;
        int i = 0;
        apply_fdf(0.0, f0v, fp0v);
        falpha_prev = f0v;
        fpalpha_prev = fp0v;
        a = 0.0;
        b = alpha;
        fa = f0v;
        fb = 0.0;
        fpa = fp0v;
        fpb = 0.0;
        while (i++ < bracket_iters) {  // bracketing
            falpha = apply_f(alpha);
            if (falpha > f0v + alpha * rho * fp0v || falpha >= falpha_prev) {
                a = alpha_prev;
                fa = falpha_prev;
                fpa = fpalpha_prev;
                b = alpha;
                fb = falpha;
                fpb = NAN;
                break;
            }
            fpalpha = apply_df(alpha);
            if (fabs(fpalpha) <= -sigma * fp0v) {
                alpha_new = alpha;
                return kSuccess;
            }
            if (fpalpha >= 0.0) {
                a = alpha;
                fa = falpha;
                fpa = fpalpha;
                b = alpha_prev;
                fb = falpha_prev;
                fpb = fpalpha_prev;
                break;
            }
            delta = alpha - alpha_prev;
            {
                const double lower = alpha + delta, upper = alpha + tau1 * delta;
                alpha_next = interpolate(alpha_prev, falpha_prev, fpalpha_prev, alpha, falpha, fpalpha, lower, upper, order);
            }
            alpha_prev = alpha;
            falpha_prev = falpha;
            fpalpha_prev = fpalpha;
            alpha = alpha_next;
        }
        while (i++ < section_iters) {  // sectioning of the bracket [a, b]
            delta = b - a;
            {
                const double lower = a + tau2 * delta, upper = b - tau3 * delta;
                alpha = interpolate(a, fa, fpa, b, fb, fpb, lower, upper, order);
            }
            falpha = apply_f(alpha);
            if ((a - alpha) * fpa <= DBL_EPSILON) return kNoProgress;  // roundoff prevents progress
            if (falpha > f0v + rho * alpha * fp0v || falpha >= fa) {
                b = alpha;
                fb = falpha;
                fpb = NAN;
            } else {
                fpalpha = apply_df(alpha);
                if (fabs(fpalpha) <= -sigma * fp0v) {
                    alpha_new = alpha;
                    return kSuccess;
                }
                if (((b - a) >= 0.0 && fpalpha >= 0.0) || ((b - a) <= 0.0 && fpalpha <= 0.0)) {
                    b = a;
                    fb = fa;
                    fpb = fpa;
                    a = alpha;
                    fa = falpha;
                    fpa = fpalpha;
                } else {
                    a = alpha;
                    fa = falpha;
                    fpa = fpalpha;
                }
            }
        }
        return kSuccess;
    }

    G_HD int minimize_one_step(double* x) {
        double alpha = 0.0, alpha1;
        const double f0v = f;
        if (pnorm == 0.0 || g0norm == 0.0 || fp0 == 0.0) return kNoProgress;
        if (delta_f < 0.0) {
            const double del = fmax(-delta_f, 10.0 * DBL_EPSILON * fabs(f0v));
            alpha1 = fmin(1.0, 2.0 * del / (-fp0));
        } else {
            alpha1 = fabs(step_size);
        }
        const int status = line_search(alpha1, alpha);
        if (status != kSuccess) return status;
        update_position(alpha, x, f, gradient);
        delta_f = f - f0v;
        // memoryless BFGS direction: p' = g1 - A dx - B dg
        for (int i = 0; i < N; ++i) {
            dx0[i] = x[i] - x0[i];
            dg0[i] = gradient[i] - g0[i];
        }
        const double dxg = dot(dx0, gradient), dgg = dot(dg0, gradient), dxdg = dot(dx0, dg0), dgnorm = norm(dg0);
        double A, B;
        if (dxdg != 0.0) {
            B = dxg / dxdg;
            A = -(1.0 + dgnorm * dgnorm / dxdg) * B + dgg / dxdg;
        } else {
            B = 0.0;
            A = 0.0;
        }
        for (int i = 0; i < N; ++i) p[i] = gradient[i] - A * dx0[i] - B * dg0[i];
        for (int i = 0; i < N; ++i) {
            g0[i] = gradient[i];
            x0[i] = x[i];
        }
        g0norm = norm(g0);
        pnorm = norm(p);
        const double dir = dot(p, gradient) > 0.0 ? -1.0 : 1.0;
        for (int i = 0; i < N; ++i) p[i] *= dir / pnorm;
        pnorm = norm(p);
        fp0 = dot(p, g0);
        change_direction();
        return kSuccess;
    }
    G_HD int test_gradient(double epsilon) const {
        if (epsilon < 0.0) return kNegativeGradientEpsilon;
        return g0norm < epsilon ? kSuccess : kRunning;
    }
};

// estimateRigidTransformationBFGS' loop (gicp_omp_impl.hpp:228-241).  x in / out; returns the solver status; *inner = steps taken
template <class Eval>
G_HD int minimize_rigid(Eval& ev, double* x, int max_inner_iterations, int* inner, int* calls3) {
    const double gradient_tol = 1e-2;
    Bfgs<Eval, 6> bfgs(ev);
    int inner_iterations = 0;
    int result = bfgs.minimize_init(x);
    result = kRunning;
    do {
        ++inner_iterations;
        result = bfgs.minimize_one_step(x);
        if (result) break;
        result = bfgs.test_gradient(gradient_tol);
    } while (result == kRunning && inner_iterations < max_inner_iterations);
    if (inner) *inner = inner_iterations;
    if (calls3) {
        calls3[0] = bfgs.n_f;
        calls3[1] = bfgs.n_df;
        calls3[2] = bfgs.n_fdf;
    }
    return result;
}

// delta of the outer loop (gicp_omp_impl.hpp:483-494); 3x4 row-major floats (the bottom rows are equal)
G_HD double transform_delta(const float* prev, const float* cur, double rotation_epsilon, double transformation_epsilon) {
    double delta = 0.0;
    for (int k = 0; k < 3; ++k)
        for (int l = 0; l < 4; ++l) {
            const double ratio = l < 3 ? 1.0 / rotation_epsilon : 1.0 / transformation_epsilon;
            const double c_delta = ratio * fabs((double)(prev[k * 4 + l] - cur[k * 4 + l]));
            if (c_delta > delta) delta = c_delta;
        }
    return delta;
}

}  // namespace gicp
}  // namespace b200
