// b200reg — per-point arithmetic of the point-to-plane measurement model.
//
// Every fp32 expression is written with explicit __f*_rn intrinsics in the evaluation order
// of the reference C++ (built -O3 without -march: SSE2 scalar/packed ops, no FMA), so a residual
// or Jacobian entry is the same IEEE value on the device as on the reference's CPU path:
//   common::esti_plane              jueying_lio/include/common_lib.h:186-243
//   Eigen::ColPivHouseholderQR      (Eigen 3.3 algorithm: pivoting with norm down-dating)
//   ObsModel residual / validity    jueying_lio/src/laser_mapping.cc:611-636
//   ObsModel Jacobian row           jueying_lio/src/laser_mapping.cc:674-697
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace b200 {

// --- arithmetic shims: one rounding per operation, never contracted -----------------------
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double fadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double fsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double fmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double fdiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double fsqrt(double a) { return __dsqrt_rn(a); }
template <class T> struct Lim;
template <> struct Lim<float> {
    __device__ static float eps() { return 1.1920928955078125e-07f; }
    __device__ static float tiny() { return 1.17549435e-38f; }
};
template <> struct Lim<double> {
    __device__ static double eps() { return 2.220446049250313e-16; }
    __device__ static double tiny() { return 2.2250738585072014e-308; }
};

// 4-wide fp32 dot in the SSE horizontal-add order (a0b0 + a2b2) + (a1b1 + a3b3)
__device__ __forceinline__ float dot4_sse(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3) {
    return fadd(fadd(fmul(a0, b0), fmul(a2, b2)), fadd(fmul(a1, b1), fmul(a3, b3)));
}

// Solve the (ROWS x 3) least-squares system A x = -1 by column-pivoting Householder QR.
// q[r][c] holds A on entry (destroyed).  ROWS is a compile-time 3, 4 or 5 so everything
// stays in registers.
template <class T, int ROWS>
__device__ __forceinline__ void qr_solve_neg1(T (&q)[ROWS][3], T (&x)[3]) {
    constexpr int COLS = 3;
    constexpr int SIZE = ROWS < COLS ? ROWS : COLS;
    T nu[3], nd[3], hc[3];
    int tr[3];
#pragma unroll
    for (int k = 0; k < COLS; ++k) {
        T s = T(0);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) s = fadd(s, fmul(q[r][k], q[r][k]));
        nd[k] = fsqrt(s);
        nu[k] = nd[k];
    }
    T maxn = nu[0];
    if (nu[1] > maxn) maxn = nu[1];
    if (nu[2] > maxn) maxn = nu[2];
    const T eps = Lim<T>::eps();
    const T th = fmul(maxn, eps);
    const T threshold_helper = fdiv(fmul(th, th), T(ROWS));
    const T downdate_thr = fsqrt(eps);
    int nonzero = SIZE;
#pragma unroll
    for (int k = 0; k < SIZE; ++k) {
        int big = k;
        T bigv = nu[k];
#pragma unroll
        for (int j = k + 1; j < COLS; ++j)
            if (nu[j] > bigv) { bigv = nu[j]; big = j; }
        if (nonzero == SIZE && fmul(bigv, bigv) < fmul(threshold_helper, T(ROWS - k))) nonzero = k;
        tr[k] = big;
        // column swap k <-> big (big >= k); written with selects to keep register indexing static
#pragma unroll
        for (int j = k + 1; j < COLS; ++j) {
            if (big == j) {
#pragma unroll
                for (int r = 0; r < ROWS; ++r) { T t = q[r][k]; q[r][k] = q[r][j]; q[r][j] = t; }
                T t = nu[k]; nu[k] = nu[j]; nu[j] = t;
                t = nd[k]; nd[k] = nd[j]; nd[j] = t;
            }
        }
        // Householder vector of q[k..][k]
        T tailSq = T(0);
#pragma unroll
        for (int r = k + 1; r < ROWS; ++r) tailSq = fadd(tailSq, fmul(q[r][k], q[r][k]));
        const T c0 = q[k][k];
        T tau, beta;
        if (ROWS - k == 1 || tailSq <= Lim<T>::tiny()) {
            tau = T(0);
            beta = c0;
#pragma unroll
            for (int r = k + 1; r < ROWS; ++r) q[r][k] = T(0);
        } else {
            beta = fsqrt(fadd(fmul(c0, c0), tailSq));
            if (c0 >= T(0)) beta = -beta;
            const T den = fsub(c0, beta);
#pragma unroll
            for (int r = k + 1; r < ROWS; ++r) q[r][k] = fdiv(q[r][k], den);
            tau = fdiv(fsub(beta, c0), beta);
        }
        hc[k] = tau;
        q[k][k] = beta;
        if (ROWS - k == 1) {
#pragma unroll
            for (int j = k + 1; j < COLS; ++j) q[k][j] = fmul(q[k][j], fsub(T(1), tau));
        } else if (tau != T(0)) {
#pragma unroll
            for (int j = k + 1; j < COLS; ++j) {
                T tmp = T(0);
#pragma unroll
                for (int r = k + 1; r < ROWS; ++r) tmp = fadd(tmp, fmul(q[r][k], q[r][j]));
                tmp = fadd(tmp, q[k][j]);
                q[k][j] = fsub(q[k][j], fmul(tau, tmp));
#pragma unroll
                for (int r = k + 1; r < ROWS; ++r) q[r][j] = fsub(q[r][j], fmul(fmul(tau, q[r][k]), tmp));
            }
        }
#pragma unroll
        for (int j = k + 1; j < COLS; ++j) {
            if (nu[j] != T(0)) {
                T temp = fdiv(fabs(q[k][j]), nu[j]);
                temp = fmul(fadd(T(1), temp), fsub(T(1), temp));
                temp = temp < T(0) ? T(0) : temp;
                const T ratio = fdiv(nu[j], nd[j]);
                const T temp2 = fmul(temp, fmul(ratio, ratio));
                if (temp2 <= downdate_thr) {
                    T s = T(0);
#pragma unroll
                    for (int r = k + 1; r < ROWS; ++r) s = fadd(s, fmul(q[r][j], q[r][j]));
                    nd[j] = fsqrt(s);
                    nu[j] = nd[j];
                } else {
                    nu[j] = fmul(nu[j], fsqrt(temp));
                }
            }
        }
    }
    // c = Q^T b with b = -1
    T c[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) c[r] = T(-1);
#pragma unroll
    for (int k = 0; k < SIZE; ++k) {
        if (k < nonzero) {
            const T tau = hc[k];
            if (ROWS - k == 1) {
                c[k] = fmul(c[k], fsub(T(1), tau));
            } else if (tau != T(0)) {
                T tmp = T(0);
#pragma unroll
                for (int r = k + 1; r < ROWS; ++r) tmp = fadd(tmp, fmul(q[r][k], c[r]));
                tmp = fadd(tmp, c[k]);
                c[k] = fsub(c[k], fmul(tau, tmp));
#pragma unroll
                for (int r = k + 1; r < ROWS; ++r) c[r] = fsub(c[r], fmul(fmul(tau, q[r][k]), tmp));
            }
        }
    }
    // back substitution on the leading nonzero x nonzero block, column oriented
#pragma unroll
    for (int i = SIZE - 1; i >= 0; --i) {
        if (i < nonzero) {
            c[i] = fdiv(c[i], q[i][i]);
#pragma unroll
            for (int r = 0; r < i; ++r) c[r] = fsub(c[r], fmul(c[i], q[r][i]));
        }
    }
    // undo the column permutation: perm = identity with transpositions applied on the right
    int perm[3] = {0, 1, 2};
#pragma unroll
    for (int k = 0; k < SIZE; ++k) {
        // swap(perm[k], perm[tr[k]]) with tr[k] >= k
#pragma unroll
        for (int j = k + 1; j < COLS; ++j)
            if (tr[k] == j) { int t = perm[k]; perm[k] = perm[j]; perm[j] = t; }
    }
    x[0] = x[1] = x[2] = T(0);
#pragma unroll
    for (int i = 0; i < SIZE; ++i) {
        if (i < nonzero) {
            const T v = c[i];
            if (perm[i] == 0) x[0] = v;
            else if (perm[i] == 1) x[1] = v;
            else x[2] = v;
        }
    }
}

template <int ROWS>
__device__ __noinline__ void plane_normal_double(const float4* nb, float (&nv)[3]) {
    double q[ROWS][3], x[3];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) { q[r][0] = nb[r].x; q[r][1] = nb[r].y; q[r][2] = nb[r].z; }
    qr_solve_neg1<double, ROWS>(q, x);
    nv[0] = (float)x[0]; nv[1] = (float)x[1]; nv[2] = (float)x[2];
}

// common::esti_plane (common_lib.h:186-243).  nb: m neighbour points (m in 3..5).  Writes the plane
// even when the threshold test fails (pca_result is filled before the loop at :235-240).
__device__ __forceinline__ bool esti_plane(const float4* nb, int m, float threshold, float4& plane) {
    float nv[3];
    if (m == 5) {
        float q[5][3];
#pragma unroll
        for (int r = 0; r < 5; ++r) { q[r][0] = nb[r].x; q[r][1] = nb[r].y; q[r][2] = nb[r].z; }
        qr_solve_neg1<float, 5>(q, nv);
    } else if (m == 4) {
        plane_normal_double<4>(nb, nv);
    } else {
        plane_normal_double<3>(nb, nv);
    }
    const float n = fsqrt(fadd(fadd(fmul(nv[0], nv[0]), fmul(nv[1], nv[1])), fmul(nv[2], nv[2])));
    plane.x = fdiv(nv[0], n);
    plane.y = fdiv(nv[1], n);
    plane.z = fdiv(nv[2], n);
    plane.w = (float)__ddiv_rn(1.0, (double)n);
    bool ok = true;
    for (int j = 0; j < m; ++j) {
        float d = dot4_sse(plane.x, plane.y, plane.z, plane.w, nb[j].x, nb[j].y, nb[j].z, 1.0f);
        if (fabsf(d) > threshold) ok = false;
    }
    return ok;
}

// constants of one ObsModel evaluation, derived from the filter state in fp64 and narrowed
// exactly where the reference narrows (laser_mapping.cc:602-603, 670-672)
struct PassConsts {
    float qx, qy, qz, qw;  // R_wl = (rot * offset_R_L_I).cast<float>()   (a quaternion)
    float tx, ty, tz;      // t_wl = (rot * offset_T_L_I + pos).cast<float>()
    float offR[9];         // offset_R_L_I.toRotationMatrix().cast<float>()
    float offt[3];         // offset_T_L_I.cast<float>()
    float Rt[9];           // rot.toRotationMatrix().transpose().cast<float>()
};

// p_w = R_wl * p_b + t_wl  (Eigen quaternion-vector product in fp32, laser_mapping.cc:611-612)
__device__ __forceinline__ float3 body_to_world(const PassConsts& c, float bx, float by, float bz) {
    float uvx = fsub(fmul(c.qy, bz), fmul(c.qz, by));
    float uvy = fsub(fmul(c.qz, bx), fmul(c.qx, bz));
    float uvz = fsub(fmul(c.qx, by), fmul(c.qy, bx));
    uvx = fadd(uvx, uvx); uvy = fadd(uvy, uvy); uvz = fadd(uvz, uvz);
    float cx = fsub(fmul(c.qy, uvz), fmul(c.qz, uvy));
    float cy = fsub(fmul(c.qz, uvx), fmul(c.qx, uvz));
    float cz = fsub(fmul(c.qx, uvy), fmul(c.qy, uvx));
    float3 w;
    w.x = fadd(fadd(fadd(bx, fmul(c.qw, uvx)), cx), c.tx);
    w.y = fadd(fadd(fadd(by, fmul(c.qw, uvy)), cy), c.ty);
    w.z = fadd(fadd(fadd(bz, fmul(c.qw, uvz)), cz), c.tz);
    return w;
}

__device__ __forceinline__ float dot3_seq(float a0, float a1, float a2, float b0, float b1, float b2) {
    return fadd(fadd(fmul(a0, b0), fmul(a1, b1)), fmul(a2, b2));
}

// Jacobian row of one effective point (laser_mapping.cc:674-697): row[0..2]=n, [3..5]=A,
// [6..8]=B, [9..11]=C (B, C only with extrinsic estimation)
__device__ __forceinline__ void jacobian_row(const PassConsts& c, float bx, float by, float bz, const float4& plane, bool ext,
                                             float (&row)[12]) {
    float pt[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) pt[r] = fadd(dot3_seq(c.offR[r * 3], c.offR[r * 3 + 1], c.offR[r * 3 + 2], bx, by, bz), c.offt[r]);
    float C[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) C[r] = dot3_seq(c.Rt[r * 3], c.Rt[r * 3 + 1], c.Rt[r * 3 + 2], plane.x, plane.y, plane.z);
    row[0] = plane.x; row[1] = plane.y; row[2] = plane.z;
    row[3] = fadd(fadd(fmul(0.0f, C[0]), fmul(-pt[2], C[1])), fmul(pt[1], C[2]));
    row[4] = fadd(fadd(fmul(pt[2], C[0]), fmul(0.0f, C[1])), fmul(-pt[0], C[2]));
    row[5] = fadd(fadd(fmul(-pt[1], C[0]), fmul(pt[0], C[1])), fmul(0.0f, C[2]));
    if (ext) {
        const float S[9] = {0.0f, -bz, by, bz, 0.0f, -bx, -by, bx, 0.0f};
        float M[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                M[r * 3 + k] = dot3_seq(S[r * 3], S[r * 3 + 1], S[r * 3 + 2], c.offR[k * 3], c.offR[k * 3 + 1], c.offR[k * 3 + 2]);
#pragma unroll
        for (int r = 0; r < 3; ++r) row[6 + r] = dot3_seq(M[r * 3], M[r * 3 + 1], M[r * 3 + 2], C[0], C[1], C[2]);
        row[9] = C[0]; row[10] = C[1]; row[11] = C[2];
    } else {
#pragma unroll
        for (int r = 6; r < 12; ++r) row[r] = 0.0f;
    }
}

}  // namespace b200
