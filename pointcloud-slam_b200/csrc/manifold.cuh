// b200reg — fp64 manifold arithmetic of the 23-DOF jueying_lio state, device side.
// Follows IKFoM's MTK: SO3 boxplus/boxminus (mtk/types/SOn.hpp:210-216,256-269), S2 with
// S2_typ == 1 and |g| = 9.809 (mtk/types/S2.hpp:131-242, use-ikfom.hpp:10), exp/log/A_matrix/
// cos_sinc_sqrt (mtk/src/mtkmath.hpp:149-287).  State layout (use-ikfom.hpp:14-15):
//   x[26]  = pos(3) rot(xyzw) offR(xyzw) offT(3) vel(3) bg(3) ba(3) grav(3)
//   dx[23] = pos rot offR offT vel bg ba grav(2)
#pragma once
#include <cuda_runtime.h>

namespace b200 {
namespace mf {

// The filter step runs these once per pass on a single thread, so their cost is the length of the dependent fp64 chain
// (divisions, sqrt, sin / cos / atan).  Out-of-line copies were measured SLOWER on B200: the first call into far code
// misses the instruction cache line by line (~30 cycles per instruction), inlined straight-line code streams.
#define MF_COLD __device__ inline

constexpr double kTol = 1e-11;
constexpr double kGravLen = 98090.0 / 10000.0;

struct Q { double x, y, z, w; };

__device__ inline Q qmul(const Q& a, const Q& b) {
    Q r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
__device__ inline Q qconj(const Q& q) { return Q{-q.x, -q.y, -q.z, q.w}; }
__device__ inline void cross(const double* a, const double* b, double* r) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}
// Eigen QuaternionBase::_transformVector
__device__ inline void qrot(const Q& q, const double* v, double* r) {
    double qv[3] = {q.x, q.y, q.z}, uv[3], c2[3];
    cross(qv, v, uv);
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    cross(qv, uv, c2);
    for (int i = 0; i < 3; ++i) r[i] = v[i] + q.w * uv[i] + c2[i];
}
// Eigen QuaternionBase::toRotationMatrix, row-major out
__device__ inline void qtoR(const Q& q, double* R) {
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
__device__ inline void hat(const double* v, double* M) {
    M[0] = 0; M[1] = -v[2]; M[2] = v[1];
    M[3] = v[2]; M[4] = 0; M[5] = -v[0];
    M[6] = -v[1]; M[7] = v[0]; M[8] = 0;
}
__device__ inline void mm3(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
__device__ inline void cos_sinc_sqrt(double x2, double& c, double& sinc) {
    const double taylor_n = 1.220703125e-4;  // sqrt(sqrt(DBL_EPSILON)) = 2^-13
    if (x2 >= taylor_n) {
        double x = sqrt(x2);
        c = cos(x);
        sinc = sin(x) / x;
        return;
    }
    const double inv[7] = {1 / 3., 1 / 4., 1 / 5., 1 / 6., 1 / 7., 1 / 8., 1 / 9.};
    double cosi = 1., s = 1.;
    double term = -1 / 2. * x2;
    for (int i = 0; i < 3; ++i) {
        cosi += term;
        term *= inv[2 * i];
        s += term;
        term *= -inv[2 * i + 1] * x2;
    }
    c = cosi;
    sinc = s;
}
MF_COLD Q so3_exp(const double* v, double scale_half) {
    double n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double c, sinc;
    cos_sinc_sqrt(scale_half * scale_half * n2, c, sinc);
    double mult = sinc * scale_half;
    return Q{mult * v[0], mult * v[1], mult * v[2], c};
}
MF_COLD void so3_log(const Q& q, double* r) {
    double nv = sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    if (nv < kTol) nv = kTol;
    double s = 2.0 / nv * atan(nv / q.w);
    r[0] = s * q.x; r[1] = s * q.y; r[2] = s * q.z;
}
MF_COLD void A_matrix(const double* v, double* res) {
    double sq = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double n = sqrt(sq);
    for (int i = 0; i < 9; ++i) res[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (n < kTol) return;
    double H[9], HH[9];
    hat(v, H);
    mm3(H, H, HH);
    double a = (1 - cos(n)) / sq, b = (1 - sin(n) / n) / sq;
    for (int i = 0; i < 9; ++i) res[i] = res[i] + a * H[i] + b * HH[i];
}
MF_COLD void S2_Bx(const double* vec, double* Bx /*3x2*/) {
    const double len = kGravLen;
    if (vec[0] + len > kTol) {
        // S2.hpp:166-200 divides ten times; the two denominators are inverted once instead (1 ulp apart, far inside
        // the parity tolerance) because each fp64 division is a ~40-instruction dependent chain on the filter's critical path
        const double rden = 1.0 / (len + vec[0]);
        constexpr double rlen = 1.0 / kGravLen;
        const double yz = -vec[2] * vec[1] * rden;
        Bx[0] = -vec[1] * rlen;
        Bx[1] = -vec[2] * rlen;
        Bx[2] = (len - vec[1] * vec[1] * rden) * rlen;
        Bx[3] = yz * rlen;
        Bx[4] = yz * rlen;
        Bx[5] = (len - vec[2] * vec[2] * rden) * rlen;
    } else {
        for (int i = 0; i < 6; ++i) Bx[i] = 0;
        Bx[3] = -1;
        Bx[4] = 1;
    }
}
MF_COLD void S2_boxplus(double* vec, const double* delta) {
    double Bx[6], Bu[3], R[9], r[3];
    S2_Bx(vec, Bx);
    for (int i = 0; i < 3; ++i) Bu[i] = Bx[i * 2] * delta[0] + Bx[i * 2 + 1] * delta[1];
    Q q = so3_exp(Bu, 0.5);
    qtoR(q, R);
    for (int i = 0; i < 3; ++i) r[i] = R[i * 3] * vec[0] + R[i * 3 + 1] * vec[1] + R[i * 3 + 2] * vec[2];
    vec[0] = r[0]; vec[1] = r[1]; vec[2] = r[2];
}
MF_COLD void S2_boxminus(const double* vec, const double* other, double* res) {
    double c[3];
    cross(vec, other, c);
    double v_sin = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    double v_cos = vec[0] * other[0] + vec[1] * other[1] + vec[2] * other[2];
    double theta = atan2(v_sin, v_cos);
    if (v_sin < kTol) {
        if (fabs(theta) > kTol) { res[0] = 3.1415926; res[1] = 0; }
        else { res[0] = 0; res[1] = 0; }
    } else {
        double Bx[6], hv[3];
        S2_Bx(other, Bx);
        cross(other, vec, hv);
        double f = theta / v_sin;
        for (int j = 0; j < 2; ++j) res[j] = (f * Bx[j]) * hv[0] + (f * Bx[2 + j]) * hv[1] + (f * Bx[4 + j]) * hv[2];
    }
}
MF_COLD void S2_Nx_yy(const double* vec, double* Nx /*2x3*/) {
    double Bx[6], H[9];
    S2_Bx(vec, Bx);
    hat(vec, H);
    double f = 1 / kGravLen / kGravLen;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j) Nx[i * 3 + j] = (f * Bx[i]) * H[j] + (f * Bx[2 + i]) * H[3 + j] + (f * Bx[4 + i]) * H[6 + j];
}
MF_COLD void S2_Mx(const double* vec, const double* delta, double* Mx /*3x2*/) {
    double Bx[6], H[9];
    S2_Bx(vec, Bx);
    hat(vec, H);
    if (sqrt(delta[0] * delta[0] + delta[1] * delta[1]) < kTol) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 2; ++j) Mx[i * 2 + j] = -H[i * 3] * Bx[j] + -H[i * 3 + 1] * Bx[2 + j] + -H[i * 3 + 2] * Bx[4 + j];
    } else {
        double Bu[3];
        for (int i = 0; i < 3; ++i) Bu[i] = Bx[i * 2] * delta[0] + Bx[i * 2 + 1] * delta[1];
        // S2.hpp:233 passes scalar(1 / 2) == 0 as the exp scale: the "rotation" is the identity
        double A[9], At[9], nH[9], T2[9];
        A_matrix(Bu, A);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) At[i * 3 + j] = A[j * 3 + i];
        for (int i = 0; i < 9; ++i) nH[i] = -H[i];
        mm3(nH, At, T2);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 2; ++j) Mx[i * 2 + j] = T2[i * 3] * Bx[j] + T2[i * 3 + 1] * Bx[2 + j] + T2[i * 3 + 2] * Bx[4 + j];
    }
}

__device__ inline Q ldq(const double* p) { return Q{p[0], p[1], p[2], p[3]}; }
__device__ inline void stq(double* p, const Q& q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }

// x (+) dx
__device__ inline void state_boxplus(double* x, const double* d) {
    for (int i = 0; i < 3; ++i) x[i] += d[i];
    stq(x + 3, qmul(ldq(x + 3), so3_exp(d + 3, 0.5)));
    stq(x + 7, qmul(ldq(x + 7), so3_exp(d + 6, 0.5)));
    for (int i = 0; i < 3; ++i) x[11 + i] += d[9 + i];
    for (int i = 0; i < 3; ++i) x[14 + i] += d[12 + i];
    for (int i = 0; i < 3; ++i) x[17 + i] += d[15 + i];
    for (int i = 0; i < 3; ++i) x[20 + i] += d[18 + i];
    S2_boxplus(x + 23, d + 21);
}
// r = a (-) b
__device__ inline void state_boxminus(const double* a, const double* b, double* r) {
    for (int i = 0; i < 3; ++i) r[i] = a[i] - b[i];
    so3_log(qmul(qconj(ldq(b + 3)), ldq(a + 3)), r + 3);
    so3_log(qmul(qconj(ldq(b + 7)), ldq(a + 7)), r + 6);
    for (int i = 0; i < 3; ++i) r[9 + i] = a[11 + i] - b[11 + i];
    for (int i = 0; i < 3; ++i) r[12 + i] = a[14 + i] - b[14 + i];
    for (int i = 0; i < 3; ++i) r[15 + i] = a[17 + i] - b[17 + i];
    for (int i = 0; i < 3; ++i) r[18 + i] = a[20 + i] - b[20 + i];
    S2_boxminus(a + 23, b + 23, r + 21);
}

}  // namespace mf
}  // namespace b200
