// TEST INFRASTRUCTURE ONLY — CPU oracle for the per-point motion compensation of a scan.
// PARITY UNPINNED (see oracle.h).  Restates the backward-propagation half of ImuProcess::UndistortPcl
// (jueying_lio/include/imu_processing.hpp:175-177,247-284) given the IMU poses its forward half produced
// (IMUpose_, :180-241 - esekf::predict per IMU sample, which stays on the host), with Exp from so3_math.h:31-49 and
// Pose6D from common_lib.h:111-123.  Quirks kept: points are first sorted by their time offset (std::sort, unstable in the
// reference: ascending time with ties in input order here); a point whose time is not after the first pose's offset is
// left as it is; the earliest point is compensated again by every earlier segment, because the inner loop's `break`
// at pcl_out.points.begin() leaves the iterator on it (:279-281).
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

namespace {
struct Qd { double x, y, z, w; };
inline void cross3(const double* a, const double* b, double* r) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}
inline void qrot(const Qd& q, const double* v, double* r) {  // Eigen QuaternionBase::_transformVector
    double qv[3] = {q.x, q.y, q.z}, uv[3], c2[3];
    cross3(qv, v, uv);
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    cross3(qv, uv, c2);
    for (int i = 0; i < 3; ++i) r[i] = v[i] + q.w * uv[i] + c2[i];
}
inline void exp_so3(const double* w, double dt, double* R) {  // so3_math.h:31-49
    const double n = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (n > 0.0000001) {
        const double a[3] = {w[0] / n, w[1] / n, w[2] / n};
        const double K[9] = {0.0, -a[2], a[1], a[2], 0.0, -a[0], -a[1], a[0], 0.0};
        const double ang = n * dt, s = std::sin(ang), c1 = 1.0 - std::cos(ang);
        double cK[9], KK[9];  // Eye3 + sin * K + ((1 - cos) * K) * K
        for (int i = 0; i < 9; ++i) cK[i] = c1 * K[i];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) KK[i * 3 + j] = cK[i * 3] * K[j] + cK[i * 3 + 1] * K[3 + j] + cK[i * 3 + 2] * K[6 + j];
        for (int i = 0; i < 9; ++i) R[i] = (R[i] + s * K[i]) + KK[i];
    }
}
}  // namespace

extern "C" {
/* poses: K records of 22 doubles {offset_time, acc[3], gyr[3], vel[3], pos[3], rot[9] row-major} (IMUpose_);
 * x_end26: the state after the final predict (pos, rot xyzw, offset_R_L_I xyzw, offset_T_L_I, ...);
 * points: float records, x y z first, time offset in ms at float index time_index (pcl curvature), intensity at
 * intensity_index (< 0: none).  out_xyzi: n x 4 floats in time-sorted order; out_order: n source indices. */
void orc_undistort(const float* pts, int64_t n, int64_t stride, int32_t time_index, int32_t intensity_index, const double* poses22, int32_t K,
                   const double* x_end26, float* out_xyzi, int32_t* out_order) {
    auto rec = [&](int64_t i) { return (const float*)((const char*)pts + i * stride); };
    std::vector<int32_t> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return rec(a)[time_index] < rec(b)[time_index]; });
    std::vector<float> X((size_t)n * 3), T((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const float* p = rec(order[i]);
        X[i * 3] = p[0]; X[i * 3 + 1] = p[1]; X[i * 3 + 2] = p[2];
        T[i] = p[time_index];
        out_xyzi[i * 4 + 3] = intensity_index >= 0 ? p[intensity_index] : 0.0f;
        if (out_order) out_order[i] = order[i];
    }
    const double* pos_end = x_end26;
    const Qd rot_end{x_end26[3], x_end26[4], x_end26[5], x_end26[6]}, offR{x_end26[7], x_end26[8], x_end26[9], x_end26[10]};
    const Qd rot_end_c{-rot_end.x, -rot_end.y, -rot_end.z, rot_end.w}, offR_c{-offR.x, -offR.y, -offR.z, offR.w};
    const double* offT = x_end26 + 11;
    if (n > 0 && K >= 2) {
        int64_t it = n - 1;
        for (int kp = K - 1; kp != 0; --kp) {
            const double* head = poses22 + 22 * (kp - 1);
            const double* tail = poses22 + 22 * kp;
            const double* R_imu = head + 13;
            const double *vel = head + 7, *pos = head + 10, *acc = tail + 1, *gyr = tail + 4;
            for (; (double)T[it] / double(1000) > head[0]; --it) {
                const double dt = (double)T[it] / double(1000) - head[0];
                double E[9], Ri[9];
                exp_so3(gyr, dt, E);
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) Ri[i * 3 + j] = R_imu[i * 3] * E[j] + R_imu[i * 3 + 1] * E[3 + j] + R_imu[i * 3 + 2] * E[6 + j];
                const double Pi[3] = {X[it * 3], X[it * 3 + 1], X[it * 3 + 2]};
                double Tei[3], a[3], b[3], c[3], d[3], e[3];
                for (int i = 0; i < 3; ++i) Tei[i] = ((pos[i] + vel[i] * dt) + 0.5 * acc[i] * dt * dt) - pos_end[i];
                qrot(offR, Pi, a);
                for (int i = 0; i < 3; ++i) a[i] += offT[i];
                for (int i = 0; i < 3; ++i) b[i] = (Ri[i * 3] * a[0] + Ri[i * 3 + 1] * a[1] + Ri[i * 3 + 2] * a[2]) + Tei[i];
                qrot(rot_end_c, b, c);
                for (int i = 0; i < 3; ++i) d[i] = c[i] - offT[i];
                qrot(offR_c, d, e);
                X[it * 3] = (float)e[0]; X[it * 3 + 1] = (float)e[1]; X[it * 3 + 2] = (float)e[2];
                if (it == 0) break;
            }
        }
    }
    for (int64_t i = 0; i < n; ++i) {
        out_xyzi[i * 4] = X[i * 3]; out_xyzi[i * 4 + 1] = X[i * 3 + 1]; out_xyzi[i * 4 + 2] = X[i * 3 + 2];
    }
}
}
