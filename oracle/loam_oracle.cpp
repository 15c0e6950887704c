// TEST INFRASTRUCTURE ONLY — CPU oracle for the LOAM-style scan-to-map optimisation of jueying_slam.
// PARITY UNPINNED (see oracle.h).  Restates jueying_slam/src/mapOptmization.cpp:
//   updatePointAssociateToMap / pointAssociateToMap / trans2Affine3f     :433-445,487-490
//   cornerOptimization                                                  :1255-1347
//   surfOptimization                                                    :1349-1419
//   combineOptimizationCoeffs                                           :1420-1440
//   LMOptimization                                                      :1442-1558
//   scan2MapOptimization (the iteration loop, without transformUpdate)  :1560-1590
// Third-party arithmetic restated from published behaviour (none of it is in the tree, versions unpinned):
//   pcl::KdTreeFLANN::nearestKSearch(5) = exact 5 nearest neighbours, ascending squared distance (brute force here);
//   pcl::getTransformation (PCL common/eigen.hpp): the explicit ZYX matrix; float sin/cos as correctly rounded values;
//   Eigen colPivHouseholderQr 5x3 float (smallmat.h, shared with the LIO oracle);
//   cv::eigen on symmetric float matrices: eigenvalues descending, eigenvectors in rows - restated as a cyclic Jacobi
//   iteration in the matrix' own precision; cv::gemm on CV_32F accumulates in double; cv::solve(DECOMP_QR) on the 6x6
//   normal equations - restated as pivoted elimination in double on the float matrix, narrowed to float; cv::Mat::inv
//   of the (orthogonal) eigenvector matrix = its transpose.
#include "oracle.h"
#include "smallmat.h"

#include <omp.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {
struct P3 { float x, y, z; };

inline float fsin(float a) { return (float)std::sin((double)a); }
inline float fcos(float a) { return (float)std::cos((double)a); }

// pcl::getTransformation(x, y, z, roll, pitch, yaw) as a row-major 3x4 float matrix
void pose_matrix(const float* t6, float* T) {
    const float roll = t6[0], pitch = t6[1], yaw = t6[2];
    const float A = fcos(yaw), B = fsin(yaw), C = fcos(pitch), D = fsin(pitch), E = fcos(roll), F = fsin(roll), DE = D * E, DF = D * F;
    T[0] = A * C; T[1] = A * DF - B * E; T[2] = B * F + A * DE; T[3] = t6[3];
    T[4] = B * C; T[5] = A * E + B * DF; T[6] = B * DE - A * F; T[7] = t6[4];
    T[8] = -D;    T[9] = C * F;          T[10] = C * E;         T[11] = t6[5];
}
inline P3 associate(const float* T, const P3& p) {
    return P3{T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3], T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7], T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11]};
}
// exact 5-NN, ascending (distance, index); returns false when fewer than 5 points exist
bool knn5(const std::vector<P3>& m, const P3& q, int* idx, float* d2) {
    if (m.size() < 5) return false;
    float bd[5] = {3.4e38f, 3.4e38f, 3.4e38f, 3.4e38f, 3.4e38f};
    int bi[5] = {-1, -1, -1, -1, -1};
    for (size_t i = 0; i < m.size(); ++i) {
        const float dx = m[i].x - q.x, dy = m[i].y - q.y, dz = m[i].z - q.z;
        const float d = (dx * dx + dy * dy) + dz * dz;
        if (d < bd[4]) {
            int k = 4;
            while (k > 0 && bd[k - 1] > d) { bd[k] = bd[k - 1]; bi[k] = bi[k - 1]; --k; }
            bd[k] = d; bi[k] = (int)i;
        }
    }
    for (int k = 0; k < 5; ++k) { idx[k] = bi[k]; d2[k] = bd[k]; }
    return true;
}
// cyclic Jacobi on a symmetric float matrix; eigenvalues descending, eigenvectors in the ROWS of V (cv::eigen layout)
template <int N>
void eigen_desc_f(const float* Ain, float* w, float* V) {
    float A[N * N], U[N * N];
    for (int i = 0; i < N * N; ++i) A[i] = Ain[i];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) U[i * N + j] = (i == j) ? 1.0f : 0.0f;
    for (int sweep = 0; sweep < 30; ++sweep) {
        float off = 0.f, diag = 0.f;
        for (int i = 0; i < N; ++i) {
            diag += A[i * N + i] * A[i * N + i];
            for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
        }
        if (off <= 1e-30f || off <= 1e-14f * diag) break;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                const float apq = A[p * N + q];
                if (apq == 0.0f) continue;
                const float theta = (A[q * N + q] - A[p * N + p]) / (2.0f * apq);
                const float t = (theta >= 0 ? 1.0f : -1.0f) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0f));
                const float c = 1.0f / std::sqrt(t * t + 1.0f), s = t * c;
                for (int k = 0; k < N; ++k) {
                    const float akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - s * akq;
                    A[k * N + q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; ++k) {
                    const float apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - s * aqk;
                    A[q * N + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; ++k) {
                    const float ukp = U[k * N + p], ukq = U[k * N + q];
                    U[k * N + p] = c * ukp - s * ukq;
                    U[k * N + q] = s * ukp + c * ukq;
                }
            }
    }
    int order[N];
    for (int i = 0; i < N; ++i) order[i] = i;
    for (int i = 0; i < N - 1; ++i) {  // descending, stable selection
        int m = i;
        for (int j = i + 1; j < N; ++j)
            if (A[order[j] * N + order[j]] > A[order[m] * N + order[m]]) m = j;
        std::swap(order[i], order[m]);
    }
    for (int i = 0; i < N; ++i) {
        w[i] = A[order[i] * N + order[i]];
        for (int k = 0; k < N; ++k) V[i * N + k] = U[k * N + order[i]];  // row i = eigenvector i
    }
}

struct Loam {
    std::vector<P3> corner_map, surf_map;
    int nthreads = 1;
    bool degenerate = false;
    float matP[36];

    // one cornerOptimization + surfOptimization + combineOptimizationCoeffs pass; rows = (point, coeff4) of the selected points
    void features(const std::vector<P3>& corner, const std::vector<P3>& surf, const float* t6, std::vector<P3>& ori, std::vector<float>& coeff,
                  std::vector<uint8_t>* flags_out) {
        float T[12];
        pose_matrix(t6, T);
        const size_t nc = corner.size(), ns = surf.size();
        std::vector<uint8_t> flag(nc + ns, 0);
        std::vector<float> co((nc + ns) * 4, 0.f);
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 16)
        for (size_t i = 0; i < nc; ++i) {
            const P3 sel = associate(T, corner[i]);
            int idx[5];
            float d2[5];
            if (!knn5(corner_map, sel, idx, d2) || !(d2[4] < 1.0)) continue;
            float cx = 0, cy = 0, cz = 0;
            for (int j = 0; j < 5; ++j) { cx += corner_map[idx[j]].x; cy += corner_map[idx[j]].y; cz += corner_map[idx[j]].z; }
            cx /= 5; cy /= 5; cz /= 5;
            float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
            for (int j = 0; j < 5; ++j) {
                const float ax = corner_map[idx[j]].x - cx, ay = corner_map[idx[j]].y - cy, az = corner_map[idx[j]].z - cz;
                a11 += ax * ax; a12 += ax * ay; a13 += ax * az; a22 += ay * ay; a23 += ay * az; a33 += az * az;
            }
            a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
            const float A[9] = {a11, a12, a13, a12, a22, a23, a13, a23, a33};
            float D[3], V[9];
            eigen_desc_f<3>(A, D, V);
            if (!(D[0] > 3 * D[1])) continue;
            const float x0 = sel.x, y0 = sel.y, z0 = sel.z;
            const float x1 = cx + 0.1 * V[0], y1 = cy + 0.1 * V[1], z1 = cz + 0.1 * V[2];
            const float x2 = cx - 0.1 * V[0], y2 = cy - 0.1 * V[1], z2 = cz - 0.1 * V[2];
            const float m11 = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1), m22 = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1),
                        m33 = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
            const float a012 = std::sqrt(m11 * m11 + m22 * m22 + m33 * m33);
            const float l12 = std::sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
            const float la = ((y1 - y2) * m11 + (z1 - z2) * m22) / a012 / l12;
            const float lb = -((x1 - x2) * m11 - (z1 - z2) * m33) / a012 / l12;
            const float lc = -((x1 - x2) * m22 + (y1 - y2) * m33) / a012 / l12;
            const float ld2 = a012 / l12;
            const float s = 1 - 0.9 * std::fabs(ld2);
            if (s > 0.1) {
                flag[i] = 1;
                co[i * 4] = s * la; co[i * 4 + 1] = s * lb; co[i * 4 + 2] = s * lc; co[i * 4 + 3] = s * ld2;
            }
        }
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 16)
        for (size_t i = 0; i < ns; ++i) {
            const P3 sel = associate(T, surf[i]);
            int idx[5];
            float d2[5];
            if (!knn5(surf_map, sel, idx, d2) || !(d2[4] < 1.0)) continue;
            float nb[15], rhs[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};
            for (int j = 0; j < 5; ++j) { nb[j * 3] = surf_map[idx[j]].x; nb[j * 3 + 1] = surf_map[idx[j]].y; nb[j * 3 + 2] = surf_map[idx[j]].z; }
            float X[3];
            orc::colpiv_qr_solve3<float>(nb, 5, rhs, X);
            float pa = X[0], pb = X[1], pc = X[2], pd = 1;
            const float ps = std::sqrt(pa * pa + pb * pb + pc * pc);
            pa /= ps; pb /= ps; pc /= ps; pd /= ps;
            bool valid = true;
            for (int j = 0; j < 5; ++j)
                if (std::fabs(pa * surf_map[idx[j]].x + pb * surf_map[idx[j]].y + pc * surf_map[idx[j]].z + pd) > 0.2) { valid = false; break; }
            if (!valid) continue;
            const float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
            const float s = 1 - 0.9 * std::fabs(pd2) / std::sqrt(std::sqrt(sel.x * sel.x + sel.y * sel.y + sel.z * sel.z));
            if (s > 0.1) {
                flag[nc + i] = 1;
                co[(nc + i) * 4] = s * pa; co[(nc + i) * 4 + 1] = s * pb; co[(nc + i) * 4 + 2] = s * pc; co[(nc + i) * 4 + 3] = s * pd2;
            }
        }
        ori.clear();
        coeff.clear();
        for (size_t i = 0; i < nc + ns; ++i)
            if (flag[i]) {
                ori.push_back(i < nc ? corner[i] : surf[i - nc]);
                for (int k = 0; k < 4; ++k) coeff.push_back(co[i * 4 + k]);
            }
        if (flags_out) *flags_out = flag;
    }

    // LMOptimization; returns true when converged.  AtA/AtB of the step are reported for parity.
    bool lm(int iterCount, const std::vector<P3>& ori, const std::vector<float>& coeff, float* t6, double* AtA_out, double* AtB_out) {
        const float srx = fsin(t6[1]), crx = fcos(t6[1]), sry = fsin(t6[2]), cry = fcos(t6[2]), srz = fsin(t6[0]), crz = fcos(t6[0]);
        const int n = (int)ori.size();
        if (n < 50) return false;
        double AtA[36], AtB[6];
        for (int i = 0; i < 36; ++i) AtA[i] = 0;
        for (int i = 0; i < 6; ++i) AtB[i] = 0;
        for (int i = 0; i < n; ++i) {
            const float px = ori[i].y, py = ori[i].z, pz = ori[i].x;
            const float cx = coeff[i * 4 + 1], cy = coeff[i * 4 + 2], cz = coeff[i * 4], ci = coeff[i * 4 + 3];
            const float arx = (crx * sry * srz * px + crx * crz * sry * py - srx * sry * pz) * cx + (-srx * srz * px - crz * srx * py - crx * pz) * cy +
                              (crx * cry * srz * px + crx * cry * crz * py - cry * srx * pz) * cz;
            const float ary = ((cry * srx * srz - crz * sry) * px + (sry * srz + cry * crz * srx) * py + crx * cry * pz) * cx +
                              ((-cry * crz - srx * sry * srz) * px + (cry * srz - crz * srx * sry) * py - crx * sry * pz) * cz;
            const float arz = ((crz * srx * sry - cry * srz) * px + (-cry * crz - srx * sry * srz) * py) * cx + (crx * crz * px - crx * srz * py) * cy +
                              ((sry * srz + cry * crz * srx) * px + (crz * sry - cry * srx * srz) * py) * cz;
            const float row[6] = {arz, arx, ary, cz, cx, cy};
            const float b = -ci;
            for (int r = 0; r < 6; ++r) {
                for (int c = 0; c < 6; ++c) AtA[r * 6 + c] += (double)row[r] * (double)row[c];
                AtB[r] += (double)row[r] * (double)b;
            }
        }
        float Af[36], Bf[6];
        for (int i = 0; i < 36; ++i) Af[i] = (float)AtA[i];
        for (int i = 0; i < 6; ++i) Bf[i] = (float)AtB[i];
        if (AtA_out) for (int i = 0; i < 36; ++i) AtA_out[i] = Af[i];
        if (AtB_out) for (int i = 0; i < 6; ++i) AtB_out[i] = Bf[i];
        // cv::solve(matAtA, matAtB, matX, DECOMP_QR)
        double M[6][7];
        for (int r = 0; r < 6; ++r) { for (int c = 0; c < 6; ++c) M[r][c] = Af[r * 6 + c]; M[r][6] = Bf[r]; }
        for (int k = 0; k < 6; ++k) {
            int piv = k;
            for (int r = k + 1; r < 6; ++r) if (std::fabs(M[r][k]) > std::fabs(M[piv][k])) piv = r;
            if (piv != k) for (int c = 0; c < 7; ++c) std::swap(M[k][c], M[piv][c]);
            for (int r = k + 1; r < 6; ++r) {
                const double f = M[r][k] / M[k][k];
                for (int c = k; c < 7; ++c) M[r][c] -= f * M[k][c];
            }
        }
        double xs[6];
        for (int r = 5; r >= 0; --r) {
            double s = M[r][6];
            for (int c = r + 1; c < 6; ++c) s -= M[r][c] * xs[c];
            xs[r] = s / M[r][r];
        }
        float X[6];
        for (int i = 0; i < 6; ++i) X[i] = (float)xs[i];
        if (iterCount == 0) {
            float E[6], V[36], V2[36];
            eigen_desc_f<6>(Af, E, V);
            std::memcpy(V2, V, sizeof V);
            degenerate = false;
            for (int i = 5; i >= 0; --i) {
                if (E[i] < 100.0f) {
                    for (int j = 0; j < 6; ++j) V2[i * 6 + j] = 0;
                    degenerate = true;
                } else break;
            }
            for (int r = 0; r < 6; ++r)  // matP = matV.inv() * matV2, V orthogonal: inverse = transpose
                for (int c = 0; c < 6; ++c) {
                    float s = 0;
                    for (int k = 0; k < 6; ++k) s += V[k * 6 + r] * V2[k * 6 + c];
                    matP[r * 6 + c] = s;
                }
        }
        if (degenerate) {
            float X2[6];
            std::memcpy(X2, X, sizeof X);
            for (int r = 0; r < 6; ++r) {
                float s = 0;
                for (int k = 0; k < 6; ++k) s += matP[r * 6 + k] * X2[k];
                X[r] = s;
            }
        }
        for (int i = 0; i < 6; ++i) t6[i] += X[i];
        const float r2d = 180.0f / 3.14159265358979323846f;
        const float deltaR = std::sqrt((X[0] * r2d) * (X[0] * r2d) + (X[1] * r2d) * (X[1] * r2d) + (X[2] * r2d) * (X[2] * r2d));
        const float deltaT = std::sqrt((X[3] * 100) * (X[3] * 100) + (X[4] * 100) * (X[4] * 100) + (X[5] * 100) * (X[5] * 100));
        return deltaR < 0.01 && deltaT < 0.05;
    }
};
}  // namespace

struct orc_loam { Loam L; };

extern "C" {
orc_loam* orc_loam_create(int32_t num_threads) {
    orc_loam* h = new orc_loam();
    h->L.nthreads = num_threads > 0 ? num_threads : omp_get_max_threads();
    return h;
}
void orc_loam_destroy(orc_loam* h) { delete h; }
static void fill(std::vector<P3>& v, const float* xyz, int64_t n, int64_t stride) {
    v.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const float* p = (const float*)((const char*)xyz + i * stride);
        v[i] = P3{p[0], p[1], p[2]};
    }
}
/* kdtreeCornerFromMap / kdtreeSurfFromMap ->setInputCloud */
void orc_loam_set_map(orc_loam* h, const float* corner, int64_t nc, int64_t sc, const float* surf, int64_t ns, int64_t ss) {
    fill(h->L.corner_map, corner, nc, sc);
    fill(h->L.surf_map, surf, ns, ss);
}
/* one cornerOptimization + surfOptimization pass at transform t6 = (roll, pitch, yaw, x, y, z): flags[nc+ns], coeff[(nc+ns)*4] */
int32_t orc_loam_features(orc_loam* h, const float* corner, int64_t nc, int64_t sc, const float* surf, int64_t ns, int64_t ss, const float* t6,
                          uint8_t* flags, float* coeff4) {
    std::vector<P3> c, s, ori;
    fill(c, corner, nc, sc);
    fill(s, surf, ns, ss);
    std::vector<float> co;
    std::vector<uint8_t> fl;
    h->L.features(c, s, t6, ori, co, &fl);
    size_t k = 0;
    for (size_t i = 0; i < fl.size(); ++i) {
        flags[i] = fl[i];
        for (int j = 0; j < 4; ++j) coeff4[i * 4 + j] = fl[i] ? co[k * 4 + j] : 0.f;
        if (fl[i]) ++k;
    }
    return (int32_t)ori.size();
}
/* scan2MapOptimization's loop: t6 in/out; returns iterations run; stats: n_sel of the last pass, converged, degenerate */
int32_t orc_loam_optimize(orc_loam* h, const float* corner, int64_t nc, int64_t sc, const float* surf, int64_t ns, int64_t ss, float* t6,
                          int32_t iter_num, int32_t* n_sel, int32_t* converged, int32_t* degenerate, double* AtA_first) {
    std::vector<P3> c, s, ori;
    fill(c, corner, nc, sc);
    fill(s, surf, ns, ss);
    std::vector<float> co;
    int it = 0;
    bool conv = false;
    h->L.degenerate = false;
    for (; it < iter_num; ++it) {
        h->L.features(c, s, t6, ori, co, nullptr);
        if (n_sel) *n_sel = (int32_t)ori.size();
        if (h->L.lm(it, ori, co, t6, it == 0 ? AtA_first : nullptr, nullptr)) { conv = true; ++it; break; }
    }
    if (converged) *converged = conv ? 1 : 0;
    if (degenerate) *degenerate = h->L.degenerate ? 1 : 0;
    return it;
}
}
