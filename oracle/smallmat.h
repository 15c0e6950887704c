// TEST INFRASTRUCTURE ONLY — part of the CPU oracle (see oracle/README.md).
// Nothing under pointcloud-slam_b200/ may include this file.
//
// Tiny dense linear algebra used by the oracle restatements.  The reference uses Eigen.  The copy vendored in the
// reference (src/pointcloud_match/fast_gicp/thirdparty/Eigen) cannot be compiled - Eigen/Core and Eigen/src/Core are
// missing from the snapshot - but the decomposition sources ARE there, and every routine below restates
// the file it cites line by line (E/ = that directory's Eigen/src/):
//   - ColPivHouseholderQR::computeInPlace / _solve_impl   E/QR/ColPivHouseholderQR.h:482-584, 587-608 + E/Householder/Householder.h:43-103
//                                                          (makeHouseholder), 116-137 (applyHouseholderOnTheLeft) (call sites common_lib.h:208,223)
//   - Matrix::inverse() for n > 4 = partialPivLu().inverse()   E/LU/InverseImpl.h:22-31, E/LU/PartialPivLU.h:358-411 (unblocked_lu: pivot
//                                                          search, row swap, column scaling, rank-1 update), :525-548 (compute),
//                                                          :196, 225-245 (inverse = solve(Identity): P, L, U)     (esekfom.hpp:1685,1706)
//                                                          For 23 x 23 Eigen takes blocked_lu (:430-499, panels of 8 columns) whose
//                                                          trailing updates and triangular solves run through Core/products kernels
//                                                          (GEBP, absent from the snapshot): same pivots and the same factors in exact
//                                                          arithmetic, rounding order of the block updates unknowable -> unblocked form here.
//   - 3x3 inverse by cofactors                             E/LU/InverseImpl.h:125-176                      (vgc_impl:355,359)
//   - SelfAdjointEigenSolver<Matrix3d>::compute            E/Eigenvalues/SelfAdjointEigenSolver.h:414-461, 498-569, 823-893,
//                                                          E/Eigenvalues/Tridiagonalization.h:459-503, E/Jacobi/Jacobi.h:231-267,331-332
//                                                                                                         (vgc_impl:333)
//   - JacobiSVD<Matrix6d>(H, FullU|FullV).solve            E/SVD/JacobiSVD.h:666-796, E/misc/RealSvd2x2.h:18-51, E/Jacobi/Jacobi.h:83-113,
//                                                          E/SVD/SVDBase.h:149-157,198-205,308-318        (ndt_omp_impl.hpp:112-114)
// What stays unknowable without Eigen/src/Core: the order in which Eigen's SSE2 packet kernels add the terms of a
// reduction (dot products, products of small matrices) and numext::hypot (restated from Eigen 3.4's published
// MathFunctionsImpl.h).  Every reduction below is a plain left-to-right loop (documented contract).
#pragma once
#include <cmath>
#include <cstring>
#include <limits>
#include <algorithm>

namespace orc {

// ---------------------------------------------------------------------------
// Column-pivoting Householder QR solve of  A x = b  for an (rows x 3) system,
// rows in [3,5].  Follows Eigen 3.3 ColPivHouseholderQR::computeInPlace and
// _solve_impl step by step (norm down-dating included).
// ---------------------------------------------------------------------------
template <class T>
inline void colpiv_qr_solve3(const T* A_rowmajor, int rows, const T* b, T x[3]) {
    const int cols = 3;
    const int size = rows < cols ? rows : cols;
    T qr[5][3];
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) qr[r][c] = A_rowmajor[r * 3 + c];

    T normsUpdated[3], normsDirect[3], hCoeffs[3];
    int transp[3];
    for (int k = 0; k < cols; ++k) {
        T s = T(0);
        for (int r = 0; r < rows; ++r) s = s + qr[r][k] * qr[r][k];
        normsDirect[k] = std::sqrt(s);
        normsUpdated[k] = normsDirect[k];
    }
    T maxn = normsUpdated[0];
    for (int k = 1; k < cols; ++k)
        if (normsUpdated[k] > maxn) maxn = normsUpdated[k];
    const T eps = std::numeric_limits<T>::epsilon();
    T th = maxn * eps;
    const T threshold_helper = (th * th) / T(rows);
    const T norm_downdate_threshold = std::sqrt(eps);
    int nonzero_pivots = size;

    for (int k = 0; k < size; ++k) {
        int biggest = k;
        T bigv = normsUpdated[k];
        for (int j = k + 1; j < cols; ++j)
            if (normsUpdated[j] > bigv) { bigv = normsUpdated[j]; biggest = j; }
        T biggest_sq = bigv * bigv;
        if (nonzero_pivots == size && biggest_sq < threshold_helper * T(rows - k)) nonzero_pivots = k;
        transp[k] = biggest;
        if (k != biggest) {
            for (int r = 0; r < rows; ++r) std::swap(qr[r][k], qr[r][biggest]);
            std::swap(normsUpdated[k], normsUpdated[biggest]);
            std::swap(normsDirect[k], normsDirect[biggest]);
        }
        // makeHouseholderInPlace on qr[k..rows-1][k]
        T tailSq = T(0);
        for (int r = k + 1; r < rows; ++r) tailSq = tailSq + qr[r][k] * qr[r][k];
        T c0 = qr[k][k];
        T tau, beta;
        if (rows - k == 1 || tailSq <= std::numeric_limits<T>::min()) {
            tau = T(0);
            beta = c0;
            for (int r = k + 1; r < rows; ++r) qr[r][k] = T(0);
        } else {
            beta = std::sqrt(c0 * c0 + tailSq);
            if (c0 >= T(0)) beta = -beta;
            T denom = c0 - beta;
            for (int r = k + 1; r < rows; ++r) qr[r][k] = qr[r][k] / denom;
            tau = (beta - c0) / beta;
        }
        hCoeffs[k] = tau;
        qr[k][k] = beta;
        // apply H_k to the trailing columns
        if (rows - k == 1) {
            for (int j = k + 1; j < cols; ++j) qr[k][j] = qr[k][j] * (T(1) - tau);
        } else if (tau != T(0)) {
            for (int j = k + 1; j < cols; ++j) {
                T tmp = T(0);
                for (int r = k + 1; r < rows; ++r) tmp = tmp + qr[r][k] * qr[r][j];
                tmp = tmp + qr[k][j];
                qr[k][j] = qr[k][j] - tau * tmp;
                for (int r = k + 1; r < rows; ++r) qr[r][j] = qr[r][j] - (tau * qr[r][k]) * tmp;
            }
        }
        // norm down-date (LAPACK xGEQPF style, Eigen 3.3)
        for (int j = k + 1; j < cols; ++j) {
            if (normsUpdated[j] != T(0)) {
                T temp = std::fabs(qr[k][j]) / normsUpdated[j];
                temp = (T(1) + temp) * (T(1) - temp);
                temp = temp < T(0) ? T(0) : temp;
                T ratio = normsUpdated[j] / normsDirect[j];
                T temp2 = temp * (ratio * ratio);
                if (temp2 <= norm_downdate_threshold) {
                    T s = T(0);
                    for (int r = k + 1; r < rows; ++r) s = s + qr[r][j] * qr[r][j];
                    normsDirect[j] = std::sqrt(s);
                    normsUpdated[j] = normsDirect[j];
                } else {
                    normsUpdated[j] = normsUpdated[j] * std::sqrt(temp);
                }
            }
        }
    }
    int perm[3] = {0, 1, 2};
    for (int k = 0; k < size; ++k) std::swap(perm[k], perm[transp[k]]);

    x[0] = x[1] = x[2] = T(0);
    if (nonzero_pivots == 0) return;
    T c[5];
    for (int r = 0; r < rows; ++r) c[r] = b[r];
    for (int k = 0; k < nonzero_pivots; ++k) {
        T tau = hCoeffs[k];
        if (rows - k == 1) {
            c[k] = c[k] * (T(1) - tau);
        } else if (tau != T(0)) {
            T tmp = T(0);
            for (int r = k + 1; r < rows; ++r) tmp = tmp + qr[r][k] * c[r];
            tmp = tmp + c[k];
            c[k] = c[k] - tau * tmp;
            for (int r = k + 1; r < rows; ++r) c[r] = c[r] - (tau * qr[r][k]) * tmp;
        }
    }
    // upper-triangular back substitution, column oriented
    for (int i = nonzero_pivots - 1; i >= 0; --i) {
        c[i] = c[i] / qr[i][i];
        for (int r = 0; r < i; ++r) c[r] = c[r] - c[i] * qr[r][i];
    }
    for (int i = 0; i < nonzero_pivots; ++i) x[perm[i]] = c[i];
}

// ---------------------------------------------------------------------------
// n x n inverse through partial-pivoting LU (row-major, in place into out).
// Returns false when a pivot is exactly zero.
// ---------------------------------------------------------------------------
inline bool lu_inverse(const double* A, int n, double* out) {
    double lu[23 * 23];
    int piv[23];
    std::memcpy(lu, A, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i) piv[i] = i;
    bool ok = true;
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = std::fabs(lu[k * n + k]);
        for (int r = k + 1; r < n; ++r) {
            double v = std::fabs(lu[r * n + k]);
            if (v > best) { best = v; p = r; }
        }
        if (best == 0.0) { ok = false; continue; }
        if (p != k) {
            for (int c = 0; c < n; ++c) std::swap(lu[k * n + c], lu[p * n + c]);
            std::swap(piv[k], piv[p]);
        }
        double d = lu[k * n + k];
        for (int r = k + 1; r < n; ++r) {
            double f = lu[r * n + k] / d;
            lu[r * n + k] = f;
            for (int c = k + 1; c < n; ++c) lu[r * n + c] -= f * lu[k * n + c];
        }
    }
    // solve LU X = P I, column by column
    for (int col = 0; col < n; ++col) {
        double y[23];
        for (int r = 0; r < n; ++r) y[r] = (piv[r] == col) ? 1.0 : 0.0;
        for (int r = 0; r < n; ++r) {
            double s = y[r];
            for (int c = 0; c < r; ++c) s -= lu[r * n + c] * y[c];
            y[r] = s;
        }
        for (int r = n - 1; r >= 0; --r) {
            double s = y[r];
            for (int c = r + 1; c < n; ++c) s -= lu[r * n + c] * y[c];
            y[r] = s / lu[r * n + r];
        }
        for (int r = 0; r < n; ++r) out[r * n + col] = y[r];
    }
    return ok;
}

// 3x3 inverse by cofactors / determinant (Eigen's compute_inverse_size3).
inline void inverse3(const double m[9], double inv[9]) {
    double c00 = m[4] * m[8] - m[5] * m[7];
    double c10 = m[5] * m[6] - m[3] * m[8];
    double c20 = m[3] * m[7] - m[4] * m[6];
    double det = m[0] * c00 + m[1] * c10 + m[2] * c20;
    double id = 1.0 / det;
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

// ---------------------------------------------------------------------------
// Cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (n <= 6).
// A (row-major) is destroyed; w = eigenvalues ascending, V columns = vectors.
// ---------------------------------------------------------------------------
inline void jacobi_eig_sym(double* A, int n, double* w, double* V) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; ++i) {
            diag += A[i * n + i] * A[i * n + i];
            for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
        }
        if (off <= 1e-300 || off <= 1e-34 * diag) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double apq = A[p * n + q];
                if (apq == 0.0) continue;
                double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
    // sort ascending (selection sort, swap columns of V)
    for (int i = 0; i < n - 1; ++i) {
        int m = i;
        for (int j = i + 1; j < n; ++j)
            if (w[j] < w[m]) m = j;
        if (m != i) {
            std::swap(w[i], w[m]);
            for (int k = 0; k < n; ++k) std::swap(V[k * n + i], V[k * n + m]);
        }
    }
}

// Solve H x = rhs for symmetric 6x6 H the way JacobiSVD(H).solve(rhs) does:
// pseudo-inverse over singular values above max(sv)*6*eps.  For symmetric H
// the SVD is the eigen-decomposition with |lambda| as singular values.
inline void svd_solve_sym6(const double H[36], const double rhs[6], double x[6]) {
    double A[36], w[6], V[36];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) A[i * 6 + j] = 0.5 * (H[i * 6 + j] + H[j * 6 + i]);
    jacobi_eig_sym(A, 6, w, V);
    double smax = 0.0;
    for (int i = 0; i < 6; ++i) smax = std::max(smax, std::fabs(w[i]));
    double thr = std::max(smax * 6.0 * std::numeric_limits<double>::epsilon(), std::numeric_limits<double>::min());
    for (int i = 0; i < 6; ++i) x[i] = 0.0;
    for (int k = 0; k < 6; ++k) {
        if (std::fabs(w[k]) <= thr) continue;
        double d = 0.0;
        for (int i = 0; i < 6; ++i) d += V[i * 6 + k] * rhs[i];
        d /= w[k];
        for (int i = 0; i < 6; ++i) x[i] += V[i * 6 + k] * d;
    }
}

// ---------------------------------------------------------------------------
// Givens rotation, real case: JacobiRotation<double>::makeGivens (E/Jacobi/Jacobi.h:231-267)
// ---------------------------------------------------------------------------
inline void make_givens(double p, double q, double& c, double& s) {
    if (q == 0.0) {
        c = p < 0.0 ? -1.0 : 1.0;
        s = 0.0;
    } else if (p == 0.0) {
        c = 0.0;
        s = q < 0.0 ? 1.0 : -1.0;
    } else if (std::fabs(p) > std::fabs(q)) {
        double t = q / p;
        double u = std::sqrt(1.0 + t * t);
        if (p < 0.0) u = -u;
        c = 1.0 / u;
        s = -t * c;
    } else {
        double t = p / q;
        double u = std::sqrt(1.0 + t * t);
        if (q < 0.0) u = -u;
        s = -1.0 / u;
        c = -t * s;
    }
}
// numext::hypot for reals (Eigen 3.4 Core/MathFunctionsImpl.h positive_real_hypot - not in the snapshot, published form)
inline double eigen_hypot(double x, double y) {
    x = std::fabs(x);
    y = std::fabs(y);
    if (std::isinf(x) || std::isinf(y)) return std::numeric_limits<double>::infinity();
    if (std::isnan(x) || std::isnan(y)) return std::numeric_limits<double>::quiet_NaN();
    double p = x > y ? x : y;
    if (p == 0.0) return 0.0;
    double qp = (y < x ? y : x) / p;
    return p * std::sqrt(1.0 + qp * qp);
}

// ---------------------------------------------------------------------------
// SelfAdjointEigenSolver<Matrix3d>::compute(A, ComputeEigenvectors): eigenvalues ascending in w, eigenvectors in
// the columns of V (row-major storage here).  Only the lower triangle of A (row-major) is read.
//   scaling                   SelfAdjointEigenSolver.h:445-449
//   3x3 tridiagonalisation    Tridiagonalization.h:459-503
//   deflation / iteration     SelfAdjointEigenSolver.h:498-550 (m_maxIterations = 30, :375)
//   implicit QR step          SelfAdjointEigenSolver.h:823-893
//   ascending sort            SelfAdjointEigenSolver.h:551-567
// Returns false on NoConvergence (the eigenvalues are then left unsorted, as in Eigen).
// ---------------------------------------------------------------------------
inline bool eigen_selfadjoint3(const double* A, double w[3], double V[9]) {
    const double dmin = std::numeric_limits<double>::min();
    double m00 = A[0], m10 = A[3], m11 = A[4], m20 = A[6], m21 = A[7], m22 = A[8];
    double scale = 0.0;
    {
        const double l[6] = {m00, m10, m11, m20, m21, m22};
        for (int i = 0; i < 6; ++i) scale = std::fabs(l[i]) > scale ? std::fabs(l[i]) : scale;  // the strict upper part is zero
    }
    if (scale == 0.0) scale = 1.0;
    m00 /= scale; m10 /= scale; m11 /= scale; m20 /= scale; m21 /= scale; m22 /= scale;
    double diag[3], sub[2];
    double Q[3][3];  // Q[row][col]
    diag[0] = m00;
    const double v1norm2 = m20 * m20;
    if (v1norm2 <= dmin) {
        diag[1] = m11; diag[2] = m22; sub[0] = m10; sub[1] = m21;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Q[i][j] = i == j ? 1.0 : 0.0;
    } else {
        const double beta = std::sqrt(m10 * m10 + v1norm2);
        const double invBeta = 1.0 / beta;
        const double m01 = m10 * invBeta, m02 = m20 * invBeta;
        const double q = 2.0 * m01 * m21 + m02 * (m22 - m11);
        diag[1] = m11 + m02 * q;
        diag[2] = m22 - m02 * q;
        sub[0] = beta;
        sub[1] = m21 - m01 * q;
        Q[0][0] = 1; Q[0][1] = 0; Q[0][2] = 0;
        Q[1][0] = 0; Q[1][1] = m01; Q[1][2] = m02;
        Q[2][0] = 0; Q[2][1] = m02; Q[2][2] = -m01;
    }
    const int n = 3, maxIterations = 30;
    int end = n - 1, start = 0, iter = 0;
    const double precision_inv = 1.0 / std::numeric_limits<double>::epsilon();
    while (end > 0) {
        for (int i = start; i < end; ++i) {
            if (std::fabs(sub[i]) < dmin) {
                sub[i] = 0.0;
            } else {
                const double scaled = precision_inv * sub[i];
                if (scaled * scaled <= (std::fabs(diag[i]) + std::fabs(diag[i + 1]))) sub[i] = 0.0;
            }
        }
        while (end > 0 && sub[end - 1] == 0.0) end--;
        if (end <= 0) break;
        iter++;
        if (iter > maxIterations * n) break;
        start = end - 1;
        while (start > 0 && sub[start - 1] != 0.0) start--;
        // tridiagonal_qr_step
        double td = (diag[end - 1] - diag[end]) * 0.5;
        double e = sub[end - 1];
        double mu = diag[end];
        if (td == 0.0) {
            mu -= std::fabs(e);
        } else if (e != 0.0) {
            const double e2 = e * e;
            const double h = eigen_hypot(td, e);
            if (e2 == 0.0) mu -= e / ((td + (td > 0.0 ? h : -h)) / e);
            else mu -= e2 / (td + (td > 0.0 ? h : -h));
        }
        double x = diag[start] - mu;
        double z = sub[start];
        for (int k = start; k < end && z != 0.0; ++k) {
            double c, s;
            make_givens(x, z, c, s);
            const double sdk = s * diag[k] + c * sub[k];
            const double dkp1 = s * sub[k] + c * diag[k + 1];
            diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
            diag[k + 1] = s * sdk + c * dkp1;
            sub[k] = c * sdk - s * dkp1;
            if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
            x = sub[k];
            if (k < end - 1) {
                z = -s * sub[k + 1];
                sub[k + 1] = c * sub[k + 1];
            }
            // Q = Q * G: q.applyOnTheRight(k, k+1, rot) = apply_rotation_in_the_plane(col k, col k+1, rot.transpose())
            if (!(c == 1.0 && s == 0.0)) {
                for (int i = 0; i < 3; ++i) {
                    const double xi = Q[i][k], yi = Q[i][k + 1];
                    Q[i][k] = c * xi - s * yi;
                    Q[i][k + 1] = s * xi + c * yi;
                }
            }
        }
    }
    const bool ok = iter <= maxIterations * n;
    if (ok) {
        for (int i = 0; i < n - 1; ++i) {
            int k = 0;
            for (int j = 1; j < n - i; ++j)
                if (diag[i + j] < diag[i + k]) k = j;
            if (k > 0) {
                std::swap(diag[i], diag[k + i]);
                for (int r = 0; r < 3; ++r) std::swap(Q[r][i], Q[r][k + i]);
            }
        }
    }
    for (int i = 0; i < 3; ++i) w[i] = diag[i] * scale;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i * 3 + j] = Q[i][j];
    return ok;
}

// ---------------------------------------------------------------------------
// JacobiSVD<Matrix<double,6,6>> sv(H, ComputeFullU | ComputeFullV); x = sv.solve(rhs)   (ndt_omp_impl.hpp:112-114)
//   two-sided Jacobi sweeps        JacobiSVD.h:666-745
//   2x2 real SVD                   misc/RealSvd2x2.h:18-51, makeJacobi Jacobi.h:83-113, rotation product Jacobi.h:53-59
//   signs / scale / sort           JacobiSVD.h:747-792
//   rank (threshold 6 eps) + solve SVDBase.h:149-157, 198-205, 308-318
// H row-major.  The dot products of solve are plain left-to-right loops.
// ---------------------------------------------------------------------------
inline void jacobi_svd_solve6(const double H[36], const double rhs[6], double x[6], double* sv_out = nullptr) {
    const int n = 6;
    const double eps = std::numeric_limits<double>::epsilon(), dmin = std::numeric_limits<double>::min();
    const double precision = 2.0 * eps;
    double scale = 0.0;
    bool nan = false;
    for (int i = 0; i < 36; ++i) {
        const double a = std::fabs(H[i]);
        if (a != a) nan = true;
        if (a > scale) scale = a;
    }
    if (nan || !std::isfinite(scale)) {  // InvalidInput: Eigen returns without a decomposition
        for (int i = 0; i < 6; ++i) x[i] = std::numeric_limits<double>::quiet_NaN();
        return;
    }
    if (scale == 0.0) scale = 1.0;
    double W[6][6], U[6][6], V[6][6];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            W[i][j] = H[i * 6 + j] / scale;
            U[i][j] = V[i][j] = i == j ? 1.0 : 0.0;
        }
    double maxDiag = 0.0;
    for (int i = 0; i < n; ++i) maxDiag = std::fabs(W[i][i]) > maxDiag ? std::fabs(W[i][i]) : maxDiag;
    bool finished = false;
    while (!finished) {
        finished = true;
        for (int p = 1; p < n; ++p) {
            for (int q = 0; q < p; ++q) {
                const double thr = std::max(dmin, precision * maxDiag);
                if (std::fabs(W[p][q]) > thr || std::fabs(W[q][p]) > thr) {
                    finished = false;
                    // real_2x2_jacobi_svd
                    double m00 = W[p][p], m01 = W[p][q], m10 = W[q][p], m11 = W[q][q];
                    double c1, s1;
                    const double t = m00 + m11, d = m10 - m01;
                    if (std::fabs(d) < dmin) {
                        s1 = 0.0; c1 = 1.0;
                    } else {
                        const double u = t / d;
                        const double tmp = std::sqrt(1.0 + u * u);
                        s1 = 1.0 / tmp;
                        c1 = u / tmp;
                    }
                    if (!(c1 == 1.0 && s1 == 0.0)) {  // m.applyOnTheLeft(0,1,rot1)
                        const double a0 = m00, a1 = m01, b0 = m10, b1 = m11;
                        m00 = c1 * a0 + s1 * b0; m01 = c1 * a1 + s1 * b1;
                        m10 = -s1 * a0 + c1 * b0; m11 = -s1 * a1 + c1 * b1;
                    }
                    double cr, sr;  // j_right.makeJacobi(m00, m01, m11)
                    {
                        const double deno = 2.0 * std::fabs(m01);
                        if (deno < dmin) {
                            cr = 1.0; sr = 0.0;
                        } else {
                            const double tau = (m00 - m11) / deno;
                            const double w = std::sqrt(tau * tau + 1.0);
                            const double tt = tau > 0.0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
                            const double sign_t = tt > 0.0 ? 1.0 : -1.0;
                            const double nn = 1.0 / std::sqrt(tt * tt + 1.0);
                            sr = -sign_t * (m01 / std::fabs(m01)) * std::fabs(tt) * nn;
                            cr = nn;
                        }
                    }
                    // j_left = rot1 * j_right.transpose()
                    const double c2 = cr, s2 = -sr;
                    const double cl = c1 * c2 - s1 * s2, sl = c1 * s2 + s1 * c2;
                    // m_workMatrix.applyOnTheLeft(p,q,j_left): rows p, q
                    if (!(cl == 1.0 && sl == 0.0)) {
                        for (int k = 0; k < n; ++k) {
                            const double xi = W[p][k], yi = W[q][k];
                            W[p][k] = cl * xi + sl * yi;
                            W[q][k] = -sl * xi + cl * yi;
                        }
                        // m_matrixU.applyOnTheRight(p,q,j_left.transpose()): columns p, q with (cl, sl)
                        for (int k = 0; k < n; ++k) {
                            const double xi = U[k][p], yi = U[k][q];
                            U[k][p] = cl * xi + sl * yi;
                            U[k][q] = -sl * xi + cl * yi;
                        }
                    }
                    // m_workMatrix.applyOnTheRight(p,q,j_right); m_matrixV.applyOnTheRight(p,q,j_right): columns with (cr, -sr)
                    if (!(cr == 1.0 && -sr == 0.0)) {
                        for (int k = 0; k < n; ++k) {
                            const double xi = W[k][p], yi = W[k][q];
                            W[k][p] = cr * xi - sr * yi;
                            W[k][q] = sr * xi + cr * yi;
                        }
                        for (int k = 0; k < n; ++k) {
                            const double xi = V[k][p], yi = V[k][q];
                            V[k][p] = cr * xi - sr * yi;
                            V[k][q] = sr * xi + cr * yi;
                        }
                    }
                    maxDiag = std::max(maxDiag, std::max(std::fabs(W[p][p]), std::fabs(W[q][q])));
                }
            }
        }
    }
    double sv[6];
    for (int i = 0; i < n; ++i) {
        const double a = W[i][i];
        sv[i] = std::fabs(a);
        if (a < 0.0)
            for (int k = 0; k < n; ++k) U[k][i] = -U[k][i];
    }
    for (int i = 0; i < n; ++i) sv[i] *= scale;
    int nonzero = n;
    for (int i = 0; i < n; ++i) {
        int pos = 0;
        double mx = sv[i];
        for (int j = 1; j < n - i; ++j)
            if (sv[i + j] > mx) { mx = sv[i + j]; pos = j; }
        if (mx == 0.0) { nonzero = i; break; }
        if (pos) {
            pos += i;
            std::swap(sv[i], sv[pos]);
            for (int k = 0; k < n; ++k) { std::swap(U[k][pos], U[k][i]); std::swap(V[k][pos], V[k][i]); }
        }
    }
    if (sv_out) for (int i = 0; i < n; ++i) sv_out[i] = sv[i];
    // rank() and _solve_impl
    const double pthr = std::max(sv[0] * (6.0 * eps), dmin);
    int r = nonzero - 1;
    while (r >= 0 && sv[r] < pthr) --r;
    const int rank = r + 1;
    double tmp[6];
    for (int k = 0; k < rank; ++k) {
        double d = 0.0;
        for (int i = 0; i < n; ++i) d += U[i][k] * rhs[i];
        tmp[k] = (1.0 / sv[k]) * d;
    }
    for (int i = 0; i < n; ++i) {
        double d = 0.0;
        for (int k = 0; k < rank; ++k) d += V[i][k] * tmp[k];
        x[i] = d;
    }
}

// ---------------------------------------------------------------------------
// JacobiSVD<Matrix<double,N,N>>(A, ComputeFullU | ComputeFullV): the decomposition itself, any small N (square: no QR
// preconditioner).  Same sources as jacobi_svd_solve6: JacobiSVD.h:666-792, misc/RealSvd2x2.h:18-51, Jacobi.h:83-113.
// A, U, V row-major; singular values descending.  Used by the GICP oracle with N = 3 (gicp_omp_impl.hpp:109-111).
// ---------------------------------------------------------------------------
template <int N>
inline bool jacobi_svd(const double* A, double* Uo, double* Vo, double* sv) {
    const double eps = std::numeric_limits<double>::epsilon(), dmin = std::numeric_limits<double>::min();
    const double precision = 2.0 * eps;
    double scale = 0.0;
    bool bad = false;
    for (int i = 0; i < N * N; ++i) {
        const double a = std::fabs(A[i]);
        if (a != a) bad = true;
        if (a > scale) scale = a;
    }
    if (bad || !std::isfinite(scale)) {
        for (int i = 0; i < N * N; ++i) Uo[i] = Vo[i] = std::numeric_limits<double>::quiet_NaN();
        for (int i = 0; i < N; ++i) sv[i] = std::numeric_limits<double>::quiet_NaN();
        return false;
    }
    if (scale == 0.0) scale = 1.0;
    double W[N][N], U[N][N], V[N][N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            W[i][j] = A[i * N + j] / scale;
            U[i][j] = V[i][j] = i == j ? 1.0 : 0.0;
        }
    double maxDiag = 0.0;
    for (int i = 0; i < N; ++i) maxDiag = std::max(maxDiag, std::fabs(W[i][i]));
    bool finished = false;
    while (!finished) {
        finished = true;
        for (int p = 1; p < N; ++p)
            for (int q = 0; q < p; ++q) {
                const double thr = std::max(dmin, precision * maxDiag);
                if (!(std::fabs(W[p][q]) > thr || std::fabs(W[q][p]) > thr)) continue;
                finished = false;
                double m[2][2] = {{W[p][p], W[p][q]}, {W[q][p], W[q][q]}};
                // real_2x2_jacobi_svd: rot1 makes m symmetric, then makeJacobi diagonalises it
                double c1 = 1.0, s1 = 0.0;
                const double t = m[0][0] + m[1][1], d = m[1][0] - m[0][1];
                if (!(std::fabs(d) < dmin)) {
                    const double u = t / d, tmp = std::sqrt(1.0 + u * u);
                    s1 = 1.0 / tmp;
                    c1 = u / tmp;
                }
                if (!(c1 == 1.0 && s1 == 0.0))
                    for (int k = 0; k < 2; ++k) {
                        const double x = m[0][k], y = m[1][k];
                        m[0][k] = c1 * x + s1 * y;
                        m[1][k] = -s1 * x + c1 * y;
                    }
                double cr = 1.0, sr = 0.0;
                const double deno = 2.0 * std::fabs(m[0][1]);
                if (!(deno < dmin)) {
                    const double tau = (m[0][0] - m[1][1]) / deno, w = std::sqrt(tau * tau + 1.0);
                    const double tt = tau > 0.0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
                    const double sign_t = tt > 0.0 ? 1.0 : -1.0, nn = 1.0 / std::sqrt(tt * tt + 1.0);
                    sr = -sign_t * (m[0][1] / std::fabs(m[0][1])) * std::fabs(tt) * nn;
                    cr = nn;
                }
                const double cl = c1 * cr - s1 * (-sr), sl = c1 * (-sr) + s1 * cr;  // j_left = rot1 * j_right.transpose()
                if (!(cl == 1.0 && sl == 0.0)) {
                    for (int k = 0; k < N; ++k) {
                        const double x = W[p][k], y = W[q][k];
                        W[p][k] = cl * x + sl * y;
                        W[q][k] = -sl * x + cl * y;
                    }
                    for (int k = 0; k < N; ++k) {
                        const double x = U[k][p], y = U[k][q];
                        U[k][p] = cl * x + sl * y;
                        U[k][q] = -sl * x + cl * y;
                    }
                }
                if (!(cr == 1.0 && -sr == 0.0)) {
                    for (int k = 0; k < N; ++k) {
                        const double x = W[k][p], y = W[k][q];
                        W[k][p] = cr * x - sr * y;
                        W[k][q] = sr * x + cr * y;
                    }
                    for (int k = 0; k < N; ++k) {
                        const double x = V[k][p], y = V[k][q];
                        V[k][p] = cr * x - sr * y;
                        V[k][q] = sr * x + cr * y;
                    }
                }
                maxDiag = std::max(maxDiag, std::max(std::fabs(W[p][p]), std::fabs(W[q][q])));
            }
    }
    for (int i = 0; i < N; ++i) {
        const double a = W[i][i];
        sv[i] = std::fabs(a);
        if (a < 0.0)
            for (int k = 0; k < N; ++k) U[k][i] = -U[k][i];
    }
    for (int i = 0; i < N; ++i) sv[i] *= scale;
    for (int i = 0; i < N; ++i) {
        int pos = 0;
        double mx = sv[i];
        for (int j = 1; j < N - i; ++j)
            if (sv[i + j] > mx) { mx = sv[i + j]; pos = j; }
        if (mx == 0.0) break;
        if (pos) {
            pos += i;
            std::swap(sv[i], sv[pos]);
            for (int k = 0; k < N; ++k) { std::swap(U[k][pos], U[k][i]); std::swap(V[k][pos], V[k][i]); }
        }
    }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) { Uo[i * N + j] = U[i][j]; Vo[i * N + j] = V[i][j]; }
    return true;
}

}  // namespace orc
