// TEST INFRASTRUCTURE ONLY — part of the CPU oracle (see oracle/README.md).
// Nothing under pointcloud-slam_b200/ may include this file.
//
// Tiny dense linear algebra used by the oracle restatements.  The reference
// uses Eigen (not vendored in a usable form); the routines here
// restate the *published* Eigen 3.3 algorithms the reference call sites pick:
//   - ColPivHouseholderQR::solve   (common_lib.h:208,223)
//   - Matrix::inverse() via PartialPivLU for n>4 (esekfom.hpp:1685,1706)
//   - 3x3 inverse by cofactors      (voxel_grid_covariance_omp_impl.hpp:355,359)
//   - SelfAdjointEigenSolver 3x3    (voxel_grid_covariance_omp_impl.hpp:333)
//   - JacobiSVD(6x6).solve          (ndt_omp_impl.hpp:112-114)
// Summation orders that Eigen's SIMD kernels would pick are unknowable here;
// every reduction below is a plain left-to-right loop (documented contract).
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

}  // namespace orc
