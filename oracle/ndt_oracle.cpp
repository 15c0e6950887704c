// TEST INFRASTRUCTURE ONLY — CPU oracle for the pclomp NDT hot path.
// PARITY UNPINNED (see oracle.h).  Never linked into the product library.
//
// Restates (paths relative to /root/reference/src/pointcloud_match/ndt_omp/include/pclomp):
//   VoxelGridCovariance::applyFilter / Leaf          voxel_grid_covariance_omp_impl.hpp:49-370, .h:98-116,208-222
//   getNeighborhoodAtPoint{,7,1}                     voxel_grid_covariance_omp_impl.hpp:374-442
//   NormalDistributionsTransform ctor / computeTransformation   ndt_omp_impl.hpp:47-156
//   computeDerivatives / computeAngleDerivatives     ndt_omp_impl.hpp:169-267,271-366
//   computePointDerivatives (float + double)         ndt_omp_impl.hpp:370-449
//   updateDerivatives / computeHessian / updateHessian           ndt_omp_impl.hpp:452-590
//   updateIntervalMT / trialValueSelectionMT / computeStepLengthMT  ndt_omp_impl.hpp:594-833
//   calculateScore                                    ndt_omp_impl.hpp:836-880
// Third-party arithmetic: Eigen's SelfAdjointEigenSolver, JacobiSVD, 3x3 inverse and eulerAngles(0,1,2) are restated
// from the Eigen sources vendored in the reference (fast_gicp/thirdparty/Eigen/Eigen/src/{Eigenvalues,SVD,LU,Geometry},
// file:line in smallmat.h and at euler_012); pcl::transformPointCloud (PCL 1.7/1.8 scalar form), pcl::getMinMax3D and the
// AngleAxis products (Eigen/src/Core and Geometry/AngleAxis arithmetic over absent Core) from published behaviour.
#include "oracle.h"
#include "smallmat.h"

#include <omp.h>
#include <cstdio>
#include <map>
#include <vector>

namespace orc {
int g_legacy_eigen = 0;  // test switch (orc_set_legacy_eigen): 1 = round-1 cyclic-Jacobi substitutes instead of the Eigen restatements

struct Leaf {
    int nr_points = 0;
    double mean[3] = {0, 0, 0};
    double cov[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // Leaf() sets cov_ to identity (.h:103-112) and the sums accumulate on top of it
    double icov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double evals[3] = {0, 0, 0};
    float centroid[3] = {0, 0, 0};  // leaf.centroid: fp32 running sum in input order, then / (float) nr_points (vgc_impl:241-242, 289)
    bool in_kdtree = false;         // pushed into voxel_centroids_ (nr_points >= min_points at that moment, vgc_impl:297-326)
};

struct F3 { float x, y, z; };

struct Ndt {
    orc_ndt_params prm;
    int nthreads = 1;
    // voxel grid
    float leaf = 1.0f, inv_leaf = 1.0f;
    int min_b[3] = {0, 0, 0}, max_b[3] = {0, 0, 0}, div_b[3] = {0, 0, 0}, divb_mul[3] = {0, 0, 0};
    std::map<size_t, Leaf> leaves;
    std::vector<F3> source, target;
    // gaussian constants
    double d1 = 0, d2 = 0, d3 = 0;
    // angle tables
    double jd[8][3];   // j_ang_a_..h_ (double)
    double hd[15][3];  // h_ang_a2_.. f3_ (double)
    float jf[8][4];    // j_ang (float 8x4)
    float hf[16][4];   // h_ang (float 16x4)
    int evals = 0, hess_evals = 0;

    void gauss() {  // ndt_omp_impl.hpp:77-81
        double c1 = 10 * (1 - prm.outlier_ratio);
        double c2 = prm.outlier_ratio / std::pow((double)prm.resolution, 3);
        d3 = -std::log(c2);
        d1 = -std::log(c1 + c2) - d3;
        d2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / d1);
    }

    int64_t set_target(const float* xyz, int64_t n, int64_t stride) {  // applyFilter
        leaves.clear();
        target.resize((size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            const float* p = (const float*)((const char*)xyz + i * stride);
            target[i] = F3{p[0], p[1], p[2]};
        }
        leaf = prm.resolution;
        inv_leaf = 1.0f / leaf;  // setLeafSize: inverse_leaf_size_ = Array4f::Ones()/leaf_size_
        float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
        float mx[3] = {-mn[0], -mn[1], -mn[2]};
        for (int64_t i = 0; i < n; ++i) {  // pcl::getMinMax3D
            const float* p = (const float*)((const char*)xyz + i * stride);
            if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
            for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], p[k]); mx[k] = std::max(mx[k], p[k]); }
        }
        int64_t dx = (int64_t)((mx[0] - mn[0]) * inv_leaf) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv_leaf) + 1,
                dz = (int64_t)((mx[2] - mn[2]) * inv_leaf) + 1;
        if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) return -1;
        for (int k = 0; k < 3; ++k) {
            min_b[k] = (int)std::floor(mn[k] * inv_leaf);
            max_b[k] = (int)std::floor(mx[k] * inv_leaf);
            div_b[k] = max_b[k] - min_b[k] + 1;
        }
        divb_mul[0] = 1; divb_mul[1] = div_b[0]; divb_mul[2] = div_b[0] * div_b[1];
        for (int64_t i = 0; i < n; ++i) {  // first pass (:209-264)
            const float* p = (const float*)((const char*)xyz + i * stride);
            if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
            int ijk0 = (int)(std::floor(p[0] * inv_leaf) - (float)min_b[0]);
            int ijk1 = (int)(std::floor(p[1] * inv_leaf) - (float)min_b[1]);
            int ijk2 = (int)(std::floor(p[2] * inv_leaf) - (float)min_b[2]);
            int idx = ijk0 * divb_mul[0] + ijk1 * divb_mul[1] + ijk2 * divb_mul[2];
            Leaf& lf = leaves[(size_t)idx];
            double pt[3] = {p[0], p[1], p[2]};
            for (int a = 0; a < 3; ++a) lf.mean[a] += pt[a];
            for (int a = 0; a < 3; ++a) lf.centroid[a] += p[a];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) lf.cov[a * 3 + b] += pt[a] * pt[b];
            ++lf.nr_points;
        }
        int64_t valid = 0;
        for (auto& kv : leaves) {  // second pass (:282-367)
            Leaf& lf = kv.second;
            double pt_sum[3] = {lf.mean[0], lf.mean[1], lf.mean[2]};
            for (int a = 0; a < 3; ++a) lf.mean[a] /= lf.nr_points;
            for (int a = 0; a < 3; ++a) lf.centroid[a] /= (float)lf.nr_points;
            if (lf.nr_points < prm.min_pts) continue;
            lf.in_kdtree = true;
            double np = lf.nr_points;
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b)
                    lf.cov[a * 3 + b] = (lf.cov[a * 3 + b] - 2 * (pt_sum[a] * lf.mean[b])) / np + lf.mean[a] * lf.mean[b];
            double f = (np - 1.0) / np;
            for (int a = 0; a < 9; ++a) lf.cov[a] *= f;
            double A[9], w[3], V[9];
            for (int a = 0; a < 3; ++a)  // SelfAdjointEigenSolver reads the lower triangle
                for (int b = 0; b < 3; ++b) A[a * 3 + b] = (a >= b) ? lf.cov[a * 3 + b] : lf.cov[b * 3 + a];
            // eigensolver.compute(leaf.cov_) (vgc_impl:333): Eigen's tridiagonalisation + implicit QR, restated from the vendored
            // sources (smallmat.h).  g_legacy_eigen: the round-1 substitute (cyclic Jacobi), kept for the cross-check test only.
            if (g_legacy_eigen) jacobi_eig_sym(A, 3, w, V);
            else eigen_selfadjoint3(A, w, V);
            if (w[0] < 0 || w[1] < 0 || w[2] <= 0) { lf.nr_points = -1; continue; }
            double min_ev = prm.eig_ratio * w[2];
            if (w[0] < min_ev) {
                w[0] = min_ev;
                if (w[1] < min_ev) w[1] = min_ev;
                double Vi[9], VL[9];
                inverse3(V, Vi);
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) VL[a * 3 + b] = V[a * 3 + b] * w[b];
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b)
                        lf.cov[a * 3 + b] = VL[a * 3] * Vi[b] + VL[a * 3 + 1] * Vi[3 + b] + VL[a * 3 + 2] * Vi[6 + b];
            }
            for (int a = 0; a < 3; ++a) lf.evals[a] = w[a];
            inverse3(lf.cov, lf.icov);
            double mxc = lf.icov[0], mnc = lf.icov[0];
            for (int a = 1; a < 9; ++a) { mxc = std::max(mxc, lf.icov[a]); mnc = std::min(mnc, lf.icov[a]); }
            if (mxc == std::numeric_limits<float>::infinity() || mnc == -std::numeric_limits<float>::infinity()) {
                lf.nr_points = -1;
                continue;
            }
            ++valid;
        }
        return valid;
    }

    // getNeighborhoodAtPoint (vgc_impl:374-404) for the DIRECT1/7/26 stencils
    int neighborhood(float x, float y, float z, const Leaf** out) const {
        static const int d7[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
        int ijk[3] = {(int)std::floor(x / leaf), (int)std::floor(y / leaf), (int)std::floor(z / leaf)};
        int cnt = 0;
        // KDTREE (search == 0): target_cells_.radiusSearch(pt, resolution_) = every kd-tree centroid with fp32 squared distance
        // < resolution^2 (voxel_grid_covariance_omp.h:477-505; FLANN L2_Simple + RadiusResultSet, third party).  A centroid lies in
        // its own cell, so the hits are found in the 27-cell block; FLANN returns them by ascending distance - only the order
        // of the per-point sums depends on that (here: block order).  Leaves dropped after entering the kd-tree (bad
        // covariance, nr_points = -1) would still be returned by the reference with an unusable covariance; skipped here.
        int nrel = prm.search == 1 ? 1 : (prm.search == 27 || prm.search == 0) ? 27 : 7;
        for (int ni = 0; ni < nrel; ++ni) {
            int d[3];
            if (nrel == 27) {  // pcl::getAllNeighborCellIndices(): i,j,k in -1..1, x slowest
                d[0] = ni / 9 - 1; d[1] = (ni / 3) % 3 - 1; d[2] = ni % 3 - 1;
            } else { d[0] = d7[ni][0]; d[1] = d7[ni][1]; d[2] = d7[ni][2]; }
            bool in = true;
            for (int k = 0; k < 3; ++k)
                if (!(min_b[k] - ijk[k] <= d[k] && max_b[k] - ijk[k] >= d[k])) in = false;
            if (!in) continue;
            int id = (ijk[0] + d[0] - min_b[0]) * divb_mul[0] + (ijk[1] + d[1] - min_b[1]) * divb_mul[1] +
                     (ijk[2] + d[2] - min_b[2]) * divb_mul[2];
            auto it = leaves.find((size_t)id);
            if (it != leaves.end() && it->second.nr_points >= prm.min_pts) {
                if (prm.search == 0) {
                    const float ax = x - it->second.centroid[0], ay = y - it->second.centroid[1], az = z - it->second.centroid[2];
                    const float d = (ax * ax + ay * ay) + az * az;
                    if (!(d < (float)((double)leaf * (double)leaf))) continue;
                }
                out[cnt++] = &it->second;
            }
        }
        return cnt;
    }

    void angle_derivatives(const double* p, bool compute_hessian = true) {  // ndt_omp_impl.hpp:271-366
        double cx, cy, cz, sx, sy, sz;
        if (std::fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = std::cos(p[3]); sx = std::sin(p[3]); }
        if (std::fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = std::cos(p[4]); sy = std::sin(p[4]); }
        if (std::fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = std::cos(p[5]); sz = std::sin(p[5]); }
        double J[8][3] = {{(-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy)},
                          {(cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy)},
                          {(-sy * cz), sy * sz, cy},
                          {sx * cy * cz, (-sx * cy * sz), sx * sy},
                          {(-cx * cy * cz), cx * cy * sz, (-cx * sy)},
                          {(-cy * sz), (-cy * cz), 0},
                          {(cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0},
                          {(sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0}};
        for (int r = 0; r < 8; ++r)
            for (int c = 0; c < 3; ++c) { jd[r][c] = J[r][c]; jf[r][c] = (float)J[r][c]; }
        for (int r = 0; r < 8; ++r) jf[r][3] = 0.0f;
        if (!compute_hessian) return;
        double Hh[15][3] = {{(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy},      // a2
                            {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy)},   // a3
                            {(cx * cy * cz), (-cx * cy * sz), (cx * sy)},                         // b2
                            {(sx * cy * cz), (-sx * cy * sz), (sx * sy)},                         // b3
                            {(-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0},             // c2
                            {(cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0},             // c3
                            {(-cy * cz), (cy * sz), (-sy)},                                       // d1 (double table: -sy, :332)
                            {(-sx * sy * cz), (sx * sy * sz), (sx * cy)},                         // d2
                            {(cx * sy * cz), (-cx * sy * sz), (-cx * cy)},                        // d3
                            {(sy * sz), (sy * cz), 0},                                            // e1
                            {(-sx * cy * sz), (-sx * cy * cz), 0},                                // e2
                            {(cx * cy * sz), (cx * cy * cz), 0},                                  // e3
                            {(-cy * cz), (cy * sz), 0},                                           // f1
                            {(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0},            // f2
                            {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0}};           // f3
        for (int r = 0; r < 15; ++r)
            for (int c = 0; c < 3; ++c) { hd[r][c] = Hh[r][c]; hf[r][c] = (float)Hh[r][c]; }
        hf[6][2] = (float)(sy);  // bug-compat: the float table's d1 row carries +sy (:354)
        for (int r = 0; r < 15; ++r) hf[r][3] = 0.0f;
        for (int c = 0; c < 4; ++c) hf[15][c] = 0.0f;
    }

    // float path: computePointDerivatives (:370-409) + updateDerivatives (:452-495) for one (point, cell)
    double update_derivatives_f(const double x[3], const double x_trans[3], const double c_inv[9], double g[6], double H[36],
                                bool compute_hessian) const {
        const float x4[4] = {(float)x[0], (float)x[1], (float)x[2], 0.0f};
        float pg[4][6];  // point_gradient_ 4x6
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 6; ++c) pg[r][c] = 0.0f;
        pg[0][0] = pg[1][1] = pg[2][2] = 1.0f;
        float xj[8];
        for (int r = 0; r < 8; ++r) xj[r] = ((jf[r][0] * x4[0] + jf[r][1] * x4[1]) + jf[r][2] * x4[2]) + jf[r][3] * x4[3];
        pg[1][3] = xj[0]; pg[2][3] = xj[1];
        pg[0][4] = xj[2]; pg[1][4] = xj[3]; pg[2][4] = xj[4];
        pg[0][5] = xj[5]; pg[1][5] = xj[6]; pg[2][5] = xj[7];
        float ph[24][6];
        if (compute_hessian) {
            for (int r = 0; r < 24; ++r)
                for (int c = 0; c < 6; ++c) ph[r][c] = 0.0f;
            float xh[16];
            for (int r = 0; r < 16; ++r) xh[r] = ((hf[r][0] * x4[0] + hf[r][1] * x4[1]) + hf[r][2] * x4[2]) + hf[r][3] * x4[3];
            const float a[4] = {0, xh[0], xh[1], 0}, b[4] = {0, xh[2], xh[3], 0}, c[4] = {0, xh[4], xh[5], 0};
            const float d[4] = {xh[6], xh[7], xh[8], 0}, e[4] = {xh[9], xh[10], xh[11], 0}, f[4] = {xh[12], xh[13], xh[14], 0};
            for (int r = 0; r < 4; ++r) {
                ph[12 + r][3] = a[r]; ph[16 + r][3] = b[r]; ph[20 + r][3] = c[r];
                ph[12 + r][4] = b[r]; ph[16 + r][4] = d[r]; ph[20 + r][4] = e[r];
                ph[12 + r][5] = c[r]; ph[16 + r][5] = e[r]; ph[20 + r][5] = f[r];
            }
        }
        const float xt[4] = {(float)x_trans[0], (float)x_trans[1], (float)x_trans[2], 0.0f};
        float ci[4][4];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) ci[r][c] = (r < 3 && c < 3) ? (float)c_inv[r * 3 + c] : 0.0f;
        const float gd2 = (float)d2;
        float xc[4];  // x_trans4 * c_inv4 (row vector times matrix)
        for (int c = 0; c < 4; ++c) xc[c] = ((xt[0] * ci[0][c] + xt[1] * ci[1][c]) + xt[2] * ci[2][c]) + xt[3] * ci[3][c];
        float q = ((xt[0] * xc[0] + xt[1] * xc[1]) + xt[2] * xc[2]) + xt[3] * xc[3];
        float arg = -gd2 * q * 0.5f;
        float e_x_cov_x = (float)std::exp((double)arg);  // exp() of a float argument; evaluated in double, narrowed
        float score_inc = (float)(-d1 * (double)e_x_cov_x);
        e_x_cov_x = gd2 * e_x_cov_x;
        if (e_x_cov_x > 1 || e_x_cov_x < 0 || e_x_cov_x != e_x_cov_x) return 0;
        e_x_cov_x = (float)((double)e_x_cov_x * d1);
        float cg[4][6];  // c_inv4 * point_gradient4
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 6; ++c) cg[r][c] = ((ci[r][0] * pg[0][c] + ci[r][1] * pg[1][c]) + ci[r][2] * pg[2][c]) + ci[r][3] * pg[3][c];
        float xg[6];  // x_trans4 * cg
        for (int c = 0; c < 6; ++c) xg[c] = ((xt[0] * cg[0][c] + xt[1] * cg[1][c]) + xt[2] * cg[2][c]) + xt[3] * cg[3][c];
        for (int c = 0; c < 6; ++c) g[c] += (double)(e_x_cov_x * xg[c]);
        if (compute_hessian) {
            float gg[6][6];  // point_gradient4^T * cg : (j,i)
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c < 6; ++c) gg[r][c] = ((pg[0][r] * cg[0][c] + pg[1][r] * cg[1][c]) + pg[2][r] * cg[2][c]) + pg[3][r] * cg[3][c];
            for (int i = 0; i < 6; ++i) {
                float xh6[6];
                for (int j = 0; j < 6; ++j)
                    xh6[j] = ((xc[0] * ph[i * 4 + 0][j] + xc[1] * ph[i * 4 + 1][j]) + xc[2] * ph[i * 4 + 2][j]) + xc[3] * ph[i * 4 + 3][j];
                for (int j = 0; j < 6; ++j) H[i * 6 + j] += (double)(e_x_cov_x * (-gd2 * xg[i] * xg[j] + xh6[j] + gg[j][i]));
            }
        }
        return (double)score_inc;
    }

    // double path: computePointDerivatives(double) (:413-449) + updateHessian (:565-590)
    void update_hessian_d(const double x[3], const double xt[3], const double ci[9], double H[36]) const {
        double pg[3][6] = {{1, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 0}};
        auto dot = [&](const double v[3]) { return x[0] * v[0] + x[1] * v[1] + x[2] * v[2]; };
        pg[1][3] = dot(jd[0]); pg[2][3] = dot(jd[1]);
        pg[0][4] = dot(jd[2]); pg[1][4] = dot(jd[3]); pg[2][4] = dot(jd[4]);
        pg[0][5] = dot(jd[5]); pg[1][5] = dot(jd[6]); pg[2][5] = dot(jd[7]);
        double ph[18][6];
        for (int r = 0; r < 18; ++r)
            for (int c = 0; c < 6; ++c) ph[r][c] = 0;
        double a[3] = {0, dot(hd[0]), dot(hd[1])}, b[3] = {0, dot(hd[2]), dot(hd[3])}, c[3] = {0, dot(hd[4]), dot(hd[5])};
        double d[3] = {dot(hd[6]), dot(hd[7]), dot(hd[8])}, e[3] = {dot(hd[9]), dot(hd[10]), dot(hd[11])},
               f[3] = {dot(hd[12]), dot(hd[13]), dot(hd[14])};
        for (int r = 0; r < 3; ++r) {
            ph[9 + r][3] = a[r]; ph[12 + r][3] = b[r]; ph[15 + r][3] = c[r];
            ph[9 + r][4] = b[r]; ph[12 + r][4] = d[r]; ph[15 + r][4] = e[r];
            ph[9 + r][5] = c[r]; ph[12 + r][5] = e[r]; ph[15 + r][5] = f[r];
        }
        auto mv = [&](const double v[3], double r[3]) {
            for (int k = 0; k < 3; ++k) r[k] = ci[k * 3] * v[0] + ci[k * 3 + 1] * v[1] + ci[k * 3 + 2] * v[2];
        };
        double cx[3];
        mv(xt, cx);
        double e_x = d2 * std::exp(-d2 * (xt[0] * cx[0] + xt[1] * cx[1] + xt[2] * cx[2]) / 2);
        if (e_x > 1 || e_x < 0 || e_x != e_x) return;
        e_x *= d1;
        for (int i = 0; i < 6; ++i) {
            double col_i[3] = {pg[0][i], pg[1][i], pg[2][i]}, cov_dxd_pi[3];
            mv(col_i, cov_dxd_pi);
            for (int j = 0; j < 6; ++j) {
                double col_j[3] = {pg[0][j], pg[1][j], pg[2][j]}, cj[3], hij[3] = {ph[3 * i][j], ph[3 * i + 1][j], ph[3 * i + 2][j]}, chij[3];
                mv(col_j, cj);
                mv(hij, chij);
                double t1 = xt[0] * cov_dxd_pi[0] + xt[1] * cov_dxd_pi[1] + xt[2] * cov_dxd_pi[2];
                double t2 = xt[0] * cj[0] + xt[1] * cj[1] + xt[2] * cj[2];
                double t3 = xt[0] * chij[0] + xt[1] * chij[1] + xt[2] * chij[2];
                double t4 = col_j[0] * cov_dxd_pi[0] + col_j[1] * cov_dxd_pi[1] + col_j[2] * cov_dxd_pi[2];
                H[i * 6 + j] += e_x * (-d2 * t1 * t2 + t3 + t4);
            }
        }
    }

    static void pose_matrix(const double* p, float M[16] /*row-major 4x4*/) {
        // Translation * AngleAxis(x) * AngleAxis(y) * AngleAxis(z) in float (:129,749-753)
        float rx = (float)p[3], ry = (float)p[4], rz = (float)p[5];
        // float sin/cos of AngleAxisf evaluated as the correctly rounded value (double evaluation, narrowed): libm's
        // sinf/cosf differ from it by one ulp for a few percent of the arguments, platform by platform
        float cx = (float)std::cos((double)rx), sx = (float)std::sin((double)rx), cy = (float)std::cos((double)ry),
              sy = (float)std::sin((double)ry), cz = (float)std::cos((double)rz), sz = (float)std::sin((double)rz);
        float Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
        float Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
        float Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
        float T1[9], T2[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) T1[i * 3 + j] = (Rx[i * 3] * Ry[j] + Rx[i * 3 + 1] * Ry[3 + j]) + Rx[i * 3 + 2] * Ry[6 + j];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) T2[i * 3 + j] = (T1[i * 3] * Rz[j] + T1[i * 3 + 1] * Rz[3 + j]) + T1[i * 3 + 2] * Rz[6 + j];
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) M[i * 4 + j] = T2[i * 3 + j];
            M[i * 4 + 3] = (float)p[i];
        }
        M[12] = M[13] = M[14] = 0.0f;
        M[15] = 1.0f;
    }
    static void transform_cloud(const std::vector<F3>& in, const float M[16], std::vector<F3>& out) {
        out.resize(in.size());
        for (size_t i = 0; i < in.size(); ++i) {  // pcl::transformPointCloud, PCL 1.7/1.8 scalar form
            const F3& p = in[i];
            out[i].x = ((M[0] * p.x + M[1] * p.y) + M[2] * p.z) + M[3];
            out[i].y = ((M[4] * p.x + M[5] * p.y) + M[6] * p.z) + M[7];
            out[i].z = ((M[8] * p.x + M[9] * p.y) + M[10] * p.z) + M[11];
        }
    }

    double compute_derivatives(double g[6], double H[36], const std::vector<F3>& trans, const double* p, bool compute_hessian = true) {
        ++evals;
        for (int i = 0; i < 6; ++i) g[i] = 0;
        for (int i = 0; i < 36; ++i) H[i] = 0;
        const size_t n = source.size();
        std::vector<double> scores(n, 0.0), gs(n * 6, 0.0), Hs(n * 36, 0.0);
        angle_derivatives(p);
#pragma omp parallel for num_threads(nthreads) schedule(guided, 8)
        for (size_t idx = 0; idx < n; ++idx) {
            const Leaf* nb[27];
            int cnt = neighborhood(trans[idx].x, trans[idx].y, trans[idx].z, nb);
            double score_pt = 0, gp[6] = {0, 0, 0, 0, 0, 0}, Hp[36];
            for (int k = 0; k < 36; ++k) Hp[k] = 0;
            for (int c = 0; c < cnt; ++c) {
                double x[3] = {source[idx].x, source[idx].y, source[idx].z};
                double xt[3] = {trans[idx].x - nb[c]->mean[0], trans[idx].y - nb[c]->mean[1], trans[idx].z - nb[c]->mean[2]};
                score_pt += update_derivatives_f(x, xt, nb[c]->icov, gp, Hp, compute_hessian);
            }
            scores[idx] = score_pt;
            for (int k = 0; k < 6; ++k) gs[idx * 6 + k] = gp[k];
            for (int k = 0; k < 36; ++k) Hs[idx * 36 + k] = Hp[k];
        }
        double score = 0;
        for (size_t i = 0; i < n; ++i) {
            score += scores[i];
            for (int k = 0; k < 6; ++k) g[k] += gs[i * 6 + k];
            for (int k = 0; k < 36; ++k) H[k] += Hs[i * 36 + k];
        }
        return score;
    }
    void compute_hessian(double H[36], const std::vector<F3>& trans) {  // :499-560 (serial)
        ++hess_evals;
        for (int i = 0; i < 36; ++i) H[i] = 0;
        for (size_t idx = 0; idx < source.size(); ++idx) {
            const Leaf* nb[27];
            int cnt = neighborhood(trans[idx].x, trans[idx].y, trans[idx].z, nb);
            for (int c = 0; c < cnt; ++c) {
                double x[3] = {source[idx].x, source[idx].y, source[idx].z};
                double xt[3] = {trans[idx].x - nb[c]->mean[0], trans[idx].y - nb[c]->mean[1], trans[idx].z - nb[c]->mean[2]};
                update_hessian_d(x, xt, nb[c]->icov, H);
            }
        }
    }
    double calculate_score(const std::vector<F3>& trans) const {  // :836-880
        double score = 0;
        for (size_t idx = 0; idx < trans.size(); ++idx) {
            const Leaf* nb[27];
            int cnt = neighborhood(trans[idx].x, trans[idx].y, trans[idx].z, nb);
            for (int c = 0; c < cnt; ++c) {
                double xt[3] = {trans[idx].x - nb[c]->mean[0], trans[idx].y - nb[c]->mean[1], trans[idx].z - nb[c]->mean[2]};
                const double* ci = nb[c]->icov;
                double cx[3];
                for (int k = 0; k < 3; ++k) cx[k] = ci[k * 3] * xt[0] + ci[k * 3 + 1] * xt[1] + ci[k * 3 + 2] * xt[2];
                double e = std::exp(-d2 * (xt[0] * cx[0] + xt[1] * cx[1] + xt[2] * cx[2]) / 2);
                double inc = -d1 * e - d3;
                score += inc / cnt;
            }
        }
        return score / (double)trans.size();
    }

    // ---- More-Thuente (:594-833)
    static bool update_interval(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t, double g_t) {
        if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
        else if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
        else if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
        else return true;
    }
    static double trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
        if (f_t > f_l) {
            double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
            double w = std::sqrt(z * z - g_t * g_l);
            double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
            double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
            if (std::fabs(a_c - a_l) < std::fabs(a_q - a_l)) return a_c;
            else return 0.5 * (a_q + a_c);
        } else if (g_t * g_l < 0) {
            double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
            double w = std::sqrt(z * z - g_t * g_l);
            double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
            double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
            if (std::fabs(a_c - a_t) >= std::fabs(a_s - a_t)) return a_c;
            else return a_s;
        } else if (std::fabs(g_t) <= std::fabs(g_l)) {
            double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
            double w = std::sqrt(z * z - g_t * g_l);
            double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
            double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
            double a_t_next = (std::fabs(a_c - a_t) < std::fabs(a_s - a_t)) ? a_c : a_s;
            if (a_t > a_l) return std::min(a_t + 0.66 * (a_u - a_t), a_t_next);
            else return std::max(a_t + 0.66 * (a_u - a_t), a_t_next);
        } else {
            double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
            double w = std::sqrt(z * z - g_t * g_u);
            return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
        }
    }
    double step_length_mt(const double x[6], double step_dir[6], double step_init, double step_max, double step_min, double& score,
                          double g[6], double H[36], std::vector<F3>& trans, float final_T[16]) {
        double phi_0 = -score;
        double d_phi_0 = 0;
        for (int i = 0; i < 6; ++i) d_phi_0 += g[i] * step_dir[i];
        d_phi_0 = -d_phi_0;
        double x_t[6];
        if (d_phi_0 >= 0) {
            if (d_phi_0 == 0) return 0;
            d_phi_0 *= -1;
            for (int i = 0; i < 6; ++i) step_dir[i] *= -1;
        }
        const int max_step_iterations = 10;
        int step_iterations = 0;
        const double mu = 1.e-4, nu = 0.9;
        double a_l = 0, a_u = 0;
        double f_l = phi_0 - phi_0 - mu * d_phi_0 * a_l, g_l = d_phi_0 - mu * d_phi_0;
        double f_u = phi_0 - phi_0 - mu * d_phi_0 * a_u, g_u = d_phi_0 - mu * d_phi_0;
        bool interval_converged = (step_max - step_min) < 0, open_interval = true;
        double a_t = step_init;
        a_t = std::min(a_t, step_max);
        a_t = std::max(a_t, step_min);
        for (int i = 0; i < 6; ++i) x_t[i] = x[i] + step_dir[i] * a_t;
        pose_matrix(x_t, final_T);
        transform_cloud(source, final_T, trans);
        score = compute_derivatives(g, H, trans, x_t, true);
        double phi_t = -score, d_phi_t = 0;
        for (int i = 0; i < 6; ++i) d_phi_t += g[i] * step_dir[i];
        d_phi_t = -d_phi_t;
        double psi_t = phi_t - phi_0 - mu * d_phi_0 * a_t, d_psi_t = d_phi_t - mu * d_phi_0;
        while (!interval_converged && step_iterations < max_step_iterations && !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
            if (open_interval) a_t = trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
            else a_t = trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
            a_t = std::min(a_t, step_max);
            a_t = std::max(a_t, step_min);
            for (int i = 0; i < 6; ++i) x_t[i] = x[i] + step_dir[i] * a_t;
            pose_matrix(x_t, final_T);
            transform_cloud(source, final_T, trans);
            score = compute_derivatives(g, H, trans, x_t, false);
            phi_t = -score;
            d_phi_t = 0;
            for (int i = 0; i < 6; ++i) d_phi_t += g[i] * step_dir[i];
            d_phi_t = -d_phi_t;
            psi_t = phi_t - phi_0 - mu * d_phi_0 * a_t;
            d_psi_t = d_phi_t - mu * d_phi_0;
            if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
                open_interval = false;
                f_l = f_l + phi_0 - mu * d_phi_0 * a_l;
                g_l = g_l + mu * d_phi_0;
                f_u = f_u + phi_0 - mu * d_phi_0 * a_u;
                g_u = g_u + mu * d_phi_0;
            }
            if (open_interval) interval_converged = update_interval(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
            else interval_converged = update_interval(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
            step_iterations++;
        }
        if (step_iterations) compute_hessian(H, trans);
        return a_t;
    }
};

// Eigen Matrix3f::eulerAngles(0,1,2): fast_gicp/thirdparty/Eigen/Eigen/src/Geometry/EulerAngles.h:36-107 (a0 != a2 branch,
// even permutation: i=0, j=1, k=2, result negated), R row-major
static void euler_012(const float R[9], float res[3]) {
    const int i = 0, j = 1, k = 2;
    auto c = [&](int r, int cc) { return R[r * 3 + cc]; };
    const float pi = (float)M_PI;
    // atan2f/sinf/cosf as correctly rounded floats (double evaluation, narrowed), see pose_matrix
    res[0] = (float)std::atan2((double)c(j, k), (double)c(k, k));
    float c2 = std::sqrt(c(i, i) * c(i, i) + c(i, j) * c(i, j));
    if (res[0] > 0.0f) {  // even permutation: flip when res[0] > 0
        if (res[0] > 0.0f) res[0] -= pi; else res[0] += pi;
        res[1] = (float)std::atan2((double)-c(i, k), (double)-c2);
    } else {
        res[1] = (float)std::atan2((double)-c(i, k), (double)c2);
    }
    float s1 = (float)std::sin((double)res[0]), c1 = (float)std::cos((double)res[0]);
    res[2] = (float)std::atan2((double)(s1 * c(k, i) - c1 * c(j, i)), (double)(c1 * c(j, j) - s1 * c(k, j)));
    res[0] = -res[0]; res[1] = -res[1]; res[2] = -res[2];
}

static int ndt_align(Ndt& N, const float* guess_cm, float* final_cm, orc_ndt_result* r) {  // ndt_omp_impl.hpp:70-156
    N.gauss();
    N.evals = N.hess_evals = 0;
    int nr_iterations = 0;
    bool converged = false;
    float G[16];  // row-major
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) G[i * 4 + j] = guess_cm[j * 4 + i];
    bool is_identity = true;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (G[i * 4 + j] != (i == j ? 1.0f : 0.0f)) is_identity = false;
    float final_T[16];
    for (int i = 0; i < 16; ++i) final_T[i] = (i % 5 == 0) ? 1.0f : 0.0f;  // align() resets final_transformation_ = I
    std::vector<F3> trans = N.source;
    if (!is_identity) {
        std::memcpy(final_T, G, sizeof G);
        Ndt::transform_cloud(N.source, G, trans);
    }
    double p[6], delta_p[6], g[6], H[36];
    float Rm[9] = {final_T[0], final_T[1], final_T[2], final_T[4], final_T[5], final_T[6], final_T[8], final_T[9], final_T[10]};
    float eul[3];
    euler_012(Rm, eul);
    p[0] = final_T[3]; p[1] = final_T[7]; p[2] = final_T[11];
    p[3] = eul[0]; p[4] = eul[1]; p[5] = eul[2];
    double score = N.compute_derivatives(g, H, trans, p);
    double trans_probability = 0;
    bool early = false;
    while (!converged) {
        double ng[6];
        for (int i = 0; i < 6; ++i) ng[i] = -g[i];
        // Eigen::JacobiSVD<Matrix<double,6,6>> sv(hessian, FullU | FullV); delta_p = sv.solve(-score_gradient)  (:112-114)
        if (g_legacy_eigen) svd_solve_sym6(H, ng, delta_p);
        else jacobi_svd_solve6(H, ng, delta_p);
        double dn = 0;
        for (int i = 0; i < 6; ++i) dn += delta_p[i] * delta_p[i];
        dn = std::sqrt(dn);
        if (dn == 0 || dn != dn) {
            trans_probability = score / (double)N.source.size();
            converged = dn == dn;
            early = true;
            break;
        }
        for (int i = 0; i < 6; ++i) delta_p[i] /= dn;
        dn = N.step_length_mt(p, delta_p, dn, N.prm.step_size, N.prm.trans_eps / 2, score, g, H, trans, final_T);
        for (int i = 0; i < 6; ++i) delta_p[i] *= dn;
        for (int i = 0; i < 6; ++i) p[i] = p[i] + delta_p[i];
        if (nr_iterations > N.prm.max_iter || (nr_iterations && (std::fabs(dn) < N.prm.trans_eps))) converged = true;
        nr_iterations++;
    }
    if (!early) trans_probability = score / (double)N.source.size();
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) final_cm[j * 4 + i] = final_T[i * 4 + j];
    if (r) {
        r->converged = converged ? 1 : 0;
        r->iters = nr_iterations;
        r->evals = N.evals;
        r->hess_evals = N.hess_evals;
        r->trans_probability = trans_probability;
        std::memcpy(r->hessian, H, sizeof H);
        r->score = score;
        std::memcpy(r->p_final, p, sizeof p);
    }
    return converged ? 0 : 2;
}

}  // namespace orc

using namespace orc;
struct orc_ndt { Ndt N; };

extern "C" {
void orc_set_legacy_eigen(int32_t on) { orc::g_legacy_eigen = on; }
/* stand-alone probes of the Eigen restatements (tests) */
int32_t orc_eigen_selfadjoint3(const double* A9, double* w3, double* V9) { return orc::eigen_selfadjoint3(A9, w3, V9) ? 1 : 0; }
void orc_jacobi_svd_solve6(const double* H36, const double* rhs6, double* x6, double* sv6) { orc::jacobi_svd_solve6(H36, rhs6, x6, sv6); }
orc_ndt* orc_ndt_create(const orc_ndt_params* p) {
    orc_ndt* h = new orc_ndt();
    h->N.prm = *p;
    h->N.nthreads = p->num_threads > 0 ? p->num_threads : omp_get_max_threads();
    h->N.gauss();
    return h;
}
void orc_ndt_destroy(orc_ndt* h) { delete h; }
int64_t orc_ndt_set_target(orc_ndt* h, const float* xyz, int64_t n, int64_t stride) { return h->N.set_target(xyz, n, stride); }
void orc_ndt_set_source(orc_ndt* h, const float* xyz, int64_t n, int64_t stride) {
    h->N.source.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const float* p = (const float*)((const char*)xyz + i * stride);
        h->N.source[i] = F3{p[0], p[1], p[2]};
    }
}
int64_t orc_ndt_num_leaves(orc_ndt* h) { return (int64_t)h->N.leaves.size(); }
int64_t orc_ndt_leaves(orc_ndt* h, int64_t max, int64_t* ids, int32_t* npts, double* mean, double* cov, double* icov) {
    int64_t k = 0;
    for (auto& kv : h->N.leaves) {
        if (kv.second.nr_points < h->N.prm.min_pts) continue;
        if (k < max) {
            ids[k] = (int64_t)kv.first;
            npts[k] = kv.second.nr_points;
            std::memcpy(mean + k * 3, kv.second.mean, 24);
            std::memcpy(cov + k * 9, kv.second.cov, 72);
            std::memcpy(icov + k * 9, kv.second.icov, 72);
        }
        ++k;
    }
    return k;
}
void orc_ndt_grid(orc_ndt* h, int32_t* min_b, int32_t* div_b) {
    for (int k = 0; k < 3; ++k) { min_b[k] = h->N.min_b[k]; div_b[k] = h->N.div_b[k]; }
}
double orc_ndt_derivatives(orc_ndt* h, const double* p6, double* g6, double* H36, int32_t compute_hessian) {
    h->N.gauss();
    float M[16];
    Ndt::pose_matrix(p6, M);
    std::vector<F3> trans;
    Ndt::transform_cloud(h->N.source, M, trans);
    return h->N.compute_derivatives(g6, H36, trans, p6, compute_hessian != 0);
}
void orc_ndt_hessian(orc_ndt* h, const double* p6, double* H36) {
    h->N.gauss();
    float M[16];
    Ndt::pose_matrix(p6, M);
    std::vector<F3> trans;
    Ndt::transform_cloud(h->N.source, M, trans);
    h->N.angle_derivatives(p6);
    h->N.compute_hessian(H36, trans);
}
int32_t orc_ndt_align(orc_ndt* h, const float* guess, float* final_T, orc_ndt_result* r) { return ndt_align(h->N, guess, final_T, r); }
void orc_ndt_score_batch(orc_ndt* h, const float* poses, int64_t np, double* scores) {
    h->N.gauss();
#pragma omp parallel for num_threads(h->N.nthreads) schedule(dynamic, 1)
    for (int64_t k = 0; k < np; ++k) {
        float M[16];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) M[i * 4 + j] = poses[k * 16 + j * 4 + i];
        std::vector<F3> trans;
        Ndt::transform_cloud(h->N.source, M, trans);
        scores[k] = h->N.calculate_score(trans);
    }
}
/* pcl::Registration::getFitnessScore(max_range) (PCL, third party): transformPointCloud(source, final_transformation),
 * kd-tree nearestKSearch(1) per point = exact nearest neighbour (brute force here), float squared distances (FLANN
 * L2_Simple: sequential float accumulation) summed in double over the points with dist <= max_range, divided by their
 * number; DBL_MAX when there is none. */
double orc_ndt_fitness(orc_ndt* h, const float* T16_colmajor, double max_range, int64_t* n_in_range) {
    float M[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) M[i * 4 + j] = T16_colmajor[j * 4 + i];
    std::vector<F3> trans;
    Ndt::transform_cloud(h->N.source, M, trans);
    const size_t n = trans.size();
    std::vector<float> best(n);
#pragma omp parallel for num_threads(h->N.nthreads) schedule(static)
    for (size_t i = 0; i < n; ++i) {
        float b = std::numeric_limits<float>::max();
        for (const F3& t : h->N.target) {
            float dx = trans[i].x - t.x, dy = trans[i].y - t.y, dz = trans[i].z - t.z;
            float d = (dx * dx + dy * dy) + dz * dz;
            if (d < b) b = d;
        }
        best[i] = b;
    }
    double sum = 0;
    int64_t nr = 0;
    for (size_t i = 0; i < n; ++i)
        if ((double)best[i] <= max_range && best[i] < std::numeric_limits<float>::max()) { sum += (double)best[i]; ++nr; }
    if (n_in_range) *n_in_range = nr;
    return nr > 0 ? sum / (double)nr : std::numeric_limits<double>::max();
}
int64_t orc_ndt_nbhd_total(orc_ndt* h, const double* p6) {
    float M[16];
    Ndt::pose_matrix(p6, M);
    std::vector<F3> trans;
    Ndt::transform_cloud(h->N.source, M, trans);
    int64_t tot = 0;
    for (auto& t : trans) {
        const Leaf* nb[27];
        tot += h->N.neighborhood(t.x, t.y, t.z, nb);
    }
    return tot;
}
void orc_euler_from_matrix(const float* m_cm, float* rpy) {
    float R[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = m_cm[j * 4 + i];
    euler_012(R, rpy);
}
void orc_matrix_from_pose(const double* p6, float* m_cm) {
    float M[16];
    Ndt::pose_matrix(p6, M);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) m_cm[j * 4 + i] = M[i * 4 + j];
}
}
