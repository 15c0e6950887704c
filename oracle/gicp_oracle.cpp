// TEST INFRASTRUCTURE ONLY — CPU oracle for pclomp::GeneralizedIterativeClosestPoint.
// PARITY UNPINNED (see oracle.h): the reference holds no fixture for this class and cannot be compiled here.
// Never linked into the product library.
//
// Restates (paths relative to /root/reference/src/pointcloud_match/ndt_omp/include/pclomp):
//   computeCovariances                                gicp_omp_impl.hpp:49-123
//   computeRDerivative / matricesInnerProd            gicp_omp_impl.hpp:127-184, gicp_omp.h:312-322
//   estimateRigidTransformationBFGS                   gicp_omp_impl.hpp:188-242
//   OptimizationFunctorWithIndices::operator() / df / fdf   gicp_omp_impl.hpp:245-368
//   computeTransformation                             gicp_omp_impl.hpp:371-516
//   applyState                                        gicp_omp_impl.hpp:518-529
//   constructor defaults                              gicp_omp.h:115-135
// Third-party code that is NOT in the reference tree, restated from its published source:
//   pcl::BFGS<Functor> (PCL registration/bfgs.h, itself a port of GSL multimin/vector_bfgs2.c + linear_minimize.c; the
//   reference builds against the system PCL, `find_package(PCL)` without a version, ndt_omp/CMakeLists.txt) — including
//   two published oddities of interpolate(): the cubic branch is guarded by `!(fpb != fpa)` and the quadratic branch tests
//   `c > a`;
//   pcl::Registration::align (sets transformation_ = Identity and data[3] = 1, then computeTransformation);
//   pcl::search::KdTree::nearestKSearch = exact k nearest neighbours, ascending distance (FLANN L2_Simple float
//   distances); ties are broken by the lower point index here (FLANN's order among equal distances is an implementation
//   detail of its heap).
// Eigen arithmetic vendored in the reference (fast_gicp/thirdparty/Eigen/Eigen/src): JacobiSVD (smallmat.h jacobi_svd),
// 3x3 inverse, AngleAxis -> Quaternion (Geometry/Quaternion.h:561-569), float quaternion product
// (Geometry/arch/Geometry_SIMD.h:33-45), toRotationMatrix (Quaternion.h:600-621).  Reductions are left-to-right loops
// (smallmat.h contract).
#include "oracle.h"
#include "smallmat.h"

#include <omp.h>
#include <cmath>
#include <cstdio>
#include <algorithm>
#include <functional>
#include <unordered_map>
#include <vector>

namespace orc {
namespace {

struct P3 { float x, y, z; };
struct M3d { double m[3][3]; };
struct M4f { float m[4][4]; };
struct V6 { double v[6]; };

inline float sqdist(const P3& a, const P3& b) {
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return (dx * dx + dy * dy) + dz * dz;  // FLANN L2_Simple: sequential float accumulation
}

// ---- exact k nearest neighbours: the role of pcl::search::KdTree.  A uniform grid searched shell by shell until no
// unexplored cell can hold a closer point; checked against brute force in tests/test_oracle_gicp.py.
struct ExactSearch {
    const std::vector<P3>* pts = nullptr;
    float leaf = 1.0f;
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    std::unordered_map<uint64_t, std::vector<int>> cells;

    static uint64_t key(int x, int y, int z) {
        return ((uint64_t)(uint32_t)(x + (1 << 20)) << 42) | ((uint64_t)(uint32_t)(y + (1 << 20)) << 21) | (uint64_t)(uint32_t)(z + (1 << 20));
    }
    void build(const std::vector<P3>& p) {
        pts = &p;
        cells.clear();
        if (p.empty()) return;
        float mn[3] = {p[0].x, p[0].y, p[0].z}, mx[3] = {p[0].x, p[0].y, p[0].z};
        for (const P3& q : p) {
            mn[0] = std::min(mn[0], q.x); mn[1] = std::min(mn[1], q.y); mn[2] = std::min(mn[2], q.z);
            mx[0] = std::max(mx[0], q.x); mx[1] = std::max(mx[1], q.y); mx[2] = std::max(mx[2], q.z);
        }
        // about eight points per cell if the cloud were a sheet spanning the two longest extents of its box
        float e[3] = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
        std::sort(e, e + 3);
        const double area = std::max((double)e[2] * (double)e[1], 1e-6);
        leaf = (float)std::max(std::sqrt(8.0 * area / (double)p.size()), (double)e[2] / 1000.0 + 1e-6);
        for (int a = 0; a < 3; ++a) { lo[a] = INT32_MAX; hi[a] = INT32_MIN; }
        for (size_t i = 0; i < p.size(); ++i) {
            const int c[3] = {(int)std::floor(p[i].x / leaf), (int)std::floor(p[i].y / leaf), (int)std::floor(p[i].z / leaf)};
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], c[a]); hi[a] = std::max(hi[a], c[a]); }
            cells[key(c[0], c[1], c[2])].push_back((int)i);
        }
    }
    // k best (distance, index) pairs, ascending, ties to the lower index
    void knn(const P3& q, int k, std::vector<std::pair<float, int>>& out) const {
        out.clear();
        if (!pts || pts->empty()) return;
        int c[3] = {(int)std::floor(q.x / leaf), (int)std::floor(q.y / leaf), (int)std::floor(q.z / leaf)};
        for (int a = 0; a < 3; ++a) c[a] = std::min(std::max(c[a], lo[a]), hi[a]);
        const float qv[3] = {q.x, q.y, q.z};
        const int rmax = std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]) + 1;
        for (int r = 0; r <= rmax; ++r) {
            int b0[3], b1[3];
            for (int a = 0; a < 3; ++a) { b0[a] = std::max(c[a] - r, lo[a]); b1[a] = std::min(c[a] + r, hi[a]); }
            for (int z = b0[2]; z <= b1[2]; ++z)
                for (int y = b0[1]; y <= b1[1]; ++y)
                    for (int x = b0[0]; x <= b1[0]; ++x) {
                        if (r > 0 && std::abs(x - c[0]) < r && std::abs(y - c[1]) < r && std::abs(z - c[2]) < r) continue;
                        auto it = cells.find(key(x, y, z));
                        if (it == cells.end()) continue;
                        for (int i : it->second) out.emplace_back(sqdist(q, (*pts)[i]), i);
                    }
            bool all = true;
            double lb = 1e300;
            for (int a = 0; a < 3; ++a) {
                if (b0[a] > lo[a]) { all = false; lb = std::min(lb, (double)qv[a] - (double)b0[a] * (double)leaf); }
                if (b1[a] < hi[a]) { all = false; lb = std::min(lb, (double)(b1[a] + 1) * (double)leaf - (double)qv[a]); }
            }
            if (all) break;
            if ((int)out.size() >= k) {
                std::nth_element(out.begin(), out.begin() + (k - 1), out.end());
                const double kth = out[k - 1].first;
                lb -= 1e-4 * leaf;
                if (lb > 0 && kth < lb * lb) break;
            }
        }
        const size_t kk = std::min<size_t>(k, out.size());
        std::partial_sort(out.begin(), out.begin() + kk, out.end());
        out.resize(kk);
    }
};

void knn_brute(const std::vector<P3>& pts, const P3& q, int k, std::vector<std::pair<float, int>>& out) {
    out.resize(pts.size());
    for (size_t i = 0; i < pts.size(); ++i) out[i] = {sqdist(q, pts[i]), (int)i};
    const size_t kk = std::min<size_t>(k, out.size());
    std::partial_sort(out.begin(), out.begin() + kk, out.end());
    out.resize(kk);
}

// ---- pcl::BFGS (see header).  Status values of BFGSSpace.
enum { NegativeGradientEpsilon = -3, NotStarted = -2, Running = -1, Success = 0, NoProgress = 1 };

struct Functor {
    std::function<double(const V6&)> f;
    std::function<void(const V6&, V6&)> df;
    std::function<void(const V6&, double&, V6&)> fdf;
};

struct BFGS {
    struct Parameters {
        int max_iters = 400, bracket_iters = 100, section_iters = 100;
        double rho = 0.01, sigma = 0.01, tau1 = 9, tau2 = 0.05, tau3 = 0.5, step_size = 1;
        int order = 3;
    } parameters;
    Functor& functor;
    int calls[3] = {0, 0, 0};
    double f = 0, delta_f = 0, fp0 = 0, g0norm = 0, pnorm = 0;
    V6 x0, dx0, dg0, g0, dx, p, gradient;
    // wrapper
    double f_alpha = 0, df_alpha = 0, f_cache_key = 0, df_cache_key = 0, x_cache_key = 0, g_cache_key = 0;
    V6 x_alpha, g_alpha;

    explicit BFGS(Functor& fn) : functor(fn) {}
    static double dot(const V6& a, const V6& b) { double s = 0; for (int i = 0; i < 6; ++i) s += a.v[i] * b.v[i]; return s; }
    static double norm(const V6& a) { return std::sqrt(dot(a, a)); }

    void moveTo(double alpha) {
        if (alpha == x_cache_key) return;
        for (int i = 0; i < 6; ++i) x_alpha.v[i] = x0.v[i] + alpha * p.v[i];
        x_cache_key = alpha;
    }
    double slope() { return dot(g_alpha, p); }
    double applyF(double alpha) {
        if (alpha == f_cache_key) return f_alpha;
        moveTo(alpha);
        f_alpha = functor.f(x_alpha); ++calls[0];
        f_cache_key = alpha;
        return f_alpha;
    }
    double applyDF(double alpha) {
        if (alpha == df_cache_key) return df_alpha;
        moveTo(alpha);
        if (alpha != g_cache_key) { functor.df(x_alpha, g_alpha); ++calls[1]; g_cache_key = alpha; }
        df_alpha = slope();
        df_cache_key = alpha;
        return df_alpha;
    }
    void applyFDF(double alpha, double& fo, double& dfo) {
        if (alpha == f_cache_key && alpha == df_cache_key) { fo = f_alpha; dfo = df_alpha; return; }
        if (alpha == f_cache_key || alpha == df_cache_key) { fo = applyF(alpha); dfo = applyDF(alpha); return; }
        moveTo(alpha);
        functor.fdf(x_alpha, f_alpha, g_alpha); ++calls[2];
        f_cache_key = alpha;
        g_cache_key = alpha;
        df_alpha = slope();
        df_cache_key = alpha;
        fo = f_alpha;
        dfo = df_alpha;
    }
    void updatePosition(double alpha, V6& x, double& fo, V6& g) {
        double fa, dfa;
        applyFDF(alpha, fa, dfa);
        fo = fa;
        x = x_alpha;
        g = g_alpha;
    }
    void changeDirection() {
        x_alpha = x0; x_cache_key = 0.0;
        f_cache_key = 0.0;
        g_alpha = g0; g_cache_key = 0.0;
        df_alpha = slope(); df_cache_key = 0.0;
    }
    int minimizeInit(V6& x) {
        delta_f = 0;
        for (int i = 0; i < 6; ++i) dx.v[i] = 0;
        functor.fdf(x, f, gradient); ++calls[2];
        x0 = x;
        g0 = gradient;
        g0norm = norm(g0);
        for (int i = 0; i < 6; ++i) p.v[i] = gradient.v[i] * (-1 / g0norm);
        pnorm = norm(p);
        fp0 = -g0norm;
        x_alpha = x0; x_cache_key = 0;
        f_alpha = f; f_cache_key = 0;
        g_alpha = g0; g_cache_key = 0;
        df_alpha = slope(); df_cache_key = 0;
        return NotStarted;
    }
    static void checkExtremum(const double c[4], double x, double& xmin, double& fmin) {
        const double y = c[0] + x * (c[1] + x * (c[2] + x * c[3]));  // Eigen::poly_eval (Horner for |x| <= 1; the reversed form for |x| > 1 is algebraically the same)
        if (y < fmin) { xmin = x; fmin = y; }
    }
    static double interpolate(double a, double fa, double fpa, double b, double fb, double fpb, double xmin, double xmax, int order) {
        double y, alpha, ymin, ymax;
        ymin = (xmin - a) / (b - a);
        ymax = (xmax - a) / (b - a);
        if (ymin > ymax) std::swap(ymin, ymax);
        if (order > 2 && !(fpb != fpa) && fpb != std::numeric_limits<double>::infinity()) {
            fpa = fpa * (b - a);
            fpb = fpb * (b - a);
            const double eta = 3 * (fb - fa) - 2 * fpa - fpb, xi = fpa + fpb - 2 * (fb - fa);
            const double c[4] = {fa, fpa, eta, xi};
            y = ymin;
            double fmin = c[0] + ymin * (c[1] + ymin * (c[2] + ymin * c[3]));
            checkExtremum(c, ymax, y, fmin);
            // PolynomialSolver<Scalar, 2> on (c1, 2 c2, 3 c3)
            const double p0 = c[1], p1 = 2 * c[2], p2 = 3 * c[3];
            const double a2 = 2 * p2, disc = p1 * p1 - 4 * p0 * p2;
            if (0 < disc) {
                const double dr = std::sqrt(disc);
                double y0 = (-p1 - dr) / a2, y1 = (-p1 + dr) / a2;
                if (y0 > y1) std::swap(y0, y1);
                if (y0 > ymin && y0 < ymax) checkExtremum(c, y0, y, fmin);
                if (y1 > ymin && y1 < ymax) checkExtremum(c, y1, y, fmin);
            } else if (0 == disc) {
                const double y0 = -p1 / a2;
                if (y0 > ymin && y0 < ymax) checkExtremum(c, y0, y, fmin);
            }
        } else {
            fpa = fpa * (b - a);
            const double fl = fa + ymin * (fpa + ymin * (fb - fa - fpa));
            const double fh = fa + ymax * (fpa + ymax * (fb - fa - fpa));
            const double c = 2 * (fb - fa - fpa);
            y = ymin;
            double fmin = fl;
            if (fh < fmin) { y = ymax; fmin = fh; }
            if (c > a) {
                const double z = -fpa / c;
                if (z > ymin && z < ymax) {
                    const double fz = fa + z * (fpa + z * (fb - fa - fpa));
                    if (fz < fmin) { y = z; fmin = fz; }
                }
            }
        }
        alpha = a + y * (b - a);
        return alpha;
    }
    int lineSearch(double rho, double sigma, double tau1, double tau2, double tau3, int order, double alpha1, double& alpha_new) {
        double f0, fp0l, falpha, falpha_prev, fpalpha, fpalpha_prev, delta, alpha_next;
        double alpha = alpha1, alpha_prev = 0.0;
        double a, b, fa, fb, fpa, fpb;
        int i = 0;
        applyFDF(0.0, f0, fp0l);
        falpha_prev = f0;
        fpalpha_prev = fp0l;
        a = 0.0; b = alpha;
        fa = f0; fb = 0.0;
        fpa = fp0l; fpb = 0.0;
        while (i++ < parameters.bracket_iters) {
            falpha = applyF(alpha);
            if (falpha > f0 + alpha * rho * fp0l || falpha >= falpha_prev) {
                a = alpha_prev; fa = falpha_prev; fpa = fpalpha_prev;
                b = alpha; fb = falpha; fpb = std::numeric_limits<double>::quiet_NaN();
                break;
            }
            fpalpha = applyDF(alpha);
            if (std::fabs(fpalpha) <= -sigma * fp0l) { alpha_new = alpha; return Success; }
            if (fpalpha >= 0) {
                a = alpha; fa = falpha; fpa = fpalpha;
                b = alpha_prev; fb = falpha_prev; fpb = fpalpha_prev;
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
        while (i++ < parameters.section_iters) {
            delta = b - a;
            {
                const double lower = a + tau2 * delta, upper = b - tau3 * delta;
                alpha = interpolate(a, fa, fpa, b, fb, fpb, lower, upper, order);
            }
            falpha = applyF(alpha);
            if ((a - alpha) * fpa <= std::numeric_limits<double>::epsilon()) return NoProgress;
            if (falpha > f0 + rho * alpha * fp0l || falpha >= fa) {
                b = alpha; fb = falpha; fpb = std::numeric_limits<double>::quiet_NaN();
            } else {
                fpalpha = applyDF(alpha);
                if (std::fabs(fpalpha) <= -sigma * fp0l) { alpha_new = alpha; return Success; }
                if (((b - a) >= 0 && fpalpha >= 0) || ((b - a) <= 0 && fpalpha <= 0)) {
                    b = a; fb = fa; fpb = fpa;
                    a = alpha; fa = falpha; fpa = fpalpha;
                } else {
                    a = alpha; fa = falpha; fpa = fpalpha;
                }
            }
        }
        return Success;
    }
    int minimizeOneStep(V6& x) {
        double alpha = 0.0, alpha1;
        const double f0 = f;
        if (pnorm == 0.0 || g0norm == 0.0 || fp0 == 0) {
            for (int i = 0; i < 6; ++i) dx.v[i] = 0;
            return NoProgress;
        }
        if (delta_f < 0) {
            const double del = std::max(-delta_f, 10 * std::numeric_limits<double>::epsilon() * std::fabs(f0));
            alpha1 = std::min(1.0, 2.0 * del / (-fp0));
        } else {
            alpha1 = std::fabs(parameters.step_size);
        }
        const int status = lineSearch(parameters.rho, parameters.sigma, parameters.tau1, parameters.tau2, parameters.tau3, parameters.order, alpha1, alpha);
        if (status != Success) return status;
        updatePosition(alpha, x, f, gradient);
        delta_f = f - f0;
        {
            for (int i = 0; i < 6; ++i) dx0.v[i] = x.v[i] - x0.v[i];
            dx = dx0;
            for (int i = 0; i < 6; ++i) dg0.v[i] = gradient.v[i] - g0.v[i];
            const double dxg = dot(dx0, gradient), dgg = dot(dg0, gradient), dxdg = dot(dx0, dg0), dgnorm = norm(dg0);
            double A, B;
            if (dxdg != 0) {
                B = dxg / dxdg;
                A = -(1.0 + dgnorm * dgnorm / dxdg) * B + dgg / dxdg;
            } else {
                B = 0;
                A = 0;
            }
            for (int i = 0; i < 6; ++i) p.v[i] = -A * dx0.v[i];
            for (int i = 0; i < 6; ++i) p.v[i] += gradient.v[i];
            for (int i = 0; i < 6; ++i) p.v[i] += -B * dg0.v[i];
        }
        g0 = gradient;
        x0 = x;
        g0norm = norm(g0);
        pnorm = norm(p);
        const double dir = (dot(p, gradient) > 0) ? -1.0 : 1.0;
        for (int i = 0; i < 6; ++i) p.v[i] *= dir / pnorm;
        pnorm = norm(p);
        fp0 = dot(p, g0);
        changeDirection();
        return Success;
    }
    int testGradient(double epsilon) {
        if (epsilon < 0) return NegativeGradientEpsilon;
        return g0norm < epsilon ? Success : Running;
    }
};

// the driving loop of estimateRigidTransformationBFGS (gicp_omp_impl.hpp:211-241)
int run_bfgs(Functor& functor, V6& x, int max_inner_iterations_, int* inner, int* calls3) {
    const double gradient_tol = 1e-2;
    BFGS bfgs(functor);
    bfgs.parameters.sigma = 0.01;
    bfgs.parameters.rho = 0.01;
    bfgs.parameters.tau1 = 9;
    bfgs.parameters.tau2 = 0.05;
    bfgs.parameters.tau3 = 0.5;
    bfgs.parameters.order = 3;
    int inner_iterations_ = 0;
    int result = bfgs.minimizeInit(x);
    result = Running;
    do {
        inner_iterations_++;
        result = bfgs.minimizeOneStep(x);
        if (result) break;
        result = bfgs.testGradient(gradient_tol);
    } while (result == Running && inner_iterations_ < max_inner_iterations_);
    if (inner) *inner = inner_iterations_;
    if (calls3) for (int i = 0; i < 3; ++i) calls3[i] = bfgs.calls[i];
    return result;
}

}  // namespace

struct Gicp {
    orc_gicp_params prm;
    int nthreads = 1;
    std::vector<P3> input, target;        // input_, target_
    ExactSearch tree, tree_reciprocal;    // tree_ (target), tree_reciprocal_ (input)
    std::vector<M3d> input_cov, target_cov;
    std::vector<M4f> mahalanobis;
    M4f base_transformation;
    // tmp_* of the functor
    const std::vector<P3>* tmp_src = nullptr;
    const std::vector<int>*tmp_idx_src = nullptr, *tmp_idx_tgt = nullptr;
    std::vector<int> last_src_idx, last_tgt_idx;
    std::vector<P3> last_output;

    static M4f identity() {
        M4f t;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) t.m[i][j] = i == j ? 1.0f : 0.0f;
        return t;
    }

    // computeCovariances (gicp_omp_impl.hpp:49-123)
    bool compute_covariances(const std::vector<P3>& cloud, const ExactSearch& kdtree, std::vector<M3d>& cloud_covariances) const {
        const int k_correspondences_ = prm.k_correspondences;
        if (k_correspondences_ > (int)cloud.size()) return false;
        cloud_covariances.resize(cloud.size());
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 64)
        for (size_t i = 0; i < cloud.size(); ++i) {
            std::vector<std::pair<float, int>> nn;
            const P3& query_point = cloud[i];
            double mean[3] = {0, 0, 0};
            M3d& cov = cloud_covariances[i];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) cov.m[a][b] = 0;
            kdtree.knn(query_point, k_correspondences_, nn);
            for (int j = 0; j < k_correspondences_; ++j) {
                const P3& pt = cloud[nn[j].second];
                mean[0] += pt.x;
                mean[1] += pt.y;
                mean[2] += pt.z;
                cov.m[0][0] += pt.x * pt.x;   // float products, accumulated in double (:86-94)
                cov.m[1][0] += pt.y * pt.x;
                cov.m[1][1] += pt.y * pt.y;
                cov.m[2][0] += pt.z * pt.x;
                cov.m[2][1] += pt.z * pt.y;
                cov.m[2][2] += pt.z * pt.z;
            }
            for (int a = 0; a < 3; ++a) mean[a] /= (double)k_correspondences_;
            for (int k = 0; k < 3; ++k)
                for (int l = 0; l <= k; ++l) {
                    cov.m[k][l] /= (double)k_correspondences_;
                    cov.m[k][l] -= mean[k] * mean[l];
                    cov.m[l][k] = cov.m[k][l];
                }
            double U[9], V[9], sv[3];
            jacobi_svd<3>(&cov.m[0][0], U, V, sv);
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) cov.m[a][b] = 0;
            for (int k = 0; k < 3; ++k) {
                const double col[3] = {U[0 * 3 + k], U[1 * 3 + k], U[2 * 3 + k]};
                double v = 1.;
                if (k == 2) v = prm.gicp_epsilon;
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) cov.m[a][b] += (v * col[a]) * col[b];
            }
        }
        return true;
    }

    // applyState (gicp_omp_impl.hpp:518-529)
    static void apply_state(M4f& t, const V6& x) {
        struct Q { float x, y, z, w; };
        auto from_aa = [](float angle, float ax, float ay, float az) {
            const float ha = 0.5f * angle;
            const float s = (float)std::sin((double)ha), c = (float)std::cos((double)ha);  // correctly rounded sinf / cosf
            return Q{s * ax, s * ay, s * az, c};
        };
        auto mul = [](const Q& a, const Q& b) {
            Q r;
            r.x = (a.x * b.w - a.z * b.y) + (a.y * b.z + a.w * b.x);
            r.y = (a.y * b.w - a.x * b.z) + (a.z * b.x + a.w * b.y);
            r.z = (a.z * b.w - a.y * b.x) + (a.x * b.y + a.w * b.z);
            r.w = (a.w * b.w - a.x * b.x) - (a.z * b.z + a.y * b.y);
            return r;
        };
        const Q q = mul(mul(from_aa((float)x.v[5], 0, 0, 1), from_aa((float)x.v[4], 0, 1, 0)), from_aa((float)x.v[3], 1, 0, 0));
        float R[3][3];
        const float tx = 2.0f * q.x, ty = 2.0f * q.y, tz = 2.0f * q.z;
        const float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
        const float txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
        const float tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
        R[0][0] = 1.0f - (tyy + tzz); R[0][1] = txy - twz; R[0][2] = txz + twy;
        R[1][0] = txy + twz; R[1][1] = 1.0f - (txx + tzz); R[1][2] = tyz - twx;
        R[2][0] = txz - twy; R[2][1] = tyz + twx; R[2][2] = 1.0f - (txx + tyy);
        float out[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) out[i][j] = (R[i][0] * t.m[0][j] + R[i][1] * t.m[1][j]) + R[i][2] * t.m[2][j];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) t.m[i][j] = out[i][j];
        t.m[0][3] += (float)x.v[0];
        t.m[1][3] += (float)x.v[1];
        t.m[2][3] += (float)x.v[2];
        t.m[3][3] += 0.0f;
    }

    // computeRDerivative (gicp_omp_impl.hpp:127-184)
    static void compute_r_derivative(const V6& x, const double R[3][3], V6& g) {
        double dR_dPhi[3][3], dR_dTheta[3][3], dR_dPsi[3][3];
        const double phi = x.v[3], theta = x.v[4], psi = x.v[5];
        const double cphi = std::cos(phi), sphi = std::sin(phi);
        const double ctheta = std::cos(theta), stheta = std::sin(theta);
        const double cpsi = std::cos(psi), spsi = std::sin(psi);
        dR_dPhi[0][0] = 0.; dR_dPhi[1][0] = 0.; dR_dPhi[2][0] = 0.;
        dR_dPhi[0][1] = sphi * spsi + cphi * cpsi * stheta;
        dR_dPhi[1][1] = -cpsi * sphi + cphi * spsi * stheta;
        dR_dPhi[2][1] = cphi * ctheta;
        dR_dPhi[0][2] = cphi * spsi - cpsi * sphi * stheta;
        dR_dPhi[1][2] = -cphi * cpsi - sphi * spsi * stheta;
        dR_dPhi[2][2] = -ctheta * sphi;
        dR_dTheta[0][0] = -cpsi * stheta;
        dR_dTheta[1][0] = -spsi * stheta;
        dR_dTheta[2][0] = -ctheta;
        dR_dTheta[0][1] = cpsi * ctheta * sphi;
        dR_dTheta[1][1] = ctheta * sphi * spsi;
        dR_dTheta[2][1] = -sphi * stheta;
        dR_dTheta[0][2] = cphi * cpsi * ctheta;
        dR_dTheta[1][2] = cphi * ctheta * spsi;
        dR_dTheta[2][2] = -cphi * stheta;
        dR_dPsi[0][0] = -ctheta * spsi;
        dR_dPsi[1][0] = cpsi * ctheta;
        dR_dPsi[2][0] = 0.;
        dR_dPsi[0][1] = -cphi * cpsi - sphi * spsi * stheta;
        dR_dPsi[1][1] = -cphi * spsi + cpsi * sphi * stheta;
        dR_dPsi[2][1] = 0.;
        dR_dPsi[0][2] = cpsi * sphi - cphi * spsi * stheta;
        dR_dPsi[1][2] = sphi * spsi + cphi * cpsi * stheta;
        dR_dPsi[2][2] = 0.;
        auto inner = [&](const double m1[3][3]) {  // matricesInnerProd (gicp_omp.h:312-322)
            double r = 0.;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) r += m1[j][i] * R[i][j];
            return r;
        };
        g.v[3] = inner(dR_dPhi);
        g.v[4] = inner(dR_dTheta);
        g.v[5] = inner(dR_dPsi);
    }

    static void mul4(const M4f& t, const float p[4], float out[4]) {  // Matrix4f * Vector4f, terms added left to right
        for (int i = 0; i < 4; ++i) out[i] = ((t.m[i][0] * p[0] + t.m[i][1] * p[1]) + t.m[i][2] * p[2]) + t.m[i][3] * p[3];
    }

    // OptimizationFunctorWithIndices::operator() (gicp_omp_impl.hpp:245-275)
    double functor_f(const V6& x) const {
        M4f transformation_matrix = base_transformation;
        apply_state(transformation_matrix, x);
        double f = 0;
        const int m = (int)tmp_idx_src->size();
        for (int i = 0; i < m; ++i) {
            const P3& s = (*tmp_src)[(*tmp_idx_src)[i]];
            const P3& t = target[(*tmp_idx_tgt)[i]];
            const float p_src[4] = {s.x, s.y, s.z, 1.0f}, p_tgt[4] = {t.x, t.y, t.z, 1.0f};
            float res[4], mr[4];
            mul4(transformation_matrix, p_src, res);
            for (int a = 0; a < 4; ++a) res[a] -= p_tgt[a];
            const M4f& maha = mahalanobis[(*tmp_idx_src)[i]];
            mul4(maha, res, mr);
            const float ret = ((res[0] * mr[0] + res[1] * mr[1]) + res[2] * mr[2]) + res[3] * mr[3];
            f += (double)ret;
        }
        return f / m;
    }
    // df (gicp_omp_impl.hpp:279-340)
    void functor_df(const V6& x, V6& g) const {
        M4f transformation_matrix = base_transformation;
        apply_state(transformation_matrix, x);
        double Rm[4][4] = {{0}}, gsum[4] = {0, 0, 0, 0};
        const int m = (int)tmp_idx_src->size();
        for (int i = 0; i < m; ++i) {
            const P3& s = (*tmp_src)[(*tmp_idx_src)[i]];
            const P3& t = target[(*tmp_idx_tgt)[i]];
            const float p_src[4] = {s.x, s.y, s.z, 1.0f}, p_tgt[4] = {t.x, t.y, t.z, 1.0f};
            float pp[4];
            mul4(transformation_matrix, p_src, pp);
            const double res[4] = {(double)(pp[0] - p_tgt[0]), (double)(pp[1] - p_tgt[1]), (double)(pp[2] - p_tgt[2]), 0.0};
            const M4f& maha = mahalanobis[(*tmp_idx_src)[i]];
            double temp[4];
            for (int a = 0; a < 4; ++a)
                temp[a] = (((double)maha.m[a][0] * res[0] + (double)maha.m[a][1] * res[1]) + (double)maha.m[a][2] * res[2]) + (double)maha.m[a][3] * res[3];
            mul4(base_transformation, p_src, pp);
            const double p_src3[4] = {(double)pp[0], (double)pp[1], (double)pp[2], 0.0};
            for (int a = 0; a < 4; ++a) gsum[a] += temp[a];
            for (int a = 0; a < 4; ++a)
                for (int b = 0; b < 4; ++b) Rm[a][b] += p_src3[a] * temp[b];
        }
        for (int a = 0; a < 6; ++a) g.v[a] = 0;
        for (int a = 0; a < 3; ++a) g.v[a] += gsum[a];
        for (int a = 0; a < 3; ++a) g.v[a] *= 2.0 / m;
        double R3[3][3];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) R3[a][b] = Rm[a][b] * (2.0 / m);
        compute_r_derivative(x, R3, g);
    }
    // fdf (gicp_omp_impl.hpp:344-368)
    void functor_fdf(const V6& x, double& f, V6& g) const {
        M4f transformation_matrix = base_transformation;
        apply_state(transformation_matrix, x);
        f = 0;
        for (int a = 0; a < 6; ++a) g.v[a] = 0;
        double R[3][3] = {{0}};
        const int m = (int)tmp_idx_src->size();
        for (int i = 0; i < m; ++i) {
            const P3& s = (*tmp_src)[(*tmp_idx_src)[i]];
            const P3& t = target[(*tmp_idx_tgt)[i]];
            const float p_src[4] = {s.x, s.y, s.z, 1.0f}, p_tgt[4] = {t.x, t.y, t.z, 1.0f};
            float pp[4];
            mul4(transformation_matrix, p_src, pp);
            const double res[3] = {(double)(pp[0] - p_tgt[0]), (double)(pp[1] - p_tgt[1]), (double)(pp[2] - p_tgt[2])};
            const M4f& maha = mahalanobis[(*tmp_idx_src)[i]];
            double temp[3];
            for (int a = 0; a < 3; ++a) temp[a] = ((double)maha.m[a][0] * res[0] + (double)maha.m[a][1] * res[1]) + (double)maha.m[a][2] * res[2];
            f += (res[0] * temp[0] + res[1] * temp[1]) + res[2] * temp[2];
            for (int a = 0; a < 3; ++a) g.v[a] += temp[a];
            mul4(base_transformation, p_src, pp);
            const double p_src3[3] = {(double)pp[0], (double)pp[1], (double)pp[2]};
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) R[a][b] += p_src3[a] * temp[b];
        }
        f /= double(m);
        for (int a = 0; a < 3; ++a) g.v[a] *= double(2.0 / m);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) R[a][b] *= 2.0 / m;
        compute_r_derivative(x, R, g);
    }

    // estimateRigidTransformationBFGS (gicp_omp_impl.hpp:188-242); false = exception thrown
    bool estimate_rigid_transformation_bfgs(const std::vector<P3>& cloud_src, const std::vector<int>& indices_src, const std::vector<int>& indices_tgt,
                                            M4f& transformation_matrix, int* inner, int* status, int* calls3) {
        if (indices_src.size() < 4) return false;  // NotEnoughPointsException
        V6 x;
        x.v[0] = transformation_matrix.m[0][3];
        x.v[1] = transformation_matrix.m[1][3];
        x.v[2] = transformation_matrix.m[2][3];
        x.v[3] = (float)std::atan2((double)transformation_matrix.m[2][1], (double)transformation_matrix.m[2][2]);  // std::atan2(float, float)
        x.v[4] = (float)std::asin((double)-transformation_matrix.m[2][0]);                                         // asin(float)
        x.v[5] = (float)std::atan2((double)transformation_matrix.m[1][0], (double)transformation_matrix.m[0][0]);
        tmp_src = &cloud_src;
        tmp_idx_src = &indices_src;
        tmp_idx_tgt = &indices_tgt;
        Functor functor;
        functor.f = [this](const V6& xx) { return functor_f(xx); };
        functor.df = [this](const V6& xx, V6& g) { functor_df(xx, g); };
        functor.fdf = [this](const V6& xx, double& f, V6& g) { functor_fdf(xx, f, g); };
        const int result = run_bfgs(functor, x, prm.max_inner_iterations, inner, calls3);
        if (status) *status = result;
        if (result == NoProgress || result == Success || *inner == prm.max_inner_iterations) {
            transformation_matrix = identity();
            apply_state(transformation_matrix, x);
            return true;
        }
        return false;  // SolverDidntConvergeException
    }

    // the correspondence part of one pass of the while loop (gicp_omp_impl.hpp:408-470)
    void correspondences(const std::vector<P3>& output, const M4f& transformation_, const M4f& guess, std::vector<int>& source_indices,
                         std::vector<int>& target_indices, std::vector<float>* dists) {
        const size_t N = output.size();
        const double dist_threshold = prm.corr_dist_threshold * prm.corr_dist_threshold;
        double transform_R[4][4] = {{0}};
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++)
                for (int k = 0; k < 4; k++) transform_R[i][j] += double(transformation_.m[i][k]) * double(guess.m[k][j]);
        double R[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[i][j] = transform_R[i][j];
        std::vector<int> nn_of(N, -1);
        std::vector<float> d_of(N, 0.f);
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 64)
        for (size_t i = 0; i < N; ++i) {
            std::vector<std::pair<float, int>> nn;
            const float q4[4] = {output[i].x, output[i].y, output[i].z, 1.0f};
            float q[4];
            mul4(transformation_, q4, q);
            tree.knn(P3{q[0], q[1], q[2]}, 1, nn);
            if (nn.empty()) continue;
            d_of[i] = nn[0].first;
            if ((double)nn[0].first < dist_threshold) {
                const M3d& C1 = input_cov[i];
                const M3d& C2 = target_cov[nn[0].second];
                double M[3][3], temp[3][3];
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) M[a][b] = (R[a][0] * C1.m[0][b] + R[a][1] * C1.m[1][b]) + R[a][2] * C1.m[2][b];
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) temp[a][b] = (M[a][0] * R[b][0] + M[a][1] * R[b][1]) + M[a][2] * R[b][2];
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) temp[a][b] += C2.m[a][b];
                double inv[9];
                inverse3(&temp[0][0], inv);
                M4f& M_ = mahalanobis[i];
                for (int a = 0; a < 4; ++a)
                    for (int b = 0; b < 4; ++b) M_.m[a][b] = 0.0f;
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) M_.m[a][b] = (float)inv[a * 3 + b];
                nn_of[i] = nn[0].second;
            }
        }
        // the reference collects (i, nn) pairs under an atomic counter and sorts them by source index (:462-470)
        source_indices.clear();
        target_indices.clear();
        for (size_t i = 0; i < N; ++i)
            if (nn_of[i] >= 0) {
                source_indices.push_back((int)i);
                target_indices.push_back(nn_of[i]);
            }
        if (dists) *dists = d_of;
    }

    bool ensure_covariances() {
        if (target_cov.empty()) {
            tree.build(target);
            if (!compute_covariances(target, tree, target_cov)) return false;
        }
        if (input_cov.empty()) {
            tree_reciprocal.build(input);
            if (!compute_covariances(input, tree_reciprocal, input_cov)) return false;
        }
        return true;
    }

    // pcl::Registration::align + computeTransformation (gicp_omp_impl.hpp:371-516)
    int align(const M4f& guess, M4f& final_transformation, orc_gicp_result* res) {
        M4f transformation_ = identity(), previous_transformation_ = identity();
        std::vector<P3> output = input;
        const size_t N = input.size();
        mahalanobis.assign(N, identity());
        if (!ensure_covariances()) return -1;
        base_transformation = identity();
        int nr_iterations_ = 0;
        bool converged_ = false;
        for (P3& p : output) {  // pcl::transformPointCloud(output, output, guess)
            const P3 q = p;
            p.x = ((guess.m[0][0] * q.x + guess.m[0][1] * q.y) + guess.m[0][2] * q.z) + guess.m[0][3];
            p.y = ((guess.m[1][0] * q.x + guess.m[1][1] * q.y) + guess.m[1][2] * q.z) + guess.m[1][3];
            p.z = ((guess.m[2][0] * q.x + guess.m[2][1] * q.y) + guess.m[2][2] * q.z) + guess.m[2][3];
        }
        int tot_calls[3] = {0, 0, 0}, last_inner = 0, last_status = 0, last_m = 0, tot_inner = 0;
        while (!converged_) {
            std::vector<int> source_indices, target_indices;
            correspondences(output, transformation_, guess, source_indices, target_indices, nullptr);
            last_m = (int)source_indices.size();
            previous_transformation_ = transformation_;
            int calls3[3] = {0, 0, 0};
            double delta = 0.;
            if (!estimate_rigid_transformation_bfgs(output, source_indices, target_indices, transformation_, &last_inner, &last_status, calls3)) break;
            for (int i = 0; i < 3; ++i) tot_calls[i] += calls3[i];
            tot_inner += last_inner;
            for (int k = 0; k < 4; k++)
                for (int l = 0; l < 4; l++) {
                    double ratio = 1;
                    if (k < 3 && l < 3) ratio = 1. / prm.rotation_epsilon;
                    else ratio = 1. / prm.transformation_epsilon;
                    const double c_delta = ratio * std::abs(previous_transformation_.m[k][l] - transformation_.m[k][l]);
                    if (c_delta > delta) delta = c_delta;
                }
            nr_iterations_++;
            if (nr_iterations_ >= prm.max_iterations || delta < 1) {
                converged_ = true;
                previous_transformation_ = transformation_;
            }
        }
        for (int i = 0; i < 4; ++i)  // final_transformation_ = previous_transformation_ * guess
            for (int j = 0; j < 4; ++j)
                final_transformation.m[i][j] = ((previous_transformation_.m[i][0] * guess.m[0][j] + previous_transformation_.m[i][1] * guess.m[1][j]) +
                                                previous_transformation_.m[i][2] * guess.m[2][j]) + previous_transformation_.m[i][3] * guess.m[3][j];
        if (res) {
            res->converged = converged_;
            res->iterations = nr_iterations_;
            res->last_m = last_m;
            res->last_inner = last_inner;
            res->last_status = last_status;
            res->inner_total = tot_inner;
            res->n_f = tot_calls[0];
            res->n_df = tot_calls[1];
            res->n_fdf = tot_calls[2];
        }
        return converged_ ? 0 : 2;
    }
};

namespace {
void load_cloud(const float* xyz, int64_t n, int64_t stride, std::vector<P3>& out) {
    out.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const float* p = (const float*)((const char*)xyz + i * stride);
        out[i] = P3{p[0], p[1], p[2]};
    }
}
M4f from_colmajor(const float* m) {
    M4f t;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) t.m[i][j] = m[j * 4 + i];
    return t;
}
void to_colmajor(const M4f& t, float* m) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) m[j * 4 + i] = t.m[i][j];
}
// the smooth test function of orc_bfgs_test: a coupled quartic bowl with its minimum at c
const double kC[6] = {0.3, -0.2, 0.5, 0.05, -0.04, 0.08};
const double kA[6] = {1.0, 2.5, 0.7, 4.0, 3.0, 1.5};
double test_f(const V6& x) {
    double s = 0;
    for (int i = 0; i < 6; ++i) { const double d = x.v[i] - kC[i]; s += kA[i] * d * d + 0.25 * d * d * d * d; }
    for (int i = 0; i < 5; ++i) s += 0.3 * (x.v[i] - kC[i]) * (x.v[i + 1] - kC[i + 1]);
    return s;
}
void test_g(const V6& x, V6& g) {
    for (int i = 0; i < 6; ++i) { const double d = x.v[i] - kC[i]; g.v[i] = 2 * kA[i] * d + d * d * d; }
    for (int i = 0; i < 5; ++i) { g.v[i] += 0.3 * (x.v[i + 1] - kC[i + 1]); g.v[i + 1] += 0.3 * (x.v[i] - kC[i]); }
}
}  // namespace
}  // namespace orc

using namespace orc;
struct orc_gicp { Gicp G; };

extern "C" {
orc_gicp* orc_gicp_create(const orc_gicp_params* p) {
    orc_gicp* h = new orc_gicp();
    h->G.prm = *p;
    h->G.nthreads = p->num_threads > 0 ? p->num_threads : omp_get_max_threads();
    h->G.base_transformation = Gicp::identity();
    return h;
}
void orc_gicp_destroy(orc_gicp* h) { delete h; }
void orc_gicp_set_target(orc_gicp* h, const float* xyz, int64_t n, int64_t stride) {
    load_cloud(xyz, n, stride, h->G.target);
    h->G.target_cov.clear();  // target_covariances_.reset() (gicp_omp.h:173-178)
}
void orc_gicp_set_source(orc_gicp* h, const float* xyz, int64_t n, int64_t stride) {
    load_cloud(xyz, n, stride, h->G.input);
    h->G.input_cov.clear();  // input_covariances_.reset() (gicp_omp.h:141-157)
}
int32_t orc_gicp_covariances(orc_gicp* h, int32_t which, double* cov9) {
    if (!h->G.ensure_covariances()) return -1;
    const std::vector<M3d>& c = which ? h->G.target_cov : h->G.input_cov;
    memcpy(cov9, c.data(), c.size() * sizeof(M3d));
    return 0;
}
int32_t orc_gicp_align(orc_gicp* h, const float* guess16_cm, float* final16_cm, orc_gicp_result* r) {
    M4f fin;
    const int rc = h->G.align(from_colmajor(guess16_cm), fin, r);
    if (rc >= 0) to_colmajor(fin, final16_cm);
    return rc;
}
int64_t orc_gicp_correspondences(orc_gicp* h, const float* trans16_cm, const float* guess16_cm, int32_t* tgt_idx, float* maha9, float* d2) {
    Gicp& G = h->G;
    if (!G.ensure_covariances()) return -1;
    const M4f guess = from_colmajor(guess16_cm), tr = from_colmajor(trans16_cm);
    G.last_output = G.input;
    for (P3& p : G.last_output) {
        const P3 q = p;
        p.x = ((guess.m[0][0] * q.x + guess.m[0][1] * q.y) + guess.m[0][2] * q.z) + guess.m[0][3];
        p.y = ((guess.m[1][0] * q.x + guess.m[1][1] * q.y) + guess.m[1][2] * q.z) + guess.m[1][3];
        p.z = ((guess.m[2][0] * q.x + guess.m[2][1] * q.y) + guess.m[2][2] * q.z) + guess.m[2][3];
    }
    G.mahalanobis.assign(G.input.size(), Gicp::identity());
    std::vector<float> dists;
    G.correspondences(G.last_output, tr, guess, G.last_src_idx, G.last_tgt_idx, &dists);
    const size_t N = G.input.size();
    if (tgt_idx) for (size_t i = 0; i < N; ++i) tgt_idx[i] = -1;
    for (size_t c = 0; c < G.last_src_idx.size(); ++c) {
        const int i = G.last_src_idx[c];
        if (tgt_idx) tgt_idx[i] = G.last_tgt_idx[c];
    }
    if (maha9)
        for (size_t i = 0; i < N; ++i)
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) maha9[i * 9 + a * 3 + b] = G.mahalanobis[i].m[a][b];
    if (d2) memcpy(d2, dists.data(), N * sizeof(float));
    return (int64_t)G.last_src_idx.size();
}
void orc_gicp_cost(orc_gicp* h, const double* x6, double* f_op, double* f_fdf, double* g_df6, double* g_fdf6) {
    Gicp& G = h->G;
    G.tmp_src = &G.last_output;
    G.tmp_idx_src = &G.last_src_idx;
    G.tmp_idx_tgt = &G.last_tgt_idx;
    V6 x, g1, g2;
    memcpy(x.v, x6, sizeof x.v);
    *f_op = G.functor_f(x);
    G.functor_df(x, g1);
    G.functor_fdf(x, *f_fdf, g2);
    memcpy(g_df6, g1.v, sizeof g1.v);
    memcpy(g_fdf6, g2.v, sizeof g2.v);
}
/* estimateRigidTransformationBFGS on the correspondences of the last orc_gicp_correspondences call, starting from trans16 */
int32_t orc_gicp_estimate(orc_gicp* h, float* trans16_cm, int32_t* inner, int32_t* status, int32_t* calls3) {
    Gicp& G = h->G;
    M4f t = from_colmajor(trans16_cm);
    const bool ok = G.estimate_rigid_transformation_bfgs(G.last_output, G.last_src_idx, G.last_tgt_idx, t, inner, status, calls3);
    to_colmajor(t, trans16_cm);
    return ok ? 0 : -1;
}
void orc_gicp_apply_state(const double* x6, float* t16_cm) {
    M4f t = Gicp::identity();
    V6 x;
    memcpy(x.v, x6, sizeof x.v);
    Gicp::apply_state(t, x);
    to_colmajor(t, t16_cm);
}
void orc_gicp_knn(const float* xyz, int64_t n, int64_t stride, const float* q, int64_t nq, int64_t qstride, int32_t k, int32_t brute,
                  int32_t* idx, float* d2) {
    std::vector<P3> pts, qs;
    load_cloud(xyz, n, stride, pts);
    load_cloud(q, nq, qstride, qs);
    ExactSearch es;
    if (!brute) es.build(pts);
#pragma omp parallel for schedule(dynamic, 32)
    for (int64_t i = 0; i < nq; ++i) {
        std::vector<std::pair<float, int>> nn;
        if (brute) knn_brute(pts, qs[i], k, nn);
        else es.knn(qs[i], k, nn);
        for (int j = 0; j < k; ++j) {
            idx[i * k + j] = j < (int)nn.size() ? nn[j].second : -1;
            d2[i * k + j] = j < (int)nn.size() ? nn[j].first : -1.0f;
        }
    }
}
int32_t orc_bfgs_test(double* x6, int32_t max_inner, int32_t* inner, int32_t* calls3) {
    Functor fn;
    fn.f = [](const V6& x) { return test_f(x); };
    fn.df = [](const V6& x, V6& g) { test_g(x, g); };
    fn.fdf = [](const V6& x, double& f, V6& g) { f = test_f(x); test_g(x, g); };
    V6 x;
    memcpy(x.v, x6, sizeof x.v);
    const int r = run_bfgs(fn, x, max_inner, inner, calls3);
    memcpy(x6, x.v, sizeof x.v);
    return r;
}
}
