// Test helper: compiles the PRODUCT's GICP arithmetic (pointcloud-slam_b200/csrc/gicp_math.cuh, the very header the CUDA
// kernels of gicp.cu include) with g++ and exposes it to pytest, so that the per-point math, the cost functor and the BFGS
// minimiser are checked against the oracle on the CPU; the GPU tests then only have to cover the kernels' indexing.
// Not a fallback: nothing in the product library links or calls this file.
#include "gicp_math.cuh"

#include <cstdint>
#include <cstring>
#include <vector>

using namespace b200::gicp;

namespace {
struct HostEval {  // the cost functor over explicit correspondence arrays, summed serially (the oracle's order)
    const float* src;  // n x 3, moved by the guess
    const float* tgt;  // n x 3
    const float* M;    // n x 9
    const int32_t* flag;
    int n, m;
    void sums(const double* x, double* tot) const {
        float T[12];
        apply_state(x, T);
        for (int a = 0; a < kAcc; ++a) tot[a] = 0.0;
        for (int i = 0; i < n; ++i)
            if (flag[i] >= 0) point_terms(T, src[3 * i], src[3 * i + 1], src[3 * i + 2], tgt[3 * i], tgt[3 * i + 1], tgt[3 * i + 2], M + 9 * i, tot);
    }
    double f(const double* x) const {
        double tot[kAcc];
        sums(x, tot);
        return cost_f(tot, m);
    }
    void df(const double* x, double* g) const {
        double tot[kAcc];
        sums(x, tot);
        cost_gradient(tot, m, x, g);
    }
    void fdf(const double* x, double& fo, double* g) const {
        double tot[kAcc];
        sums(x, tot);
        fo = cost_f_fdf(tot, m);
        cost_gradient(tot, m, x, g);
    }
};
// the same smooth 6-D function as oracle/gicp_oracle.cpp's orc_bfgs_test (written independently there)
struct TestEval {
    static constexpr double c[6] = {0.3, -0.2, 0.5, 0.05, -0.04, 0.08};
    static constexpr double a[6] = {1.0, 2.5, 0.7, 4.0, 3.0, 1.5};
    double f(const double* x) const {
        double s = 0;
        for (int i = 0; i < 6; ++i) {
            const double d = x[i] - c[i];
            s += a[i] * d * d + 0.25 * d * d * d * d;
        }
        for (int i = 0; i < 5; ++i) s += 0.3 * (x[i] - c[i]) * (x[i + 1] - c[i + 1]);
        return s;
    }
    void df(const double* x, double* g) const {
        for (int i = 0; i < 6; ++i) {
            const double d = x[i] - c[i];
            g[i] = 2 * a[i] * d + d * d * d;
        }
        for (int i = 0; i < 5; ++i) {
            g[i] += 0.3 * (x[i + 1] - c[i + 1]);
            g[i + 1] += 0.3 * (x[i] - c[i]);
        }
    }
    void fdf(const double* x, double& fo, double* g) const {
        fo = f(x);
        df(x, g);
    }
};
constexpr double TestEval::c[6];
constexpr double TestEval::a[6];
int count(const int32_t* flag, int n) {
    int m = 0;
    for (int i = 0; i < n; ++i) m += flag[i] >= 0;
    return m;
}
}  // namespace

extern "C" {
void hh_apply_state(const double* x6, float* T12) { apply_state(x6, T12); }
void hh_state_from_transform(const float* T12, double* x6) { state_from_transform(T12, x6); }
void hh_cov_regularize(const double* mean_sum3, const double* c6, int32_t k, double eps, double* out9) { cov_regularize(mean_sum3, c6, k, eps, out9); }
void hh_svd3_u(const double* A9, double* U9, double* sv3) { jacobi_svd3_u(A9, U9, sv3); }
void hh_mahalanobis(const float* T12, const float* G12, const double* C1, const double* C2, float* M9) {
    double R[9];
    rotation_of_product(T12, G12, R);
    mahalanobis3(R, C1, C2, M9);
}
void hh_cost(const float* src, const float* tgt, const float* M, const int32_t* flag, int32_t n, const double* x6, double* f_op, double* f_fdf, double* g_df6,
             double* g_fdf6) {
    HostEval ev{src, tgt, M, flag, n, count(flag, n)};
    *f_op = ev.f(x6);
    ev.df(x6, g_df6);
    ev.fdf(x6, *f_fdf, g_fdf6);
}
// estimateRigidTransformationBFGS as k_g_bfgs runs it: T12 in / out (3x4 row-major); returns the solver status
int32_t hh_estimate(const float* src, const float* tgt, const float* M, const int32_t* flag, int32_t n, int32_t max_inner, float* T12, int32_t* inner,
                    int32_t* calls3) {
    HostEval ev{src, tgt, M, flag, n, count(flag, n)};
    double x[6];
    state_from_transform(T12, x);
    const int r = minimize_rigid(ev, x, max_inner, inner, calls3);
    apply_state(x, T12);
    return r;
}
int32_t hh_bfgs_test(double* x6, int32_t max_inner, int32_t* inner, int32_t* calls3) {
    TestEval ev;
    return minimize_rigid(ev, x6, max_inner, inner, calls3);
}
double hh_transform_delta(const float* prev12, const float* cur12, double rot_eps, double trans_eps) { return transform_delta(prev12, cur12, rot_eps, trans_eps); }
}
