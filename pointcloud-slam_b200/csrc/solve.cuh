// b200reg — the filter step of one IEKF pass, executed by ONE thread block (>= 96 threads).
//
// Follows esekf::update_iterated_dyn_share_modified (IKFoM_toolkit/esekfom/esekfom.hpp:1546-1831):
// dx = x (-) x_prop, re-linearised covariance P = J P_prop J^T, gain, x (+)= dx_, convergence
// bookkeeping, final covariance.
//
// One deliberate algebraic change, equal to the reference within fp64 rounding (IEKF):
// the reference forms  P_inv = ((P/R)^-1 + E HTH E^T)^-1  with two dense 23x23 inversions
// (esekfom.hpp:1685-1706; E selects the first 12 tangent coordinates) and then only uses
// P_inv[:, :12] * HTH.  HTH is non-zero only in its leading m x m block A (m = 6, or 12 with
// extrinsic estimation), so only P_inv[:, :m] matters, and by the conditional-Gaussian identities,
// with B = P/R split into the measured coordinates a (first m) and the rest b:
//        P_inv[a, a] = (B_aa^-1 + A)^-1            P_inv[:, a] = B[:, a] * B_aa^-1 * P_inv[a, a]
// i.e. two m x m SPD inversions (done in registers by one warp) instead of two 23 x 23 ones, with
// the same conditioning as the reference's information form.  It also covers the reference's
// small-N branch (N_eff < 23, esekfom.hpp:1618-1651), which is the same gain in its dual form.
#pragma once
#include "manifold.cuh"
#include "pointmath.cuh"

namespace b200 {

constexpr int NS = 23;     // state DOF
constexpr int NPART = 91;  // 78 unique HTH entries + 12 HTh entries + effective-point count
constexpr int MAXB = 160;  // upper bound on k_obs blocks (partials rows)

struct Ctl {
    // inputs (H2D header)
    double x[26];
    double P[NS * NS];
    int n, prev_n;
    int pad0[2];
    // loop state
    double x_prop[26];
    double P_prop[NS * NS];
    int iter;      // loop variable i of esekfom.hpp:1539
    int converge;  // dyn_share.converge
    int done;
    int t;
    unsigned int ticket;  // blocks of the current k_obs launch that have published their partials
    int fault;            // the filter block's watchdog fired: worker tickets never arrived
    int pad1[2];
    // stats
    int passes, knn_passes, any_valid, converged;
    int n_eff[B200_MAX_PASSES], knn[B200_MAX_PASSES];
    PassConsts pc;
    double x_in[B200_MAX_PASSES][26];
    double HtH[B200_MAX_PASSES][144];
    double Hth[B200_MAX_PASSES][12];
    long long dbg[B200_MAX_PASSES][16];  // clock64() stage stamps (profiling aid)
};

static __constant__ unsigned char c_pair_a[78];
static __constant__ unsigned char c_pair_b[78];
// Active accumulator columns.  Without extrinsic estimation only the leading 6 x 6 block of h_x^T h_x and the first 6
// entries of h_x^T h are non-zero (laser_mapping.cc:687-694), so a pass accumulates, publishes and reduces
// NCOL6 = 21 + 6 + 1 columns instead of NPART = 78 + 12 + 1.  Column c multiplies row entries (a, b): b == 12 is the
// residual h, a == 255 marks the effective-point counter.
constexpr int NCOL6 = 28;
static __constant__ unsigned char c_col_a[2][NPART];
static __constant__ unsigned char c_col_b[2][NPART];

__device__ inline void make_pass_consts(const double* x, PassConsts& pc) {
    using namespace mf;
    const Q rot = ldq(x + 3), offR = ldq(x + 7);
    const Q qd = qmul(rot, offR);
    pc.qx = (float)qd.x; pc.qy = (float)qd.y; pc.qz = (float)qd.z; pc.qw = (float)qd.w;
    double td[3];
    qrot(rot, x + 11, td);
    pc.tx = (float)(td[0] + x[0]); pc.ty = (float)(td[1] + x[1]); pc.tz = (float)(td[2] + x[2]);
    double Ro[9], Rr[9];
    qtoR(offR, Ro);
    qtoR(rot, Rr);
    for (int i = 0; i < 9; ++i) pc.offR[i] = (float)Ro[i];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) pc.Rt[i * 3 + j] = (float)Rr[j * 3 + i];
    for (int i = 0; i < 3; ++i) pc.offt[i] = (float)x[11 + i];
}

// the same constants for the next pass, the four independent pieces on lane 0 of four warps (block-wide call)
__device__ inline void make_pass_consts_par(const double* x, PassConsts& pc) {
    using namespace mf;
    const int tid = threadIdx.x;
    if (tid == 0) {
        const Q qd = qmul(ldq(x + 3), ldq(x + 7));
        pc.qx = (float)qd.x; pc.qy = (float)qd.y; pc.qz = (float)qd.z; pc.qw = (float)qd.w;
    } else if (tid == 32) {
        double td[3];
        qrot(ldq(x + 3), x + 11, td);
        pc.tx = (float)(td[0] + x[0]); pc.ty = (float)(td[1] + x[1]); pc.tz = (float)(td[2] + x[2]);
        for (int i = 0; i < 3; ++i) pc.offt[i] = (float)x[11 + i];
    } else if (tid == 64) {
        double Ro[9];
        qtoR(ldq(x + 7), Ro);
        for (int i = 0; i < 9; ++i) pc.offR[i] = (float)Ro[i];
    } else if (tid == 96) {
        double Rr[9];
        qtoR(ldq(x + 3), Rr);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) pc.Rt[i * 3 + j] = (float)Rr[j * 3 + i];
    }
}

struct SolveSmem {
    double P[NS * NS];
    double L[NS * NS];
    double B12[NS * 12];          // (P / R)[:, :12]
    double BaaInv[12 * 12];       // B_aa^-1
    double Minv[12 * 12];         // (B_aa^-1 + A)^-1
    double G[12 * 12];            // B_aa^-1 (B_aa^-1 + A)^-1
    double Pinv12[NS * 12];       // P_inv[:, :m]
    double HTH[144];
    double HTh[12];
    double Kx[NS * 12];
    double Kh[NS];
    double vw[12], vy[12], vu[12];  // m-vectors of the gain product
    double x[26], xp[26];
    double dx[NS], dxn[NS], dxu[NS];
    double J3[2][9];  // A(dx)^T for rot / offR
    double J2[4];     // Nx * Mx for grav
    int piv;
    int n_eff;
    int finalize;
    int conv;
};

// Stage stamps (clock64 into ctl->dbg) are a development aid, compiled in with -DB200_STAMPS only: the global stores and
// the ordering they impose were measured to cost several microseconds per pass on the filter block's critical path.
#ifdef B200_STAMPS
#define STAMP(i) do { if (threadIdx.x == 0 && pass < B200_MAX_PASSES) ctl->dbg[pass][i] = clock64(); } while (0)
#else
#define STAMP(i) do { } while (0)
#endif

// In-register inverse of a symmetric positive-definite M x M matrix by ONE warp: Gauss-Jordan on
// [A | I] without pivoting (safe for SPD); lane c < 2M owns column c of the augmented matrix.
// src/dst are row-major with stride ld.
// 1/x to fp64 accuracy without the IEEE division sequence: hardware seed (about 20 bits) + two Newton steps.  Only used
// on pivots of SPD matrices (positive, normal range).
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

template <int M>
__device__ inline void warp_inverse_spd(const double* src, double* dst, int ld) {
    const int lane = threadIdx.x & 31;
    double col[M];
#pragma unroll
    for (int r = 0; r < M; ++r) {
        double v = 0.0;
        if (lane < M) v = src[r * ld + lane];
        else if (lane - M == r) v = 1.0;
        col[r] = v;
    }
#pragma unroll
    for (int k = 0; k < M; ++k) {
        const double pivot = __shfl_sync(0xffffffffu, col[k], k);
        const double prow = col[k] * fast_rcp(pivot);
#pragma unroll
        for (int r = 0; r < M; ++r) {
            if (r != k) {
                const double f = __shfl_sync(0xffffffffu, col[r], k);
                col[r] = __fma_rn(-f, prow, col[r]);
            }
        }
        col[k] = prow;
    }
    if (lane >= M && lane < 2 * M) {
#pragma unroll
        for (int r = 0; r < M; ++r) dst[r * ld + (lane - M)] = col[r];
    }
}
// m = 12 (extrinsic estimation, one shipped yaml) is kept out of line: its fully unrolled elimination is a fifth of k_obs'
// instructions and would otherwise sit twice in the middle of the filter block's straight-line path
static __device__ __noinline__ void warp_inverse_spd12(const double* src, double* dst, int ld) { warp_inverse_spd<12>(src, dst, ld); }
__device__ inline void warp_inverse_spd_m(const double* src, double* dst, int m, int ld) {
    if (m == 6) warp_inverse_spd<6>(src, dst, ld);
    else warp_inverse_spd12(src, dst, ld);
}

// rows idx.. of dst <- J * rows of src (block `which`: 0 rot, 1 offR, 2 grav); one thread per column
__device__ inline void project_rows(double* dst, const double* src, const SolveSmem& s, int which, int ncols, int stride) {
    const int c = threadIdx.x;
    if (c >= ncols) return;
    if (which < 2) {
        const int idx = which == 0 ? 3 : 6;
        const double* J = s.J3[which];
        const double a = src[idx * stride + c], b = src[(idx + 1) * stride + c], d = src[(idx + 2) * stride + c];
        for (int r = 0; r < 3; ++r) dst[(idx + r) * stride + c] = J[r * 3] * a + J[r * 3 + 1] * b + J[r * 3 + 2] * d;
    } else {
        const double a = src[21 * stride + c], b = src[22 * stride + c];
        dst[21 * stride + c] = s.J2[0] * a + s.J2[1] * b;
        dst[22 * stride + c] = s.J2[2] * a + s.J2[3] * b;
    }
}
// columns idx.. of M <- M * J^T ; one thread per row (threads 32.. so it can overlap project_rows users)
__device__ inline void project_cols(double* M, const SolveSmem& s, int which) {
    const int i = threadIdx.x;
    if (i >= NS) return;
    if (which < 2) {
        const int idx = which == 0 ? 3 : 6;
        const double* J = s.J3[which];
        const double a = M[i * NS + idx], b = M[i * NS + idx + 1], d = M[i * NS + idx + 2];
        for (int r = 0; r < 3; ++r) M[i * NS + idx + r] = a * J[r * 3] + b * J[r * 3 + 1] + d * J[r * 3 + 2];
    } else {
        const double a = M[i * NS + 21], b = M[i * NS + 22];
        M[i * NS + 21] = a * s.J2[0] + b * s.J2[1];
        M[i * NS + 22] = a * s.J2[2] + b * s.J2[3];
    }
}

// Row i of the block-diagonal re-linearisation Jacobian J = diag(I3, J_rot, J_offR, I12, J_grav): the block's first
// index, its width and the row of coefficients.
__device__ __forceinline__ void jrow(const SolveSmem& s, int i, int& base, double& c0, double& c1, double& c2) {
    if (i >= 3 && i < 9) {
        const int k = i < 6 ? 0 : 1;
        base = k == 0 ? 3 : 6;
        const double* J = s.J3[k] + (i - base) * 3;
        c0 = J[0]; c1 = J[1]; c2 = J[2];
    } else if (i >= 21) {
        base = 21;
        c0 = s.J2[(i - 21) * 2]; c1 = s.J2[(i - 21) * 2 + 1]; c2 = 0.0;
    } else {
        base = i;
        c0 = 1.0; c1 = 0.0; c2 = 0.0;
    }
}
// (J M J^T)[i][j] for a 23 x 23 row-major M: the three manifold blocks at once (the reference applies them one after
// the other, esekfom.hpp:1561-1601; the blocks are disjoint, so the product is the same up to rounding).  Always a
// 3 x 3 stencil with zero coefficients (and clamped indices) outside the block: no data-dependent loops.
__device__ __forceinline__ double project_elem(const SolveSmem& s, const double* M, int i, int j) {
    int bi, bj;
    double a0, a1, a2, b0, b1, b2;
    jrow(s, i, bi, a0, a1, a2);
    jrow(s, j, bj, b0, b1, b2);
    const int r0 = bi, r1 = min(bi + 1, NS - 1), r2 = min(bi + 2, NS - 1);
    const int q0 = bj, q1 = min(bj + 1, NS - 1), q2 = min(bj + 2, NS - 1);
    const double t0 = fma(M[r0 * NS + q2], b2, fma(M[r0 * NS + q1], b1, M[r0 * NS + q0] * b0));
    const double t1 = fma(M[r1 * NS + q2], b2, fma(M[r1 * NS + q1], b1, M[r1 * NS + q0] * b0));
    const double t2 = fma(M[r2 * NS + q2], b2, fma(M[r2 * NS + q1], b1, M[r2 * NS + q0] * b0));
    return fma(a2, t2, fma(a1, t1, a0 * t0));
}

// Jacobians of the (+)/(-) re-linearisation for the tangent increment d (esekfom.hpp:1561-1601, 1739-1789).
// Three independent manifold blocks -> three warps (lane 0 of warps 0, 1, 2).
__device__ inline void make_projection_par(SolveSmem& s, const double* d, const double* x_cur, const double* x_prop) {
    const int tid = threadIdx.x;
    if (tid == 0 || tid == 32) {
        const int k = tid == 0 ? 0 : 1;
        double A[9];
        mf::A_matrix(d + (k == 0 ? 3 : 6), A);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) s.J3[k][r * 3 + c] = A[c * 3 + r];
    } else if (tid == 64) {
        double Nx[6], Mx[6];
        mf::S2_Nx_yy(x_cur + 23, Nx);
        mf::S2_Mx(x_prop + 23, d + 21, Mx);
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 2; ++c) s.J2[r * 2 + c] = Nx[r * 3] * Mx[c] + Nx[r * 3 + 1] * Mx[2 + c] + Nx[r * 3 + 2] * Mx[4 + c];
    }
}

// ---- filter step, part 1: everything that depends on the state only (esekfom.hpp:1556-1601 and
// the B_aa^-1 factor).  Runs on the filter block WHILE the other blocks measure the scan.
__device__ inline void iekf_presolve(Ctl* ctl, double Rcov, int ext, SolveSmem& s) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid < 26) { s.x[tid] = ctl->x[tid]; s.xp[tid] = ctl->x_prop[tid]; }
    for (int i = tid; i < NS * NS; i += nt) s.P[i] = ctl->P_prop[i];
    __syncthreads();
    // dx = x (-) x_prop (esekfom.hpp:1556): three manifold blocks on three warps
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) s.dx[i] = s.x[i] - s.xp[i];
        mf::so3_log(mf::qmul(mf::qconj(mf::ldq(s.xp + 3)), mf::ldq(s.x + 3)), s.dx + 3);
        for (int i = 0; i < 3; ++i) s.dx[9 + i] = s.x[11 + i] - s.xp[11 + i];
        for (int i = 0; i < 3; ++i) s.dx[12 + i] = s.x[14 + i] - s.xp[14 + i];
    } else if (tid == 32) {
        mf::so3_log(mf::qmul(mf::qconj(mf::ldq(s.xp + 7)), mf::ldq(s.x + 7)), s.dx + 6);
        for (int i = 0; i < 3; ++i) s.dx[15 + i] = s.x[17 + i] - s.xp[17 + i];
        for (int i = 0; i < 3; ++i) s.dx[18 + i] = s.x[20 + i] - s.xp[20 + i];
    } else if (tid == 64) {
        mf::S2_boxminus(s.x + 23, s.xp + 23, s.dx + 21);
    }
    __syncthreads();
    make_projection_par(s, s.dx, s.x, s.xp);
    __syncthreads();
    if (tid < NS) {  // dx_new = J dx
        double v = s.dx[tid];
        if (tid >= 3 && tid < 9) {
            const int k = tid < 6 ? 0 : 1, idx = k == 0 ? 3 : 6, r = tid - idx;
            v = s.J3[k][r * 3] * s.dx[idx] + s.J3[k][r * 3 + 1] * s.dx[idx + 1] + s.J3[k][r * 3 + 2] * s.dx[idx + 2];
        } else if (tid >= 21) {
            const int r = tid - 21;
            v = s.J2[r * 2] * s.dx[21] + s.J2[r * 2 + 1] * s.dx[22];
        }
        s.dxn[tid] = v;
    }
    // P = J P_prop J^T, block by block as the reference does (rows then columns of each block)
    for (int which = 0; which < 3; ++which) {
        project_rows(s.P, s.P, s, which, NS, NS);
        __syncthreads();
        project_cols(s.P, s, which);
        __syncthreads();
    }
    const int m = ext ? 12 : 6;
    for (int i = tid; i < NS * 12; i += nt) s.B12[i] = s.P[(i / 12) * NS + (i % 12)] / Rcov;
    __syncthreads();
    if (warp == 0) warp_inverse_spd_m(s.B12, s.BaaInv, m, 12);  // B_aa^-1
    (void)lane;
    __syncthreads();
}

// ---- filter step, part 2: needs the measurement sums.  Returns after updating the control block.
__device__ inline void iekf_postsolve(Ctl* ctl, const double* partials, int nblocks, int max_iter, const double* __restrict__ limit,
                                      int ext, int single_pass, SolveSmem& s) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const int pass = ctl->passes;
    STAMP(0);
    // 1. deterministic reduction of the per-block partial sums (active columns only): warp w owns columns w and
    //    w + nwarp, ...; lanes stride over blocks in a fixed order, then a fixed shuffle tree.
    const int ncol = ext ? NPART : NCOL6;
    for (int i = tid; i < 144 + 12; i += nt) {
        if (i < 144) s.HTH[i] = 0.0; else s.HTh[i - 144] = 0.0;
    }
    __syncthreads();
    for (int col0 = warp; col0 < ncol; col0 += 2 * nwarp) {
        double v[2][MAXB / 32];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int col = col0 + u * nwarp;
#pragma unroll
            for (int j = 0; j < MAXB / 32; ++j) {
                const int b = lane + 32 * j;
                v[u][j] = (col < ncol && b < nblocks) ? __ldcg(partials + (size_t)col * nblocks + b) : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int col = col0 + u * nwarp;
            double sum = 0.0;
#pragma unroll
            for (int j = 0; j < MAXB / 32; ++j) sum += v[u][j];
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0 && col < ncol) {
                const int a = c_col_a[ext ? 1 : 0][col], b = c_col_b[ext ? 1 : 0][col];
                if (a == 255) {
                    s.n_eff = (int)(sum + 0.5);
                } else if (b == 12) {
                    s.HTh[a] = sum;
                } else {
                    s.HTH[a * 12 + b] = sum;
                    s.HTH[b * 12 + a] = sum;
                }
            }
        }
    }
    __syncthreads();
    STAMP(1);
    const int conv_in = ctl->converge;
    const int iter = ctl->iter;
    if (pass < B200_MAX_PASSES) {
        for (int i = tid; i < 144; i += nt) ctl->HtH[pass][i] = s.HTH[i];
        if (tid < 12) ctl->Hth[pass][tid] = s.HTh[tid];
        if (tid < 26) ctl->x_in[pass][tid] = s.x[tid];
        if (tid == 0) { ctl->n_eff[pass] = s.n_eff; ctl->knn[pass] = conv_in; }
    }
    if (tid == 0) {
        ctl->passes = pass + 1;
        ctl->knn_passes += conv_in ? 1 : 0;
    }
    if (single_pass) {  // parity primitive: one ObsModel evaluation, no filter step
        if (tid == 0) ctl->done = 1;
        return;
    }
    if (s.n_eff < 1) {  // ekfom_data.valid == false -> `continue` (esekfom.hpp:1543-1545)
        if (tid == 0) {
            ctl->iter = iter + 1;
            if (iter + 1 >= max_iter) ctl->done = 1;
        }
        return;
    }
    if (tid == 0) ctl->any_valid = 1;
    // 4.-6. dx_ = K_h + (K_x - I) dx_new (:1708-1719) without forming the gain matrices: with K = P_inv[:, :m] = B[:, :m] G,
    //    G = B_aa^-1 (B_aa^-1 + A)^-1 (header comment),  K_h + K_x dx_new = K (H^T h + A dx_new[:m]), so the step needs three
    //    m-vectors and one 23 x m product.  One warp does it with warp barriers only (the other warps wait at the block
    //    barrier below); K_x itself is only needed for the final covariance and is built in the finalising pass.
    const int m = ext ? 12 : 6;
    if (warp == 0) {
        for (int i = lane; i < m * m; i += 32) {
            const int r = i / m, c = i % m;
            s.Minv[r * 12 + c] = s.BaaInv[r * 12 + c] + s.HTH[r * 12 + c];
        }
        __syncwarp();
        warp_inverse_spd_m(s.Minv, s.Minv, m, 12);  // (B_aa^-1 + A)^-1 = P_inv[a, a]
        if (lane < m) {  // w = H^T h + A dx_new[:m]
            double v = s.HTh[lane];
            for (int k = 0; k < m; ++k) v = fma(s.HTH[lane * 12 + k], s.dxn[k], v);
            s.vw[lane] = v;
        }
        __syncwarp();
        if (lane < m) {  // y = (B_aa^-1 + A)^-1 w
            double v = 0.0;
            for (int k = 0; k < m; ++k) v = fma(s.Minv[lane * 12 + k], s.vw[k], v);
            s.vy[lane] = v;
        }
        __syncwarp();
        if (lane < m) {  // u = B_aa^-1 y
            double v = 0.0;
            for (int k = 0; k < m; ++k) v = fma(s.BaaInv[lane * 12 + k], s.vy[k], v);
            s.vu[lane] = v;
        }
        __syncwarp();
        if (lane < NS) {  // dx_ = B[:, :m] u - dx_new
            double v = -s.dxn[lane];
            for (int k = 0; k < m; ++k) v = fma(s.B12[lane * 12 + k], s.vu[k], v);
            s.dxu[lane] = v;
        }
    }
    __syncthreads();
    STAMP(4);
    // 7. x (+)= dx_ ; convergence bookkeeping (:1720-1735); again one warp per manifold block
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) s.x[i] += s.dxu[i];
        mf::stq(s.x + 3, mf::qmul(mf::ldq(s.x + 3), mf::so3_exp(s.dxu + 3, 0.5)));
        for (int i = 0; i < 3; ++i) s.x[11 + i] += s.dxu[9 + i];
        for (int i = 0; i < 3; ++i) s.x[14 + i] += s.dxu[12 + i];
    } else if (tid == 32) {
        mf::stq(s.x + 7, mf::qmul(mf::ldq(s.x + 7), mf::so3_exp(s.dxu + 6, 0.5)));
        for (int i = 0; i < 3; ++i) s.x[17 + i] += s.dxu[15 + i];
        for (int i = 0; i < 3; ++i) s.x[20 + i] += s.dxu[18 + i];
    } else if (tid == 64) {
        mf::S2_boxplus(s.x + 23, s.dxu + 21);
    } else if (tid == 96) {
        int conv = 1;
        for (int i = 0; i < NS; ++i)
            if (fabs(s.dxu[i]) > limit[i]) { conv = 0; break; }
        int t = ctl->t;
        if (conv) t++;
        if (!t && iter == max_iter - 2) conv = 1;
        ctl->t = t;
        s.conv = conv;
        s.finalize = (t > 1 || iter == max_iter - 1) ? 1 : 0;
    }
    __syncthreads();
    STAMP(5);
    if (tid < 26) ctl->x[tid] = s.x[tid];
    if (s.finalize) {  // :1735-1831
        make_projection_par(s, s.dxu, s.x, s.xp);
        // K_x[:, :12] = P_inv[:, :12] HTH = B[:, :m] G A   (esekfom.hpp:1713), only needed here
        for (int idx = tid; idx < m * m; idx += nt) {  // G = B_aa^-1 * P_inv[a,a]
            const int r = idx / m, c = idx % m;
            double v = 0.0;
            for (int k = 0; k < m; ++k) v = fma(s.BaaInv[r * 12 + k], s.Minv[k * 12 + c], v);
            s.G[r * 12 + c] = v;
        }
        __syncthreads();
        for (int idx = tid; idx < NS * 12; idx += nt) {  // P_inv[:, :m] = B[:, :m] G ; columns >= m are never used
            const int r = idx / 12, c = idx % 12;
            double v = 0.0;
            if (c < m)
                for (int k = 0; k < m; ++k) v = fma(s.B12[r * 12 + k], s.G[k * 12 + c], v);
            s.Pinv12[idx] = v;
        }
        __syncthreads();
        for (int idx = tid; idx < NS * 12; idx += nt) {
            const int r = idx / 12, c = idx % 12;
            double sum = 0.0;
            for (int k = 0; k < m; ++k) sum = fma(s.Pinv12[r * 12 + k], s.HTH[k * 12 + c], sum);
            s.Kx[idx] = sum;
        }
        __syncthreads();
        STAMP(6);
        // L = J P J^T and K_x <- J K_x in one step (esekfom.hpp:1739-1789 applies the blocks one after the other)
        {
            double v[2] = {0.0, 0.0}, kx = 0.0;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int e = tid + u * nt;
                if (e < NS * NS) v[u] = project_elem(s, s.P, e / NS, e % NS);
            }
            if (tid < NS * 12) {
                int bi;
                double a0, a1, a2;
                jrow(s, tid / 12, bi, a0, a1, a2);
                const int c = tid % 12;
                kx = fma(a2, s.Kx[min(bi + 2, NS - 1) * 12 + c], fma(a1, s.Kx[min(bi + 1, NS - 1) * 12 + c], a0 * s.Kx[bi * 12 + c]));
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int e = tid + u * nt;
                if (e < NS * NS) s.L[e] = v[u];
            }
            if (tid < NS * 12) s.Kx[tid] = kx;
            __syncthreads();
        }
        // P_final = L - K_x (P J^T)[:12, :]
        for (int idx = tid; idx < NS * NS; idx += nt) {
            const int r = idx / NS, c = idx % NS;
            int bc;
            double b0, b1, b2;
            jrow(s, c, bc, b0, b1, b2);
            const int q1 = min(bc + 1, NS - 1), q2 = min(bc + 2, NS - 1);
            double sum0 = 0.0, sum1 = 0.0;  // two independent chains
#pragma unroll
            for (int k = 0; k < 12; k += 2) {
                const double pj0 = fma(s.P[k * NS + q2], b2, fma(s.P[k * NS + q1], b1, s.P[k * NS + bc] * b0));
                const double pj1 = fma(s.P[(k + 1) * NS + q2], b2, fma(s.P[(k + 1) * NS + q1], b1, s.P[(k + 1) * NS + bc] * b0));
                sum0 = fma(s.Kx[r * 12 + k], pj0, sum0);
                sum1 = fma(s.Kx[r * 12 + k + 1], pj1, sum1);
            }
            ctl->P[idx] = s.L[idx] - (sum0 + sum1);
        }
        if (tid == 0) {
            ctl->converge = s.conv;
            ctl->done = 1;
            ctl->converged = ctl->t > 1 ? 1 : 0;
        }
        STAMP(7);
    } else {
        // the covariance the filter holds when the loop ends without finalising is the projected P_
        for (int i = tid; i < NS * NS; i += nt) ctl->P[i] = s.P[i];
        if (tid == 0) {
            ctl->converge = s.conv;
            ctl->iter = iter + 1;
            if (iter + 1 >= max_iter) ctl->done = 1;
        }
        make_pass_consts_par(s.x, ctl->pc);
        STAMP(7);
    }
}

}  // namespace b200
