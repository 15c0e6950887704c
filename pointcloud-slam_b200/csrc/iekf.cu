// b200reg — iterated-EKF point-to-plane measurement update, fully device resident.
//
// Replaces esekf::update_iterated_dyn_share_modified (IKFoM_toolkit/esekfom/esekfom.hpp:1526-1834)
// driven with LaserMapping::ObsModel (jueying_lio/src/laser_mapping.cc:592-701), plus
// LaserMapping::MapIncremental (laser_mapping.cc:525-583).
//
// One update = one H2D copy (state + scan), then per pass two kernels enqueued back to back with
// no host round trip:
//   k_obs    transform -> (k=5 stencil search -> plane fit) -> residual/validity -> Jacobian row ->
//            per-block fp64 partial sums of h_x^T h_x (78 unique) and h_x^T h (12)
//   k_solve  one CTA: deterministic reduction of the partials, (-) / P projection, two 23x23
//            inversions, gain, (+), convergence logic, next-pass constants
// Whether a pass searches the map (dyn_share.converge) and whether it runs at all (early exit)
// is decided on the device through the control block, so the host just enqueues max_iter+1 pairs.
#include "map.cuh"
#include "manifold.cuh"
#include "pointmath.cuh"

#include <cub/cub.cuh>
#include <vector>

namespace b200 {

constexpr int NS = 23;          // state DOF
constexpr int NPART = 91;       // 78 + 12 + count
constexpr int OBS_THREADS = 256;
constexpr int KNN_G = 8;        // lanes per query in a search pass
constexpr int KNN_TILE = OBS_THREADS / KNN_G;

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
    // stats
    int passes, knn_passes, any_valid, converged;
    int n_eff[B200_MAX_PASSES], knn[B200_MAX_PASSES];
    PassConsts pc;
    double x_in[B200_MAX_PASSES][26];
    double HtH[B200_MAX_PASSES][144];
    double Hth[B200_MAX_PASSES][12];
};

struct PointState {  // per-point arrays that persist across passes and scans (laser_mapping.cc:335-339)
    float4* plane;   // plane_coef_
    float* resid;    // residuals_
    uint8_t* sel;    // point_selected_surf_
    uint8_t* nn_cnt; // nearest_points_[i].size()
    float4* nn;      // nearest_points_[i][0..4] (xyz + ordinal)
};

__constant__ unsigned char c_pair_a[78];
__constant__ unsigned char c_pair_b[78];

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

// ------------------------------------------------------------------ init
// hdr: the {x, P, n, prev_n} header as it arrived from the host (either inside the staged scan block or
// already in the control block itself)
__global__ void k_iekf_init(Ctl* ctl, const Ctl* hdr, PointState ps, int force_converge) {
    const int tid = threadIdx.x + blockIdx.x * blockDim.x;
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < NS * NS; i += blockDim.x) {
            const double v = hdr->P[i];
            ctl->P[i] = v;
            ctl->P_prop[i] = v;
        }
        if (threadIdx.x < 26) {
            const double v = hdr->x[threadIdx.x];
            ctl->x[threadIdx.x] = v;
            ctl->x_prop[threadIdx.x] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            ctl->n = hdr->n;
            ctl->prev_n = hdr->prev_n;
            ctl->iter = -1;
            ctl->converge = force_converge;
            ctl->done = 0;
            ctl->t = 0;
            ctl->passes = ctl->knn_passes = ctl->any_valid = ctl->converged = 0;
            for (int i = 0; i < B200_MAX_PASSES; ++i) { ctl->n_eff[i] = 0; ctl->knn[i] = 0; }
            make_pass_consts(ctl->x, ctl->pc);
        }
    }
    // vector::resize(cur_pts, default) semantics: slots at or beyond the previous scan's size are fresh
    const int n = hdr->n, prev = hdr->prev_n;
    for (int i = prev + tid; i < n; i += gridDim.x * blockDim.x) {
        ps.plane[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        ps.resid[i] = 0.0f;
        ps.sel[i] = 1;
        ps.nn_cnt[i] = 0;
    }
}

// ------------------------------------------------------------------ ObsModel
// Residual / validity / Jacobian row of point i given its (possibly refreshed) neighbour set.
// Returns whether the point is an effective feature; fills row[0..11] and h.
__device__ __forceinline__ bool point_measure(const PassConsts& pc, const float4 pb, const float3 pw, bool searched, int m_new,
                                              const float4* nb_new, PointState& ps, int i, float thr, bool ext,
                                              float (&row)[12], float& h) {
    bool sel = ps.sel[i] != 0;
    float4 plane = ps.plane[i];
    if (searched) {  // laser_mapping.cc:616-624
        float4 nb[5];
        int m = m_new;
        if (m_new > 0) {
            for (int k = 0; k < 5; ++k) nb[k] = nb_new[k];
            ps.nn_cnt[i] = (uint8_t)m_new;
            for (int k = 0; k < m_new; ++k) ps.nn[(size_t)i * 5 + k] = nb[k];
        } else {
            // GetClosestPoint returned before clearing its output (ivox3d.h:151-153): the slot keeps
            // the neighbours it had before
            m = ps.nn_cnt[i];
            for (int k = 0; k < m; ++k) nb[k] = ps.nn[(size_t)i * 5 + k];
        }
        sel = m >= 3;
        if (sel) {
            sel = esti_plane(nb, m, thr, plane);
            ps.plane[i] = plane;
        }
    }
    float resid = ps.resid[i];
    if (sel) {  // laser_mapping.cc:626-636
        const float pd2 = dot4_sse(plane.x, plane.y, plane.z, plane.w, pw.x, pw.y, pw.z, 1.0f);
        const float bn = fsqrt(fadd(fadd(fmul(pb.x, pb.x), fmul(pb.y, pb.y)), fmul(pb.z, pb.z)));
        if (bn > fmul(fmul(81.0f, pd2), pd2)) {
            resid = pd2;
            ps.resid[i] = pd2;
        }
    }
    ps.sel[i] = sel ? 1 : 0;
    if (sel) {
        jacobian_row(pc, pb.x, pb.y, pb.z, plane, ext, row);
        h = -resid;
    }
    return sel;
}

struct ObsSmem {
    float rows[OBS_THREADS][13];
    unsigned char eff[OBS_THREADS];
    float4 nb[KNN_TILE][5];
    int nbc[KNN_TILE];
};

__global__ void __launch_bounds__(OBS_THREADS) k_obs(MapView map, const float4* __restrict__ scan, PointState ps, const Ctl* __restrict__ ctl,
                                                     float thr, int ext, double* __restrict__ partials) {
    if (ctl->done) return;
    __shared__ ObsSmem sm;
    const int tid = threadIdx.x;
    const int n = ctl->n;
    const bool searched = ctl->converge != 0;
    __shared__ PassConsts pc;
    if (tid < (int)(sizeof(PassConsts) / 4)) ((float*)&pc)[tid] = ((const float*)&ctl->pc)[tid];
    __syncthreads();

    // accumulator of this thread's (a,b) pair / h column across all tiles of the block
    double acc = 0.0;
    int pa = 0, pb_ = 0;
    const int npairs = 78;
    if (tid < npairs) { pa = c_pair_a[tid]; pb_ = c_pair_b[tid]; }
    else if (tid < 90) { pa = tid - 78; pb_ = 12; }
    const bool pair_active = tid < 90 && (ext || (pa < 6 && (pb_ < 6 || pb_ == 12)));
    int cnt_acc = 0;

    const int tile = searched ? KNN_TILE : OBS_THREADS;
    for (int base = blockIdx.x * tile; base < n; base += gridDim.x * tile) {
        int q_here;  // queries in this tile
        if (searched) {
            // phase 1: 8 lanes per query search the stencil
            const int ql = tid / KNN_G, lg = tid % KNN_G;
            const int i = base + ql;
            if (i < n) {
                const unsigned gmask = ((1u << KNN_G) - 1u) << ((tid & 31) / KNN_G * KNN_G);
                const float4 pbody = __ldg(scan + i);
                const float3 pw = body_to_world(pc, pbody.x, pbody.y, pbody.z);
                uint64_t win[5];
                float4 mine;
                const int c = knn5_group<KNN_G>(map, pw.x, pw.y, pw.z, lg, gmask, win, mine);
                if (lg < 5) sm.nb[ql][lg] = mine;
                if (lg == 0) sm.nbc[ql] = c;
            }
            __syncthreads();
            q_here = min(KNN_TILE, n - base);
        } else {
            q_here = min(OBS_THREADS, n - base);
        }
        // phase 2: one thread per query
        if (tid < q_here) {
            const int i = base + tid;
            const float4 pbody = __ldg(scan + i);
            const float3 pw = body_to_world(pc, pbody.x, pbody.y, pbody.z);
            float row[12], h = 0.f;
            const bool eff = point_measure(pc, pbody, pw, searched, searched ? sm.nbc[tid] : 0, sm.nb[searched ? tid : 0], ps, i, thr,
                                           ext != 0, row, h);
            sm.eff[tid] = eff ? 1 : 0;
            if (eff) {
#pragma unroll
                for (int k = 0; k < 12; ++k) sm.rows[tid][k] = row[k];
                sm.rows[tid][12] = h;
            }
        }
        __syncthreads();
        // phase 3: fp64 accumulation in query order (deterministic)
        if (pair_active) {
            for (int q = 0; q < q_here; ++q)
                if (sm.eff[q]) acc += (double)sm.rows[q][pa] * (double)sm.rows[q][pb_];
        } else if (tid == 90) {
            for (int q = 0; q < q_here; ++q) cnt_acc += sm.eff[q];
        }
        __syncthreads();
    }
    // column-major by block so the solve kernel reads each column coalesced
    if (tid < 90) partials[(size_t)tid * gridDim.x + blockIdx.x] = acc;
    if (tid == 90) partials[(size_t)90 * gridDim.x + blockIdx.x] = (double)cnt_acc;
}

// ------------------------------------------------------------------ solve
struct SolveSmem {
    double P[NS * NS];
    double L[NS * NS];
    double aug[NS * 2 * NS];
    double HTH[144];
    double HTh[12];
    double Kx[NS * 12];
    double Kh[NS];
    double dx[NS], dxn[NS], dxu[NS];
    double prow[2 * NS], scol[NS];
    double J3[2][9];  // A(dx)^T for rot / offR
    double J2[4];     // Nx * Mx for grav
    int piv;
    int n_eff;
    int finalize;
};

// in-place inverse of the NS x NS matrix M (row-major, shared) by Gauss-Jordan with partial pivoting
__device__ void block_inverse(double* M, SolveSmem& s) {
    const int tid = threadIdx.x, nt = blockDim.x;
    constexpr int W = 2 * NS;
    for (int idx = tid; idx < NS * W; idx += nt) {
        int r = idx / W, c = idx % W;
        s.aug[idx] = c < NS ? M[r * NS + c] : (c - NS == r ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int k = 0; k < NS; ++k) {
        if (tid < 32) {
            double v = -1.0;
            int r = k + tid;
            if (r < NS) v = fabs(s.aug[r * W + k]);
            for (int o = 16; o > 0; o >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, v, o);
                int orr = __shfl_xor_sync(0xffffffffu, r, o);
                if (ov > v || (ov == v && orr < r)) { v = ov; r = orr; }
            }
            if (tid == 0) s.piv = r;
        }
        __syncthreads();
        const int p = s.piv;
        if (p != k && tid < W) {
            double a = s.aug[k * W + tid];
            s.aug[k * W + tid] = s.aug[p * W + tid];
            s.aug[p * W + tid] = a;
        }
        __syncthreads();
        const double pivot = s.aug[k * W + k];
        if (tid < W) s.prow[tid] = s.aug[k * W + tid] / pivot;
        else if (tid >= 64 && tid < 64 + NS) s.scol[tid - 64] = s.aug[(tid - 64) * W + k];
        __syncthreads();
        for (int idx = tid; idx < NS * W; idx += nt) {
            int r = idx / W, c = idx % W;
            s.aug[idx] = (r == k) ? s.prow[c] : s.aug[idx] - s.scol[r] * s.prow[c];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < NS * NS; idx += nt) {
        int r = idx / NS, c = idx % NS;
        M[idx] = s.aug[r * W + NS + c];
    }
    __syncthreads();
}

// M <- J M (rows idx..idx+d-1) for the block-diagonal J made of J3[0] (3), J3[1] (6), J2 (21)
__device__ void project_rows(double* dst, const double* src, const SolveSmem& s, int which /*0,1 SO3; 2 S2*/, int ncols_stride) {
    const int tid = threadIdx.x;
    if (tid >= NS) return;
    const int c = tid;
    if (which < 2) {
        const int idx = which == 0 ? 3 : 6;
        const double* J = s.J3[which];
        double a = src[idx * ncols_stride + c], b = src[(idx + 1) * ncols_stride + c], d = src[(idx + 2) * ncols_stride + c];
        for (int r = 0; r < 3; ++r) dst[(idx + r) * ncols_stride + c] = J[r * 3] * a + J[r * 3 + 1] * b + J[r * 3 + 2] * d;
    } else {
        const int idx = 21;
        double a = src[idx * ncols_stride + c], b = src[(idx + 1) * ncols_stride + c];
        dst[idx * ncols_stride + c] = s.J2[0] * a + s.J2[1] * b;
        dst[(idx + 1) * ncols_stride + c] = s.J2[2] * a + s.J2[3] * b;
    }
}
// M <- M J^T (columns idx..)
__device__ void project_cols(double* M, const SolveSmem& s, int which) {
    const int tid = threadIdx.x;
    if (tid >= NS) return;
    const int i = tid;
    if (which < 2) {
        const int idx = which == 0 ? 3 : 6;
        const double* J = s.J3[which];
        double a = M[i * NS + idx], b = M[i * NS + idx + 1], d = M[i * NS + idx + 2];
        for (int r = 0; r < 3; ++r) M[i * NS + idx + r] = a * J[r * 3] + b * J[r * 3 + 1] + d * J[r * 3 + 2];
    } else {
        const int idx = 21;
        double a = M[i * NS + idx], b = M[i * NS + idx + 1];
        M[i * NS + idx] = a * s.J2[0] + b * s.J2[1];
        M[i * NS + idx + 1] = a * s.J2[2] + b * s.J2[3];
    }
}

// thread 0: Jacobians of the (+)/(-) re-linearisation for a tangent increment d (esekfom.hpp:1561-1601 / 1739-1789)
__device__ void make_projection(SolveSmem& s, const double* d, const double* x_cur, const double* x_prop) {
    for (int k = 0; k < 2; ++k) {
        double A[9];
        mf::A_matrix(d + (k == 0 ? 3 : 6), A);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) s.J3[k][r * 3 + c] = A[c * 3 + r];
    }
    double Nx[6], Mx[6];
    mf::S2_Nx_yy(x_cur + 23, Nx);
    mf::S2_Mx(x_prop + 23, d + 21, Mx);
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c) s.J2[r * 2 + c] = Nx[r * 3] * Mx[c] + Nx[r * 3 + 1] * Mx[2 + c] + Nx[r * 3 + 2] * Mx[4 + c];
}

__global__ void __launch_bounds__(256) k_solve(Ctl* ctl, const double* __restrict__ partials, int nblocks, int max_iter, double Rcov,
                                                const double* __restrict__ limit, int single_pass) {
    if (ctl->done) return;
    __shared__ SolveSmem s;
    const int tid = threadIdx.x, nt = blockDim.x;
    // 1. deterministic reduction of the per-block partial sums
    //    (warp w owns columns w, w+8, ...; lane-strided sums in block order, then a fixed shuffle tree)
    for (int col = tid / 32; col < NPART; col += nt / 32) {
        double sum = 0.0;
        for (int b = tid % 32; b < nblocks; b += 32) sum += partials[(size_t)col * nblocks + b];
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (tid % 32 == 0) {
            if (col < 78) {
                int a = c_pair_a[col], b = c_pair_b[col];
                s.HTH[a * 12 + b] = sum;
                s.HTH[b * 12 + a] = sum;
            } else if (col < 90) {
                s.HTh[col - 78] = sum;
            } else {
                s.n_eff = (int)(sum + 0.5);
            }
        }
    }
    __syncthreads();
    const int pass = ctl->passes;
    const int conv_in = ctl->converge;
    const int iter = ctl->iter;
    if (pass < B200_MAX_PASSES) {
        for (int i = tid; i < 144; i += nt) ctl->HtH[pass][i] = s.HTH[i];
        if (tid < 12) ctl->Hth[pass][tid] = s.HTh[tid];
        if (tid < 26) ctl->x_in[pass][tid] = ctl->x[tid];
        if (tid == 0) { ctl->n_eff[pass] = s.n_eff; ctl->knn[pass] = conv_in; }
    }
    __syncthreads();
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
    // 2. dx = x (-) x_prop and the projection Jacobians
    if (tid == 0) {
        ctl->any_valid = 1;
        mf::state_boxminus(ctl->x, ctl->x_prop, s.dx);
        make_projection(s, s.dx, ctl->x, ctl->x_prop);
        for (int i = 0; i < NS; ++i) s.dxn[i] = s.dx[i];
        for (int k = 0; k < 2; ++k) {
            const int idx = k == 0 ? 3 : 6;
            double a = s.dxn[idx], b = s.dxn[idx + 1], c = s.dxn[idx + 2];
            for (int r = 0; r < 3; ++r) s.dxn[idx + r] = s.J3[k][r * 3] * a + s.J3[k][r * 3 + 1] * b + s.J3[k][r * 3 + 2] * c;
        }
        double a = s.dxn[21], b = s.dxn[22];
        s.dxn[21] = s.J2[0] * a + s.J2[1] * b;
        s.dxn[22] = s.J2[2] * a + s.J2[3] * b;
    }
    for (int i = tid; i < NS * NS; i += nt) s.P[i] = ctl->P_prop[i];
    __syncthreads();
    // 3. P = J P_prop J^T, block by block as the reference does (rows then columns of each block)
    for (int which = 0; which < 3; ++which) {
        project_rows(s.P, s.P, s, which, NS);
        __syncthreads();
        project_cols(s.P, s, which);
        __syncthreads();
    }
    // 4. P_inv = ((P / R)^-1 + [HTH 0; 0 0])^-1   (esekfom.hpp:1685-1706)
    for (int i = tid; i < NS * NS; i += nt) s.L[i] = s.P[i] / Rcov;
    __syncthreads();
    block_inverse(s.L, s);
    for (int i = tid; i < 144; i += nt) s.L[(i / 12) * NS + (i % 12)] += s.HTH[i];
    __syncthreads();
    block_inverse(s.L, s);
    // 5. K_h = P_inv[:, :12] H^T h ; K_x[:, :12] = P_inv[:, :12] HTH   (:1708-1713)
    for (int idx = tid; idx < NS * 13; idx += nt) {
        const int r = idx / 13, c = idx % 13;
        double sum = 0.0;
        if (c < 12) {
            for (int k = 0; k < 12; ++k) sum += s.L[r * NS + k] * s.HTH[k * 12 + c];
            s.Kx[r * 12 + c] = sum;
        } else {
            for (int k = 0; k < 12; ++k) sum += s.L[r * NS + k] * s.HTh[k];
            s.Kh[r] = sum;
        }
    }
    __syncthreads();
    // 6. dx_ = K_h + (K_x - I) dx_new   (:1719)
    if (tid < NS) {
        double sum = 0.0;
        for (int c = 0; c < NS; ++c) {
            double kx = c < 12 ? s.Kx[tid * 12 + c] : 0.0;
            sum += (kx - (c == tid ? 1.0 : 0.0)) * s.dxn[c];
        }
        s.dxu[tid] = s.Kh[tid] + sum;
    }
    __syncthreads();
    // 7. x (+)= dx_ ; convergence bookkeeping (:1720-1735)
    if (tid == 0) {
        mf::state_boxplus(ctl->x, s.dxu);
        int conv = 1;
        for (int i = 0; i < NS; ++i)
            if (fabs(s.dxu[i]) > limit[i]) { conv = 0; break; }
        int t = ctl->t;
        if (conv) t++;
        if (!t && iter == max_iter - 2) conv = 1;
        ctl->t = t;
        ctl->converge = conv;
        s.finalize = (t > 1 || iter == max_iter - 1) ? 1 : 0;
        if (s.finalize) make_projection(s, s.dxu, ctl->x, ctl->x_prop);
    }
    __syncthreads();
    if (s.finalize) {  // :1735-1831
        for (int i = tid; i < NS * NS; i += nt) s.L[i] = s.P[i];
        __syncthreads();
        for (int which = 0; which < 3; ++which) {
            project_rows(s.L, s.P, s, which, NS);  // L rows <- J * P rows
            if (tid >= 32 && tid < 32 + 12) {      // K_x rows <- J * K_x rows
                const int c = tid - 32;
                if (which < 2) {
                    const int idx = which == 0 ? 3 : 6;
                    const double* J = s.J3[which];
                    double a = s.Kx[idx * 12 + c], b = s.Kx[(idx + 1) * 12 + c], d = s.Kx[(idx + 2) * 12 + c];
                    for (int r = 0; r < 3; ++r) s.Kx[(idx + r) * 12 + c] = J[r * 3] * a + J[r * 3 + 1] * b + J[r * 3 + 2] * d;
                } else {
                    double a = s.Kx[21 * 12 + c], b = s.Kx[22 * 12 + c];
                    s.Kx[21 * 12 + c] = s.J2[0] * a + s.J2[1] * b;
                    s.Kx[22 * 12 + c] = s.J2[2] * a + s.J2[3] * b;
                }
            }
            __syncthreads();
            project_cols(s.L, s, which);
            __syncthreads();
            project_cols(s.P, s, which);
            __syncthreads();
        }
        for (int idx = tid; idx < NS * NS; idx += nt) {
            const int r = idx / NS, c = idx % NS;
            double sum = 0.0;
            for (int k = 0; k < 12; ++k) sum += s.Kx[r * 12 + k] * s.P[k * NS + c];
            ctl->P[idx] = s.L[idx] - sum;
        }
        if (tid == 0) {
            ctl->done = 1;
            ctl->converged = ctl->t > 1 ? 1 : 0;
        }
    } else {
        // the covariance the filter holds when the loop ends without finalising is the projected P_
        for (int i = tid; i < NS * NS; i += nt) ctl->P[i] = s.P[i];
        if (tid == 0) {
            ctl->iter = iter + 1;
            make_pass_consts(ctl->x, ctl->pc);
            if (iter + 1 >= max_iter) ctl->done = 1;
        }
    }
}

// ------------------------------------------------------------------ MapIncremental (laser_mapping.cc:525-583)
// flag: 0 = drop, 1 = points_to_add, 2 = point_no_need_downsample
__global__ void k_map_incremental_flags(const float4* __restrict__ scan, int n, const double* __restrict__ x, PointState ps, int ekf_inited,
                                        double fs, float4* __restrict__ world, uint8_t* __restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    using namespace mf;
    const float4 pb = scan[i];
    // PointBodyToWorld (:855-864): fp64 quaternion arithmetic, narrowed on store
    double pbd[3] = {pb.x, pb.y, pb.z}, t1[3], t2[3];
    qrot(ldq(x + 7), pbd, t1);
    for (int k = 0; k < 3; ++k) t1[k] = t1[k] + x[11 + k];
    qrot(ldq(x + 3), t1, t2);
    const float w[3] = {(float)(t2[0] + x[0]), (float)(t2[1] + x[1]), (float)(t2[2] + x[2])};
    world[i] = make_float4(w[0], w[1], w[2], 0.f);
    const int m = ps.nn_cnt[i];
    uint8_t f = 1;
    if (m > 0 && ekf_inited) {
        const float fsf = (float)fs;
        float center[3];
        for (int k = 0; k < 3; ++k) center[k] = fmul(fadd(floorf(fdiv(w[k], fsf)), 0.5f), fsf);
        const float4 n0 = ps.nn[(size_t)i * 5];
        const float d0 = fsub(n0.x, center[0]), d1 = fsub(n0.y, center[1]), d2 = fsub(n0.z, center[2]);
        if (fabs((double)d0) > 0.5 * fs && fabs((double)d1) > 0.5 * fs && fabs((double)d2) > 0.5 * fs) {
            f = 2;
        } else {
            const float ex = fsub(w[0], center[0]), ey = fsub(w[1], center[1]), ez = fsub(w[2], center[2]);
            const float dist = fadd(fadd(fmul(ex, ex), fmul(ey, ey)), fmul(ez, ez));
            if (m >= 5) {
                for (int k = 0; k < 5; ++k) {
                    const float4 q = ps.nn[(size_t)i * 5 + k];
                    const float ax = fsub(q.x, center[0]), ay = fsub(q.y, center[1]), az = fsub(q.z, center[2]);
                    const float dk = fadd(fadd(fmul(ax, ax), fmul(ay, ay)), fmul(az, az));
                    if ((double)dk < (double)dist + 1e-6) { f = 0; break; }
                }
            }
        }
    }
    flag[i] = f;
}
__global__ void k_flag_eq(const uint8_t* __restrict__ flag, int n, uint8_t v, uint8_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = flag[i] == v ? 1 : 0;
}

// ------------------------------------------------------------------ host object
struct Iekf {
    b200_iekf_params prm;
    Map* map = nullptr;
    cudaStream_t stream = nullptr;
    Ctl* d_ctl = nullptr;
    double* d_limit = nullptr;
    double* d_partials = nullptr;
    int nblocks = 0;
    PointState ps{};
    size_t ps_cap = 0;
    DevBuf<float4> d_scan;
    int last_n = 0;       // size the per-point arrays were last resized to
    const float4* last_scan = nullptr;
    PinnedBuf<uint8_t> h_stage;  // [Ctl header | float4 points]
    PinnedBuf<Ctl> h_out;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // map-incremental scratch
    DevBuf<float4> d_world, d_sel_pts;
    DevBuf<uint8_t> d_flag, d_flag2, cub_tmp;
    DevBuf<int32_t> d_count;
    DevBuf<double> d_x;
    PinnedBuf<int32_t> h_count;
    PinnedBuf<double> h_x;

    int32_t init(const b200_iekf_params* p, Map* m);
    void destroy();
    int32_t ensure_points(size_t n);
    int32_t run(const float4* d_pts, int n, const Ctl* d_hdr, double* x, double* P, b200_iekf_stats* st, int single_pass,
                int force_converge);
};

static void init_pair_tables() {
    unsigned char a[78], b[78];
    int k = 0;
    for (int i = 0; i < 12; ++i)
        for (int j = i; j < 12; ++j) { a[k] = (unsigned char)i; b[k] = (unsigned char)j; ++k; }
    cudaMemcpyToSymbol(c_pair_a, a, sizeof a);
    cudaMemcpyToSymbol(c_pair_b, b, sizeof b);
}

int32_t Iekf::init(const b200_iekf_params* p, Map* m) {
    prm = *p;
    map = m;
    if (prm.max_iter < 1 || prm.max_iter + 1 > B200_MAX_PASSES) B200_FAIL(B200_ERR_ARG, "max_iter must be in [1, 7]");
    CUDA_TRY(cudaSetDevice(m->device));
    stream = m->stream;  // one stream per map/filter pair: inserts and updates are naturally ordered
    init_pair_tables();
    CUDA_TRY(cudaMalloc(&d_ctl, sizeof(Ctl)));
    CUDA_TRY(cudaMemset(d_ctl, 0, sizeof(Ctl)));
    CUDA_TRY(cudaMalloc(&d_limit, sizeof(double) * NS));
    CUDA_TRY(cudaMemcpy(d_limit, prm.limit, sizeof(double) * NS, cudaMemcpyHostToDevice));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, m->device));
    nblocks = prop.multiProcessorCount * 4;
    CUDA_TRY(cudaMalloc(&d_partials, sizeof(double) * NPART * nblocks));
    CUDA_TRY(h_out.reserve(1));
    CUDA_TRY(cudaEventCreate(&ev0));
    CUDA_TRY(cudaEventCreate(&ev1));
    CUDA_TRY(h_count.reserve(4));
    CUDA_TRY(h_x.reserve(32));
    CUDA_TRY(d_x.reserve(32));
    CUDA_TRY(d_count.reserve(4));
    return B200_OK;
}

void Iekf::destroy() {
    if (map) cudaSetDevice(map->device);
    if (stream) cudaStreamSynchronize(stream);
    cudaFree(d_ctl); cudaFree(d_limit); cudaFree(d_partials);
    cudaFree(ps.plane); cudaFree(ps.resid); cudaFree(ps.sel); cudaFree(ps.nn_cnt); cudaFree(ps.nn);
    d_scan.release(); h_stage.release(); h_out.release();
    d_world.release(); d_sel_pts.release(); d_flag.release(); d_flag2.release(); cub_tmp.release(); d_count.release(); d_x.release();
    h_count.release(); h_x.release();
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
}

int32_t Iekf::ensure_points(size_t n) {
    if (n <= ps_cap) return B200_OK;
    size_t ncap = n + n / 2 + 1024;
    PointState np{};
    CUDA_TRY(cudaMalloc(&np.plane, ncap * sizeof(float4)));
    CUDA_TRY(cudaMalloc(&np.resid, ncap * sizeof(float)));
    CUDA_TRY(cudaMalloc(&np.sel, ncap));
    CUDA_TRY(cudaMalloc(&np.nn_cnt, ncap));
    CUDA_TRY(cudaMalloc(&np.nn, ncap * 5 * sizeof(float4)));
    if (ps_cap) {
        CUDA_TRY(cudaMemcpyAsync(np.plane, ps.plane, ps_cap * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(np.resid, ps.resid, ps_cap * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(np.sel, ps.sel, ps_cap, cudaMemcpyDeviceToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(np.nn_cnt, ps.nn_cnt, ps_cap, cudaMemcpyDeviceToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(np.nn, ps.nn, ps_cap * 5 * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        cudaFree(ps.plane); cudaFree(ps.resid); cudaFree(ps.sel); cudaFree(ps.nn_cnt); cudaFree(ps.nn);
    }
    ps = np;
    ps_cap = ncap;
    return B200_OK;
}

// d_pts: device scan (float4).  d_hdr: device copy of the Ctl header (x, P, n, prev_n) staged with the
// scan, or null to upload it from x/P here.
int32_t Iekf::run(const float4* d_pts, int n, const Ctl* d_hdr, double* x, double* P, b200_iekf_stats* st, int single_pass,
                  int force_converge) {
    CUDA_TRY(cudaSetDevice(map->device));
    int32_t rc = ensure_points((size_t)n);
    if (rc) return rc;
    if (!d_hdr) {
        CUDA_TRY(h_stage.reserve(offsetof(Ctl, x_prop)));
        Ctl* hc = (Ctl*)h_stage.p;
        memcpy(hc->x, x, sizeof(double) * 26);
        if (P) memcpy(hc->P, P, sizeof(double) * NS * NS); else memset(hc->P, 0, sizeof(double) * NS * NS);
        hc->n = n;
        hc->prev_n = last_n < n ? last_n : n;
        CUDA_TRY(cudaMemcpyAsync(d_ctl, hc, offsetof(Ctl, x_prop), cudaMemcpyHostToDevice, stream));
        d_hdr = d_ctl;
    }
    CUDA_TRY(cudaEventRecord(ev0, stream));
    k_iekf_init<<<8, 256, 0, stream>>>(d_ctl, d_hdr, ps, force_converge);
    const MapView mv = map->view();
    const int npass = single_pass ? 1 : prm.max_iter + 1;
    for (int it = 0; it < npass; ++it) {
        k_obs<<<nblocks, OBS_THREADS, 0, stream>>>(mv, d_pts, ps, d_ctl, prm.plane_thr, prm.extrinsic_est_en, d_partials);
        k_solve<<<1, 256, 0, stream>>>(d_ctl, d_partials, nblocks, prm.max_iter, prm.R, d_limit, single_pass);
    }
    LAUNCH_COUNT(1 + 2 * npass);
    CUDA_TRY(cudaEventRecord(ev1, stream));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_out.p, d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    last_n = n;
    last_scan = d_pts;
    const Ctl& o = *h_out.p;
    if (!single_pass) {
        memcpy(x, o.x, sizeof(double) * 26);
        if (P) memcpy(P, o.P, sizeof(double) * NS * NS);
    }
    int32_t status = o.any_valid || single_pass ? B200_OK : B200_NO_EFFECTIVE_POINTS;
    if (st) {
        memset(st, 0, sizeof *st);
        st->status = status;
        st->passes = o.passes;
        st->knn_passes = o.knn_passes;
        st->converged = o.converged;
        for (int i = 0; i < B200_MAX_PASSES; ++i) { st->n_eff[i] = o.n_eff[i]; st->knn[i] = o.knn[i]; }
        cudaEventElapsedTime(&st->gpu_ms, ev0, ev1);
    }
    return status;
}

}  // namespace b200

// ------------------------------------------------------------------ C ABI (B2)
using namespace b200;
struct b200_iekf { Iekf k; };

extern "C" {

int32_t b200_iekf_create(const b200_iekf_params* params, b200_map* map, b200_iekf** out) {
    if (!params || !map || !out) B200_FAIL(B200_ERR_ARG, "null argument");
    b200_iekf* h = new b200_iekf();
    int32_t rc = h->k.init(params, &map->m);
    if (rc != B200_OK) { h->k.destroy(); delete h; return rc; }
    *out = h;
    return B200_OK;
}
int32_t b200_iekf_destroy(b200_iekf* ekf) {
    if (!ekf) return B200_OK;
    ekf->k.destroy();
    delete ekf;
    return B200_OK;
}

static int32_t stage_scan(Iekf& k, const float* xyz, int64_t n, int64_t stride, const double* x, const double* P) {
    // one pinned block [Ctl header | points] -> one H2D copy
    const size_t hdr = offsetof(Ctl, x_prop);
    const size_t hdr_pad = (hdr + 255) / 256 * 256;
    CUDA_TRY(k.h_stage.reserve(hdr_pad + (size_t)n * sizeof(float4)));
    CUDA_TRY(k.d_scan.reserve((size_t)n + hdr_pad / sizeof(float4)));
    Ctl* hc = (Ctl*)k.h_stage.p;
    memcpy(hc->x, x, sizeof(double) * 26);
    if (P) memcpy(hc->P, P, sizeof(double) * NS * NS); else memset(hc->P, 0, sizeof(double) * NS * NS);
    hc->n = (int)n;
    hc->prev_n = k.last_n < (int)n ? k.last_n : (int)n;
    pack_xyz_float4(xyz, n, stride, (float4*)(k.h_stage.p + hdr_pad));
    // the device scan buffer mirrors the pinned layout; k_iekf_init forwards the header to the control block
    CUDA_TRY(cudaMemcpyAsync(k.d_scan.p, k.h_stage.p, hdr_pad + (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, k.stream));
    return B200_OK;
}

int32_t b200_iekf_update(b200_iekf* ekf, const float* scan, int64_t n, int64_t stride, double* x26, double* P, b200_iekf_stats* stats) {
    if (!ekf || !scan || !x26 || !P || n < 1 || stride < 12 || n > (1 << 26)) B200_FAIL(B200_ERR_ARG, "bad argument");
    Iekf& k = ekf->k;
    CUDA_TRY(cudaSetDevice(k.map->device));
    int32_t rc = stage_scan(k, scan, n, stride, x26, P);
    if (rc) return rc;
    const size_t hdr_pad = (offsetof(Ctl, x_prop) + 255) / 256 * 256;
    return k.run((const float4*)((const uint8_t*)k.d_scan.p + hdr_pad), (int)n, (const Ctl*)k.d_scan.p, x26, P, stats, 0, 1);
}

int32_t b200_iekf_update_device(b200_iekf* ekf, const void* d_scan_float4, int64_t n, double* x26, double* P, b200_iekf_stats* stats) {
    if (!ekf || !d_scan_float4 || !x26 || !P || n < 1 || n > (1 << 26)) B200_FAIL(B200_ERR_ARG, "bad argument");
    return ekf->k.run((const float4*)d_scan_float4, (int)n, nullptr, x26, P, stats, 0, 1);
}

int32_t b200_iekf_last_HtH(b200_iekf* ekf, int32_t pass, double* HtH, double* Hth, double* x_in) {
    if (!ekf || pass < 0 || pass >= B200_MAX_PASSES) B200_FAIL(B200_ERR_ARG, "bad argument");
    const Ctl& o = *ekf->k.h_out.p;
    if (pass >= o.passes) B200_FAIL(B200_ERR_ARG, "pass was not executed");
    if (HtH) memcpy(HtH, o.HtH[pass], sizeof(double) * 144);
    if (Hth) memcpy(Hth, o.Hth[pass], sizeof(double) * 12);
    if (x_in) memcpy(x_in, o.x_in[pass], sizeof(double) * 26);
    return B200_OK;
}

int32_t b200_iekf_obs_model(b200_iekf* ekf, const float* scan, int64_t n, int64_t stride, const double* x26, int32_t converge,
                            double* HtH, double* Hth, int32_t* n_eff) {
    if (!ekf || !scan || !x26 || n < 1 || stride < 12) B200_FAIL(B200_ERR_ARG, "bad argument");
    Iekf& k = ekf->k;
    CUDA_TRY(cudaSetDevice(k.map->device));
    int32_t rc = stage_scan(k, scan, n, stride, x26, nullptr);
    if (rc) return rc;
    const size_t hdr_pad = (offsetof(Ctl, x_prop) + 255) / 256 * 256;
    double xtmp[26];
    memcpy(xtmp, x26, sizeof xtmp);
    b200_iekf_stats st;
    rc = k.run((const float4*)((const uint8_t*)k.d_scan.p + hdr_pad), (int)n, (const Ctl*)k.d_scan.p, xtmp, nullptr, &st, 1, converge ? 1 : 0);
    if (rc < 0) return rc;
    const Ctl& o = *k.h_out.p;
    if (HtH) memcpy(HtH, o.HtH[0], sizeof(double) * 144);
    if (Hth) memcpy(Hth, o.Hth[0], sizeof(double) * 12);
    if (n_eff) *n_eff = o.n_eff[0];
    return o.n_eff[0] > 0 ? B200_OK : B200_NO_EFFECTIVE_POINTS;
}

int32_t b200_iekf_point_state(b200_iekf* ekf, int64_t n, float* plane4, float* residual, uint8_t* selected, int32_t* nn_idx5,
                              int32_t* nn_count) {
    if (!ekf || n < 0 || (size_t)n > ekf->k.ps_cap) B200_FAIL(B200_ERR_ARG, "bad argument");
    Iekf& k = ekf->k;
    CUDA_TRY(cudaSetDevice(k.map->device));
    CUDA_TRY(cudaStreamSynchronize(k.stream));
    if (plane4) CUDA_TRY(cudaMemcpy(plane4, k.ps.plane, n * sizeof(float4), cudaMemcpyDeviceToHost));
    if (residual) CUDA_TRY(cudaMemcpy(residual, k.ps.resid, n * sizeof(float), cudaMemcpyDeviceToHost));
    if (selected) CUDA_TRY(cudaMemcpy(selected, k.ps.sel, n, cudaMemcpyDeviceToHost));
    if (nn_idx5 || nn_count) {
        std::vector<char> cnt((size_t)n);
        std::vector<char> nn((size_t)n * 5 * sizeof(float4));
        CUDA_TRY(cudaMemcpy(&cnt[0], k.ps.nn_cnt, n, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(&nn[0], k.ps.nn, n * 5 * sizeof(float4), cudaMemcpyDeviceToHost));
        const float4* p = (const float4*)nn.data();
        for (int64_t i = 0; i < n; ++i) {
            int c = (unsigned char)cnt[i];
            if (nn_count) nn_count[i] = c;
            if (nn_idx5)
                for (int j = 0; j < 5; ++j) {
                    int32_t ord;
                    memcpy(&ord, &p[i * 5 + j].w, 4);
                    nn_idx5[i * 5 + j] = j < c ? ord : -1;
                }
        }
    }
    return B200_OK;
}

int32_t b200_iekf_map_incremental(b200_iekf* ekf, const double* x26, int32_t ekf_inited, int32_t* n_added, int32_t* n_no_downsample) {
    if (!ekf || !x26) B200_FAIL(B200_ERR_ARG, "bad argument");
    Iekf& k = ekf->k;
    const int n = k.last_n;
    if (n < 1 || !k.last_scan) B200_FAIL(B200_ERR_ARG, "no scan has been processed");
    CUDA_TRY(cudaSetDevice(k.map->device));
    CUDA_TRY(k.d_world.reserve(n)); CUDA_TRY(k.d_sel_pts.reserve(2 * (size_t)n));
    CUDA_TRY(k.d_flag.reserve(n)); CUDA_TRY(k.d_flag2.reserve(n));
    memcpy(k.h_x.p, x26, sizeof(double) * 26);
    CUDA_TRY(cudaMemcpyAsync(k.d_x.p, k.h_x.p, sizeof(double) * 26, cudaMemcpyHostToDevice, k.stream));
    const int nb = (n + 255) / 256;
    k_map_incremental_flags<<<nb, 256, 0, k.stream>>>(k.last_scan, n, k.d_x.p, k.ps, ekf_inited, k.prm.filter_size_map, k.d_world.p, k.d_flag.p);
    size_t tmp = 0;
    cub::DeviceSelect::Flagged(nullptr, tmp, k.d_world.p, k.d_flag2.p, k.d_sel_pts.p, k.d_count.p, n, k.stream);
    CUDA_TRY(k.cub_tmp.reserve(tmp));
    // points_to_add first, then point_no_need_downsample, each in scan order (:579-580)
    k_flag_eq<<<nb, 256, 0, k.stream>>>(k.d_flag.p, n, 1, k.d_flag2.p);
    CUDA_TRY(cub::DeviceSelect::Flagged(k.cub_tmp.p, tmp, k.d_world.p, k.d_flag2.p, k.d_sel_pts.p, k.d_count.p, n, k.stream));
    CUDA_TRY(cudaMemcpyAsync(k.h_count.p, k.d_count.p, sizeof(int32_t), cudaMemcpyDeviceToHost, k.stream));
    CUDA_TRY(cudaStreamSynchronize(k.stream));
    const int na = k.h_count.p[0];
    k_flag_eq<<<nb, 256, 0, k.stream>>>(k.d_flag.p, n, 2, k.d_flag2.p);
    CUDA_TRY(cub::DeviceSelect::Flagged(k.cub_tmp.p, tmp, k.d_world.p, k.d_flag2.p, k.d_sel_pts.p + na, k.d_count.p, n, k.stream));
    CUDA_TRY(cudaMemcpyAsync(k.h_count.p, k.d_count.p, sizeof(int32_t), cudaMemcpyDeviceToHost, k.stream));
    CUDA_TRY(cudaStreamSynchronize(k.stream));
    const int nd = k.h_count.p[0];
    LAUNCH_COUNT(3);
    if (n_added) *n_added = na;
    if (n_no_downsample) *n_no_downsample = nd;
    return k.map->insert_device(k.d_sel_pts.p, (int64_t)na + nd);
}

}  // extern "C"
