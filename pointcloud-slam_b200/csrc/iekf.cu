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
#ifndef B200_KNN_BLOCK
#define B200_KNN_BLOCK 256
#endif
#include "map.cuh"
#include "manifold.cuh"
#include "pointmath.cuh"
#include "solve.cuh"

#include <cub/cub.cuh>
#include <chrono>
#include <vector>

namespace b200 {

constexpr int OBS_THREADS = 512;
constexpr int KNN_G = 8;        // lanes per query in a search pass
constexpr int KNN_TILE = OBS_THREADS / KNN_G;

struct PointState {  // per-point arrays that persist across passes and scans (laser_mapping.cc:335-339)
    float4* plane;   // plane_coef_
    float* resid;    // residuals_
    uint8_t* sel;    // point_selected_surf_
    uint8_t* nn_cnt; // nearest_points_[i].size()
    float4* nn;      // nearest_points_[i][0..4] (xyz + ordinal)
};

// ------------------------------------------------------------------ init
// hdr: the {x, P, n, prev_n} header as it arrived from the host (either inside the staged scan block or
// already in the control block itself)
__global__ void k_iekf_init(Ctl* ctl, const Ctl* hdr, PointState ps, int force_converge) {
    pdl_trigger();  // the first search may be scheduled while this kernel runs (it waits in griddepcontrol.wait)
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
            ctl->ticket = 0;
            ctl->fault = 0;
            ctl->passes = ctl->knn_passes = ctl->any_valid = ctl->converged = 0;
            for (int i = 0; i < B200_MAX_PASSES; ++i) { ctl->n_eff[i] = 0; ctl->knn[i] = 0; }
        }
        make_pass_consts_par(hdr->x, ctl->pc);  // four independent pieces on four warps
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

// strided xyz records (12-byte packed, 16-byte, 48-byte PointXYZINormal ...) -> float4, for scans that arrived unpacked
// from page-locked caller memory
__global__ void k_unpack_xyz(const uint8_t* __restrict__ raw, int64_t stride, int n, float4* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
    dst[i] = make_float4(p[0], p[1], p[2], 0.0f);
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

constexpr int QCAP = 256;  // queries a worker block handles per round (20k-point scan / 147 workers = 137)
struct ObsSmem {
    float rows[QCAP][13];
    unsigned char eff[QCAP];
};

// Stencil search of a pass (dyn_share.converge == true): 8 lanes per scan point, one wave over the
// whole scan at high occupancy.  Results go to a scratch array that k_obs consumes; the kernel is a
// no-op when the pass reuses the previous neighbours (decided on the device).
constexpr int KNN_BLOCK = B200_KNN_BLOCK;  // threads per search block
template <int MODE>
__global__ void __launch_bounds__(KNN_BLOCK, (MODE == 0 ? 1280 : MODE == 5 ? 1152 : 1024) / KNN_BLOCK) k_search(MapView map, const float4* __restrict__ scan, const Ctl* __restrict__ ctl,
                                                float4* __restrict__ nb_out, unsigned char* __restrict__ nbc_out) {
    pdl_trigger();
    pdl_wait();
    if (ctl->done || !ctl->converge) return;
    __shared__ PassConsts pc;
    const int tid = threadIdx.x;
    if (tid < (int)(sizeof(PassConsts) / 4)) ((float*)&pc)[tid] = ((const float*)&ctl->pc)[tid];
    __syncthreads();
    const int n = ctl->n;
    const int q = (blockIdx.x * blockDim.x + tid) / KNN_G, lg = tid % KNN_G;
    if (q >= n) return;
    const unsigned gmask = ((1u << KNN_G) - 1u) << ((tid & 31) / KNN_G * KNN_G);
    const float4 pbody = __ldg(scan + q);
    const float3 pw = body_to_world(pc, pbody.x, pbody.y, pbody.z);
    uint64_t wkey;
    float4 mine;
    __shared__ uint2 s_flat[MODE >= 5 ? (KNN_BLOCK / KNN_G) * kFlatStride : 1];
    const int c = knn5_group<KNN_G, MODE>(map, pw.x, pw.y, pw.z, lg, gmask, lane_stencil<KNN_G>(lg, map.nstencil), wkey, mine,
                                          MODE >= 5 ? s_flat + (tid / KNN_G) * kFlatStride : nullptr);
    if (lg < 5) nb_out[(size_t)q * 5 + lg] = mine;
    if (lg == 0) nbc_out[q] = (unsigned char)c;
}

// warp-per-query variant (knn5_warp, map.cuh): one warp per scan point
__global__ void __launch_bounds__(256) k_search_w(MapView map, const float4* __restrict__ scan, const Ctl* __restrict__ ctl,
                                                  float4* __restrict__ nb_out, unsigned char* __restrict__ nbc_out) {
    pdl_trigger();
    pdl_wait();
    if (ctl->done || !ctl->converge) return;
    __shared__ PassConsts pc;
    const int tid = threadIdx.x;
    if (tid < (int)(sizeof(PassConsts) / 4)) ((float*)&pc)[tid] = ((const float*)&ctl->pc)[tid];
    __syncthreads();
    const int q = (blockIdx.x * blockDim.x + tid) >> 5, lane = tid & 31;
    if (q >= ctl->n) return;
    const float4 pbody = __ldg(scan + q);
    const float3 pw = body_to_world(pc, pbody.x, pbody.y, pbody.z);
    float4 mine;
    uint32_t key;
    const int c = knn5_warp(map, pw.x, pw.y, pw.z, lane, mine, key);
    if (lane < 5) nb_out[(size_t)q * 5 + lane] = mine;
    if (lane == 0) nbc_out[q] = (unsigned char)c;
}

// warp-per-query with the candidate runs staged in shared memory by 1-D TMA bulk copies (knn5_warp_t<true>, B200_KNN_MODE=9)
__global__ void __launch_bounds__(256) k_search_t(MapView map, const float4* __restrict__ scan, const Ctl* __restrict__ ctl,
                                                  float4* __restrict__ nb_out, unsigned char* __restrict__ nbc_out) {
    pdl_trigger();
    pdl_wait();
    if (ctl->done || !ctl->converge) return;
    __shared__ PassConsts pc;
    __shared__ __align__(16) float4 s_stage[8][kStageCap];
    __shared__ __align__(8) uint64_t s_bar[8];
    const int tid = threadIdx.x;
    if (tid < (int)(sizeof(PassConsts) / 4)) ((float*)&pc)[tid] = ((const float*)&ctl->pc)[tid];
    __syncthreads();
    const int q = (blockIdx.x * blockDim.x + tid) >> 5, lane = tid & 31;
    if (q >= ctl->n) return;
    const float4 pbody = __ldg(scan + q);
    const float3 pw = body_to_world(pc, pbody.x, pbody.y, pbody.z);
    float4 mine;
    uint32_t key;
    const int c = knn5_warp_t<true>(map, pw.x, pw.y, pw.z, lane, mine, key, s_stage[tid >> 5], s_bar + (tid >> 5));
    if (lane < 5) nb_out[(size_t)q * 5 + lane] = mine;
    if (lane == 0) nbc_out[q] = (unsigned char)c;
}

// balanced 8-lanes-per-query variant (knn5_g8p, map.cuh)
__global__ void __launch_bounds__(256, 5) k_search_p(MapView map, const float4* __restrict__ scan, const Ctl* __restrict__ ctl,
                                                  float4* __restrict__ nb_out, unsigned char* __restrict__ nbc_out) {
    pdl_trigger();
    pdl_wait();
    if (ctl->done || !ctl->converge) return;
    __shared__ PassConsts pc;
    __shared__ __align__(16) uint32_t s_q[(256 / 8) * kG8pWords];
    const int tid = threadIdx.x;
    if (tid < (int)(sizeof(PassConsts) / 4)) ((float*)&pc)[tid] = ((const float*)&ctl->pc)[tid];
    __syncthreads();
    const int q = (blockIdx.x * blockDim.x + tid) >> 3, lg = tid & 7;
    const bool active = q < ctl->n;   // groups past the end stay with their warp (full-mask collectives inside)
    float3 pw = make_float3(0.f, 0.f, 0.f);
    if (active) {
        const float4 pbody = __ldg(scan + q);
        pw = body_to_world(pc, pbody.x, pbody.y, pbody.z);
    }
    float4 mine;
    uint32_t key;
    const int c = knn5_g8p(map, active, pw.x, pw.y, pw.z, lg, s_q + (tid >> 3) * kG8pWords, mine, key);
    if (!active) return;
    if (lg < 5) nb_out[(size_t)q * 5 + lg] = mine;
    if (lg == 0) nbc_out[q] = (unsigned char)c;
}

union ObsSolveSmem {
    ObsSmem obs;
    SolveSmem solve;
};

// One IEKF pass = one launch.  Block 0 is the filter block: it prepares the state-only half of the
// filter step while blocks 1.. measure the scan and publish fp64 partial sums, waits for their
// tickets, then finishes the step.  (Blocks 1.. never wait on block 0, so there is no deadlock; with
// one block per SM every block of the grid is resident.)
__global__ void __launch_bounds__(OBS_THREADS, 1) k_obs(const float4* __restrict__ scan, const float4* __restrict__ nb_new,
                                                        const unsigned char* __restrict__ nbc_new, PointState ps, Ctl* ctl, float thr,
                                                        int ext, double* partials, int max_iter, double Rcov,
                                                        const double* __restrict__ limit, int single_pass) {
    pdl_trigger();
    pdl_wait();
    if (ctl->done) return;
    __shared__ ObsSolveSmem smu;
    if (blockIdx.x == 0) {
        const int nworkers = (int)gridDim.x - 1;
#ifdef B200_STAMPS
        const int pass_f = ctl->passes;
        const long long tf0 = clock64();
#endif
        if (!single_pass) iekf_presolve(ctl, Rcov, ext, smu.solve);
        else if (threadIdx.x < 26) smu.solve.x[threadIdx.x] = ctl->x[threadIdx.x];
#ifdef B200_STAMPS
        const long long tf1 = clock64();
#endif
        if (threadIdx.x == 0) {
            // wait for the workers' tickets, with a watchdog (~0.25 s): worker blocks never wait on this block, so the only way
            // to starve here is a grid larger than the device can hold or a fault in a worker - report it instead of hanging
            volatile unsigned int* tk = &ctl->ticket;
            unsigned int spins = 0;
            while (*tk < (unsigned)nworkers && ++spins < (1u << 22)) __nanosleep(40);
            if (*tk < (unsigned)nworkers) { ctl->fault = 1; ctl->done = 1; }
            ctl->ticket = 0;
#ifdef B200_STAMPS
            if (pass_f < B200_MAX_PASSES) {
                ctl->dbg[pass_f][8] = tf1 - tf0;         // presolve cycles
                ctl->dbg[pass_f][9] = clock64() - tf1;   // wait-for-workers cycles
            }
#endif
            __threadfence();  // acquire: the partials are read with ld.cg after the barrier below
        }
        __syncthreads();
        if (ctl->fault) return;
        iekf_postsolve(ctl, partials, nworkers, max_iter, limit, ext, single_pass, smu.solve);
        return;
    }
    const int wblock = (int)blockIdx.x - 1, wgrid = (int)gridDim.x - 1;  // worker index / count
    ObsSmem& sm = smu.obs;
    const int tid = threadIdx.x;
    const int n = ctl->n;
    const bool searched = ctl->converge != 0;
    __shared__ PassConsts pc;
    if (tid < (int)(sizeof(PassConsts) / 4)) ((float*)&pc)[tid] = ((const float*)&ctl->pc)[tid];
    __syncthreads();

    // Accumulators: the block's queries are cut into nslice contiguous slices per tile; thread (slice, col) owns active
    // column col (solve.cuh: 21 + 6 + 1 columns without extrinsic estimation, 78 + 12 + 1 with) of its slice across
    // all tiles.  Fixed slice boundaries and query order: the sums are deterministic.
    const int ncol = ext ? NPART : NCOL6;
    const int nslice = OBS_THREADS / ncol;  // 5 or 18
    const int slice = tid / ncol, col = tid % ncol;
    double acc = 0.0;
    const int pa = c_col_a[ext ? 1 : 0][col], pb_ = c_col_b[ext ? 1 : 0][col];
    const bool slice_active = slice < nslice;
    const bool pair_active = slice_active && pa != 255;

#ifdef B200_STAMPS
    long long tw0 = clock64(), tw1 = tw0, tw2 = tw0, tw3 = tw0;
#define WSTAMP(v) v = clock64()
#else
#define WSTAMP(v) do { } while (0)
#endif
    // each worker owns one contiguous chunk of the scan, processed in rounds of <= QCAP queries
    const int chunk = (n + wgrid - 1) / wgrid;
    const int c_begin = wblock * chunk, c_end = min(n, c_begin + chunk);
    for (int base = c_begin; base < c_end; base += QCAP) {
        const int q_here = min(QCAP, c_end - base);
        WSTAMP(tw1);
        // phase 2: one thread per query
        if (tid < q_here) {
            const int i = base + tid;
            const float4 pbody = __ldg(scan + i);
            const float3 pw = body_to_world(pc, pbody.x, pbody.y, pbody.z);
            float row[12], h = 0.f;
            const bool eff = point_measure(pc, pbody, pw, searched, searched ? (int)nbc_new[i] : 0, nb_new + (size_t)i * 5, ps, i, thr,
                                           ext != 0, row, h);
            sm.eff[tid] = eff ? 1 : 0;
            if (eff) {
#pragma unroll
                for (int k = 0; k < 12; ++k) sm.rows[tid][k] = row[k];
                sm.rows[tid][12] = h;
            }
        }
        __syncthreads();
        WSTAMP(tw2);
        // phase 3: fp64 accumulation, fixed slice boundaries and query order (deterministic)
        {
            const int per = (q_here + nslice - 1) / nslice;
            const int q0 = slice * per, q1 = min(q_here, q0 + per);
            if (pair_active) {
                for (int q = q0; q < q1; ++q)
                    if (sm.eff[q]) acc += (double)sm.rows[q][pa] * (double)sm.rows[q][pb_];
            } else if (slice_active) {
                for (int q = q0; q < q1; ++q) acc += (double)sm.eff[q];
            }
        }
        __syncthreads();
    }
    // combine the slices in a fixed order, then publish column-major by block so the filter block
    // reads each column coalesced
    double* red = (double*)&sm.rows[0][0];
    if (slice_active) red[slice * ncol + col] = acc;
    __syncthreads();
    if (tid < ncol) {
        double v = red[tid];
        for (int sl = 1; sl < nslice; ++sl) v += red[sl * ncol + tid];
        __stcg(partials + (size_t)tid * wgrid + wblock, v);
        __threadfence();  // release: only the publishing threads need it
    }
    __syncthreads();
#ifdef B200_STAMPS
    tw3 = clock64();
    if (tid == 0 && wblock == 0) {
        const int pw_ = ctl->passes;
        if (pw_ < B200_MAX_PASSES) {
            ctl->dbg[pw_][10] = tw1 - tw0;  // search phase (last round)
            ctl->dbg[pw_][11] = tw2 - tw1;  // measure phase
            ctl->dbg[pw_][12] = tw3 - tw2;  // accumulate + publish
        }
    }
#endif
    if (tid == 0) atomicAdd(&ctl->ticket, 1u);
}

// ------------------------------------------------------------------ forward propagation (esekf::predict, esekfom.hpp:269-374)
// K IMU intervals in ONE launch of one block: per step the state-dependent pieces of the process model (use-ikfom.hpp:36-77:
// get_f, df_dx, df_dw) on three lanes, F = F_x1 + f_x_final dt and G = f_w_final assembled in shared memory, then
// P = F P F^T + (dt G) Q (dt G)^T with one thread per covariance entry.  Quirks of the reference kept: the SO3 / S2 blocks of
// F_x1 come from MTK::exp(.., scalar(1 / 2)) - an INTEGER quotient, scale 0, the identity rotation (esekfom.hpp:307,331).
// steps: K x 8 {dt, offs_t, acc_avr[3], angvel_avr[3]} (imu_processing.hpp:190-241); poses22: K x 22 IMUpose_ entries
// {offs_t, acc_s_last, angvel_last, vel, pos, R} (:225-236) for the backward undistortion pass (b200_scan_undistort).
struct PredictSmem {
    double x[26], xb[26];
    double F[NS * NS], P[NS * NS], T[NS * NS];
    double G[NS * 12];
    double R[9], RH[9], A[9], Mx0[6], Nx[6], omega[3], am[3], a_in[3];
};
__global__ void __launch_bounds__(544, 1) k_iekf_predict(const double* __restrict__ steps, int K, const double* __restrict__ Q12,
                                                         double* x_io, double* P_io, double* __restrict__ poses22) {
    __shared__ PredictSmem s;
    using namespace mf;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid < 26) s.x[tid] = x_io[tid];
    for (int i = tid; i < NS * NS; i += nt) s.P[i] = P_io[i];
    __syncthreads();
    for (int k = 0; k < K; ++k) {
        const double dt = steps[k * 8], offs_t = steps[k * 8 + 1];
        const double* acc = steps + k * 8 + 2;
        const double* gyr = steps + k * 8 + 5;
        for (int i = tid; i < NS * NS; i += nt) s.F[i] = (i / NS == i % NS) ? 1.0 : 0.0;
        for (int i = tid; i < NS * 12; i += nt) s.G[i] = 0.0;
        if (tid < 26) s.xb[tid] = s.x[tid];
        __syncthreads();
        if (tid == 0) {  // rotation chain of the state (get_f, df_dx blocks, x_.oplus)
            for (int i = 0; i < 3; ++i) { s.omega[i] = gyr[i] - s.xb[17 + i]; s.am[i] = acc[i] - s.xb[20 + i]; }
            const Q rot = ldq(s.xb + 3);
            qrot(rot, s.am, s.a_in);
            qtoR(rot, s.R);
            double H[9];
            hat(s.am, H);
            mm3(s.R, H, s.RH);
            for (int i = 0; i < 3; ++i) s.x[i] = s.xb[i] + dt * s.xb[14 + i];
            stq(s.x + 3, qmul(rot, so3_exp(s.omega, dt / 2)));
            const double z3[3] = {0.0, 0.0, 0.0};
            stq(s.x + 7, qmul(ldq(s.xb + 7), so3_exp(z3, dt / 2)));
            for (int i = 0; i < 3; ++i) s.x[14 + i] = s.xb[14 + i] + dt * (s.a_in[i] + s.xb[23 + i]);
        } else if (tid == 32) {  // gravity manifold pieces (the gravity vector itself does not move: its rate is zero)
            const double zero2[2] = {0.0, 0.0};
            S2_Mx(s.xb + 23, zero2, s.Mx0);
            S2_Nx_yy(s.xb + 23, s.Nx);
        } else if (tid == 64) {
            double seg[3];
            for (int i = 0; i < 3; ++i) seg[i] = -1 * (gyr[i] - s.xb[17 + i]) * dt;
            A_matrix(seg, s.A);
        }
        __syncthreads();
        if (tid < 9) {  // 3 x 3 blocks, one entry per thread
            const int r = tid / 3, c = tid % 3;
            s.F[(3 + r) * NS + 15 + c] += (-s.A[r * 3 + c]) * dt;
            s.G[(3 + r) * 12 + c] = -s.A[r * 3 + c];
            s.F[(12 + r) * NS + 3 + c] += (-s.RH[r * 3 + c]) * dt;
            s.F[(12 + r) * NS + 18 + c] += (-s.R[r * 3 + c]) * dt;
            s.G[(12 + r) * 12 + 3 + c] = -s.R[r * 3 + c];
            if (c < 2) s.F[(12 + r) * NS + 21 + c] += s.Mx0[r * 2 + c] * dt;
            if (r == c) {
                s.F[r * NS + 12 + r] += 1.0 * dt;
                s.G[(15 + r) * 12 + 6 + r] = 1.0;
                s.G[(18 + r) * 12 + 9 + r] = 1.0;
            }
            if (r < 2 && c < 2) s.F[(21 + r) * NS + 21 + c] = s.Nx[r * 3] * s.Mx0[c] + s.Nx[r * 3 + 1] * s.Mx0[2 + c] + s.Nx[r * 3 + 2] * s.Mx0[4 + c];
        }
        __syncthreads();
        if (tid < NS * NS) {  // T = F P
            const int i = tid / NS, j = tid % NS;
            double v = 0.0;
            for (int c = 0; c < NS; ++c) v += s.F[i * NS + c] * s.P[c * NS + j];
            s.T[tid] = v;
        }
        __syncthreads();
        double pn = 0.0;
        if (tid < NS * NS) {  // P = T F^T + (dt G) Q (dt G)^T
            const int i = tid / NS, j = tid % NS;
            double v = 0.0, w = 0.0;
            for (int c = 0; c < NS; ++c) v += s.T[i * NS + c] * s.F[j * NS + c];
            for (int c = 0; c < 12; ++c) w += ((dt * s.G[i * 12 + c]) * Q12[c]) * (dt * s.G[j * 12 + c]);
            pn = v + w;
        }
        __syncthreads();
        if (tid < NS * NS) s.P[tid] = pn;
        if (poses22 && tid == 0) {  // imu_processing.hpp:225-236, from the propagated state
            double* o = poses22 + (size_t)k * 22;
            double am2[3], as[3], R2[9];
            for (int i = 0; i < 3; ++i) am2[i] = acc[i] - s.x[20 + i];
            const Q rot = ldq(s.x + 3);
            qrot(rot, am2, as);
            qtoR(rot, R2);
            o[0] = offs_t;
            for (int i = 0; i < 3; ++i) { o[1 + i] = as[i] + s.x[23 + i]; o[4 + i] = gyr[i] - s.x[17 + i]; o[7 + i] = s.x[14 + i]; o[10 + i] = s.x[i]; }
            for (int i = 0; i < 9; ++i) o[13 + i] = R2[i];
        }
        __syncthreads();
    }
    if (tid < 26) x_io[tid] = s.x[tid];
    for (int i = tid; i < NS * NS; i += nt) P_io[i] = s.P[i];
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
// PointBodyToWorld (laser_mapping.cc:855-864) of the whole scan: fp64 quaternion arithmetic, narrowed on store - the cloud
// PublishFrameWorld sends as /cloud_registered (:747-773).  Records are written at `stride` bytes (x, y, z first; 48 = the
// PointXYZINormal layout of the message, the other bytes are left as they are).
__global__ void k_body_to_world(const float4* __restrict__ scan, int n, const double* __restrict__ x, uint8_t* __restrict__ out, int64_t stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    using namespace mf;
    const float4 pb = scan[i];
    double pbd[3] = {pb.x, pb.y, pb.z}, t1[3], t2[3];
    qrot(ldq(x + 7), pbd, t1);
    for (int k = 0; k < 3; ++k) t1[k] = t1[k] + x[11 + k];
    qrot(ldq(x + 3), t1, t2);
    float* o = reinterpret_cast<float*>(out + (size_t)i * stride);
    o[0] = (float)(t2[0] + x[0]);
    o[1] = (float)(t2[1] + x[1]);
    o[2] = (float)(t2[2] + x[2]);
}

// Stable three-way partition of the scan by flag, one block: points_to_add (flag 1) first, then point_no_need_downsample
// (flag 2), each in scan order (laser_mapping.cc:579-580) - the order the insertion ordinals follow.  counts = {na, nd, na + nd}.
// Thread t owns the contiguous slice [t * per, (t + 1) * per) of the scan: two counts per thread, one block scan, then every
// thread writes its slice behind its offsets - one pass, no host round trip.
__global__ void __launch_bounds__(1024) k_partition3(const uint8_t* __restrict__ flag, const float4* __restrict__ world, int n,
                                                     float4* __restrict__ out, int32_t* __restrict__ counts) {
    __shared__ int s_w1[32], s_w2[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + 1023) / 1024;
    const int i0 = min(n, tid * per), i1 = min(n, i0 + per);
    int c1 = 0, c2 = 0;
    for (int i = i0; i < i1; ++i) {
        const int f = flag[i];
        c1 += f == 1 ? 1 : 0;
        c2 += f == 2 ? 1 : 0;
    }
    int x1 = c1, x2 = c2;  // inclusive scan inside the warp
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, x1, o), b = __shfl_up_sync(0xffffffffu, x2, o);
        if (lane >= o) { x1 += a; x2 += b; }
    }
    if (lane == 31) { s_w1[warp] = x1; s_w2[warp] = x2; }
    __syncthreads();
    if (warp == 0) {  // scan of the 32 warp totals
        int a = s_w1[lane], b = s_w2[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, a, o), v = __shfl_up_sync(0xffffffffu, b, o);
            if (lane >= o) { a += u; b += v; }
        }
        s_w1[lane] = a;
        s_w2[lane] = b;
    }
    __syncthreads();
    const int na = s_w1[31], nd = s_w2[31];
    int p1 = x1 - c1 + (warp ? s_w1[warp - 1] : 0), p2 = na + x2 - c2 + (warp ? s_w2[warp - 1] : 0);
    for (int i = i0; i < i1; ++i) {
        const int f = flag[i];
        if (f == 1) out[p1++] = world[i];
        else if (f == 2) out[p2++] = world[i];
    }
    if (tid == 0) { counts[0] = na; counts[1] = nd; counts[2] = na + nd; }
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
    DevBuf<float4> d_scan, d_nb;
    DevBuf<uint8_t> d_raw;  // unpacked scan bytes when they come straight from page-locked caller memory
    DevBuf<unsigned char> d_nbc;
    int last_n = 0;       // size the per-point arrays were last resized to
    const float4* last_scan = nullptr;
    PinnedBuf<uint8_t> h_stage;  // [Ctl header | float4 points]
    PinnedBuf<Ctl> h_out;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // the kernel sequence of one update is captured once into a CUDA graph and replayed (the control flow
    // lives on the device, so the sequence never changes); re-captured only when a buffer moves
    struct GraphKey {
        const void *pts, *hdr, *ent, *pool, *plane, *nb, *nbc;   // every pointer baked into the captured launches
        unsigned search_grid;
        int single, force, knn_mode;
        bool operator==(const GraphKey& o) const {
            return pts == o.pts && hdr == o.hdr && ent == o.ent && pool == o.pool && plane == o.plane && nb == o.nb && nbc == o.nbc &&
                   search_grid == o.search_grid && single == o.single && force == o.force && knn_mode == o.knn_mode;
        }
    };
    GraphKey gkey{}, pending_key{};
    cudaGraphExec_t gexec = nullptr;
    int graph_captures = 0, graph_replays = 0, direct_runs = 0;
    int use_graph = 1;
    int profiling = 0;  // record an event after every kernel (disables the graph path)
    cudaEvent_t evk[2 * B200_MAX_PASSES + 2] = {};
    float kernel_ms[2 * B200_MAX_PASSES + 1] = {};
    int n_kernels = 0;
    int32_t enqueue(const float4* d_pts, const Ctl* d_hdr, unsigned search_grid, int single_pass, int force_converge, bool events);
    // map-incremental scratch
    DevBuf<float4> d_world, d_sel_pts;
    DevBuf<uint8_t> d_flag, d_flag2, cub_tmp;
    DevBuf<int32_t> d_count;
    DevBuf<double> d_x;
    PinnedBuf<int32_t> h_count;
    PinnedBuf<double> h_x, h_pred;
    DevBuf<double> d_pred;

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
    // active accumulator columns: [0] without extrinsic estimation (leading 6 x 6 block), [1] with (all 12 x 12)
    unsigned char ca[2][NPART], cb[2][NPART];
    memset(ca, 255, sizeof ca);
    memset(cb, 255, sizeof cb);
    for (int mode = 0; mode < 2; ++mode) {
        const int m = mode ? 12 : 6;
        int c = 0;
        for (int i = 0; i < m; ++i)
            for (int j = i; j < m; ++j) { ca[mode][c] = (unsigned char)i; cb[mode][c] = (unsigned char)j; ++c; }
        for (int i = 0; i < m; ++i) { ca[mode][c] = (unsigned char)i; cb[mode][c] = 12; ++c; }
        // column c (= 27 or 90) stays (255, 255): the effective-point counter
    }
    cudaMemcpyToSymbol(c_col_a, ca, sizeof ca);
    cudaMemcpyToSymbol(c_col_b, cb, sizeof cb);
}

int32_t Iekf::init(const b200_iekf_params* p, Map* m) {
    prm = *p;
    map = m;
    if (prm.max_iter < 1 || prm.max_iter + 1 > B200_MAX_PASSES) B200_FAIL(B200_ERR_ARG, "max_iter must be in [1, 7]");
    CUDA_SET_DEVICE(m->device);
    stream = m->stream;  // one stream per map/filter pair: inserts and updates are naturally ordered
    init_pair_tables();
    CUDA_TRY(cudaMalloc(&d_ctl, sizeof(Ctl)));
    CUDA_TRY(cudaMemset(d_ctl, 0, sizeof(Ctl)));
    CUDA_TRY(cudaMalloc(&d_limit, sizeof(double) * NS));
    CUDA_TRY(cudaMemcpy(d_limit, prm.limit, sizeof(double) * NS, cudaMemcpyHostToDevice));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, m->device));
    nblocks = prop.multiProcessorCount;
    if (nblocks > MAXB) nblocks = MAXB;
    CUDA_TRY(cudaMalloc(&d_partials, sizeof(double) * NPART * nblocks));
    CUDA_TRY(h_out.reserve(1));
    CUDA_TRY(cudaEventCreate(&ev0));
    CUDA_TRY(cudaEventCreate(&ev1));
    for (auto& e : evk) CUDA_TRY(cudaEventCreate(&e));
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
    d_scan.release(); d_raw.release(); d_nb.release(); d_nbc.release(); h_stage.release(); h_out.release();
    d_world.release(); d_sel_pts.release(); d_flag.release(); d_flag2.release(); cub_tmp.release(); d_count.release(); d_x.release();
    h_count.release(); h_x.release(); h_pred.release(); d_pred.release();
    if (gexec) cudaGraphExecDestroy(gexec);
    for (auto& e : evk) if (e) cudaEventDestroy(e);
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
int32_t Iekf::enqueue(const float4* d_pts, const Ctl* d_hdr, unsigned search_grid, int single_pass, int force_converge, bool events) {
    int e = 0;
    if (events) CUDA_TRY(cudaEventRecord(evk[e++], stream));
    k_iekf_init<<<8, 256, 0, stream>>>(d_ctl, d_hdr, ps, force_converge);
    if (events) CUDA_TRY(cudaEventRecord(evk[e++], stream));
    const MapView mv = map->view();
    const int npass = single_pass ? 1 : prm.max_iter + 1;
    for (int it = 0; it < npass; ++it) {
        const int mode = map->knn_mode();
        const bool pdl = !events;  // kernel -> kernel edges only (an event record in between makes it a full dependency anyway)
        if (mode == 8) {
            CUDA_TRY(launch_k(k_search_p, dim3(search_grid * KNN_BLOCK / 256), dim3(256), stream, pdl, mv, d_pts, (const Ctl*)d_ctl, d_nb.p, d_nbc.p));
        } else if (mode == 9) {
            CUDA_TRY(launch_k(k_search_t, dim3(search_grid * (32 / KNN_G) * KNN_BLOCK / 256), dim3(256), stream, pdl, mv, d_pts, (const Ctl*)d_ctl, d_nb.p, d_nbc.p));
        } else if (mode == 7) {
            CUDA_TRY(launch_k(k_search_w, dim3(search_grid * (32 / KNN_G) * KNN_BLOCK / 256), dim3(256), stream, pdl, mv, d_pts, (const Ctl*)d_ctl, d_nb.p, d_nbc.p));
        } else {
            auto ks = mode == 5 ? k_search<5> : mode == 6 ? k_search<6> : mode == 4 ? k_search<4> : mode == 1 ? k_search<1> : k_search<0>;
            CUDA_TRY(launch_k(ks, dim3(search_grid), dim3(KNN_BLOCK), stream, pdl, mv, d_pts, (const Ctl*)d_ctl, d_nb.p, d_nbc.p));
        }
        if (events) CUDA_TRY(cudaEventRecord(evk[e++], stream));
        CUDA_TRY(launch_k(k_obs, dim3(nblocks), dim3(OBS_THREADS), stream, pdl, d_pts, (const float4*)d_nb.p, (const unsigned char*)d_nbc.p, ps,
                          d_ctl, prm.plane_thr, (int)prm.extrinsic_est_en, d_partials, (int)prm.max_iter, prm.R, (const double*)d_limit,
                          single_pass));
        if (events) CUDA_TRY(cudaEventRecord(evk[e++], stream));
    }
    n_kernels = events ? e - 1 : 0;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}

int32_t Iekf::run(const float4* d_pts, int n, const Ctl* d_hdr, double* x, double* P, b200_iekf_stats* st, int single_pass,
                  int force_converge) {
    CUDA_SET_DEVICE(map->device);
    int32_t rc = ensure_points((size_t)n);
    if (rc) return rc;
    CUDA_TRY(d_nb.reserve((size_t)n * 5));
    CUDA_TRY(d_nbc.reserve((size_t)n));
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
    // search grid sized for the scan rounded up to 8192 points (blocks past ctl->n exit at once), so scans of
    // similar size replay the same graph
    const unsigned search_grid = (unsigned)((((size_t)n + 8191) / 8192 * 8192 * KNN_G + KNN_BLOCK - 1) / KNN_BLOCK);
    const int npass = single_pass ? 1 : prm.max_iter + 1;
    CUDA_TRY(cudaEventRecord(ev0, stream));
    if (profiling || !use_graph) {
        rc = enqueue(d_pts, d_hdr, search_grid, single_pass, force_converge, profiling != 0);
        if (rc) return rc;
    } else {
        const MapView mv0 = map->view();
        GraphKey key{d_pts, d_hdr, mv0.ent, mv0.pool, ps.plane, d_nb.p, d_nbc.p, search_grid, single_pass, force_converge, map->knn_mode()};
        if (gexec && key == gkey) {
            CUDA_TRY(cudaGraphLaunch(gexec, stream));
            ++graph_replays;
        } else if (key == pending_key) {  // the same buffers twice in a row: worth a graph from now on
            if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
            cudaGraph_t graph = nullptr;
            CUDA_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            rc = enqueue(d_pts, d_hdr, search_grid, single_pass, force_converge, false);
            cudaError_t ce = cudaStreamEndCapture(stream, &graph);
            if (rc) return rc;
            CUDA_TRY(ce);
            CUDA_TRY(cudaGraphInstantiate(&gexec, graph, 0));
            cudaGraphDestroy(graph);
            gkey = key;
            ++graph_captures;
            CUDA_TRY(cudaGraphLaunch(gexec, stream));
        } else {  // a buffer moved (map grew, scan size class changed): plain launches, no instantiation cost
            pending_key = key;
            rc = enqueue(d_pts, d_hdr, search_grid, single_pass, force_converge, false);
            if (rc) return rc;
            ++direct_runs;
        }
    }
    LAUNCH_COUNT(1 + 2 * npass);
    CUDA_TRY(cudaEventRecord(ev1, stream));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_out.p, d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    last_n = n;
    last_scan = d_pts;
    const Ctl& o = *h_out.p;
    if (o.fault) B200_FAIL(B200_ERR_CUDA, "IEKF update: the measurement blocks did not report within the watchdog period");
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
    for (int i = 0; i < n_kernels; ++i) cudaEventElapsedTime(&kernel_ms[i], evk[i], evk[i + 1]);
    return status;
}

}  // namespace b200

// ------------------------------------------------------------------ C ABI (B2)
using namespace b200;
struct b200_iekf { Iekf k; b200_map* owner = nullptr; };

extern "C" {

int32_t b200_host_alloc(size_t bytes, void** out) {
    if (!out || bytes == 0) B200_FAIL(B200_ERR_ARG, "bad argument");
    CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return B200_OK;
}
int32_t b200_host_free(void* p) {
    if (p) CUDA_TRY(cudaFreeHost(p));
    return B200_OK;
}

int32_t b200_iekf_create(const b200_iekf_params* params, b200_map* map, b200_iekf** out) {
    if (!params || !map || !out) B200_FAIL(B200_ERR_ARG, "null argument");
    b200_iekf* h = new b200_iekf();
    int32_t rc = h->k.init(params, &map->m);
    if (rc != B200_OK) { h->k.destroy(); delete h; return rc; }
    h->owner = map;
    ++map->refs;  // the filter shares the map's stream and tables: the map outlives it even if destroyed first
    *out = h;
    return B200_OK;
}
int32_t b200_iekf_destroy(b200_iekf* ekf) {
    if (!ekf) return B200_OK;
    ekf->k.destroy();
    b200_map* map = ekf->owner;
    delete ekf;
    if (map && --map->refs == 0 && map->zombie) {
        map->m.destroy();
        delete map;
    }
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
    // page-locked caller memory: the raw records cross PCIe as they are and are unpacked on the device
    cudaPointerAttributes at{};
    if (stride % 4 == 0 && cudaPointerGetAttributes(&at, xyz) == cudaSuccess && at.type == cudaMemoryTypeHost) {
        const size_t raw_bytes = (size_t)(n - 1) * stride + 12;
        CUDA_TRY(k.d_raw.reserve(raw_bytes));
        CUDA_TRY(cudaMemcpyAsync(k.d_scan.p, k.h_stage.p, hdr_pad, cudaMemcpyHostToDevice, k.stream));
        CUDA_TRY(cudaMemcpyAsync(k.d_raw.p, xyz, raw_bytes, cudaMemcpyHostToDevice, k.stream));
        k_unpack_xyz<<<(unsigned)((n + 255) / 256), 256, 0, k.stream>>>(k.d_raw.p, stride, (int)n,
                                                                         (float4*)((uint8_t*)k.d_scan.p + hdr_pad));
        LAUNCH_COUNT(1);
        return B200_OK;
    }
    cudaGetLastError();  // cudaPointerGetAttributes on plain malloc memory may leave an error behind on old drivers
    pack_xyz_float4(xyz, n, stride, (float4*)(k.h_stage.p + hdr_pad));
    // the device scan buffer mirrors the pinned layout; k_iekf_init forwards the header to the control block
    CUDA_TRY(cudaMemcpyAsync(k.d_scan.p, k.h_stage.p, hdr_pad + (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, k.stream));
    return B200_OK;
}

int32_t b200_iekf_update(b200_iekf* ekf, const float* scan, int64_t n, int64_t stride, double* x26, double* P, b200_iekf_stats* stats) {
    if (!ekf || !scan || !x26 || !P || n < 1 || stride < 12 || n > (1 << 26)) B200_FAIL(B200_ERR_ARG, "bad argument");
    Iekf& k = ekf->k;
    static const bool timing = getenv("B200_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    CUDA_SET_DEVICE(k.map->device);
    int32_t rc = stage_scan(k, scan, n, stride, x26, P);
    if (rc) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    const size_t hdr_pad = (offsetof(Ctl, x_prop) + 255) / 256 * 256;
    rc = k.run((const float4*)((const uint8_t*)k.d_scan.p + hdr_pad), (int)n, (const Ctl*)k.d_scan.p, x26, P, stats, 0, 1);
    if (timing) {
        const auto t2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[b200_iekf_update] stage (pack + H2D enqueue) %.1f us, run (launch + kernels + D2H + sync) %.1f us, device %.1f us\n",
                std::chrono::duration<double, std::micro>(t1 - t0).count(), std::chrono::duration<double, std::micro>(t2 - t1).count(),
                stats ? stats->gpu_ms * 1e3 : 0.0);
    }
    return rc;
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

/* profiling aid: on = record a CUDA event after every kernel of an update (and launch without the graph) */
int32_t b200_iekf_set_profiling(b200_iekf* ekf, int32_t on) {
    if (!ekf) B200_FAIL(B200_ERR_ARG, "bad argument");
    ekf->k.profiling = on;
    return B200_OK;
}
int32_t b200_iekf_set_graph(b200_iekf* ekf, int32_t on) {
    if (!ekf) B200_FAIL(B200_ERR_ARG, "bad argument");
    ekf->k.use_graph = on;
    return B200_OK;
}
/* durations (ms) of the kernels of the last profiled update: init, then (search, obs) per pass */
int32_t b200_iekf_kernel_times(b200_iekf* ekf, float* ms, int32_t max) {
    if (!ekf || !ms) B200_FAIL(B200_ERR_ARG, "bad argument");
    int n = ekf->k.n_kernels < max ? ekf->k.n_kernels : max;
    for (int i = 0; i < n; ++i) ms[i] = ekf->k.kernel_ms[i];
    return n;
}

/* bytes moved per b200_iekf_update call for an n-point scan (bench.py "e2e") */
int32_t b200_iekf_io_bytes(b200_iekf* ekf, int64_t n, int64_t* h2d, int64_t* d2h) {
    if (!ekf) B200_FAIL(B200_ERR_ARG, "bad argument");
    const size_t hdr_pad = (offsetof(Ctl, x_prop) + 255) / 256 * 256;
    if (h2d) *h2d = (int64_t)(hdr_pad + (size_t)n * sizeof(float4));
    if (d2h) *d2h = (int64_t)sizeof(Ctl);
    return B200_OK;
}

/* how the updates were launched so far: graph captures / graph replays / plain launch sequences */
int32_t b200_iekf_launch_modes(b200_iekf* ekf, int32_t* captures, int32_t* replays, int32_t* direct) {
    if (!ekf) B200_FAIL(B200_ERR_ARG, "bad argument");
    if (captures) *captures = ekf->k.graph_captures;
    if (replays) *replays = ekf->k.graph_replays;
    if (direct) *direct = ekf->k.direct_runs;
    return B200_OK;
}

int32_t b200_iekf_debug_stamps(b200_iekf* ekf, long long* out /*[B200_MAX_PASSES*16]*/) {
    if (!ekf || !out) B200_FAIL(B200_ERR_ARG, "bad argument");
    memcpy(out, ekf->k.h_out.p->dbg, sizeof(long long) * B200_MAX_PASSES * 16);
    return B200_OK;
}

int32_t b200_iekf_obs_model(b200_iekf* ekf, const float* scan, int64_t n, int64_t stride, const double* x26, int32_t converge,
                            double* HtH, double* Hth, int32_t* n_eff) {
    if (!ekf || !scan || !x26 || n < 1 || stride < 12) B200_FAIL(B200_ERR_ARG, "bad argument");
    Iekf& k = ekf->k;
    CUDA_SET_DEVICE(k.map->device);
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
    CUDA_SET_DEVICE(k.map->device);
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

/* esekf::predict over K IMU intervals on the device (see k_iekf_predict) */
int32_t b200_iekf_predict(b200_iekf* ekf, const double* steps8, int32_t K, const double* Q12, double* x26, double* P, double* poses22) {
    if (!ekf || !steps8 || !Q12 || !x26 || !P || K < 1 || K > 4096) B200_FAIL(B200_ERR_ARG, "bad argument");
    Iekf& k = ekf->k;
    CUDA_SET_DEVICE(k.map->device);
    const size_t n_in = (size_t)K * 8 + 12 + 26 + NS * NS, n_out = 26 + NS * NS + (size_t)K * 22;
    CUDA_TRY(k.h_pred.reserve(n_in + n_out));
    CUDA_TRY(k.d_pred.reserve(n_in + (size_t)K * 22));
    double* h = k.h_pred.p;
    memcpy(h, steps8, sizeof(double) * K * 8);
    memcpy(h + (size_t)K * 8, Q12, sizeof(double) * 12);
    memcpy(h + (size_t)K * 8 + 12, x26, sizeof(double) * 26);
    memcpy(h + (size_t)K * 8 + 38, P, sizeof(double) * NS * NS);
    double* d = k.d_pred.p;
    CUDA_TRY(cudaMemcpyAsync(d, h, n_in * sizeof(double), cudaMemcpyHostToDevice, k.stream));
    double *d_steps = d, *d_Q = d + (size_t)K * 8, *d_x = d_Q + 12, *d_P = d_x + 26, *d_poses = d + n_in;
    k_iekf_predict<<<1, 544, 0, k.stream>>>(d_steps, K, d_Q, d_x, d_P, d_poses);
    LAUNCH_COUNT(1);
    double* ho = h + n_in;
    CUDA_TRY(cudaMemcpyAsync(ho, d_x, (26 + NS * NS) * sizeof(double), cudaMemcpyDeviceToHost, k.stream));
    CUDA_TRY(cudaMemcpyAsync(ho + 26 + NS * NS, d_poses, (size_t)K * 22 * sizeof(double), cudaMemcpyDeviceToHost, k.stream));
    CUDA_TRY(cudaStreamSynchronize(k.stream));
    CUDA_TRY(cudaGetLastError());
    memcpy(x26, ho, sizeof(double) * 26);
    memcpy(P, ho + 26, sizeof(double) * NS * NS);
    if (poses22) memcpy(poses22, ho + 26 + NS * NS, sizeof(double) * K * 22);
    return B200_OK;
}

/* the last scan in the world frame at state x (laserCloudWorld of PublishFrameWorld, laser_mapping.cc:747-773) */
int32_t b200_iekf_world_scan(b200_iekf* ekf, const double* x26, float* out_xyz, int64_t stride_bytes, int64_t max_points, int64_t* n_out) {
    if (!ekf || !x26 || !out_xyz || stride_bytes < 12 || stride_bytes % 4) B200_FAIL(B200_ERR_ARG, "bad argument");
    Iekf& k = ekf->k;
    const int n = k.last_n;
    if (n < 1 || !k.last_scan) B200_FAIL(B200_ERR_ARG, "no scan has been processed");
    if (max_points < n) B200_FAIL(B200_ERR_ARG, "output buffer too small");
    CUDA_SET_DEVICE(k.map->device);
    CUDA_TRY(k.d_raw.reserve((size_t)n * stride_bytes));
    memcpy(k.h_x.p, x26, sizeof(double) * 26);
    CUDA_TRY(cudaMemcpyAsync(k.d_x.p, k.h_x.p, sizeof(double) * 26, cudaMemcpyHostToDevice, k.stream));
    CUDA_TRY(cudaMemsetAsync(k.d_raw.p, 0, (size_t)n * stride_bytes, k.stream));
    k_body_to_world<<<(n + 255) / 256, 256, 0, k.stream>>>(k.last_scan, n, k.d_x.p, k.d_raw.p, stride_bytes);
    LAUNCH_COUNT(1);
    CUDA_TRY(cudaMemcpyAsync(out_xyz, k.d_raw.p, (size_t)n * stride_bytes, cudaMemcpyDeviceToHost, k.stream));
    CUDA_TRY(cudaStreamSynchronize(k.stream));
    CUDA_TRY(cudaGetLastError());
    if (n_out) *n_out = n;
    return B200_OK;
}

int32_t b200_iekf_map_incremental(b200_iekf* ekf, const double* x26, int32_t ekf_inited, int32_t* n_added, int32_t* n_no_downsample) {
    if (!ekf || !x26) B200_FAIL(B200_ERR_ARG, "bad argument");
    Iekf& k = ekf->k;
    const int n = k.last_n;
    if (n < 1 || !k.last_scan) B200_FAIL(B200_ERR_ARG, "no scan has been processed");
    CUDA_SET_DEVICE(k.map->device);
    CUDA_TRY(k.d_world.reserve(n)); CUDA_TRY(k.d_sel_pts.reserve(2 * (size_t)n));
    CUDA_TRY(k.d_flag.reserve(n)); CUDA_TRY(k.d_flag2.reserve(n));
    memcpy(k.h_x.p, x26, sizeof(double) * 26);
    CUDA_TRY(cudaMemcpyAsync(k.d_x.p, k.h_x.p, sizeof(double) * 26, cudaMemcpyHostToDevice, k.stream));
    const int nb = (n + 255) / 256;
    k_map_incremental_flags<<<nb, 256, 0, k.stream>>>(k.last_scan, n, k.d_x.p, k.ps, ekf_inited, k.prm.filter_size_map, k.d_world.p, k.d_flag.p);
    // points_to_add first, then point_no_need_downsample, each in scan order (:579-580); the counts stay on the device: the
    // insert is enqueued right behind on the scan's size with the real count read by its kernels, so the whole
    // MapIncremental costs one stream synchronisation (the one at the end of the insert)
    k_partition3<<<1, 1024, 0, k.stream>>>(k.d_flag.p, k.d_world.p, n, k.d_sel_pts.p, k.d_count.p);
    LAUNCH_COUNT(2);
    CUDA_TRY(cudaMemcpyAsync(k.h_count.p, k.d_count.p, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, k.stream));
    const int32_t rc = k.map->insert_device(k.d_sel_pts.p, (int64_t)n, k.d_count.p + 2, k.h_count.p + 2);
    if (n_added) *n_added = k.h_count.p[0];
    if (n_no_downsample) *n_no_downsample = k.h_count.p[1];
    return rc;
}

}  // extern "C"
