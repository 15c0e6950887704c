// b200reg — Generalized ICP on the device.  Replaces pclomp::GeneralizedIterativeClosestPoint
// (pointcloud_match/ndt_omp/include/pclomp/gicp_omp.h, gicp_omp_impl.hpp; selected as "GICP_OMP" by
// jueying_slam/src/localization.cpp:175-177 and driven through pcl::Registration::align at :323-328).
//
// exact k-NN index  the role of pcl::search::KdTree (nearestKSearch: exact, ascending distance).  The cloud is sorted by the
//                   cell of a dense uniform grid (cell size chosen from the cloud's own density, about 4-8 points per occupied
//                   cell); a warp searches cube shells of growing radius around the query until no unexplored cell can hold
//                   a closer point than its current k-th best.  One table load resolves a cell to its run of points, the
//                   32 lanes read a run 32 points at a time (float4, coalesced).
// k_g_knn_cov       computeCovariances (:49-123): the k = 20 nearest neighbours of every point of a cloud in a lane-resident
//                   sorted list (lane j = j-th best, 64-bit keys: fp32 distance bits << 32 | point index, so ties go to
//                   the lower index), sums in the neighbours' order in fp64, JacobiSVD of the 3x3 covariance in registers,
//                   singular values replaced by (1, 1, gicp_epsilon).
// k_g_correspond    the correspondence half of computeTransformation's loop (:408-470): exact nearest target point of every
//                   moved source point, distance gate, Mahalanobis matrix (R C1 R^T + C2)^-1 in fp64 -> ONE 64-byte record
//                   per source point {target xyz, matched flag, float 3x3}.
// k_g_bfgs          estimateRigidTransformationBFGS + the cost functor (:188-368) + the convergence test (:483-506) in ONE
//                   thread-block cluster of 8 CTAs: every thread runs the scalar BFGS / line-search logic of gicp_math.cuh on
//                   identical reduced sums; a functor evaluation is a reduction over the correspondence records (L2
//                   resident) split over the cluster, combined through distributed shared memory behind one hardware
//                   cluster barrier, so an outer iteration needs no host round trip and the host only polls a 200-byte
//                   control block.
#include "common.cuh"
#include "gicp_math.cuh"

#include <cooperative_groups.h>
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace cg = cooperative_groups;

namespace b200 {
namespace gicp {

constexpr int kMaxK = 32;
constexpr int kBfgsThreads = 512;

struct IndexView {
    const float4* sorted;  // x, y, z, original index (int bits), in cell order
    const int32_t* cell2run;
    const int32_t* run_off;
    const int32_t* run_cnt;
    int min_b[3], div_b[3];
    float leaf, inv_leaf;
};

struct __align__(64) Corr {  // one source point's correspondence
    float tx, ty, tz;
    int32_t tgt;  // matched target point, -1 = none
    float M[9];
    float d2;
    float pad[2];
};
static_assert(sizeof(Corr) == 64, "correspondence record");

struct GCtl {  // control block of one align, lives on the device; the host keeps a pinned mirror
    float T[12];      // transformation_
    float prev[12];   // previous_transformation_
    float guess[12];
    double rotation_epsilon, transformation_epsilon, delta, last_f;
    int32_t max_iterations, max_inner_iterations;
    int32_t nr_iterations, converged, failed;
    int32_t last_m, last_inner, last_status, inner_total, n_f, n_df, n_fdf;
};

// ------------------------------------------------------------------ index build
__device__ __forceinline__ int ordered_int(float f) {
    const int a = __float_as_int(f);
    return a >= 0 ? a : a ^ 0x7fffffff;
}
__global__ void k_g_bbox_init(int* mm) {
    if (threadIdx.x < 3) mm[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) mm[threadIdx.x] = (int)0x80000000;
}
__global__ void __launch_bounds__(256) k_g_bbox(const float4* __restrict__ pts, int64_t n, int* __restrict__ mm) {
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 p = __ldg(pts + i);
        if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) continue;
        const int v[3] = {ordered_int(p.x), ordered_int(p.y), ordered_int(p.z)};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            lo[a] = min(lo[a], v[a]);
            hi[a] = max(hi[a], v[a]);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(mm + a, lo[a]);
            atomicMax(mm + 3 + a, hi[a]);
        }
    }
}
struct GridSpec {
    int min_b[3], div_b[3];
    float inv_leaf;
};
__global__ void __launch_bounds__(256) k_g_keys(const float4* __restrict__ pts, int n, GridSpec g, uint32_t sentinel, uint32_t* __restrict__ keys,
                                                int32_t* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = __ldg(pts + i);
    uint32_t key = sentinel;
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        const int cx = (int)floorf(p.x * g.inv_leaf) - g.min_b[0], cy = (int)floorf(p.y * g.inv_leaf) - g.min_b[1],
                  cz = (int)floorf(p.z * g.inv_leaf) - g.min_b[2];
        if ((unsigned)cx < (unsigned)g.div_b[0] && (unsigned)cy < (unsigned)g.div_b[1] && (unsigned)cz < (unsigned)g.div_b[2])
            key = ((uint32_t)cz * (uint32_t)g.div_b[1] + (uint32_t)cy) * (uint32_t)g.div_b[0] + (uint32_t)cx;
    }
    keys[i] = key;
    vals[i] = i;
}
__global__ void k_g_cell_runs(const uint32_t* __restrict__ uniq, const int32_t* __restrict__ nruns, uint32_t sentinel, int32_t* __restrict__ cell2run) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *nruns) return;
    if (uniq[r] != sentinel) cell2run[uniq[r]] = r;
}
__global__ void k_g_gather(const float4* __restrict__ pts, const int32_t* __restrict__ order, int n, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = __ldg(order + i);
    const float4 p = __ldg(pts + j);
    out[i] = make_float4(p.x, p.y, p.z, __int_as_float(j));
}

// ------------------------------------------------------------------ shell walk helpers
// lower bound on the distance from (x, y, z) to anything outside the explored box [lo, hi] (cells); faces on the grid boundary
// have nothing behind them.  all_grid = the box is the whole grid.
__device__ __forceinline__ float outside_bound(const IndexView& ix, const int* lo, const int* hi, const float* qv, bool& all_grid) {
    float lb = 3.402823466e+38f;
    all_grid = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (lo[a] > 0) {
            all_grid = false;
            lb = fminf(lb, qv[a] - (float)(lo[a] + ix.min_b[a]) * ix.leaf);
        }
        if (hi[a] < ix.div_b[a] - 1) {
            all_grid = false;
            lb = fminf(lb, (float)(hi[a] + 1 + ix.min_b[a]) * ix.leaf - qv[a]);
        }
    }
    return lb - 1e-4f * ix.leaf;  // a small margin absorbs the rounding of the face coordinates
}
__device__ __forceinline__ void start_cell(const IndexView& ix, float x, float y, float z, int* c) {
    c[0] = (int)floorf(x * ix.inv_leaf) - ix.min_b[0];
    c[1] = (int)floorf(y * ix.inv_leaf) - ix.min_b[1];
    c[2] = (int)floorf(z * ix.inv_leaf) - ix.min_b[2];
#pragma unroll
    for (int a = 0; a < 3; ++a) c[a] = min(max(c[a], 0), ix.div_b[a] - 1);
}
__device__ __forceinline__ unsigned long long dist_key(float d2, int idx) {
    return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)(uint32_t)idx;
}

// ------------------------------------------------------------------ computeCovariances
// One warp per point of the cloud.  list (lanes 0..k-1) = the k best keys so far, ascending.
__global__ void __launch_bounds__(256) k_g_knn_cov(IndexView ix, const float4* __restrict__ pts, int n, int k, double gicp_epsilon,
                                                   double* __restrict__ cov_out, int32_t* __restrict__ knn_out) {
    const int lane = threadIdx.x & 31;
    const int q = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (q >= n) return;
    const float4 p = __ldg(pts + q);
    const float x = p.x, y = p.y, z = p.z;
    unsigned long long list = ~0ull;
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
        int c[3];
        start_cell(ix, x, y, z, c);
        const float qv[3] = {x, y, z};
        const int rmax = max(max(ix.div_b[0], ix.div_b[1]), ix.div_b[2]);
        for (int r = 0; r <= rmax; ++r) {
            int lo[3], hi[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                lo[a] = max(c[a] - r, 0);
                hi[a] = min(c[a] + r, ix.div_b[a] - 1);
            }
            for (int iz = lo[2]; iz <= hi[2]; ++iz) {
                const bool ez = abs(iz - c[2]) == r;
                for (int iy = lo[1]; iy <= hi[1]; ++iy) {
                    const bool face = ez || abs(iy - c[1]) == r;
                    const int step = face ? 1 : max(2 * r, 1);  // interior rows of the shell: only the two end cells
                    for (int jx = c[0] - r; jx <= c[0] + r; jx += step) {
                        if (jx < 0 || jx >= ix.div_b[0]) continue;
                        const int run = __ldg(ix.cell2run + ((size_t)iz * ix.div_b[1] + iy) * ix.div_b[0] + jx);
                        if (run < 0) continue;
                        const int off = __ldg(ix.run_off + run), cnt = __ldg(ix.run_cnt + run);
                        for (int j0 = 0; j0 < cnt; j0 += 32) {
                            const int j = j0 + lane;
                            unsigned long long key = ~0ull;
                            if (j < cnt) {
                                const float4 t = __ldg(ix.sorted + off + j);
                                const float dx = x - t.x, dy = y - t.y, dz = z - t.z;
                                key = dist_key((dx * dx + dy * dy) + dz * dz, __float_as_int(t.w));  // FLANN L2_Simple order
                            }
                            unsigned long long kth = __shfl_sync(0xffffffffu, list, k - 1);
                            unsigned mask = __ballot_sync(0xffffffffu, key < kth);
                            while (mask) {
                                const int src = __ffs(mask) - 1;
                                mask &= mask - 1;
                                const unsigned long long cand = __shfl_sync(0xffffffffu, key, src);
                                kth = __shfl_sync(0xffffffffu, list, k - 1);
                                if (cand >= kth) continue;
                                const int pos = __popc(__ballot_sync(0xffffffffu, lane < k && list < cand));
                                const unsigned long long up = __shfl_up_sync(0xffffffffu, list, 1);
                                if (lane < k) {
                                    if (lane > pos) list = up;
                                    else if (lane == pos) list = cand;
                                }
                            }
                        }
                    }
                }
            }
            bool all_grid;
            const float lb = outside_bound(ix, lo, hi, qv, all_grid);
            if (all_grid) break;
            const unsigned long long kth = __shfl_sync(0xffffffffu, list, k - 1);
            if (kth != ~0ull && lb > 0.f && __uint_as_float((uint32_t)(kth >> 32)) < lb * lb) break;
        }
    }
    // sums over the neighbours in list order (gicp_omp_impl.hpp:78-95): float products, fp64 accumulation
    const int my = (lane < k && list != ~0ull) ? (int)(uint32_t)list : -1;
    if (knn_out && lane < k) knn_out[(size_t)q * k + lane] = my;
    float4 nb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (my >= 0) nb = __ldg(pts + my);
    const unsigned found = __ballot_sync(0xffffffffu, my >= 0);
    double mean[3] = {0.0, 0.0, 0.0}, cs[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int j = 0; j < k; ++j) {
        const float px = __shfl_sync(0xffffffffu, nb.x, j), py = __shfl_sync(0xffffffffu, nb.y, j), pz = __shfl_sync(0xffffffffu, nb.z, j);
        mean[0] += (double)px;
        mean[1] += (double)py;
        mean[2] += (double)pz;
        cs[0] += (double)(px * px);
        cs[1] += (double)(py * px);
        cs[2] += (double)(py * py);
        cs[3] += (double)(pz * px);
        cs[4] += (double)(pz * py);
        cs[5] += (double)(pz * pz);
    }
    double cov[9];
    if (__popc(found) == k) {
        cov_regularize(mean, cs, k, gicp_epsilon, cov);
    } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) cov[i] = 0.0;  // a non-finite point: no covariance
    }
    // lanes 0..8 store one entry each (every lane holds the same nine values)
    double v = cov[0];
#pragma unroll
    for (int i = 1; i < 9; ++i)
        if (lane == i) v = cov[i];
    if (lane < 9) cov_out[(size_t)q * 9 + lane] = v;
}

// ------------------------------------------------------------------ correspondences
__global__ void k_g_transform(const float4* __restrict__ src, int n, const GCtl* __restrict__ ctl, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* G = ctl->guess;
    const float4 p = __ldg(src + i);
    // pcl::transformPointCloud(output, output, guess) (gicp_omp_impl.hpp:400)
    out[i] = make_float4(((G[0] * p.x + G[1] * p.y) + G[2] * p.z) + G[3], ((G[4] * p.x + G[5] * p.y) + G[6] * p.z) + G[7],
                         ((G[8] * p.x + G[9] * p.y) + G[10] * p.z) + G[11], 1.0f);
}

// exact nearest neighbour of (x, y, z): returns the key (distance bits << 32 | index), ~0 when the index is empty.  With
// give_up2 > 0 the search stops once nothing closer than sqrt(give_up2) can exist (the caller rejects such matches anyway).
__device__ __forceinline__ unsigned long long nearest1(const IndexView& ix, float x, float y, float z, int lane, float give_up2) {
    unsigned long long best = ~0ull;
    int c[3];
    start_cell(ix, x, y, z, c);
    const float qv[3] = {x, y, z};
    const int rmax = max(max(ix.div_b[0], ix.div_b[1]), ix.div_b[2]);
    for (int r = 0; r <= rmax; ++r) {
        int lo[3], hi[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            lo[a] = max(c[a] - r, 0);
            hi[a] = min(c[a] + r, ix.div_b[a] - 1);
        }
        for (int iz = lo[2]; iz <= hi[2]; ++iz) {
            const bool ez = abs(iz - c[2]) == r;
            for (int iy = lo[1]; iy <= hi[1]; ++iy) {
                const bool face = ez || abs(iy - c[1]) == r;
                const int step = face ? 1 : max(2 * r, 1);
                for (int jx = c[0] - r; jx <= c[0] + r; jx += step) {
                    if (jx < 0 || jx >= ix.div_b[0]) continue;
                    const int run = __ldg(ix.cell2run + ((size_t)iz * ix.div_b[1] + iy) * ix.div_b[0] + jx);
                    if (run < 0) continue;
                    const int off = __ldg(ix.run_off + run), cnt = __ldg(ix.run_cnt + run);
                    for (int j = lane; j < cnt; j += 32) {
                        const float4 t = __ldg(ix.sorted + off + j);
                        const float dx = x - t.x, dy = y - t.y, dz = z - t.z;
                        const unsigned long long key = dist_key((dx * dx + dy * dy) + dz * dz, __float_as_int(t.w));
                        best = key < best ? key : best;
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other < best ? other : best;
        }
        bool all_grid;
        const float lb = outside_bound(ix, lo, hi, qv, all_grid);
        if (all_grid) break;
        if (lb > 0.f) {
            if (best != ~0ull && __uint_as_float((uint32_t)(best >> 32)) < lb * lb) break;
            if (give_up2 > 0.f && lb * lb >= give_up2) break;
        }
    }
    return best;
}

// One warp per source point (gicp_omp_impl.hpp:422-458).  out = the source moved by the guess; the query is transformation_ * out.
__global__ void __launch_bounds__(256) k_g_correspond(IndexView ix, const float4* __restrict__ out, int n, const float4* __restrict__ tgt,
                                                      const double* __restrict__ cov_src, const double* __restrict__ cov_tgt,
                                                      const GCtl* __restrict__ ctl, double dist_threshold, Corr* __restrict__ corr) {
    if (ctl->converged | ctl->failed) return;
    const int lane = threadIdx.x & 31;
    const int i = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    const float* T = ctl->T;
    const float4 p = __ldg(out + i);
    // Matrix4f * Vector4f with w = 1
    const float x = ((T[0] * p.x + T[1] * p.y) + T[2] * p.z) + T[3] * 1.0f;
    const float y = ((T[4] * p.x + T[5] * p.y) + T[6] * p.z) + T[7] * 1.0f;
    const float z = ((T[8] * p.x + T[9] * p.y) + T[10] * p.z) + T[11] * 1.0f;
    unsigned long long best = ~0ull;
    if (isfinite(x) && isfinite(y) && isfinite(z)) best = nearest1(ix, x, y, z, lane, (float)fmin(dist_threshold * 1.0001 + 1e-6, 3.0e38));
    Corr c;
    c.tgt = -1;
    c.tx = c.ty = c.tz = 0.f;
    c.d2 = 3.402823466e+38f;
    c.pad[0] = c.pad[1] = 0.f;
#pragma unroll
    for (int a = 0; a < 9; ++a) c.M[a] = (a % 4 == 0) ? 1.0f : 0.0f;  // mahalanobis_ starts as the identity (:381)
    if (best != ~0ull) {
        const float d2 = __uint_as_float((uint32_t)(best >> 32));
        const int nn = (int)(uint32_t)best;
        c.d2 = d2;
        if ((double)d2 < dist_threshold) {
            double R[9], C1[9], C2[9];
            rotation_of_product(T, ctl->guess, R);
#pragma unroll
            for (int a = 0; a < 9; ++a) {
                C1[a] = __ldg(cov_src + (size_t)i * 9 + a);
                C2[a] = __ldg(cov_tgt + (size_t)nn * 9 + a);
            }
            mahalanobis3(R, C1, C2, c.M);
            const float4 t = __ldg(tgt + nn);
            c.tx = t.x;
            c.ty = t.y;
            c.tz = t.z;
            c.tgt = nn;
        }
    }
    if (lane == 0) corr[i] = c;
}

// ------------------------------------------------------------------ the optimiser cluster
// One thread-block cluster of kCluster CTAs (one per SM) runs the whole BFGS of a pass.  Every thread of every CTA executes
// the same scalar control flow on the same reduced sums; a functor evaluation is: each CTA adds the terms of its share of the
// correspondence records (strided over the cluster), reduces them to kAcc partial sums in its own shared memory, the cluster
// barrier publishes them, and every CTA adds the kCluster partials in rank order through distributed shared memory - so all
// CTAs hold bit-identical totals and take the same branches.  The partials are double-buffered: one hardware cluster barrier
// per evaluation is enough (a CTA can run at most one evaluation ahead of the slowest reader).
constexpr int kCluster = 8;  // portable cluster size
struct ClusterEval {
    const float4* out;
    const Corr* corr;
    int n, m;
    double (*sh)[kAcc];    // [warps][kAcc] block scratch
    double (*part)[kAcc];  // [2][kAcc] this CTA's partial sums, read by the whole cluster
    double* tot_sh;        // [kAcc]
    int phase;
    __device__ void sums(const double* x, double* tot) {
        cg::cluster_group cluster = cg::this_cluster();
        const unsigned rank = cluster.block_rank(), nblk = cluster.num_blocks();
        float T[12];
        apply_state(x, T);
        double acc[kAcc];
#pragma unroll
        for (int a = 0; a < kAcc; ++a) acc[a] = 0.0;
        for (int i = rank * blockDim.x + threadIdx.x; i < n; i += nblk * blockDim.x) {
            const float4* rec = reinterpret_cast<const float4*>(corr + i);
            const float4 r0 = rec[0];
            if (__float_as_int(r0.w) < 0) continue;
            const float4 r1 = rec[1], r2 = rec[2], r3 = rec[3];
            const float M[9] = {r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x};
            const float4 s = out[i];
            point_terms(T, s.x, s.y, s.z, r0.x, r0.y, r0.z, M, acc);
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
        for (int a = 0; a < kAcc; ++a) {
            double v = acc[a];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) sh[warp][a] = v;
        }
        __syncthreads();
        if (threadIdx.x < kAcc) {
            double v = 0.0;
            for (int w = 0; w < nw; ++w) v += sh[w][threadIdx.x];
            part[phase][threadIdx.x] = v;
        }
        cluster.sync();  // every CTA's partials of this evaluation are visible cluster-wide
        if (threadIdx.x < kAcc) {
            double v = 0.0;
            for (unsigned r = 0; r < nblk; ++r) v += cluster.map_shared_rank(&part[phase][0], r)[threadIdx.x];
            tot_sh[threadIdx.x] = v;
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < kAcc; ++a) tot[a] = tot_sh[a];
        __syncthreads();  // tot_sh / sh may be rewritten by the next evaluation
        phase ^= 1;
    }
    __device__ double f(const double* x) {
        double tot[kAcc];
        sums(x, tot);
        return cost_f(tot, m);
    }
    __device__ void df(const double* x, double* g) {
        double tot[kAcc];
        sums(x, tot);
        cost_gradient(tot, m, x, g);
    }
    __device__ void fdf(const double* x, double& fo, double* g) {
        double tot[kAcc];
        sums(x, tot);
        fo = cost_f_fdf(tot, m);
        cost_gradient(tot, m, x, g);
    }
};

// matched correspondences, counted by the whole cluster; two barriers: publish, and nobody leaves while its count is being read
__device__ int cluster_count_matches(const Corr* corr, int n, int* sh_i, int* cnt_part) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank(), nblk = cluster.num_blocks();
    int c = 0;
    for (int i = rank * blockDim.x + threadIdx.x; i < n; i += nblk * blockDim.x) c += corr[i].tgt >= 0 ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) sh_i[warp] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < nw; ++w) t += sh_i[w];
        *cnt_part = t;
    }
    cluster.sync();
    int tot = 0;
    for (unsigned r = 0; r < nblk; ++r) tot += *cluster.map_shared_rank(cnt_part, r);
    cluster.sync();
    return tot;
}

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kBfgsThreads) k_g_bfgs(const float4* __restrict__ out, const Corr* __restrict__ corr, int n, GCtl* ctl) {
    __shared__ double sh[kBfgsThreads / 32][kAcc];
    __shared__ double part[2][kAcc];
    __shared__ double tot_sh[kAcc];
    __shared__ int sh_i[kBfgsThreads / 32];
    __shared__ int cnt_part;
    if (ctl->converged | ctl->failed) return;  // the same for every CTA of the cluster: nobody is left waiting at a barrier
    cg::cluster_group cluster = cg::this_cluster();
    const bool writer = cluster.block_rank() == 0 && threadIdx.x == 0;
    const int m = cluster_count_matches(corr, n, sh_i, &cnt_part);
    float T[12];
#pragma unroll
    for (int a = 0; a < 12; ++a) T[a] = ctl->T[a];
    const int max_inner = ctl->max_inner_iterations;
    cluster.sync();  // every thread of the cluster has read transformation_ before the writer replaces it
    if (m < 4) {     // NotEnoughPointsException -> the loop breaks, converged_ stays false (:206-211, 496-500)
        if (writer) {
            for (int a = 0; a < 12; ++a) ctl->prev[a] = T[a];
            ctl->last_m = m;
            ctl->failed = 1;
        }
        return;
    }
    ClusterEval ev{out, corr, n, m, sh, part, tot_sh, 0};
    double x[6];
    state_from_transform(T, x);
    int inner = 0, calls[3];
    const int result = minimize_rigid(ev, x, max_inner, &inner, calls);
    const bool ok = result == kNoProgress || result == kSuccess || inner == max_inner;
    float Tn[12];
    apply_state(x, Tn);  // transformation_matrix.setIdentity(); applyState(transformation_matrix, x)
    if (writer) {
        for (int a = 0; a < 12; ++a) ctl->prev[a] = T[a];  // previous_transformation_ = transformation_ (:475)
        ctl->last_m = m;
        ctl->last_inner = inner;
        ctl->last_status = result;
        ctl->inner_total += inner;
        ctl->n_f += calls[0];
        ctl->n_df += calls[1];
        ctl->n_fdf += calls[2];
        if (!ok) {
            ctl->failed = 1;  // SolverDidntConvergeException
        } else {
            const double delta = transform_delta(T, Tn, ctl->rotation_epsilon, ctl->transformation_epsilon);
            for (int a = 0; a < 12; ++a) ctl->T[a] = Tn[a];
            ctl->delta = delta;
            const int it = ctl->nr_iterations + 1;
            ctl->nr_iterations = it;
            if (it >= ctl->max_iterations || delta < 1.0) {
                ctl->converged = 1;
                for (int a = 0; a < 12; ++a) ctl->prev[a] = Tn[a];
            }
        }
    }
    cluster.sync();  // no CTA exits while a peer may still read its partial sums
}

// parity probe: the functor at x on the current correspondences -> {operator(), fdf's f, gradient[6], m}
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kBfgsThreads) k_g_cost(const float4* __restrict__ out, const Corr* __restrict__ corr, int n,
                                                                                          const double* __restrict__ x6, double* __restrict__ res) {
    __shared__ double sh[kBfgsThreads / 32][kAcc];
    __shared__ double part[2][kAcc];
    __shared__ double tot_sh[kAcc];
    __shared__ int sh_i[kBfgsThreads / 32];
    __shared__ int cnt_part;
    cg::cluster_group cluster = cg::this_cluster();
    const int m = cluster_count_matches(corr, n, sh_i, &cnt_part);
    ClusterEval ev{out, corr, n, m, sh, part, tot_sh, 0};
    double x[6], g[6], f1;
    for (int a = 0; a < 6; ++a) x[a] = x6[a];
    const double f0 = ev.f(x);
    ev.fdf(x, f1, g);
    if (cluster.block_rank() == 0 && threadIdx.x == 0) {
        res[0] = f0;
        res[1] = f1;
        for (int a = 0; a < 6; ++a) res[2 + a] = g[a];
        res[8] = (double)m;
    }
    cluster.sync();
}

// getFitnessScore: squared distance of every moved source point to its exact nearest target point
__global__ void __launch_bounds__(256) k_g_nearest_d2(IndexView ix, const float4* __restrict__ src, int n, const float* __restrict__ M12, double* __restrict__ d2_out) {
    const int lane = threadIdx.x & 31;
    const int i = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    const float4 p = __ldg(src + i);
    const float x = ((M12[0] * p.x + M12[1] * p.y) + M12[2] * p.z) + M12[3];
    const float y = ((M12[4] * p.x + M12[5] * p.y) + M12[6] * p.z) + M12[7];
    const float z = ((M12[8] * p.x + M12[9] * p.y) + M12[10] * p.z) + M12[11];
    unsigned long long best = ~0ull;
    if (isfinite(x) && isfinite(y) && isfinite(z)) best = nearest1(ix, x, y, z, lane, 0.f);
    if (lane == 0) d2_out[i] = best != ~0ull ? (double)__uint_as_float((uint32_t)(best >> 32)) : 3.402823466e+38;
}
__global__ void k_g_fitness_reduce(const double* __restrict__ d2, int n, double max_range, double* __restrict__ outv /*sum, count*/) {
    __shared__ double ssum[256], scnt[256];  // single block, fixed order: deterministic
    double s = 0.0, c = 0.0;
    const int per = (n + blockDim.x - 1) / blockDim.x;
    const int b = threadIdx.x * per, e = min(n, b + per);
    for (int i = b; i < e; ++i) {
        const double v = d2[i];
        if (v <= max_range && v < 3.0e38) {
            s += v;
            c += 1.0;
        }
    }
    ssum[threadIdx.x] = s;
    scnt[threadIdx.x] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        double S = 0.0, C = 0.0;
        for (int i = 0; i < (int)blockDim.x; ++i) {
            S += ssum[i];
            C += scnt[i];
        }
        outv[0] = S;
        outv[1] = C;
    }
}
__global__ void k_g_corr_export(const Corr* __restrict__ corr, int n, int32_t* __restrict__ tgt, float* __restrict__ maha9, float* __restrict__ d2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Corr c = corr[i];
    tgt[i] = c.tgt;
    d2[i] = c.d2;
    for (int a = 0; a < 9; ++a) maha9[(size_t)i * 9 + a] = c.M[a];
}

// ------------------------------------------------------------------ host side
struct CloudIndex {
    DevBuf<float4> d_sorted;
    DevBuf<int32_t> d_cell2run, d_run_off, d_run_cnt, d_vals_in, d_vals_out;
    DevBuf<uint32_t> d_keys_in, d_keys_out, d_uniq;
    GridSpec g{};
    float leaf = 0.f;
    int nruns = 0;
    int64_t n = 0, ncells = 0;
    int rounds = 0;
    void release() {
        d_sorted.release(); d_cell2run.release(); d_run_off.release(); d_run_cnt.release(); d_vals_in.release(); d_vals_out.release();
        d_keys_in.release(); d_keys_out.release(); d_uniq.release();
    }
    IndexView view() const {
        IndexView v;
        v.sorted = d_sorted.p; v.cell2run = d_cell2run.p; v.run_off = d_run_off.p; v.run_cnt = d_run_cnt.p;
        for (int a = 0; a < 3; ++a) { v.min_b[a] = g.min_b[a]; v.div_b[a] = g.div_b[a]; }
        v.leaf = leaf;
        v.inv_leaf = g.inv_leaf;
        return v;
    }
};

constexpr int64_t kMaxCells = (int64_t)1 << 27;

struct Gicp {
    b200_gicp_params prm;
    int device = 0, sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DevBuf<float4> d_tgt, d_src, d_out;
    DevBuf<double> d_cov_tgt, d_cov_src, d_scratch;
    DevBuf<int32_t> d_small, d_knn;
    DevBuf<uint8_t> cub_tmp;
    DevBuf<Corr> d_corr;
    DevBuf<GCtl> d_ctl;
    DevBuf<float> d_f;
    PinnedBuf<float4> h_stage;
    PinnedBuf<int32_t> h_small;
    PinnedBuf<GCtl> h_ctl;
    PinnedBuf<double> h_d;
    CloudIndex ix_tgt, ix_src;
    int64_t n_tgt = 0, n_src = 0;
    bool have_tgt = false, have_src = false, have_cov_tgt = false, have_cov_src = false, have_corr = false;
    float final_T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};  // final_transformation_ (3x4 row-major)

    int32_t init(const b200_gicp_params* p, int dev) {
        prm = *p;
        if (prm.k_correspondences <= 0) prm.k_correspondences = 20;
        if (prm.k_correspondences > kMaxK) B200_FAIL(B200_ERR_ARG, "k_correspondences > 32 is not supported");
        if (!(prm.gicp_epsilon > 0)) prm.gicp_epsilon = 0.001;
        if (!(prm.rotation_epsilon > 0)) prm.rotation_epsilon = 2e-3;
        if (!(prm.transformation_epsilon > 0)) prm.transformation_epsilon = 5e-4;
        if (!(prm.corr_dist_threshold > 0)) prm.corr_dist_threshold = 5.0;
        if (prm.max_iterations <= 0) prm.max_iterations = 200;
        if (prm.max_inner_iterations <= 0) prm.max_inner_iterations = 20;
        device = dev;
        CUDA_SET_DEVICE(dev);
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
        sm_count = prop.multiProcessorCount;
        CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreate(&ev0));
        CUDA_TRY(cudaEventCreate(&ev1));
        CUDA_TRY(d_small.reserve(16));
        CUDA_TRY(h_small.reserve(16));
        CUDA_TRY(d_ctl.reserve(1));
        CUDA_TRY(h_ctl.reserve(1));
        CUDA_TRY(d_scratch.reserve(32));
        CUDA_TRY(h_d.reserve(32));
        CUDA_TRY(d_f.reserve(16));
        return B200_OK;
    }
    void destroy() {
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        d_tgt.release(); d_src.release(); d_out.release(); d_cov_tgt.release(); d_cov_src.release(); d_scratch.release();
        d_small.release(); d_knn.release(); cub_tmp.release(); d_corr.release(); d_ctl.release(); d_f.release();
        h_stage.release(); h_small.release(); h_ctl.release(); h_d.release();
        ix_tgt.release(); ix_src.release();
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }

    int32_t upload(const float* xyz, int64_t n, int64_t stride, DevBuf<float4>& dst) {
        CUDA_TRY(dst.reserve((size_t)n));
        CUDA_TRY(h_stage.reserve((size_t)n));
        CUDA_TRY(cudaStreamSynchronize(stream));  // the stage may still feed an earlier copy
        pack_xyz_float4(xyz, n, stride, h_stage.p);
        CUDA_TRY(cudaMemcpyAsync(dst.p, h_stage.p, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, stream));
        return B200_OK;
    }

    // dense-grid index over n device points: bounding box -> cell size from the density -> sort by cell -> runs
    int32_t build_index(const float4* d_pts, int64_t n, CloudIndex& ix) {
        int* mm = d_small.p;
        k_g_bbox_init<<<1, 32, 0, stream>>>(mm);
        k_g_bbox<<<sm_count * 4, 256, 0, stream>>>(d_pts, n, mm);
        LAUNCH_COUNT(2);
        CUDA_TRY(cudaMemcpyAsync(h_small.p, mm, 6 * sizeof(int), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        float mn[3], mx[3];
        for (int a = 0; a < 3; ++a) {
            int lo = h_small.p[a], hi = h_small.p[3 + a];
            lo = lo >= 0 ? lo : lo ^ 0x7fffffff;
            hi = hi >= 0 ? hi : hi ^ 0x7fffffff;
            memcpy(&mn[a], &lo, 4);
            memcpy(&mx[a], &hi, 4);
        }
        if (!(mn[0] <= mx[0])) B200_FAIL(B200_ERR_ARG, "cloud has no finite point");
        float e[3] = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
        std::sort(e, e + 3);
        // first guess: eight points per cell if the cloud were a sheet over the two longest extents of its box
        const double area = std::max((double)e[2] * (double)e[1], 1e-6);
        float leaf = (float)std::max(std::sqrt(8.0 * area / (double)n), (double)e[2] / 1000.0 + 1e-6);
        CUDA_TRY(ix.d_keys_in.reserve(n)); CUDA_TRY(ix.d_keys_out.reserve(n)); CUDA_TRY(ix.d_uniq.reserve(n));
        CUDA_TRY(ix.d_vals_in.reserve(n)); CUDA_TRY(ix.d_vals_out.reserve(n)); CUDA_TRY(ix.d_run_cnt.reserve(n)); CUDA_TRY(ix.d_run_off.reserve(n));
        CUDA_TRY(ix.d_sorted.reserve(n));
        ix.rounds = 0;
        for (int round = 0; round < 6; ++round) {
            GridSpec g;
            int64_t cells;
            for (;;) {
                g.inv_leaf = 1.0f / leaf;
                cells = 1;
                for (int a = 0; a < 3; ++a) {
                    g.min_b[a] = (int)std::floor(mn[a] * g.inv_leaf);
                    g.div_b[a] = (int)std::floor(mx[a] * g.inv_leaf) - g.min_b[a] + 1;
                    cells *= g.div_b[a];
                }
                if (cells <= kMaxCells) break;
                leaf *= 1.3f;
            }
            ix.g = g;
            ix.leaf = leaf;
            ix.ncells = cells;
            ix.n = n;
            CUDA_TRY(ix.d_cell2run.reserve((size_t)cells));
            CUDA_TRY(cudaMemsetAsync(ix.d_cell2run.p, 0xFF, (size_t)cells * sizeof(int32_t), stream));
            const uint32_t sentinel = (uint32_t)cells;
            k_g_keys<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_pts, (int)n, g, sentinel, ix.d_keys_in.p, ix.d_vals_in.p);
            int end_bit = 1;
            while (end_bit < 32 && ((int64_t)1 << end_bit) <= cells) ++end_bit;
            size_t t1 = 0, t2 = 0, t3 = 0;
            int32_t* d_nruns = d_small.p + 8;
            cub::DeviceRadixSort::SortPairs(nullptr, t1, ix.d_keys_in.p, ix.d_keys_out.p, ix.d_vals_in.p, ix.d_vals_out.p, (int)n, 0, end_bit, stream);
            cub::DeviceRunLengthEncode::Encode(nullptr, t2, ix.d_keys_out.p, ix.d_uniq.p, ix.d_run_cnt.p, d_nruns, (int)n, stream);
            cub::DeviceScan::ExclusiveSum(nullptr, t3, ix.d_run_cnt.p, ix.d_run_off.p, (int)n, stream);
            size_t tmp = std::max(t1, std::max(t2, t3));
            CUDA_TRY(cub_tmp.reserve(tmp));
            CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp, ix.d_keys_in.p, ix.d_keys_out.p, ix.d_vals_in.p, ix.d_vals_out.p, (int)n, 0, end_bit, stream));
            CUDA_TRY(cub::DeviceRunLengthEncode::Encode(cub_tmp.p, tmp, ix.d_keys_out.p, ix.d_uniq.p, ix.d_run_cnt.p, d_nruns, (int)n, stream));
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp.p, tmp, ix.d_run_cnt.p, ix.d_run_off.p, (int)n, stream));
            LAUNCH_COUNT(4);
            CUDA_TRY(cudaMemcpyAsync(h_small.p + 8, d_nruns, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            ix.nruns = h_small.p[8];
            ix.rounds = round + 1;
            // fewer than ~4 points per occupied cell: the shells would be mostly table probes; a coarser grid is cheaper
            if ((double)n / std::max(ix.nruns, 1) >= 4.0 || round == 5) break;
            leaf *= 1.6f;
        }
        k_g_cell_runs<<<(unsigned)((ix.nruns + 255) / 256), 256, 0, stream>>>(ix.d_uniq.p, d_small.p + 8, (uint32_t)ix.ncells, ix.d_cell2run.p);
        k_g_gather<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_pts, ix.d_vals_out.p, (int)n, ix.d_sorted.p);
        LAUNCH_COUNT(2);
        CUDA_TRY(cudaGetLastError());
        return B200_OK;
    }

    int32_t set_target(const float* xyz, int64_t n, int64_t stride) {
        if (n < 1 || !xyz || stride < 12 || n > (int64_t)0x7fffff00) B200_FAIL(B200_ERR_ARG, "bad target cloud");
        CUDA_SET_DEVICE(device);
        have_tgt = have_cov_tgt = have_corr = false;  // target_covariances_.reset() (gicp_omp.h:173-178)
        int32_t rc = upload(xyz, n, stride, d_tgt);
        if (rc) return rc;
        n_tgt = n;
        rc = build_index(d_tgt.p, n, ix_tgt);
        if (rc) return rc;
        have_tgt = true;
        return B200_OK;
    }
    int32_t set_source(const float* xyz, int64_t n, int64_t stride) {
        if (n < 1 || !xyz || stride < 12 || n > (int64_t)0x3fffff00) B200_FAIL(B200_ERR_ARG, "bad source cloud");
        CUDA_SET_DEVICE(device);
        have_src = have_cov_src = have_corr = false;  // input_covariances_.reset() (gicp_omp.h:141-157)
        int32_t rc = upload(xyz, n, stride, d_src);
        if (rc) return rc;
        n_src = n;
        rc = build_index(d_src.p, n, ix_src);  // tree_reciprocal_
        if (rc) return rc;
        have_src = true;
        return B200_OK;
    }

    int32_t covariances_of(const CloudIndex& ix, const float4* pts, int64_t n, DevBuf<double>& cov, bool want_knn) {
        if (prm.k_correspondences > n) B200_FAIL(B200_ERR_ARG, "cloud has fewer points than k_correspondences");  // (:54-58)
        CUDA_TRY(cov.reserve((size_t)n * 9));
        if (want_knn) CUDA_TRY(d_knn.reserve((size_t)n * prm.k_correspondences));
        k_g_knn_cov<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, stream>>>(ix.view(), pts, (int)n, prm.k_correspondences, prm.gicp_epsilon, cov.p,
                                                                                  want_knn ? d_knn.p : nullptr);
        LAUNCH_COUNT(1);
        CUDA_TRY(cudaGetLastError());
        return B200_OK;
    }
    int32_t ensure_covariances() {  // (:383-394)
        if (!have_tgt || !have_src) B200_FAIL(B200_ERR_ARG, "set_target and set_source first");
        if (!have_cov_tgt) {
            const int32_t rc = covariances_of(ix_tgt, d_tgt.p, n_tgt, d_cov_tgt, false);
            if (rc) return rc;
            have_cov_tgt = true;
        }
        if (!have_cov_src) {
            const int32_t rc = covariances_of(ix_src, d_src.p, n_src, d_cov_src, false);
            if (rc) return rc;
            have_cov_src = true;
        }
        return B200_OK;
    }

    static void rows_from_colmajor(const float* m16, float* T12) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 4; ++j) T12[i * 4 + j] = m16[j * 4 + i];
    }
    static void colmajor_from_rows(const float* T12, float* m16) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 4; ++j) m16[j * 4 + i] = T12[i * 4 + j];
        m16[3] = m16[7] = m16[11] = 0.f;
        m16[15] = 1.f;
    }
    void init_ctl(const float* T12, const float* guess12) {
        GCtl& c = *h_ctl.p;
        memset(&c, 0, sizeof c);
        for (int a = 0; a < 12; ++a) {
            c.T[a] = T12[a];
            c.prev[a] = T12[a];
            c.guess[a] = guess12[a];
        }
        c.rotation_epsilon = prm.rotation_epsilon;
        c.transformation_epsilon = prm.transformation_epsilon;
        c.max_iterations = prm.max_iterations;
        c.max_inner_iterations = prm.max_inner_iterations;
    }
    void launch_correspond() {
        const double thr = prm.corr_dist_threshold * prm.corr_dist_threshold;
        k_g_correspond<<<(unsigned)(((size_t)n_src * 32 + 255) / 256), 256, 0, stream>>>(ix_tgt.view(), d_out.p, (int)n_src, d_tgt.p, d_cov_src.p, d_cov_tgt.p,
                                                                                          d_ctl.p, thr, d_corr.p);
        LAUNCH_COUNT(1);
    }

    // pcl::Registration::align + computeTransformation (:371-516)
    int32_t align(const float* guess16, float* final16, b200_gicp_result* res) {
        CUDA_SET_DEVICE(device);
        int32_t rc = ensure_covariances();
        if (rc) return rc;
        float guess12[12];
        static const float ident16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        rows_from_colmajor(guess16 ? guess16 : ident16, guess12);
        static const float ident12[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        CUDA_TRY(cudaStreamSynchronize(stream));  // the pinned control block may still be in flight
        init_ctl(ident12, guess12);                // transformation_ = previous_transformation_ = Identity (pcl::Registration::align)
        CUDA_TRY(d_out.reserve((size_t)n_src));
        CUDA_TRY(d_corr.reserve((size_t)n_src));
        CUDA_TRY(cudaEventRecord(ev0, stream));
        CUDA_TRY(cudaMemcpyAsync(d_ctl.p, h_ctl.p, sizeof(GCtl), cudaMemcpyHostToDevice, stream));
        k_g_transform<<<(unsigned)((n_src + 255) / 256), 256, 0, stream>>>(d_src.p, (int)n_src, d_ctl.p, d_out.p);
        LAUNCH_COUNT(1);
        const int ahead = 2;  // outer iterations enqueued per poll of the control block; kernels of a finished align exit at once
        int enq = 0;
        for (;;) {
            for (int a = 0; a < ahead; ++a) {
                launch_correspond();
                k_g_bfgs<<<kCluster, kBfgsThreads, 0, stream>>>(d_out.p, d_corr.p, (int)n_src, d_ctl.p);
                LAUNCH_COUNT(1);
                ++enq;
            }
            CUDA_TRY(cudaMemcpyAsync(h_ctl.p, d_ctl.p, sizeof(GCtl), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            CUDA_TRY(cudaGetLastError());
            if (h_ctl.p->converged || h_ctl.p->failed) break;
            if (enq > prm.max_iterations + ahead) B200_FAIL(B200_ERR_CUDA, "GICP outer loop did not terminate");
        }
        have_corr = true;
        CUDA_TRY(cudaEventRecord(ev1, stream));
        CUDA_TRY(cudaEventSynchronize(ev1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        const GCtl& c = *h_ctl.p;
        // final_transformation_ = previous_transformation_ * guess (:513), float, terms added left to right
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 4; ++j) {
                float s = (c.prev[i * 4] * guess12[j] + c.prev[i * 4 + 1] * guess12[4 + j]) + c.prev[i * 4 + 2] * guess12[8 + j];
                s += c.prev[i * 4 + 3] * (j == 3 ? 1.0f : 0.0f);
                final_T[i * 4 + j] = s;
            }
        if (final16) colmajor_from_rows(final_T, final16);
        if (res) {
            res->converged = c.converged;
            res->iterations = c.nr_iterations;
            res->last_m = c.last_m;
            res->last_inner = c.last_inner;
            res->last_status = c.last_status;
            res->inner_total = c.inner_total;
            res->n_f = c.n_f;
            res->n_df = c.n_df;
            res->n_fdf = c.n_fdf;
            res->delta = c.delta;
            res->gpu_ms = ms;
        }
        return c.converged ? B200_OK : B200_NOT_CONVERGED;
    }

    int32_t correspondences(const float* trans16, const float* guess16, int32_t* tgt_idx, float* maha9, float* d2, int64_t* m_out) {
        CUDA_SET_DEVICE(device);
        int32_t rc = ensure_covariances();
        if (rc) return rc;
        float T12[12], G12[12];
        rows_from_colmajor(trans16, T12);
        rows_from_colmajor(guess16, G12);
        CUDA_TRY(cudaStreamSynchronize(stream));
        init_ctl(T12, G12);
        CUDA_TRY(d_out.reserve((size_t)n_src));
        CUDA_TRY(d_corr.reserve((size_t)n_src));
        CUDA_TRY(cudaMemcpyAsync(d_ctl.p, h_ctl.p, sizeof(GCtl), cudaMemcpyHostToDevice, stream));
        k_g_transform<<<(unsigned)((n_src + 255) / 256), 256, 0, stream>>>(d_src.p, (int)n_src, d_ctl.p, d_out.p);
        LAUNCH_COUNT(1);
        launch_correspond();
        have_corr = true;
        DevBuf<int32_t> t;
        DevBuf<float> m, d;
        CUDA_TRY(t.reserve(n_src));
        CUDA_TRY(m.reserve((size_t)n_src * 9));
        CUDA_TRY(d.reserve(n_src));
        k_g_corr_export<<<(unsigned)((n_src + 255) / 256), 256, 0, stream>>>(d_corr.p, (int)n_src, t.p, m.p, d.p);
        LAUNCH_COUNT(1);
        std::vector<int32_t> ht((size_t)n_src);
        CUDA_TRY(cudaMemcpyAsync(ht.data(), t.p, (size_t)n_src * 4, cudaMemcpyDeviceToHost, stream));
        if (maha9) CUDA_TRY(cudaMemcpyAsync(maha9, m.p, (size_t)n_src * 36, cudaMemcpyDeviceToHost, stream));
        if (d2) CUDA_TRY(cudaMemcpyAsync(d2, d.p, (size_t)n_src * 4, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        t.release(); m.release(); d.release();
        int64_t cnt = 0;
        for (int64_t i = 0; i < n_src; ++i) cnt += ht[i] >= 0;
        if (tgt_idx) memcpy(tgt_idx, ht.data(), (size_t)n_src * 4);
        if (m_out) *m_out = cnt;
        return B200_OK;
    }

    int32_t cost(const double* x6, double* f_op, double* f_fdf, double* g6, int64_t* m) {
        CUDA_SET_DEVICE(device);
        if (!have_corr) B200_FAIL(B200_ERR_ARG, "no correspondences yet: call align or correspondences first");
        CUDA_TRY(cudaStreamSynchronize(stream));
        memcpy(h_d.p, x6, 6 * sizeof(double));
        CUDA_TRY(cudaMemcpyAsync(d_scratch.p, h_d.p, 6 * sizeof(double), cudaMemcpyHostToDevice, stream));
        k_g_cost<<<kCluster, kBfgsThreads, 0, stream>>>(d_out.p, d_corr.p, (int)n_src, d_scratch.p, d_scratch.p + 8);
        LAUNCH_COUNT(1);
        CUDA_TRY(cudaMemcpyAsync(h_d.p + 8, d_scratch.p + 8, 9 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        CUDA_TRY(cudaGetLastError());
        if (f_op) *f_op = h_d.p[8];
        if (f_fdf) *f_fdf = h_d.p[9];
        if (g6) memcpy(g6, h_d.p + 10, 6 * sizeof(double));
        if (m) *m = (int64_t)h_d.p[16];
        return B200_OK;
    }

    int32_t fitness(const float* T16, double max_range, double* score, int64_t* nr) {
        CUDA_SET_DEVICE(device);
        if (!have_tgt || !have_src) B200_FAIL(B200_ERR_ARG, "set_target and set_source first");
        float T12[12];
        if (T16) rows_from_colmajor(T16, T12);
        else memcpy(T12, final_T, sizeof T12);
        CUDA_TRY(cudaStreamSynchronize(stream));
        memcpy(h_d.p, T12, sizeof T12);
        CUDA_TRY(cudaMemcpyAsync(d_f.p, h_d.p, sizeof T12, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(d_scratch.reserve((size_t)n_src + 32));
        k_g_nearest_d2<<<(unsigned)(((size_t)n_src * 32 + 255) / 256), 256, 0, stream>>>(ix_tgt.view(), d_src.p, (int)n_src, d_f.p, d_scratch.p);
        k_g_fitness_reduce<<<1, 256, 0, stream>>>(d_scratch.p, (int)n_src, max_range, d_scratch.p + n_src);
        LAUNCH_COUNT(2);
        CUDA_TRY(cudaMemcpyAsync(h_d.p + 16, d_scratch.p + n_src, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        CUDA_TRY(cudaGetLastError());
        const double S = h_d.p[16], C = h_d.p[17];
        if (nr) *nr = (int64_t)C;
        if (score) *score = C > 0 ? S / C : 1.7976931348623157e308;
        return B200_OK;
    }
};

}  // namespace gicp
}  // namespace b200

struct b200_gicp {
    b200::gicp::Gicp g;
};

extern "C" {
int32_t b200_gicp_create(const b200_gicp_params* params, int32_t device, b200_gicp** out) {
    if (!params || !out) B200_FAIL(B200_ERR_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        B200_FAIL(B200_ERR_CUDA, "no such CUDA device (there is no CPU fallback)");
    }
    b200_gicp* h = new (std::nothrow) b200_gicp();
    if (!h) B200_FAIL(B200_ERR_NOMEM, "out of host memory");
    const int32_t rc = h->g.init(params, device);
    if (rc) {
        h->g.destroy();
        delete h;
        return rc;
    }
    *out = h;
    return B200_OK;
}
int32_t b200_gicp_destroy(b200_gicp* h) {
    if (!h) return B200_OK;
    h->g.destroy();
    delete h;
    return B200_OK;
}
int32_t b200_gicp_set_target(b200_gicp* h, const float* xyz, int64_t n, int64_t stride_bytes) {
    if (!h) B200_FAIL(B200_ERR_ARG, "null handle");
    return h->g.set_target(xyz, n, stride_bytes);
}
int32_t b200_gicp_set_source(b200_gicp* h, const float* xyz, int64_t n, int64_t stride_bytes) {
    if (!h) B200_FAIL(B200_ERR_ARG, "null handle");
    return h->g.set_source(xyz, n, stride_bytes);
}
int32_t b200_gicp_align(b200_gicp* h, const float* guess16, float* final16, b200_gicp_result* result) {
    if (!h) B200_FAIL(B200_ERR_ARG, "null handle");
    return h->g.align(guess16, final16, result);
}
int32_t b200_gicp_covariances(b200_gicp* h, int32_t which, double* cov9, int32_t* knn_idx) {
    if (!h || !cov9) B200_FAIL(B200_ERR_ARG, "null argument");
    b200::gicp::Gicp& g = h->g;
    CUDA_SET_DEVICE(g.device);
    if (which ? !g.have_tgt : !g.have_src) B200_FAIL(B200_ERR_ARG, "cloud not set");
    const int64_t n = which ? g.n_tgt : g.n_src;
    b200::DevBuf<double>& cov = which ? g.d_cov_tgt : g.d_cov_src;
    const int32_t rc = g.covariances_of(which ? g.ix_tgt : g.ix_src, which ? g.d_tgt.p : g.d_src.p, n, cov, knn_idx != nullptr);
    if (rc) return rc;
    (which ? g.have_cov_tgt : g.have_cov_src) = true;
    CUDA_TRY(cudaMemcpyAsync(cov9, cov.p, (size_t)n * 9 * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
    if (knn_idx) CUDA_TRY(cudaMemcpyAsync(knn_idx, g.d_knn.p, (size_t)n * g.prm.k_correspondences * sizeof(int32_t), cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
}
int32_t b200_gicp_correspondences(b200_gicp* h, const float* trans16, const float* guess16, int32_t* tgt_idx, float* maha9, float* d2, int64_t* m) {
    if (!h || !trans16 || !guess16) B200_FAIL(B200_ERR_ARG, "null argument");
    return h->g.correspondences(trans16, guess16, tgt_idx, maha9, d2, m);
}
int32_t b200_gicp_cost(b200_gicp* h, const double* x6, double* f_op, double* f_fdf, double* g6, int64_t* m) {
    if (!h || !x6) B200_FAIL(B200_ERR_ARG, "null argument");
    return h->g.cost(x6, f_op, f_fdf, g6, m);
}
int32_t b200_gicp_fitness_score(b200_gicp* h, const float* T16, double max_range, double* score, int64_t* n_in_range) {
    if (!h) B200_FAIL(B200_ERR_ARG, "null handle");
    return h->g.fitness(T16, max_range, score, n_in_range);
}
int32_t b200_gicp_index_info(b200_gicp* h, int32_t which, float* leaf, int64_t* cells, int64_t* occupied) {
    if (!h) B200_FAIL(B200_ERR_ARG, "null handle");
    const b200::gicp::CloudIndex& ix = which ? h->g.ix_tgt : h->g.ix_src;
    if (leaf) *leaf = ix.leaf;
    if (cells) *cells = ix.ncells;
    if (occupied) *occupied = ix.nruns;
    return B200_OK;
}
}
