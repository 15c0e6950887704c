// b200reg — GPU-resident local map index (replaces jueying_lio::IVox, ivox3d.h:53-286).
//
// Layout in HBM
//   ent[T]    16 B     open-addressing table (T = pow2 >= 2*capacity): {packed voxel key, start, count};
//                      one 16-byte load resolves a stencil probe to the voxel's run in the point pool
//   aux[T]    int2     {cap, stamp}: slot capacity / ordinal of the last point that touched the voxel (LRU recency),
//                      only read by insert
//   pool[]    float4   points, voxel-contiguous, in-voxel order = insertion order;
//                      .w carries the global insertion ordinal (int bits)
// A voxel is one contiguous, 16-byte-aligned run, so a stencil search is <= 27 table probes
// followed by <= 27 contiguous gathers.  Insert = radix sort of the batch by key + run-length
// segmentation + per-run upsert (bump allocation, relocation with doubling when a run outgrows
// its slot); the pool is compacted when it fills up.
#pragma once
#include "common.cuh"

namespace b200 {

struct MapCounters {
    unsigned long long pool_top;
    unsigned int num_voxels;
    unsigned int err_range;    // a point fell outside the key range
    unsigned int err_pool;     // pool exhausted during insert (host compacts / grows and retries)
    unsigned int err_capacity; // voxel capacity reached
    unsigned long long num_points;
    unsigned long long live_points;  // sum of run counts (== num_points while nothing is evicted)
    unsigned int max_count;          // largest run any voxel has ever had (never decreases: picks the search walk)
    unsigned int pad;
};

struct __align__(16) MapEntry {
    uint64_t key;
    int start;
    int count;
};

struct MapView {  // what kernels see
    const MapEntry* ent;
    const float4* pool;
    uint32_t tmask;
    float inv_res;
    int nstencil;
    float max_range2;
};

// stencil enumeration order of IVox::GenerateNearbyGrids (ivox3d.h:213-231): NEARBY6 is a prefix
// of NEARBY18 which is a prefix of NEARBY26.
static __constant__ signed char c_stencil[27][4] = {
    {0, 0, 0, 0},   {-1, 0, 0, 0}, {1, 0, 0, 0},   {0, 1, 0, 0},   {0, -1, 0, 0},  {0, 0, -1, 0},  {0, 0, 1, 0},
    {1, 1, 0, 0},   {-1, 1, 0, 0}, {1, -1, 0, 0},  {-1, -1, 0, 0}, {1, 0, 1, 0},   {-1, 0, 1, 0},  {1, 0, -1, 0},
    {-1, 0, -1, 0}, {0, 1, 1, 0},  {0, -1, 1, 0},  {0, 1, -1, 0},  {0, -1, -1, 0}, {1, 1, 1, 0},   {-1, 1, 1, 0},
    {1, -1, 1, 0},  {1, 1, -1, 0}, {-1, -1, 1, 0}, {-1, 1, -1, 0}, {1, -1, -1, 0}, {-1, -1, -1, 0}};

// IVox::Pos2Grid (ivox3d.h:284-286): round(p * inv_res) with std::round semantics.
__device__ __forceinline__ int pos2cell(float v, float inv_res) { return (int)roundf(__fmul_rn(v, inv_res)); }

__device__ __forceinline__ MapEntry ld_entry(const MapEntry* e) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(e));
    MapEntry r;
    r.key = ((uint64_t)v.y << 32) | v.x;
    r.start = (int)v.z;
    r.count = (int)v.w;
    return r;
}

constexpr uint64_t kInfKey = 0xFFFFFFFFFFFFFFFFull;
constexpr uint64_t kTombKey = 0xFFFFFFFFFFFFFFFEull;  // evicted voxel: keeps probe chains intact until the next rehash
constexpr uint64_t kDropKey = 0x8000000000000000ull;  // insert: key of a non-finite / out-of-range point (sorts behind every voxel key)
constexpr int kRankBits = 26;  // in-voxel index bits inside the tie-break rank

// sorted insert of k into ascending t[0..4]
__device__ __forceinline__ void top5_insert(uint64_t (&t)[5], uint64_t k) {
    if (k < t[4]) {
        t[4] = k;
#pragma unroll
        for (int i = 4; i > 0; --i) {
            if (t[i] < t[i - 1]) {
                uint64_t a = t[i];
                t[i] = t[i - 1];
                t[i - 1] = a;
            }
        }
    }
}

// One query handled by a group of G lanes (G = 8): IVox::GetClosestPoint(pt, out, 5, max_range)
// (ivox3d.h:132-204) + IVoxNode::KNNPointByCondition (ivox3d_node.hpp:140-205).
// Candidate total order = (float d2 bits, stencil index, in-voxel index) — the "stable selection"
// contract of SURVEY.md §7.  Returns the number found (0..5); lane r < count holds the r-th winner: its
// composite key in wkey and its pool entry (x, y, z, ordinal) in mine.
//
// Memory-level parallelism and coalescing are the whole game here (a probe and a gather are both dependent
// L2/HBM round trips): a lane first issues the table loads of all its stencil cells, then the group walks
// the occupied runs together, G consecutive points per step (one or two 128-byte lines per query instead
// of G scattered 16-byte gathers), four steps in flight; short runs stay with the probing lane.  Every lane keeps a
// sorted top-5 of 64-bit keys and the G lists are merged by five group-min rounds.  (A group-shared running top-5
// with a ballot per step was measured slower on B200: the ballots serialise the four groups of a warp.)
// packed stencil offsets of lane lg: byte t of the result = (dx+1) | (dy+1)<<2 | (dz+1)<<4 for cell lg + G*t,
// 0xFF when that cell is beyond the stencil.  Computed once per thread.
template <int G>
__device__ __forceinline__ uint32_t lane_stencil(int lg, int nstencil) {
    uint32_t r = 0;
#pragma unroll
    for (int t = 0; t < (27 + G - 1) / G; ++t) {
        const int s = lg + G * t;
        uint32_t b = 0xFFu;
        if (s < nstencil) b = (uint32_t)(c_stencil[s][0] + 1) | ((uint32_t)(c_stencil[s][1] + 1) << 2) | ((uint32_t)(c_stencil[s][2] + 1) << 4);
        r |= b << (8 * t);
    }
    return r;
}

#ifndef B200_KNN_MODE
#define B200_KNN_MODE 0
#endif
// MODE 5 ("flattened"): the (cell, in-voxel index) pairs of all occupied stencil cells of a query are written as one list
// into shared memory (kFlatCap entries per query) and the G lanes then take the list G entries at a time, four steps in
// flight: every lane handles total/G candidates whatever the occupancy pattern of its own cells (two thirds of the cells
// of a surface map are empty), so the divergent per-cell loops of MODE 0 disappear and all gathers of a query are
// independent loads.  Queries with more than kFlatCap candidates (dense, long-lived maps) fall back to the cooperative
// walk of MODE 1, decided per query.
constexpr int kFlatCap = 64;
constexpr int kFlatStride = kFlatCap + 1;  // uint2 entries per query list; odd stride spreads the groups of a warp over the banks
template <int G, int MODE = B200_KNN_MODE>
__device__ __forceinline__ int knn5_group(const MapView& m, float qx, float qy, float qz, int lg, unsigned gmask, uint32_t lst,
                                          uint64_t& wkey, float4& mine, uint2* flat = nullptr) {
    constexpr int SLOTS = (27 + G - 1) / G;
    static_assert(SLOTS <= 4, "lane_stencil packs four cells per lane");
    static_assert(G >= 8, "lanes 0..4 of the group hold the running top-5");
    constexpr unsigned GM = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
    int cstart[SLOTS], ccount[SLOTS];
    {
        uint64_t ckey[SLOTS];
        uint32_t cslot[SLOTS];
        MapEntry ce[SLOTS];
        const int kx = pos2cell(qx, m.inv_res), ky = pos2cell(qy, m.inv_res), kz = pos2cell(qz, m.inv_res);
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) {  // all first probes in flight together
            const uint32_t b = (lst >> (8 * t)) & 0xFFu;
            ckey[t] = kEmptyKey;
            ce[t].key = kEmptyKey;
            ce[t].start = 0;
            ce[t].count = 0;
            cslot[t] = 0;
            if (b != 0xFFu) {
                const int cx = kx + (int)(b & 3u) - 1, cy = ky + (int)((b >> 2) & 3u) - 1, cz = kz + (int)((b >> 4) & 3u) - 1;
                if (cell_in_range(cx, cy, cz)) {
                    ckey[t] = pack_key(cx, cy, cz);
                    cslot[t] = hash_key(ckey[t]) & m.tmask;
                    ce[t] = ld_entry(m.ent + cslot[t]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) {  // collisions: keep probing linearly (rare at load factor <= 0.5)
            while (ce[t].key != ckey[t] && ce[t].key != kEmptyKey) {
                cslot[t] = (cslot[t] + 1) & m.tmask;
                ce[t] = ld_entry(m.ent + cslot[t]);
            }
            const bool hit = ckey[t] != kEmptyKey && ce[t].key == ckey[t];
            cstart[t] = ce[t].start;
            ccount[t] = hit ? ce[t].count : 0;
        }
    }
    const int lane0 = ((threadIdx.x & 31) / G) * G;  // first lane of this group inside the warp
    uint64_t top[5] = {kInfKey, kInfKey, kInfKey, kInfKey, kInfKey};  // this lane's sorted best keys
    auto consider = [&](const float4& p, uint32_t rank) {
        // distance2 (ivox3d_node.hpp:13-15): (map point - query).squaredNorm() in fp32
        const float dx = __fsub_rn(p.x, qx), dy = __fsub_rn(p.y, qy), dz = __fsub_rn(p.z, qz);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        if (d2 < m.max_range2) top5_insert(top, ((uint64_t)__float_as_uint(d2) << 32) | rank);
    };
    // MODE 0: every lane walks its own cells one after the other, four gathers in flight (scattered 16-byte loads)
    // MODE 1: the group walks every occupied cell together, G consecutive points per step
    // MODE 2/3: runs of <= SHORT points stay with the probing lane (2: two points of every cell of the lane in flight
    //           together, 3: cell by cell, four points in flight), longer runs are walked by the group
    // MODE 4: like 3 when the query's stencil holds more than DENSE candidates in total, like 0 otherwise (decided per query)
    int SHORT = MODE == 0 ? (1 << 30) : (MODE == 1 || MODE >= 5) ? 0 : MODE == 2 ? 2 : 4;
    bool flat_done = false;
    if (MODE >= 5) {
        int mytot = 0;
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) mytot += ccount[t];
        int incl = mytot;
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
            const int v = __shfl_up_sync(gmask, incl, o, G);
            if (lg >= o) incl += v;
        }
        const int total = __shfl_sync(gmask, incl, lane0 + G - 1);
        if (total <= kFlatCap) {  // uniform inside the group
            flat_done = true;
            int w = incl - mytot;
#pragma unroll
            for (int t = 0; t < SLOTS; ++t) {
                const uint32_t rbase = (uint32_t)(lg + G * t) << kRankBits;
                for (int j = 0; j < ccount[t]; ++j) flat[w++] = make_uint2((uint32_t)(cstart[t] + j), rbase | (uint32_t)j);
            }
            __syncwarp(gmask);
            for (int c0 = lg; c0 < total; c0 += 4 * G) {
                uint2 e[4];
                float4 p[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (c0 + u * G < total) e[u] = flat[c0 + u * G];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (c0 + u * G < total) p[u] = __ldg(m.pool + e[u].x);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (c0 + u * G < total) consider(p[u], e[u].y);
            }
        }
    }
    if (MODE == 4) {
        constexpr int DENSE = 80;
        int tot = 0;
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) tot += ccount[t];
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) tot += __shfl_xor_sync(gmask, tot, o);
        if (tot <= DENSE) SHORT = 1 << 30;
    }
    if (MODE == 0 || MODE == 3 || MODE == 4) {
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) {
            const int cnt = ccount[t] <= SHORT ? ccount[t] : 0;
            const float4* run = m.pool + cstart[t];
            const uint32_t rbase = (uint32_t)(lg + G * t) << kRankBits;
            for (int j0 = 0; j0 < cnt; j0 += 4) {
                float4 p[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (j0 + u < cnt) p[u] = __ldg(run + j0 + u);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (j0 + u < cnt) consider(p[u], rbase | (uint32_t)(j0 + u));
            }
        }
    }
    if (MODE == 2) {
        float4 p[SLOTS][2];
#pragma unroll
        for (int t = 0; t < SLOTS; ++t)
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (ccount[t] <= SHORT && u < ccount[t]) p[t][u] = __ldg(m.pool + cstart[t] + u);
#pragma unroll
        for (int t = 0; t < SLOTS; ++t)
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (ccount[t] <= SHORT && u < ccount[t]) consider(p[t][u], ((uint32_t)(lg + G * t) << kRankBits) | (uint32_t)u);
    }
    if (MODE != 0 && !flat_done) {
        // longer runs are walked by the whole group: G consecutive points per step (one or two 128-byte lines per query
        // instead of G scattered gathers), four steps in flight.  Which lane sees a candidate does not matter: the key
        // carries the full enumeration rank.
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) {
            unsigned occ = (__ballot_sync(gmask, ccount[t] > SHORT) >> lane0) & GM;
            while (occ) {  // uniform inside the group
                const int o = __ffs(occ) - 1;
                occ &= occ - 1;
                const int cnt = __shfl_sync(gmask, ccount[t], lane0 + o);
                const float4* run = m.pool + __shfl_sync(gmask, cstart[t], lane0 + o);
                const uint32_t rbase = (uint32_t)(o + G * t) << kRankBits;
                for (int j0 = lg; j0 < cnt; j0 += 4 * G) {
                    float4 p[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (j0 + u * G < cnt) p[u] = __ldg(run + j0 + u * G);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (j0 + u * G < cnt) consider(p[u], rbase | (uint32_t)(j0 + u * G));
                }
            }
        }
    }
    // merge the G sorted lists: 5 rounds of group-min + pop; lane r keeps winner r
    int count = 0;
    uint64_t best = kInfKey;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        uint64_t mn = top[0];
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            const uint64_t other = __shfl_xor_sync(gmask, mn, o);
            mn = other < mn ? other : mn;
        }
        if (lg == r) best = mn;
        if (mn != kInfKey) {
            ++count;
            if (top[0] == mn) {
                top[0] = top[1]; top[1] = top[2]; top[2] = top[3]; top[3] = top[4]; top[4] = kInfKey;
            }
        }
    }
    wkey = best;
    // fetch the winners: lane r loads winner r
    mine = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    {
        const uint32_t lo = (uint32_t)best;
        const int s = (int)(lo >> kRankBits), j = (int)(lo & ((1u << kRankBits) - 1));
        const int owner = s % G, slot = s / G;
        int st = 0;
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) {  // every lane takes part in the shuffles
            const int v = __shfl_sync(gmask, cstart[t], lane0 + (owner & (G - 1)));
            if (slot == t) st = v;
        }
        if (best != kInfKey) mine = __ldg(m.pool + st + j);
    }
    return count;
}

// ------------------------------------------------------------------ host-side object
struct Map {
    b200_map_params prm;
    int device = 0;
    cudaStream_t stream = nullptr;
    float inv_res = 0.f;
    int nstencil = 19;
    uint32_t tsize = 0;
    MapEntry* d_ent = nullptr;
    int2* d_aux = nullptr;
    float4* d_pool = nullptr;
    uint64_t pool_cap = 0;
    MapCounters* d_ctr = nullptr;
    MapCounters h_ctr{};   // mirror after the last insert
    int64_t next_ord = 0;
    uint64_t tombstones = 0, evicted_total = 0;
    uint64_t dropped_last = 0, dropped_total = 0;  // non-finite / out-of-range points skipped by the last insert / so far
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_knn_ms = 0.f;
    // scratch
    DevBuf<float4> in_pts;
    DevBuf<uint64_t> k_in, k_out, k_uniq;
    DevBuf<int32_t> v_in, v_out, run_cnt, run_off, run_dst, run_reloc, d_nruns;
    DevBuf<uint8_t> cub_tmp;
    PinnedBuf<float4> h_stage;
    PinnedBuf<MapCounters> h_ctr_pin;
    // LRU eviction scratch
    DevBuf<uint64_t> lru_in, lru_out;
    DevBuf<int32_t> victims_dev;
    PinnedBuf<int32_t> h_runs, h_small;
    PinnedBuf<uint64_t> h_lru;
    // knn scratch
    DevBuf<int32_t> q_idx, q_cnt;
    DevBuf<float> q_d2;

    // Candidate-walk variant of the search kernels, picked per launch from the map's density: lane-owned cells (0) win while
    // every voxel holds a few points (0.2 m voxels: 27 vs 47 us per 20k queries at 3 points/voxel), the cooperative walk (1)
    // as soon as the map has dense voxels - on average (100 vs 52 us at 25 points/voxel) or just somewhere: in the sliding-map
    // sequence a handful of voxels near the sensor path collect hundreds of points while the mean is still below 6, and
    // a lane walking such a run alone made the update 1.8 ms instead of 0.4 ms.
    int knn_mode() const {
        static const char* env = getenv("B200_KNN_MODE");
        if (env) return atoi(env);
        if (h_ctr.num_voxels == 0) return 0;
        return (h_ctr.live_points > 6ull * h_ctr.num_voxels || h_ctr.max_count > 32u) ? 1 : 0;
    }
    MapView view() const {
        MapView v;
        v.ent = d_ent; v.pool = d_pool; v.tmask = tsize - 1; v.inv_res = inv_res;
        v.nstencil = nstencil; v.max_range2 = prm.max_range * prm.max_range;
        return v;
    }
    int32_t init(const b200_map_params* p, int dev);
    void destroy();
    int32_t clear();
    int32_t insert_device(const float4* d_pts, int64_t n);  // points already on the device (x,y,z,*)
    int32_t insert_host(const float* xyz, int64_t n, int64_t stride);
    int32_t knn5_host(const float* xyz, int64_t n, int64_t stride, int32_t* idx, float* d2, int32_t* cnt);
    int32_t grow_pool(uint64_t min_cap);
    int32_t evict_for_batch(int64_t n);
    int32_t rehash();
};

}  // namespace b200

struct b200_map {
    b200::Map m;
    int refs = 0;         // b200_iekf handles built on this map
    bool zombie = false;  // b200_map_destroy was called while filters were still attached
};
