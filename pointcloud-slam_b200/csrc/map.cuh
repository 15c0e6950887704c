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

// The same order packed one byte per cell, (dx+1) | (dy+1) << 2 | (dz+1) << 4, in GLOBAL memory: a lane reading ITS cell's
// offsets is a lane-indexed access, which the constant cache serialises address by address (ncu r02: a third of the stall
// samples of the first warp-per-query kernel sat on three 27-way divergent LDC instructions); through L1 it is one
// coalesced 32-byte request.
static __device__ const unsigned char g_stencil_code[32] = {
    21, 20, 22, 25, 17, 5, 37,  26, 24, 18, 16, 38, 36, 6, 4,  41, 33, 9, 1,  42, 40, 34, 10, 32, 8, 2, 0,  255, 255, 255, 255, 255};

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
        if (s < nstencil) b = (uint32_t)__ldg(g_stencil_code + s);
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
                cslot[t] = next_slot(cslot[t], m.tmask);
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

// ------------------------------------------------------------------ warp-per-query search ("W1")
// The 8-lanes-per-query body above spends most of its instructions on per-lane sorted top-5 lists that see ~3.6 candidates
// each (ncu r01: 425 warp instructions per query, 43 % of them the 64-bit insertion, 15 of 32 lanes active).  This body gives
// one query to a whole warp and keeps every phase lane-parallel:
//   1. lane s probes stencil cell s (all 27 table loads of the query in flight together, 8 block lines at most);
//   2. a warp scan of the run lengths gives every candidate of the query a slot = its position in the reference's
//      enumeration order (stencil order, then in-voxel order - the tie-break rank);
//   3. lane i takes slot base + i: five shuffle steps find its cell, ONE 16-byte load fetches the point (consecutive lanes
//      read consecutive points of a run), fp32 distance in the reference's operation order;
//   4. selection on the 32-bit distance bits: five rounds of warp-min (one REDUX instruction) + ballot, ties go to the
//      lowest lane = the lowest enumeration rank; winners move to lanes 0..4 and, when the query has more candidates than
//      fit, ride along as lanes 0..4 of the next chunk (they precede every later candidate in enumeration order).
// Returns the number found; lane r < count holds winner r: the point (x, y, z, ordinal bits) in `mine` and the bits of its
// fp32 squared distance in `d2bits`.  Results are bit-identical to knn5_group (same arithmetic, same total order).
//
// TMA variant (STAGED = true, "W1T", B200_KNN_MODE=9): the recipe BASELINE.json's north_star names for the gather - after the
// probe and the scan, lane s hands ITS run to the TMA engine as one 1-D bulk copy (cp.async.bulk.shared.global, cnt x 16
// bytes, completion counted in bytes on a per-warp mbarrier) into the warp's staging area at the run's slot offset, so the
// candidates of the query land in shared memory already in enumeration order.  Step 3 then reads slot i with one LDS.128:
// no five-step cell search, no dependent global gather.  A query with more candidates than the staging area holds
// (kStageCap) or whose copies do not complete within the watchdog takes the gather path below; results are bit-identical
// either way (same candidates in the same order through the same selection).
constexpr int kStageCap = 160;  // candidates per warp: 2.5 KB of shared memory, 20 KB per 256-thread block
static __device__ unsigned int g_knn_tma_timeouts = 0;  // queries whose bulk copies hit the watchdog (expected: 0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // the async proxy must see the initialised barrier
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

template <bool STAGED>
__device__ __forceinline__ int knn5_warp_t(const MapView& m, float qx, float qy, float qz, int lane, float4& mine, uint32_t& d2bits,
                                           float4* stage, uint64_t* bar) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t INF = 0xffffffffu;
    if (STAGED) {  // one phase per warp and launch: the barrier is armed once, parity 0
        if (lane == 0) mbar_init(bar, 1);
        __syncwarp();
    }
    // ---- 1. probe
    int start = 0, cnt = 0;
    {
        const int kx = pos2cell(qx, m.inv_res), ky = pos2cell(qy, m.inv_res), kz = pos2cell(qz, m.inv_res);
        if (lane < m.nstencil) {
            const uint32_t code = __ldg(g_stencil_code + lane);
            const int cx = kx + (int)(code & 3u) - 1, cy = ky + (int)((code >> 2) & 3u) - 1, cz = kz + (int)((code >> 4) & 3u) - 1;
            if (cell_in_range(cx, cy, cz)) {
                const uint64_t key = pack_key(cx, cy, cz);
                uint32_t slot = hash_key(key) & m.tmask;
                MapEntry e = ld_entry(m.ent + slot);
                while (e.key != key && e.key != kEmptyKey) {  // collisions: rare at load factor <= 0.5
                    slot = next_slot(slot, m.tmask);
                    e = ld_entry(m.ent + slot);
                }
                if (e.key == key) { start = e.start; cnt = e.count; }
            }
        }
    }
    // ---- 2. slots
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    const int excl = incl - cnt;
    const int total = __shfl_sync(FULL, incl, 31);
    mine = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    d2bits = INF;
    int found = 0;
    bool staged = false;  // warp-uniform
    if (STAGED && total > 0 && total <= kStageCap) {
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)total * 16u);
        __syncwarp();
        if (cnt > 0) bulk_g2s(stage + excl, m.pool + start, (uint32_t)cnt * 16u, bar);
        const long long t0 = clock64();
        bool done = mbar_try_wait(bar, 0);
        while (!done && clock64() - t0 < 200000000ll) done = mbar_try_wait(bar, 0);  // ~0.1 s watchdog, then the gather path
        staged = __all_sync(FULL, done);
        if (!staged && lane == 0) atomicAdd(&g_knn_tma_timeouts, 1u);
    }
    // ---- 3 + 4. chunks: the first takes 32 candidates, later ones 27 new ones next to the 5 carried winners
    for (int base = 0; base < total;) {
        const int first_new = base == 0 ? 0 : 5;
        const int i = base + lane - first_new;
        const bool fresh = lane >= first_new && i < total;
        int pos = 0;
        const int ii = fresh ? i : 0;
        int addr = 0;
        if (!STAGED || !staged) {
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {  // largest cell with excl <= slot (its run is not empty); all lanes shuffle
                const int v = __shfl_sync(FULL, excl, pos + step);
                if (v <= ii) pos += step;
            }
            addr = __shfl_sync(FULL, start, pos) + (ii - __shfl_sync(FULL, excl, pos));
        }
        float4 p = mine;              // lanes 0..4 of a later chunk keep the carried winner
        uint32_t key = lane >= first_new ? INF : d2bits;
        if (fresh) {
            p = (STAGED && staged) ? stage[ii] : __ldg(m.pool + addr);
            // distance2 (ivox3d_node.hpp:13-15): (map point - query).squaredNorm() in fp32
            const float dx = __fsub_rn(p.x, qx), dy = __fsub_rn(p.y, qy), dz = __fsub_rn(p.z, qz);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (d2 < m.max_range2) key = __float_as_uint(d2);  // d2 >= +0: the unsigned order of the bits is the float order
        }
        base += 32 - first_new;
        // a later chunk only matters if one of its candidates beats the carried fifth-best (an equal distance loses to the
        // carried winner, which has the lower enumeration rank): dense voxels then cost a gather and a vote per chunk
        if (first_new) {
            const uint32_t fifth = __shfl_sync(FULL, d2bits, 4);
            if (!__any_sync(FULL, lane >= first_new && key < fifth)) continue;
        }
        // selection: winner r = lowest lane among the smallest remaining keys
        int src = lane;      // lane r < 5 learns which lane holds winner r
        uint32_t k = key;
        found = 0;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const uint32_t mn = __reduce_min_sync(FULL, k);
            if (mn == INF) break;  // uniform
            const int w = __ffs(__ballot_sync(FULL, k == mn)) - 1;
            if (lane == w) k = INF;
            if (lane == r) src = w;
            ++found;
        }
        const float4 q4 = make_float4(__shfl_sync(FULL, p.x, src), __shfl_sync(FULL, p.y, src), __shfl_sync(FULL, p.z, src),
                                      __shfl_sync(FULL, p.w, src));
        const uint32_t kk = __shfl_sync(FULL, key, src);
        if (lane < found) { mine = q4; d2bits = kk; }
        else if (lane < 5) { mine = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)); d2bits = INF; }
    }
    return found;
}
__device__ __forceinline__ int knn5_warp(const MapView& m, float qx, float qy, float qz, int lane, float4& mine, uint32_t& d2bits) {
    return knn5_warp_t<false>(m, qx, qy, qz, lane, mine, d2bits, nullptr, nullptr);
}

// ------------------------------------------------------------------ 8 lanes per query, balanced through shared memory ("G8P")
// One wave holds a whole 20k-point scan at 8 lanes per query (a warp per query needs two), and four queries share every
// instruction a warp issues - as long as the code does not diverge: the first version of this body lost a fifth of its
// instructions to BSSY / BSYNC / BRA around per-lane loops (ncu r02), so everything below is straight-line and predicated.
//   probes   lane lg probes cells lg, lg+8, lg+16, lg+24 of the stencil, four table loads in flight;
//   slots    a 16-bit packed group scan numbers the candidates of the query in the reference's enumeration order
//            (stencil order, then in-voxel order); a chunk is 32 consecutive slots;
//   push     every run that overlaps the chunk sets a head bit at its first slot; the OR of the group's head bits makes
//            "slot -> run" a popcount, and the owner stores ONE word per run (pool address minus slot) at the run's ordinal;
//   pull     lane lg takes slots 4lg..4lg+3: address = slot + delta[popc(heads below)], four gathers in flight, fp32
//            distance in the reference's operation order;
//   select   on the 32-bit distance bits: five rounds of lane-min, group-min (three butterfly shuffles), ballot; the
//            lowest hit lane wins and knocks out its first matching register - slots are in enumeration order, so ties
//            go to the lower rank; winners' addresses are staged in shared memory and land on lanes 0..4, which carry
//            them into the next chunk when the query has more than 32 candidates (a carried winner precedes every later one).
// Bit-identical to knn5_group / knn5_warp.  smem: kG8pWords 32-bit words per query.
constexpr int kG8pCap = 32;               // candidate slots per chunk and query
constexpr int kG8pWords = kG8pCap + 16;   // run deltas + staged winners: 5 addresses (padded to 8), 5 keys (padded to 8)
static __device__ const uint32_t g_stencil_code4[8] = {  // byte t = g_stencil_code[lg + 8 t]
    21u | 24u << 8 | 33u << 16 | 8u << 24,    20u | 18u << 8 | 9u << 16 | 2u << 24,     22u | 16u << 8 | 1u << 16 | 0u << 24,
    25u | 38u << 8 | 42u << 16 | 255u << 24,  17u | 36u << 8 | 40u << 16 | 255u << 24,  5u | 6u << 8 | 34u << 16 | 255u << 24,
    37u | 4u << 8 | 10u << 16 | 255u << 24,   26u | 41u << 8 | 32u << 16 | 255u << 24};

// Every lane of the warp must call this together (groups without a query pass active = false): all collectives use the
// full mask - with per-group masks the compiler wraps each of them in WARPSYNC / ENDCOLLECTIVE / BSSY / BSYNC.
__device__ __forceinline__ int knn5_g8p(const MapView& m, bool active, float qx, float qy, float qz, int lg, uint32_t* sq,
                                        float4& mine, uint32_t& d2bits) {
    constexpr int G = 8, SLOTS = 4;
    constexpr uint32_t INF = 0xffffffffu;
    constexpr unsigned gmask = 0xffffffffu;
    const int lane0 = (threadIdx.x & 31) & ~(G - 1);
    // ---- probes
    int cstart[SLOTS], ccount[SLOTS];
    {
        const uint32_t codes = __ldg(g_stencil_code4 + lg);
        const int kx = pos2cell(qx, m.inv_res), ky = pos2cell(qy, m.inv_res), kz = pos2cell(qz, m.inv_res);
        uint32_t klo[SLOTS], khi[SLOTS], cslot[SLOTS];
        uint4 ce[SLOTS];
        bool pend = false;
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) {
            const uint32_t b = (codes >> (8 * t)) & 0xFFu;
            const int cx = kx + (int)(b & 3u) - 1, cy = ky + (int)((b >> 2) & 3u) - 1, cz = kz + (int)((b >> 4) & 3u) - 1;
            const bool ok = active && lg + G * t < m.nstencil && cell_in_range(cx, cy, cz);
            const uint64_t key = ok ? pack_key(cx, cy, cz) : kEmptyKey;
            klo[t] = (uint32_t)key;
            khi[t] = (uint32_t)(key >> 32);
            cslot[t] = hash_key(key) & m.tmask;
            ce[t] = make_uint4(INF, INF, 0u, 0u);
            if (ok) ce[t] = __ldg(reinterpret_cast<const uint4*>(m.ent + cslot[t]));
        }
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) pend |= (ce[t].x != klo[t] || ce[t].y != khi[t]) && (ce[t].x & ce[t].y) != INF;
        while (pend) {  // collisions: one shared loop for the four probes of the lane
            pend = false;
#pragma unroll
            for (int t = 0; t < SLOTS; ++t) {
                if ((ce[t].x != klo[t] || ce[t].y != khi[t]) && (ce[t].x & ce[t].y) != INF) {
                    cslot[t] = next_slot(cslot[t], m.tmask);
                    ce[t] = __ldg(reinterpret_cast<const uint4*>(m.ent + cslot[t]));
                    pend |= (ce[t].x != klo[t] || ce[t].y != khi[t]) && (ce[t].x & ce[t].y) != INF;
                }
            }
        }
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) {
            const bool hit = ce[t].x == klo[t] && ce[t].y == khi[t] && (klo[t] & khi[t]) != INF;
            cstart[t] = (int)ce[t].z;
            ccount[t] = hit ? (int)ce[t].w : 0;
        }
    }
    // ---- enumeration-order slot numbers: exclusive prefix over (t, lg); two 16-bit counters per word unless a run is huge
    int excl[SLOTS], total;
    const bool huge = __any_sync(gmask, max(max(ccount[0], ccount[1]), max(ccount[2], ccount[3])) > 8191);  // warp-uniform
    if (!huge) {
        uint32_t w01 = (uint32_t)ccount[0] | ((uint32_t)ccount[1] << 16), w23 = (uint32_t)ccount[2] | ((uint32_t)ccount[3] << 16);
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
            const uint32_t a = __shfl_up_sync(gmask, w01, o, G), b = __shfl_up_sync(gmask, w23, o, G);
            if (lg >= o) { w01 += a; w23 += b; }
        }
        const uint32_t t01 = __shfl_sync(gmask, w01, G - 1, G), t23 = __shfl_sync(gmask, w23, G - 1, G);
        const int T0 = (int)(t01 & 0xFFFFu), T1 = (int)(t01 >> 16), T2 = (int)(t23 & 0xFFFFu), T3 = (int)(t23 >> 16);
        excl[0] = (int)(w01 & 0xFFFFu) - ccount[0];
        excl[1] = T0 + (int)(w01 >> 16) - ccount[1];
        excl[2] = T0 + T1 + (int)(w23 & 0xFFFFu) - ccount[2];
        excl[3] = T0 + T1 + T2 + (int)(w23 >> 16) - ccount[3];
        total = T0 + T1 + T2 + T3;
    } else {
        total = 0;
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) {
            int incl = ccount[t];
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                const int v = __shfl_up_sync(gmask, incl, o, G);
                if (lg >= o) incl += v;
            }
            excl[t] = total + incl - ccount[t];
            total += __shfl_sync(gmask, incl, G - 1, G);
        }
    }
    uint32_t* s_delta = sq;                // [kG8pCap] pool address minus slot, one word per run of the chunk
    uint32_t* s_waddr = sq + kG8pCap;      // [5] winners' pool addresses
    uint32_t* s_wkey = sq + kG8pCap + 8;   // [5] and distance bits
    uint32_t kc = INF, ac = 0;             // carried winner of lanes 0..4 (distance bits, pool address)
    int found = 0;
    for (int base = 0; __any_sync(gmask, base < total); base += kG8pCap) {  // warp-uniform trip count: a group that is done repeats its (idempotent) selection
        // push: head bits and one delta per run
        uint32_t heads = 0;
        int s0[SLOTS];
        bool in[SLOTS];
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) {
            s0[t] = min(max(excl[t] - base, 0), 31);  // first slot of the run inside this chunk
            in[t] = ccount[t] > 0 && excl[t] + ccount[t] > base && excl[t] < base + kG8pCap;
            heads |= in[t] ? (1u << s0[t]) : 0u;
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) heads |= __shfl_xor_sync(gmask, heads, o, G);
#pragma unroll
        for (int t = 0; t < SLOTS; ++t)
            if (in[t]) s_delta[__popc(heads & ((1u << s0[t]) - 1u))] = (uint32_t)(cstart[t] + (base + s0[t] - excl[t]) - s0[t]);
        __syncwarp();
        // pull: slots 4 lg .. 4 lg + 3
        const int navail = min(kG8pCap, total - base) - 4 * lg;  // how many of this lane's four slots hold a candidate
        uint32_t ad[4], k[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int slot = 4 * lg + u;
            ad[u] = (uint32_t)slot + s_delta[max(__popc(heads & (0xFFFFFFFFu >> (31 - slot))) - 1, 0)];
        }
        {
            float4 p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (u < navail) p[u] = __ldg(m.pool + ad[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                // distance2 (ivox3d_node.hpp:13-15): (map point - query).squaredNorm() in fp32
                const float dx = __fsub_rn(p[u].x, qx), dy = __fsub_rn(p[u].y, qy), dz = __fsub_rn(p[u].z, qz);
                const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                k[u] = (u < navail && d2 < m.max_range2) ? __float_as_uint(d2) : INF;  // d2 >= +0: unsigned order of the bits = float order
            }
        }
        int wr[4] = {7, 7, 7, 7}, wrc = 7;  // round in which the register won
        found = 0;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const uint32_t lmk = min(min(k[0], k[1]), min(k[2], k[3]));
            uint32_t gm = min(lmk, kc);
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) gm = min(gm, __shfl_xor_sync(gmask, gm, o, G));
            // equal distances: a carried winner first (lowest lane = best of them), then the lowest slot = lowest enumeration rank
            const unsigned hc = (__ballot_sync(gmask, kc == gm) >> lane0) & 0xFFu;
            const unsigned hk = (__ballot_sync(gmask, lmk == gm) >> lane0) & 0xFFu;
            const bool alive = gm != INF;  // uniform inside the group
            found += alive ? 1 : 0;
            const bool me = alive && lg == __ffs(hc ? hc : hk) - 1;
            const bool tc = me && hc != 0u, tk = me && hc == 0u;
            const bool h0 = tk && k[0] == gm, h1 = tk && !h0 && k[1] == gm, h2 = tk && !h0 && !h1 && k[2] == gm, h3 = tk && !h0 && !h1 && !h2;
            if (tc) { kc = INF; wrc = r; }
            if (h0) { k[0] = INF; wr[0] = r; }
            if (h1) { k[1] = INF; wr[1] = r; }
            if (h2) { k[2] = INF; wr[2] = r; }
            if (h3) { k[3] = INF; wr[3] = r; }
            if (me) s_wkey[r] = gm;
        }
        if (wrc < 5) s_waddr[wrc] = ac;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (wr[u] < 5) s_waddr[wr[u]] = ad[u];
        __syncwarp();
        kc = INF;
        if (lg < found) { kc = s_wkey[lg]; ac = s_waddr[lg]; }
        __syncwarp();
    }
    mine = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    d2bits = kc;
    if (lg < found) mine = __ldg(m.pool + ac);
    return found;
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
    int32_t* d_bcnt = nullptr;  // [tsize] batch points per voxel of the insert in flight (sort-free path), all zero between inserts
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
    DevBuf<float4> q_nb;

    // Candidate-walk variant of the search kernels, picked per launch from the map's density: lane-owned cells (0) win while
    // every voxel holds a few points (0.2 m voxels: 27 vs 47 us per 20k queries at 3 points/voxel), the cooperative walk (1)
    // as soon as the map has dense voxels - on average (100 vs 52 us at 25 points/voxel) or just somewhere: in the sliding-map
    // sequence a handful of voxels near the sensor path collect hundreds of points while the mean is still below 6, and
    // a lane walking such a run alone made the update 1.8 ms instead of 0.4 ms.
    int knn_mode() const {
        static const char* env = getenv("B200_KNN_MODE");
        if (env) return atoi(env);
        return 7;  // 7 = warp per query (knn5_warp); 9 = the same with TMA-staged candidates (knn5_warp_t<true>); 8 = balanced 8-lanes-per-query body (knn5_g8p); 0 / 1 / 4 / 5 = the round-1 walks (A/B timing)
    }
    int knn_mode_g8() const {  // density-driven choice among the 8-lanes-per-query walks (LOAM front end, A/B timing)
        static const char* env = getenv("B200_KNN_MODE");
        if (env && atoi(env) < 7) return atoi(env);
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
    // points already on the device (x,y,z,*).  d_count / h_count (optional): the batch is the first *d_count (<= n) points,
    // a count that is still being produced on the stream; h_count is its pinned host copy, valid after the call
    int32_t insert_device(const float4* d_pts, int64_t n, const int32_t* d_count = nullptr, const int32_t* h_count = nullptr);
    int32_t insert_host(const float* xyz, int64_t n, int64_t stride);
    int32_t knn5_host(const float* xyz, int64_t n, int64_t stride, int32_t* idx, float* d2, int32_t* cnt, float* nb_xyz = nullptr);
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
