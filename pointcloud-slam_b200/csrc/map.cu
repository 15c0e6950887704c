// b200reg — local map: build / incremental insert / k=5 stencil search.  See map.cuh for the layout.
#include "map.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <queue>
#include <unordered_map>
#include <vector>

namespace b200 {

thread_local std::string g_last_error;
std::atomic<int64_t> g_kernel_launches{0};

// ------------------------------------------------------------------ insert kernels
// 1. voxel key of every incoming point (IVox::Pos2Grid, ivox3d.h:284-286)
__global__ void k_point_keys(const float4* __restrict__ pts, int n, const int32_t* __restrict__ d_count, float inv_res,
                             uint64_t* __restrict__ keys, int32_t* __restrict__ vals, MapCounters* ctr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (d_count && i >= *d_count) {  // past the real end of a batch whose size is only known on the device: skipped like a dropped point
        keys[i] = kDropKey;
        vals[i] = i;
        return;
    }
    float4 p = pts[i];
    int cx = pos2cell(p.x, inv_res), cy = pos2cell(p.y, inv_res), cz = pos2cell(p.z, inv_res);
    // a non-finite or out-of-range point is dropped, not inserted: its key sorts behind every voxel key and the run of
    // such keys is skipped by the upsert; the count is reported as a non-fatal status (b200_map_dropped)
    const bool bad = !cell_in_range(cx, cy, cz) || !(isfinite(p.x) && isfinite(p.y) && isfinite(p.z));
    if (bad) atomicAdd(&ctr->err_range, 1u);
    keys[i] = bad ? kDropKey : pack_key(cx, cy, cz);
    vals[i] = i;
}

// 2. one thread per run of equal keys: find-or-create the voxel, reserve room for the run.
//    run_dst[r]   = pool index where the run's first new point goes
//    run_reloc[r] = old start if the voxel had to move (its old points are copied by k_relocate), else -1
__global__ void k_upsert_runs(const uint64_t* __restrict__ uniq, const int32_t* __restrict__ cnt, const int32_t* __restrict__ nruns,
                              const int32_t* __restrict__ run_off, const int32_t* __restrict__ sorted_vals, int base_ord,
                              MapEntry* ent, int2* aux, uint32_t tmask, uint64_t pool_cap, uint32_t capacity_voxels,
                              MapCounters* ctr, int32_t* __restrict__ run_dst, int32_t* __restrict__ run_reloc,
                              int32_t* __restrict__ run_oldcnt) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *nruns) return;
    const uint64_t key = uniq[r];
    const int c = cnt[r];
    if (key == kDropKey) {  // the dropped points of this batch
        run_dst[r] = -1;
        run_reloc[r] = -1;
        run_oldcnt[r] = 0;
        return;
    }
    // LRU recency = ordinal of the last point that touched the voxel (the sort is stable: last element of the run)
    const int stamp = base_ord + sorted_vals[run_off[r] + c - 1];
    uint32_t slot = hash_key(key) & tmask;
    bool created = false;
    while (true) {
        uint64_t k = ent[slot].key;
        if (k == key) break;
        if (k == kEmptyKey) {
            uint64_t old = atomicCAS((unsigned long long*)&ent[slot].key, (unsigned long long)kEmptyKey, (unsigned long long)key);
            if (old == kEmptyKey) { created = true; break; }
            if (old == key) break;
        }
        slot = next_slot(slot, tmask);
    }
    int4 v = created ? make_int4(0, 0, 0, 0) : make_int4(ent[slot].start, ent[slot].count, aux[slot].x, aux[slot].y);
    if (created) {
        unsigned nv = atomicAdd(&ctr->num_voxels, 1u) + 1u;
        if (nv >= capacity_voxels) atomicAdd(&ctr->err_capacity, 1u);
    }
    const int newcount = v.y + c;
    int reloc = -1;
    if (newcount > v.z) {
        int newcap = v.z == 0 ? newcount : max(2 * v.z, newcount);
        unsigned long long st = atomicAdd(&ctr->pool_top, (unsigned long long)newcap);
        if (st + (unsigned long long)newcap > pool_cap) {
            atomicAdd(&ctr->err_pool, 1u);
            run_dst[r] = -1;
            run_reloc[r] = -1;
            return;
        }
        if (v.y > 0) reloc = v.x;
        v.x = (int)st;
        v.z = newcap;
    }
    run_dst[r] = v.x + v.y;
    run_reloc[r] = reloc;
    run_oldcnt[r] = v.y;
    atomicAdd(&ctr->live_points, (unsigned long long)c);
    if ((unsigned)newcount > ctr->max_count) atomicMax(&ctr->max_count, (unsigned)newcount);
    v.y = newcount;
    v.w = stamp;
    ent[slot].start = v.x;
    ent[slot].count = v.y;
    aux[slot] = make_int2(v.z, v.w);
}

// 3. voxels that moved: copy their old points to the new slot (one thread per run, old runs are short)
__global__ void k_relocate(const int32_t* __restrict__ nruns, const int32_t* __restrict__ run_dst, const int32_t* __restrict__ run_reloc,
                            const int32_t* __restrict__ run_oldcnt, float4* pool) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *nruns) return;
    int src = run_reloc[r];
    if (src < 0) return;
    int oc = run_oldcnt[r];
    int dst = run_dst[r] - oc;
    for (int j = 0; j < oc; ++j) pool[dst + j] = pool[src + j];
}

// 4. scatter the sorted batch into the pool: element i of the sorted order belongs to run
//    r = upper_bound(run_off, i) - 1 and lands at run_dst[r] + (i - run_off[r]).
__global__ void k_scatter_points(const float4* __restrict__ pts, const int32_t* __restrict__ sorted_vals, int n,
                                 const int32_t* __restrict__ nruns, const int32_t* __restrict__ run_off,
                                 const int32_t* __restrict__ run_dst, int base_ord, float4* pool) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = *nruns;  // find last r with run_off[r] <= i
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(run_off + mid) <= i) lo = mid; else hi = mid;
    }
    int dst0 = run_dst[lo];
    if (dst0 < 0) return;
    int src = sorted_vals[i];
    float4 p = pts[src];
    p.w = __int_as_float(base_ord + src);
    pool[dst0 + (i - run_off[lo])] = p;
}

// ---- sort-free insert for small batches (a scan's MapIncremental): four short kernels, no radix sort.
//   k_ins_slots    every point finds or creates its voxel (atomicCAS on the key) and takes an arrival position among the
//                  batch points of that voxel (bcnt[slot]++, all zero between batches); position 0 = the voxel's leader
//   k_ins_reserve  leaders make room for old + new points (bump allocation, doubling; old points move along)
//   k_ins_scatter  every point is written behind the voxel's old points at its arrival position
//   k_ins_finish   leaders put the new segment into insertion order (sort by ordinal: arrival order is not deterministic,
//                  the result is), publish the new count and clear bcnt
// The final table and pool hold exactly what the sorted path produces: voxel-contiguous runs in insertion order.
__global__ void k_ins_slots(const float4* __restrict__ pts, int n, const int32_t* __restrict__ d_count, float inv_res, int base_ord,
                            MapEntry* ent, int2* aux, uint32_t tmask, uint32_t capacity_voxels, int32_t* bcnt, MapCounters* ctr,
                            int32_t* __restrict__ p_slot, int32_t* __restrict__ p_pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    p_slot[i] = -1;
    if (d_count && i >= *d_count) return;
    const float4 p = pts[i];
    const int cx = pos2cell(p.x, inv_res), cy = pos2cell(p.y, inv_res), cz = pos2cell(p.z, inv_res);
    if (!cell_in_range(cx, cy, cz) || !(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) {
        atomicAdd(&ctr->err_range, 1u);
        return;
    }
    const uint64_t key = pack_key(cx, cy, cz);
    uint32_t slot = hash_key(key) & tmask;
    while (true) {
        const uint64_t k = ent[slot].key;
        if (k == key) break;
        if (k == kEmptyKey) {
            const uint64_t old = atomicCAS((unsigned long long*)&ent[slot].key, (unsigned long long)kEmptyKey, (unsigned long long)key);
            if (old == kEmptyKey) {
                const unsigned nv = atomicAdd(&ctr->num_voxels, 1u) + 1u;
                if (nv >= capacity_voxels) atomicAdd(&ctr->err_capacity, 1u);
                break;
            }
            if (old == key) break;
        }
        slot = next_slot(slot, tmask);
    }
    p_slot[i] = (int)slot;
    p_pos[i] = atomicAdd(&bcnt[slot], 1);
    atomicMax(&aux[slot].y, base_ord + i);  // LRU recency = ordinal of the last point that touched the voxel
}
__global__ void k_ins_reserve(int n, const int32_t* __restrict__ p_slot, const int32_t* __restrict__ p_pos, MapEntry* ent, int2* aux,
                              const int32_t* __restrict__ bcnt, uint64_t pool_cap, MapCounters* ctr, float4* pool) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int slot = p_slot[i];
    if (slot < 0 || p_pos[i] != 0) return;
    const int c = bcnt[slot];
    const int start = ent[slot].start, count = ent[slot].count, cap = aux[slot].x;
    const int newcount = count + c;
    if (newcount > cap) {
        const int newcap = cap == 0 ? newcount : max(2 * cap, newcount);
        const unsigned long long st = atomicAdd(&ctr->pool_top, (unsigned long long)newcap);
        if (st + (unsigned long long)newcap > pool_cap) {
            atomicAdd(&ctr->err_pool, 1u);
            return;
        }
        for (int j = 0; j < count; ++j) pool[st + j] = pool[start + j];
        ent[slot].start = (int)st;
        aux[slot].x = newcap;
    }
    atomicAdd(&ctr->live_points, (unsigned long long)c);
    if ((unsigned)newcount > ctr->max_count) atomicMax(&ctr->max_count, (unsigned)newcount);
}
__global__ void k_ins_scatter(const float4* __restrict__ pts, int n, const int32_t* __restrict__ p_slot, const int32_t* __restrict__ p_pos,
                              const MapEntry* __restrict__ ent, int base_ord, float4* pool) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int slot = p_slot[i];
    if (slot < 0) return;
    float4 p = pts[i];
    p.w = __int_as_float(base_ord + i);
    pool[ent[slot].start + ent[slot].count + p_pos[i]] = p;
}
constexpr int kInsSmallRun = 16;  // new points per voxel up to which the leader orders the segment itself
__global__ void k_ins_finish(int n, const int32_t* __restrict__ p_slot, const int32_t* __restrict__ p_pos, MapEntry* ent, int32_t* bcnt,
                             float4* pool, int32_t* __restrict__ big, int32_t* __restrict__ nbig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int slot = p_slot[i];
    if (slot < 0 || p_pos[i] != 0) return;
    const int c = bcnt[slot];
    if (c > kInsSmallRun) {  // a voxel that took many points of this batch: ordered by a whole block (k_ins_finish_big)
        big[atomicAdd(nbig, 1)] = slot;
        return;
    }
    const int count = ent[slot].count;
    float4* seg = pool + ent[slot].start + count;
    for (int a = 1; a < c; ++a) {  // insertion sort by ordinal
        const float4 v = seg[a];
        const int ov = __float_as_int(v.w);
        int b = a - 1;
        while (b >= 0 && __float_as_int(seg[b].w) > ov) { seg[b + 1] = seg[b]; --b; }
        seg[b + 1] = v;
    }
    ent[slot].count = count + c;
    bcnt[slot] = 0;
}
// One block per listed voxel: the rank of a new point inside its voxel = the number of batch points of the voxel with a
// smaller batch index; a bitmap over the batch indices (shared memory) turns that into a prefix popcount.
constexpr int kInsBitmapWords = 65536 / 32;
__global__ void __launch_bounds__(256) k_ins_finish_big(const int32_t* __restrict__ big, const int32_t* __restrict__ nbig, MapEntry* ent,
                                                        int32_t* bcnt, float4* pool, float4* __restrict__ scratch, int base_ord) {
    __shared__ uint32_t bits[kInsBitmapWords];
    __shared__ int pre[kInsBitmapWords];
    __shared__ int wsum[8];
    const int tid = threadIdx.x;
    for (int v = blockIdx.x; v < *nbig; v += gridDim.x) {
        const int slot = big[v];
        const int c = bcnt[slot], count = ent[slot].count;
        float4* seg = pool + ent[slot].start + count;
        float4* tmp = scratch + (size_t)blockIdx.x * 65536;  // per block: a block handles one listed voxel at a time
        for (int w = tid; w < kInsBitmapWords; w += 256) bits[w] = 0u;
        __syncthreads();
        for (int e = tid; e < c; e += 256) {
            const float4 p = seg[e];
            tmp[e] = p;
            const int b = __float_as_int(p.w) - base_ord;
            atomicOr(&bits[b >> 5], 1u << (b & 31));
        }
        __syncthreads();
        {  // exclusive prefix of the word popcounts: 8 consecutive words per thread
            int loc[8], sum = 0;
            for (int k = 0; k < 8; ++k) { loc[k] = sum; sum += __popc(bits[tid * 8 + k]); }
            int incl = sum;
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, o);
                if ((tid & 31) >= o) incl += u;
            }
            if ((tid & 31) == 31) wsum[tid >> 5] = incl;
            __syncthreads();
            int woff = 0;
            for (int w = 0; w < (tid >> 5); ++w) woff += wsum[w];
            const int base = woff + incl - sum;
            for (int k = 0; k < 8; ++k) pre[tid * 8 + k] = base + loc[k];
        }
        __syncthreads();
        for (int e = tid; e < c; e += 256) {
            const float4 p = tmp[e];
            const int b = __float_as_int(p.w) - base_ord;
            seg[pre[b >> 5] + __popc(bits[b >> 5] & ((1u << (b & 31)) - 1u))] = p;
        }
        __syncthreads();
        if (tid == 0) {
            ent[slot].count = count + c;
            bcnt[slot] = 0;
        }
        __syncthreads();
    }
}

// ---- LRU eviction (IVox::AddPoints, ivox3d.h:268-275): a new voxel that brings the map to `capacity_` voxels evicts the
// least recently touched one.  Only batches that can reach the capacity take this path.
// run_slot[r] = table slot of the run's voxel or -1 when the voxel does not exist yet; run_first[r] = index (in the batch)
// of the run's first point = the time the voxel is first touched.
__global__ void k_lookup_runs(const uint64_t* __restrict__ uniq, const int32_t* __restrict__ run_off, const int32_t* __restrict__ sorted_vals,
                              const int32_t* __restrict__ nruns, const MapEntry* __restrict__ ent, uint32_t tmask,
                              int32_t* __restrict__ run_slot, int32_t* __restrict__ run_first) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *nruns) return;
    const uint64_t key = uniq[r];
    uint32_t slot = hash_key(key) & tmask;
    int found = -1;
    if (key == kDropKey) {  // dropped points create nothing and touch nothing
        run_slot[r] = -2;
        run_first[r] = 0;
        return;
    }
    while (true) {
        const uint64_t k = ent[slot].key;
        if (k == key) { found = (int)slot; break; }
        if (k == kEmptyKey) break;
        slot = next_slot(slot, tmask);
    }
    run_slot[r] = found;
    run_first[r] = sorted_vals[run_off[r]];
}
// (stamp << 32 | slot) of every live voxel
__global__ void k_collect_live(const MapEntry* __restrict__ ent, const int2* __restrict__ aux, uint32_t tsize, uint64_t* __restrict__ out,
                               unsigned int* __restrict__ n_out) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const uint64_t k = s < tsize ? ent[s].key : kEmptyKey;
    const bool is_live = k != kEmptyKey && k != kTombKey;
    const unsigned live = __ballot_sync(0xffffffffu, is_live);  // one atomic per warp, not per voxel
    if (!live) return;
    const int leader = __ffs(live) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(n_out, (unsigned int)__popc(live));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!is_live) return;
    const unsigned int i = base + (unsigned int)__popc(live & ((1u << lane) - 1u));
    out[i] = ((uint64_t)(uint32_t)aux[s].y << 32) | (uint64_t)s;
}
__global__ void k_evict(const int32_t* __restrict__ victims, int nv, MapEntry* ent, int2* aux, MapCounters* ctr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    const int s = victims[i];
    atomicAdd(&ctr->live_points, (unsigned long long)(-(long long)ent[s].count));
    atomicSub(&ctr->num_voxels, 1u);
    ent[s].key = kTombKey;   // probe chains through this slot stay intact; the pool run is reclaimed by the next compaction
    ent[s].start = 0;
    ent[s].count = 0;
    aux[s] = make_int2(0, 0);
}
// move every live voxel into a fresh table (drops the tombstones)
__global__ void k_rehash(const MapEntry* __restrict__ old_ent, const int2* __restrict__ old_aux, uint32_t tsize, MapEntry* ent, int2* aux) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= tsize) return;
    const MapEntry e = old_ent[s];
    if (e.key == kEmptyKey || e.key == kTombKey) return;
    uint32_t slot = hash_key(e.key) & (tsize - 1);
    while (atomicCAS((unsigned long long*)&ent[slot].key, (unsigned long long)kEmptyKey, (unsigned long long)e.key) != kEmptyKey)
        slot = next_slot(slot, tsize - 1);
    ent[slot].start = e.start;
    ent[slot].count = e.count;
    aux[slot] = old_aux[s];
}

__global__ void k_fill_keys(MapEntry* ent, int2* aux, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        MapEntry e;
        e.key = kEmptyKey;
        e.start = 0;
        e.count = 0;
        ent[i] = e;
        aux[i] = make_int2(0, 0);
    }
}

// pool compaction: every live voxel gets a fresh exact-fit run in a new pool
__global__ void k_compact_plan(MapEntry* ent, int2* aux, uint32_t tsize, unsigned long long* top,
                               const float4* __restrict__ old_pool, float4* new_pool) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= tsize) return;
    if (ent[s].key == kEmptyKey || ent[s].key == kTombKey) return;
    int4 v = make_int4(ent[s].start, ent[s].count, aux[s].x, aux[s].y);
    if (v.y == 0) return;
    int slack = v.y < 4 ? v.y : v.y / 2;  // leave growth room so the next insert does not move everything again
    unsigned long long st = atomicAdd(top, (unsigned long long)(v.y + slack));
    for (int j = 0; j < v.y; ++j) new_pool[st + j] = old_pool[v.x + j];
    ent[s].start = (int)st;
    aux[s].x = v.y + slack;
}

// ------------------------------------------------------------------ standalone k=5 search
template <int G, int MODE>
__global__ void __launch_bounds__(256) k_knn5(MapView m, const float4* __restrict__ q, int n, int32_t* __restrict__ idx,
                                               float* __restrict__ d2, int32_t* __restrict__ cnt) {
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lg = threadIdx.x % G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    if (gid >= n) return;  // whole groups leave together
    float4 p = __ldg(q + gid);
    uint64_t w;
    float4 mine;
    __shared__ uint2 s_flat[MODE == 5 ? (256 / G) * kFlatStride : 1];
    int c = knn5_group<G, MODE>(m, p.x, p.y, p.z, lg, gmask, lane_stencil<G>(lg, m.nstencil), w, mine,
                                MODE == 5 ? s_flat + (threadIdx.x / G) * kFlatStride : nullptr);
    if (lg < 5) {
        idx[gid * 5 + lg] = __float_as_int(mine.w);
        d2[gid * 5 + lg] = (w == kInfKey) ? 0.0f : __uint_as_float((uint32_t)(w >> 32));
    }
    if (lg == 0) cnt[gid] = c;
}

// 8 lanes per query, balanced through shared memory (knn5_g8p)
__global__ void __launch_bounds__(256, 5) k_knn5_p(MapView m, const float4* __restrict__ q, int n, int32_t* __restrict__ idx,
                                                float* __restrict__ d2, int32_t* __restrict__ cnt) {
    __shared__ __align__(16) uint32_t s_q[(256 / 8) * kG8pWords];
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, lg = threadIdx.x & 7;
    const bool active = gid < n;   // groups past the end stay with their warp (full-mask collectives inside)
    const float4 p = active ? __ldg(q + gid) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 mine;
    uint32_t key;
    const int c = knn5_g8p(m, active, p.x, p.y, p.z, lg, s_q + (threadIdx.x >> 3) * kG8pWords, mine, key);
    if (!active) return;
    if (lg < 5) {
        idx[gid * 5 + lg] = __float_as_int(mine.w);
        d2[gid * 5 + lg] = (key == 0xffffffffu) ? 0.0f : __uint_as_float(key);
    }
    if (lg == 0) cnt[gid] = c;
}

// warp-per-query variant (knn5_warp)
__global__ void __launch_bounds__(256) k_knn5_w(MapView m, const float4* __restrict__ q, int n, int32_t* __restrict__ idx,
                                                float* __restrict__ d2, int32_t* __restrict__ cnt) {
    const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (qi >= n) return;  // whole warps leave together
    const float4 p = __ldg(q + qi);
    float4 mine;
    uint32_t key;
    const int c = knn5_warp(m, p.x, p.y, p.z, lane, mine, key);
    if (lane < 5) {
        idx[qi * 5 + lane] = __float_as_int(mine.w);
        d2[qi * 5 + lane] = (key == 0xffffffffu) ? 0.0f : __uint_as_float(key);
    }
    if (lane == 0) cnt[qi] = c;
}

// the same search, also returning the neighbours themselves (x, y, z, ordinal): b200_map_knn5_points
__global__ void __launch_bounds__(256) k_knn5_wx(MapView m, const float4* __restrict__ q, int n, int32_t* __restrict__ idx,
                                                 float* __restrict__ d2, int32_t* __restrict__ cnt, float4* __restrict__ nb) {
    const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (qi >= n) return;
    const float4 p = __ldg(q + qi);
    float4 mine;
    uint32_t key;
    const int c = knn5_warp(m, p.x, p.y, p.z, lane, mine, key);
    if (lane < 5) {
        idx[qi * 5 + lane] = __float_as_int(mine.w);
        d2[qi * 5 + lane] = (key == 0xffffffffu) ? 0.0f : __uint_as_float(key);
        nb[(size_t)qi * 5 + lane] = mine;
    }
    if (lane == 0) cnt[qi] = c;
}

// warp-per-query with TMA staging (knn5_warp_t<true>): every warp owns kStageCap float4 of shared memory and one mbarrier
__global__ void __launch_bounds__(256) k_knn5_t(MapView m, const float4* __restrict__ q, int n, int32_t* __restrict__ idx,
                                                float* __restrict__ d2, int32_t* __restrict__ cnt) {
    __shared__ __align__(16) float4 s_stage[8][kStageCap];
    __shared__ __align__(8) uint64_t s_bar[8];
    const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (qi >= n) return;  // whole warps leave together
    const float4 p = __ldg(q + qi);
    float4 mine;
    uint32_t key;
    const int c = knn5_warp_t<true>(m, p.x, p.y, p.z, lane, mine, key, s_stage[w], s_bar + w);
    if (lane < 5) {
        idx[qi * 5 + lane] = __float_as_int(mine.w);
        d2[qi * 5 + lane] = (key == 0xffffffffu) ? 0.0f : __uint_as_float(key);
    }
    if (lane == 0) cnt[qi] = c;
}

// number of map points resident in the occupied stencil cells of every query (the sum C_i of the roofline's
// algorithmic-byte formula), and the number of occupied cells
__global__ void k_stencil_points(MapView m, const float4* __restrict__ q, int n, unsigned long long* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long pts = 0, cells = 0;
    if (i < n) {
        const float4 p = q[i];
        const int kx = pos2cell(p.x, m.inv_res), ky = pos2cell(p.y, m.inv_res), kz = pos2cell(p.z, m.inv_res);
        for (int s = 0; s < m.nstencil; ++s) {
            const int cx = kx + c_stencil[s][0], cy = ky + c_stencil[s][1], cz = kz + c_stencil[s][2];
            if (!cell_in_range(cx, cy, cz)) continue;
            const uint64_t key = pack_key(cx, cy, cz);
            uint32_t slot = hash_key(key) & m.tmask;
            MapEntry e = ld_entry(m.ent + slot);
            while (e.key != key && e.key != kEmptyKey) {
                slot = next_slot(slot, m.tmask);
                e = ld_entry(m.ent + slot);
            }
            if (e.key == key) { pts += (unsigned long long)e.count; cells += 1; }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        pts += __shfl_xor_sync(0xffffffffu, pts, o);
        cells += __shfl_xor_sync(0xffffffffu, cells, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, pts);
        atomicAdd(out + 1, cells);
    }
}

__global__ void k_flush_fill(float4* buf, size_t n, float v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = make_float4(v, v, v, v);
}

// ------------------------------------------------------------------ host side
static uint32_t next_pow2(uint64_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

int32_t Map::init(const b200_map_params* p, int dev) {
    prm = *p;
    if (!(prm.resolution > 0.f)) B200_FAIL(B200_ERR_ARG, "resolution must be > 0");
    if (prm.max_range <= 0.f) prm.max_range = 5.0f;
    if (prm.capacity_voxels == 0) prm.capacity_voxels = 1000000;
    if (prm.max_points == 0) prm.max_points = 8u << 20;
    device = dev;
    CUDA_SET_DEVICE(dev);
    CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    inv_res = (float)(1.0 / (double)prm.resolution);  // ivox3d.h:65
    nstencil = prm.nearby == 0 ? 1 : prm.nearby == 6 ? 7 : prm.nearby == 26 ? 27 : 19;
    tsize = next_pow2(2 * prm.capacity_voxels);
    CUDA_TRY(cudaMalloc(&d_ent, (size_t)tsize * sizeof(MapEntry)));
    CUDA_TRY(cudaMalloc(&d_aux, (size_t)tsize * sizeof(int2)));
    pool_cap = 4 * prm.max_points;
    CUDA_TRY(cudaMalloc(&d_pool, pool_cap * sizeof(float4)));
    CUDA_TRY(cudaMalloc(&d_ctr, sizeof(MapCounters)));
    CUDA_TRY(cudaMemsetAsync(d_ctr, 0, sizeof(MapCounters), stream));
    k_fill_keys<<<(tsize + 255) / 256, 256, 0, stream>>>(d_ent, d_aux, tsize);
    LAUNCH_COUNT(1);
    CUDA_TRY(h_ctr_pin.reserve(1));
    CUDA_TRY(d_nruns.reserve(4));
    CUDA_TRY(h_small.reserve(4));
    CUDA_TRY(cudaStreamSynchronize(stream));
    memset(&h_ctr, 0, sizeof h_ctr);
    return B200_OK;
}

// IVox has no clear(); the LOAM front end rebuilds its two maps for every scan (kd-tree setInputCloud)
int32_t Map::clear() {
    CUDA_SET_DEVICE(device);
    CUDA_TRY(cudaMemsetAsync(d_ctr, 0, sizeof(MapCounters), stream));
    k_fill_keys<<<(tsize + 255) / 256, 256, 0, stream>>>(d_ent, d_aux, tsize);
    LAUNCH_COUNT(1);
    memset(&h_ctr, 0, sizeof h_ctr);
    next_ord = 0;
    return B200_OK;
}

void Map::destroy() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    cudaFree(d_ent); cudaFree(d_aux); cudaFree(d_pool); cudaFree(d_ctr); cudaFree(d_bcnt);
    in_pts.release(); k_in.release(); k_out.release(); k_uniq.release();
    v_in.release(); v_out.release(); run_cnt.release(); run_off.release(); run_dst.release(); run_reloc.release();
    d_nruns.release(); cub_tmp.release(); h_stage.release(); h_ctr_pin.release();
    lru_in.release(); lru_out.release(); victims_dev.release(); h_runs.release(); h_lru.release(); h_small.release();
    q_idx.release(); q_cnt.release(); q_d2.release(); q_nb.release();
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
    stream = nullptr;
}

int32_t Map::grow_pool(uint64_t min_cap) {
    // compaction into a new (possibly larger) pool
    uint64_t live = h_ctr.live_points;
    uint64_t ncap = 4 * live + 4 * min_cap + (1u << 20);
    if (ncap < pool_cap) ncap = pool_cap;
    if (ncap > (uint64_t)INT32_MAX) {  // voxel runs address the pool with 32-bit offsets (MapEntry::start)
        ncap = (uint64_t)INT32_MAX;
        if (live + 2 * (live + min_cap) > ncap) B200_FAIL(B200_ERR_NOMEM, "point pool would exceed 2^31 entries");
    }
    float4* np = nullptr;
    CUDA_TRY(cudaMalloc(&np, ncap * sizeof(float4)));
    CUDA_TRY(cudaMemsetAsync(&d_ctr->pool_top, 0, sizeof(unsigned long long), stream));
    k_compact_plan<<<(tsize + 255) / 256, 256, 0, stream>>>(d_ent, d_aux, tsize, &d_ctr->pool_top, d_pool, np);
    LAUNCH_COUNT(1);
    CUDA_TRY(cudaMemcpyAsync(h_ctr_pin.p, d_ctr, sizeof(MapCounters), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    h_ctr.pool_top = h_ctr_pin.p->pool_top;
    cudaFree(d_pool);
    d_pool = np;
    pool_cap = ncap;
    return B200_OK;
}

constexpr int32_t kSplitBatch = 100;  // internal: evict_for_batch needs the batch replayed in smaller pieces
constexpr int64_t kFastInsertMax = 65536;  // batches up to this size take the sort-free insert path
constexpr int kBigBlocks = 16;             // blocks of its big-voxel ordering pass

int32_t Map::rehash() {
    MapEntry* ne = nullptr;
    int2* na = nullptr;
    CUDA_TRY(cudaMalloc(&ne, (size_t)tsize * sizeof(MapEntry)));
    CUDA_TRY(cudaMalloc(&na, (size_t)tsize * sizeof(int2)));
    k_fill_keys<<<(tsize + 255) / 256, 256, 0, stream>>>(ne, na, tsize);
    k_rehash<<<(tsize + 255) / 256, 256, 0, stream>>>(d_ent, d_aux, tsize, ne, na);
    LAUNCH_COUNT(2);
    CUDA_TRY(cudaStreamSynchronize(stream));
    cudaFree(d_ent);
    cudaFree(d_aux);
    d_ent = ne;
    d_aux = na;
    tombstones = 0;
    return B200_OK;
}

// Decides, with the reference's sequential semantics, which voxels the batch that has just been sorted into runs evicts.
// Walks the creation events of the batch in time order on the host (a few hundred per scan) against the LRU order of the
// live voxels (sorted on the device).  A voxel that is evicted before its first touch in this batch is re-created by that
// touch with only the new points, exactly like the reference.
int32_t Map::evict_for_batch(int64_t n) {
    CUDA_TRY(cudaMemcpyAsync(h_small.p, d_nruns.p, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    const int nruns = h_small.p[0];
    const int nb = (nruns + 255) / 256;
    // run_dst / run_reloc are free until the upsert: borrow them for (slot, first)
    k_lookup_runs<<<nb, 256, 0, stream>>>(k_uniq.p, run_off.p, v_out.p, d_nruns.p, d_ent, tsize - 1, run_dst.p, run_reloc.p);
    LAUNCH_COUNT(1);
    CUDA_TRY(h_runs.reserve(2 * (size_t)nruns));
    CUDA_TRY(cudaMemcpyAsync(h_runs.p, run_dst.p, nruns * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(h_runs.p + nruns, run_reloc.p, nruns * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    const int32_t *slot = h_runs.p, *first = h_runs.p + nruns;
    int64_t n_new = 0;
    for (int r = 0; r < nruns; ++r) n_new += slot[r] == -1;
    const int64_t capacity = (int64_t)prm.capacity_voxels;
    if ((int64_t)h_ctr.num_voxels + n_new < capacity) return B200_OK;
    // LRU order of the live voxels
    const size_t live = h_ctr.num_voxels;
    CUDA_TRY(lru_in.reserve(live + 1)); CUDA_TRY(lru_out.reserve(live + 1));
    unsigned int* d_cnt = (unsigned int*)(d_nruns.p + 1);
    CUDA_TRY(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned int), stream));
    k_collect_live<<<(tsize + 255) / 256, 256, 0, stream>>>(d_ent, d_aux, tsize, lru_in.p, d_cnt);
    LAUNCH_COUNT(1);
    size_t tmp = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp, lru_in.p, lru_out.p, (int)live, 0, 64, stream);
    CUDA_TRY(cub_tmp.reserve(tmp));
    CUDA_TRY(cub::DeviceRadixSort::SortKeys(cub_tmp.p, tmp, lru_in.p, lru_out.p, (int)live, 0, 64, stream));
    const size_t m = std::min(live, (size_t)2 * nruns + 16);
    CUDA_TRY(h_lru.reserve(m + 1));
    CUDA_TRY(cudaMemcpyAsync(h_lru.p, lru_out.p, m * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    std::unordered_map<int32_t, std::pair<int, int64_t>> touched;  // old voxel slot -> (run, first touch time)
    using Ev = std::pair<int64_t, int>;                            // (time, run)
    std::priority_queue<Ev, std::vector<Ev>, std::greater<Ev>> heap;
    for (int r = 0; r < nruns; ++r) {
        if (slot[r] == -1) heap.push({(int64_t)first[r], r});
        else if (slot[r] >= 0) touched[slot[r]] = {r, (int64_t)first[r]};
    }
    std::vector<int32_t> victims;
    int64_t size = (int64_t)h_ctr.num_voxels;
    size_t i = 0;
    while (!heap.empty()) {
        const int64_t t = heap.top().first;
        heap.pop();
        size += 1;
        if (size < capacity) continue;
        while (i < m) {  // voxels touched earlier in this batch have already moved to the front of the LRU list
            auto it = touched.find((int32_t)(h_lru.p[i] & 0xFFFFFFFFu));
            if (it != touched.end() && it->second.second < t) { ++i; continue; }
            break;
        }
        if (i >= m) return kSplitBatch;  // every older voxel is gone: the victim is a voxel of this very batch -> smaller batches
        const int32_t victim = (int32_t)(h_lru.p[i] & 0xFFFFFFFFu);
        ++i;
        victims.push_back(victim);
        size -= 1;
        auto it = touched.find(victim);
        if (it != touched.end()) heap.push({it->second.second, it->second.first});  // re-created by its first touch
    }
    if (victims.empty()) return B200_OK;
    CUDA_TRY(h_runs.reserve(victims.size()));
    memcpy(h_runs.p, victims.data(), victims.size() * sizeof(int32_t));
    CUDA_TRY(victims_dev.reserve(victims.size()));
    CUDA_TRY(cudaMemcpyAsync(victims_dev.p, h_runs.p, victims.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    k_evict<<<(unsigned)((victims.size() + 255) / 256), 256, 0, stream>>>(victims_dev.p, (int)victims.size(), d_ent, d_aux, d_ctr);
    LAUNCH_COUNT(1);
    CUDA_TRY(cudaStreamSynchronize(stream));
    tombstones += victims.size();
    evicted_total += victims.size();
    if (h_ctr.num_voxels + tombstones + (uint64_t)n > (uint64_t)(0.7 * tsize)) return rehash();
    return B200_OK;
}

int32_t Map::insert_device(const float4* d_pts, int64_t n, const int32_t* d_count, const int32_t* h_count) {
    if (n == 0) return B200_OK;
    if (n > (int64_t)0x3fffffff) B200_FAIL(B200_ERR_ARG, "batch too large");
    // insertion ordinals and LRU stamps are 32-bit on the device (pool.w, aux.y): refuse before they wrap (hours of continuous
    // LIO at 200k points/s) instead of handing out negative ordinals; b200_map_clear / a fresh map restarts the count
    if (next_ord + n > (int64_t)INT32_MAX) B200_FAIL(B200_ERR_RANGE, "insertion ordinals exhausted (2^31 points inserted): rebuild the map");
    CUDA_SET_DEVICE(device);
    // worst case for this batch: every touched voxel relocates and doubles -> <= 2*(live + n) new slots
    if (h_ctr.pool_top + 2 * (h_ctr.live_points + (uint64_t)n) > pool_cap) {
        int32_t rc = grow_pool((uint64_t)n);
        if (rc) return rc;
    }
    // small batch that cannot reach the voxel capacity: the sort-free path (four kernels instead of ~17)
    static const bool fast_ok = !(getenv("B200_INSERT_FAST") && atoi(getenv("B200_INSERT_FAST")) == 0);
    if (fast_ok && n <= kFastInsertMax && h_ctr.num_voxels + tombstones + (uint64_t)n < prm.capacity_voxels) {
        CUDA_TRY(v_in.reserve(n)); CUDA_TRY(v_out.reserve(n));
        if (!d_bcnt) {
            CUDA_TRY(cudaMalloc(&d_bcnt, (size_t)tsize * sizeof(int32_t)));
            CUDA_TRY(cudaMemsetAsync(d_bcnt, 0, (size_t)tsize * sizeof(int32_t), stream));
        }
        const int nbf = (int)((n + 255) / 256);
        CUDA_TRY(cudaMemsetAsync(&d_ctr->err_range, 0, 3 * sizeof(unsigned int), stream));
        k_ins_slots<<<nbf, 256, 0, stream>>>(d_pts, (int)n, d_count, inv_res, (int)next_ord, d_ent, d_aux, tsize - 1, (uint32_t)prm.capacity_voxels,
                                             d_bcnt, d_ctr, v_in.p, v_out.p);
        k_ins_reserve<<<nbf, 256, 0, stream>>>((int)n, v_in.p, v_out.p, d_ent, d_aux, d_bcnt, pool_cap, d_ctr, d_pool);
        k_ins_scatter<<<nbf, 256, 0, stream>>>(d_pts, (int)n, v_in.p, v_out.p, d_ent, (int)next_ord, d_pool);
        int32_t* d_nbig = d_nruns.p + 2;
        CUDA_TRY(run_dst.reserve(n));
        CUDA_TRY(k_in.reserve((size_t)kBigBlocks * 65536 * 2));  // scratch of the big-voxel pass: 65536 float4 per block
        CUDA_TRY(cudaMemsetAsync(d_nbig, 0, sizeof(int32_t), stream));
        k_ins_finish<<<nbf, 256, 0, stream>>>((int)n, v_in.p, v_out.p, d_ent, d_bcnt, d_pool, run_dst.p, d_nbig);
        k_ins_finish_big<<<kBigBlocks, 256, 0, stream>>>(run_dst.p, d_nbig, d_ent, d_bcnt, d_pool, (float4*)k_in.p, (int)next_ord);
        LAUNCH_COUNT(5);
        CUDA_TRY(cudaMemcpyAsync(h_ctr_pin.p, d_ctr, sizeof(MapCounters), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        CUDA_TRY(cudaGetLastError());
        h_ctr = *h_ctr_pin.p;
        next_ord += d_count ? (int64_t)*h_count : n;
        h_ctr.num_points = (unsigned long long)next_ord;
        dropped_last = h_ctr.err_range;
        dropped_total += h_ctr.err_range;
        if (h_ctr.err_pool) B200_FAIL(B200_ERR_NOMEM, "point pool exhausted");
        if (h_ctr.err_capacity) B200_FAIL(B200_ERR_CAPACITY, "voxel capacity exceeded (internal: fast insert taken too close to the capacity)");
        return B200_OK;
    }
    CUDA_TRY(k_in.reserve(n)); CUDA_TRY(k_out.reserve(n)); CUDA_TRY(k_uniq.reserve(n));
    CUDA_TRY(v_in.reserve(n)); CUDA_TRY(v_out.reserve(n));
    CUDA_TRY(run_cnt.reserve(n)); CUDA_TRY(run_off.reserve(n)); CUDA_TRY(run_dst.reserve(n)); CUDA_TRY(run_reloc.reserve(2 * n));
    const int nb = (int)((n + 255) / 256);
    // the three error counters describe ONE batch (they used to be cumulative: one bad point failed every later insert)
    CUDA_TRY(cudaMemsetAsync(&d_ctr->err_range, 0, 3 * sizeof(unsigned int), stream));
    k_point_keys<<<nb, 256, 0, stream>>>(d_pts, (int)n, d_count, inv_res, k_in.p, v_in.p, d_ctr);
    size_t t1 = 0, t2 = 0, t3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, k_in.p, k_out.p, v_in.p, v_out.p, (int)n, 0, 64, stream);
    cub::DeviceRunLengthEncode::Encode(nullptr, t2, k_out.p, k_uniq.p, run_cnt.p, d_nruns.p, (int)n, stream);
    cub::DeviceScan::ExclusiveSum(nullptr, t3, run_cnt.p, run_off.p, (int)n, stream);
    size_t tmp = t1 > t2 ? t1 : t2;
    tmp = tmp > t3 ? tmp : t3;
    CUDA_TRY(cub_tmp.reserve(tmp));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp, k_in.p, k_out.p, v_in.p, v_out.p, (int)n, 0, 64, stream));
    CUDA_TRY(cub::DeviceRunLengthEncode::Encode(cub_tmp.p, tmp, k_out.p, k_uniq.p, run_cnt.p, d_nruns.p, (int)n, stream));
    // exclusive scan over all n slots (entries past nruns are garbage and never read)
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp.p, tmp, run_cnt.p, run_off.p, (int)n, stream));
    int32_t* run_oldcnt = run_reloc.p + n;
    if (h_ctr.num_voxels + (uint64_t)n >= prm.capacity_voxels) {  // this batch may reach the voxel capacity: LRU eviction
        int32_t rc = evict_for_batch(n);
        if (rc == kSplitBatch) {  // nothing has been modified yet: replay the batch as two halves (exact, down to single points)
            if (n == 1) B200_FAIL(B200_ERR_CAPACITY, "voxel capacity too small");
            if (d_count) {  // the split needs the real size on the host
                CUDA_TRY(cudaStreamSynchronize(stream));
                n = *h_count;
                if (n <= 1) B200_FAIL(B200_ERR_CAPACITY, "voxel capacity too small");
            }
            rc = insert_device(d_pts, n / 2);
            return rc ? rc : insert_device(d_pts + n / 2, n - n / 2);
        }
        if (rc) return rc;
    }
    k_upsert_runs<<<nb, 256, 0, stream>>>(k_uniq.p, run_cnt.p, d_nruns.p, run_off.p, v_out.p, (int)next_ord, d_ent, d_aux, tsize - 1, pool_cap,
                                          (uint32_t)prm.capacity_voxels, d_ctr, run_dst.p, run_reloc.p, run_oldcnt);
    k_relocate<<<nb, 256, 0, stream>>>(d_nruns.p, run_dst.p, run_reloc.p, run_oldcnt, d_pool);
    k_scatter_points<<<nb, 256, 0, stream>>>(d_pts, v_out.p, (int)n, d_nruns.p, run_off.p, run_dst.p, (int)next_ord, d_pool);
    LAUNCH_COUNT(4);  // own kernels (cub's sort/RLE/scan passes are not counted)
    CUDA_TRY(cudaMemcpyAsync(h_ctr_pin.p, d_ctr, sizeof(MapCounters), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    h_ctr = *h_ctr_pin.p;
    next_ord += d_count ? (int64_t)*h_count : n;
    h_ctr.num_points = (unsigned long long)next_ord;
    dropped_last = h_ctr.err_range;  // non-fatal: the rest of the batch is in the map
    dropped_total += h_ctr.err_range;
    if (h_ctr.err_pool) B200_FAIL(B200_ERR_NOMEM, "point pool exhausted");
    if (h_ctr.err_capacity) B200_FAIL(B200_ERR_CAPACITY, "voxel capacity exceeded (internal: eviction pre-pass missed a batch)");
    return B200_OK;
}

int32_t Map::insert_host(const float* xyz, int64_t n, int64_t stride) {
    if (n == 0) return B200_OK;
    if (n < 0 || !xyz || stride < 12) B200_FAIL(B200_ERR_ARG, "bad point buffer");
    CUDA_SET_DEVICE(device);
    CUDA_TRY(h_stage.reserve(n));
    CUDA_TRY(in_pts.reserve(n));
    pack_xyz_float4(xyz, n, stride, h_stage.p);
    CUDA_TRY(cudaMemcpyAsync(in_pts.p, h_stage.p, n * sizeof(float4), cudaMemcpyHostToDevice, stream));
    return insert_device(in_pts.p, n);
}

int32_t Map::knn5_host(const float* xyz, int64_t n, int64_t stride, int32_t* idx, float* d2, int32_t* cnt, float* nb_xyz) {
    if (n == 0) return B200_OK;
    if (n < 0 || !xyz || !idx || !d2 || !cnt || stride < 12) B200_FAIL(B200_ERR_ARG, "bad query buffer");
    CUDA_SET_DEVICE(device);
    CUDA_TRY(h_stage.reserve(nb_xyz ? n * 5 : n));  // queries in, and (when asked for) 5 neighbour records per query out
    CUDA_TRY(in_pts.reserve(n));
    CUDA_TRY(q_idx.reserve(n * 5)); CUDA_TRY(q_d2.reserve(n * 5)); CUDA_TRY(q_cnt.reserve(n));
    pack_xyz_float4(xyz, n, stride, h_stage.p);
    CUDA_TRY(cudaMemcpyAsync(in_pts.p, h_stage.p, n * sizeof(float4), cudaMemcpyHostToDevice, stream));
    constexpr int G = 8;
    if (!ev0) { CUDA_TRY(cudaEventCreate(&ev0)); CUDA_TRY(cudaEventCreate(&ev1)); }
    CUDA_TRY(cudaEventRecord(ev0, stream));
    {
        int mode = knn_mode();
        if (nb_xyz) {  // neighbours' coordinates wanted as well: the warp-per-query body with one more store
            CUDA_TRY(q_nb.reserve(n * 5));
            k_knn5_wx<<<(unsigned)((n * 32 + 255) / 256), 256, 0, stream>>>(view(), in_pts.p, (int)n, q_idx.p, q_d2.p, q_cnt.p, q_nb.p);
        } else if (mode == 8) {  // 8 lanes per query, candidates balanced through shared memory
            k_knn5_p<<<(unsigned)((n * 8 + 255) / 256), 256, 0, stream>>>(view(), in_pts.p, (int)n, q_idx.p, q_d2.p, q_cnt.p);
        } else if (mode == 9) {  // one warp per query, runs staged in shared memory by 1-D TMA bulk copies
            k_knn5_t<<<(unsigned)((n * 32 + 255) / 256), 256, 0, stream>>>(view(), in_pts.p, (int)n, q_idx.p, q_d2.p, q_cnt.p);
        } else if (mode == 7) {  // one warp per query
            k_knn5_w<<<(unsigned)((n * 32 + 255) / 256), 256, 0, stream>>>(view(), in_pts.p, (int)n, q_idx.p, q_d2.p, q_cnt.p);
        } else {
            const unsigned grid = (unsigned)((n * G + 255) / 256);
            if (mode == 0 && n >= 200000 && !getenv("B200_KNN_MODE")) mode = 5;
            if (mode >= 5) k_knn5<G, 5><<<grid, 256, 0, stream>>>(view(), in_pts.p, (int)n, q_idx.p, q_d2.p, q_cnt.p);
            else if (mode == 4) k_knn5<G, 4><<<grid, 256, 0, stream>>>(view(), in_pts.p, (int)n, q_idx.p, q_d2.p, q_cnt.p);
            else if (mode == 1) k_knn5<G, 1><<<grid, 256, 0, stream>>>(view(), in_pts.p, (int)n, q_idx.p, q_d2.p, q_cnt.p);
            else k_knn5<G, 0><<<grid, 256, 0, stream>>>(view(), in_pts.p, (int)n, q_idx.p, q_d2.p, q_cnt.p);
        }
    }
    CUDA_TRY(cudaEventRecord(ev1, stream));
    LAUNCH_COUNT(1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(idx, q_idx.p, n * 5 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(d2, q_d2.p, n * 5 * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(cnt, q_cnt.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    if (nb_xyz) {
        // the H2D copy of the queries that used the stage completes in stream order before this copy starts
        CUDA_TRY(cudaMemcpyAsync(h_stage.p, q_nb.p, n * 5 * sizeof(float4), cudaMemcpyDeviceToHost, stream));
    }
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (nb_xyz)
        for (int64_t i = 0; i < n * 5; ++i) {
            nb_xyz[i * 3] = h_stage.p[i].x;
            nb_xyz[i * 3 + 1] = h_stage.p[i].y;
            nb_xyz[i * 3 + 2] = h_stage.p[i].z;
        }
    cudaEventElapsedTime(&last_knn_ms, ev0, ev1);
    return B200_OK;
}

}  // namespace b200

// ------------------------------------------------------------------ C ABI (B1)

extern "C" {

const char* b200_version(void) { return "b200reg 0.1 (sm_100a)"; }
const char* b200_last_error(void) { return b200::g_last_error.c_str(); }
int64_t b200_kernel_launches(void) { return b200::g_kernel_launches.load(); }

int32_t b200_map_create(const b200_map_params* params, int32_t device, b200_map** out) {
    if (!params || !out) B200_FAIL(B200_ERR_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) B200_FAIL(B200_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) B200_FAIL(B200_ERR_ARG, "bad device ordinal");
    b200_map* h = new b200_map();
    int32_t rc = h->m.init(params, device);
    if (rc != B200_OK) { h->m.destroy(); delete h; return rc; }
    *out = h;
    return B200_OK;
}
int32_t b200_map_destroy(b200_map* map) {
    if (!map) return B200_OK;
    if (map->refs > 0) {  // filters still attached: the last b200_iekf_destroy frees the map
        map->zombie = true;
        return B200_OK;
    }
    map->m.destroy();
    delete map;
    return B200_OK;
}
int32_t b200_map_insert(b200_map* map, const float* xyz, int64_t n, int64_t stride_bytes) {
    if (!map) B200_FAIL(B200_ERR_ARG, "null map");
    return map->m.insert_host(xyz, n, stride_bytes);
}
int32_t b200_map_knn5(b200_map* map, const float* xyz_world, int64_t n, int64_t stride_bytes, int32_t* idx, float* sqdist, int32_t* count) {
    if (!map) B200_FAIL(B200_ERR_ARG, "null map");
    return map->m.knn5_host(xyz_world, n, stride_bytes, idx, sqdist, count, nullptr);
}
int32_t b200_map_knn5_points(b200_map* map, const float* xyz_world, int64_t n, int64_t stride_bytes, int32_t* idx, float* sqdist, int32_t* count,
                             float* neighbours_xyz) {
    if (!map || !neighbours_xyz) B200_FAIL(B200_ERR_ARG, "null argument");
    return map->m.knn5_host(xyz_world, n, stride_bytes, idx, sqdist, count, neighbours_xyz);
}
/* bench/roofline helper: total map points and occupied cells in the stencils of n queries */
int32_t b200_map_stencil_points(b200_map* map, const float* xyz_world, int64_t n, int64_t stride_bytes, int64_t* points, int64_t* cells) {
    if (!map || !xyz_world || n < 1 || stride_bytes < 12) B200_FAIL(B200_ERR_ARG, "bad argument");
    b200::Map& m = map->m;
    CUDA_SET_DEVICE(m.device);
    CUDA_TRY(m.h_stage.reserve(n));
    CUDA_TRY(m.in_pts.reserve(n));
    b200::pack_xyz_float4(xyz_world, n, stride_bytes, m.h_stage.p);
    CUDA_TRY(cudaMemcpyAsync(m.in_pts.p, m.h_stage.p, n * sizeof(float4), cudaMemcpyHostToDevice, m.stream));
    unsigned long long* d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, 16));
    CUDA_TRY(cudaMemsetAsync(d_out, 0, 16, m.stream));
    b200::k_stencil_points<<<(unsigned)((n + 255) / 256), 256, 0, m.stream>>>(m.view(), m.in_pts.p, (int)n, d_out);
    unsigned long long h[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(h, d_out, 16, cudaMemcpyDeviceToHost, m.stream));
    CUDA_TRY(cudaStreamSynchronize(m.stream));
    cudaFree(d_out);
    if (points) *points = (int64_t)h[0];
    if (cells) *cells = (int64_t)h[1];
    return B200_OK;
}

/* bench helper: evict the L2 by streaming a buffer larger than it (256 MiB) on `device` */
int32_t b200_flush_l2(int32_t device) {
    static float4* buf[16] = {};
    const size_t n = (256u << 20) / sizeof(float4);
    if (device < 0 || device >= 16) B200_FAIL(B200_ERR_ARG, "bad device");
    CUDA_SET_DEVICE(device);
    if (!buf[device]) CUDA_TRY(cudaMalloc(&buf[device], n * sizeof(float4)));
    static float v = 0.f;
    v += 1.f;
    b200::k_flush_fill<<<1184, 256>>>(buf[device], n, v);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return B200_OK;
}

/* device time (CUDA events) of the search kernel of the last b200_map_knn5 call */
float b200_map_last_knn_ms(b200_map* map) { return map ? map->m.last_knn_ms : 0.f; }
int64_t b200_map_tma_timeouts(void) {
    unsigned int v = 0;
    if (cudaMemcpyFromSymbol(&v, b200::g_knn_tma_timeouts, sizeof v) != cudaSuccess) return -1;
    return (int64_t)v;
}
int64_t b200_map_evicted(b200_map* map) { return map ? (int64_t)map->m.evicted_total : 0; }
int64_t b200_map_dropped(b200_map* map, int64_t* last_batch) {
    if (!map) return 0;
    if (last_batch) *last_batch = (int64_t)map->m.dropped_last;
    return (int64_t)map->m.dropped_total;
}
int64_t b200_map_num_voxels(b200_map* map) { return map ? (int64_t)map->m.h_ctr.num_voxels : 0; }
int64_t b200_map_num_points(b200_map* map) { return map ? (int64_t)map->m.h_ctr.live_points : 0; }

}  // extern "C"
