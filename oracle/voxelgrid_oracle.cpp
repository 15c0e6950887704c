// TEST INFRASTRUCTURE ONLY — CPU oracle for pcl::VoxelGrid and the keyframe-merge map builder.
// PARITY UNPINNED (see oracle.h): PCL is a third-party dependency that is not in the tree (find_package(PCL 1.8),
// jueying_lio/cmake/packages.cmake:49); its VoxelGrid::applyFilter is restated from the vendored near-copy
// jueying_slam/include/voxel_grid_large.cpp:25-258 (same arithmetic, plus a split for huge clouds that is not needed
// here).  construct_full_map's sources are absent from the reference: only its command line
// (scripts/construct_full_map.sh:6: <poses.txt> <frames_dir> <out.pcd> <leaf 0.1>) and the on-disk formats are
// known, so "transform every keyframe by its pose, concatenate, VoxelGrid(leaf)" is this oracle's definition.
//
// Call sites restated: jueying_lio/src/laser_mapping.cc:323-328 (scan downsample, PointXYZINormal, leaf =
// filter_size_surf).  std::sort in applyFilter is not stable, so the order in which a voxel's points are summed is
// unspecified in the reference; the contract here is ascending input index (what a stable sort yields).
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct Pt4 { float x, y, z, i; };

// applyFilter with downsample_all_data_ = true (the constructor default, voxel_grid_large.h:61) on (x, y, z, intensity):
// pcl::CentroidPoint accumulates every field in float and divides by the count.  Output order = ascending leaf index.
int64_t voxel_grid(const std::vector<Pt4>& in, float leaf, unsigned min_points, std::vector<Pt4>& out, std::vector<int32_t>* counts) {
    out.clear();
    if (counts) counts->clear();
    const float inv = 1.0f / leaf;
    float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    float mx[3] = {-mn[0], -mn[1], -mn[2]};
    for (const Pt4& p : in) {
        if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
        mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
        mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
    }
    if (!(mn[0] <= mx[0])) return 0;
    const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1, dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
    // pcl::VoxelGrid warns and returns the input when dx*dy*dz overflows int32; VoxelGridLarge splits the cloud.  The
    // voxel partition itself (floor(p * inv) per axis) does not depend on that, so 64-bit indices are used here.
    (void)dx; (void)dy; (void)dz;
    int64_t min_b[3], div_b[3];
    for (int k = 0; k < 3; ++k) {
        min_b[k] = (int64_t)std::floor(mn[k] * inv);
        div_b[k] = (int64_t)std::floor(mx[k] * inv) - min_b[k] + 1;
    }
    struct Idx { int64_t idx; int32_t pt; };
    std::vector<Idx> iv;
    iv.reserve(in.size());
    for (size_t n = 0; n < in.size(); ++n) {
        const Pt4& p = in[n];
        if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
        const int64_t i0 = (int64_t)(std::floor(p.x * inv) - (float)min_b[0]);
        const int64_t i1 = (int64_t)(std::floor(p.y * inv) - (float)min_b[1]);
        const int64_t i2 = (int64_t)(std::floor(p.z * inv) - (float)min_b[2]);
        iv.push_back(Idx{i0 + i1 * div_b[0] + i2 * div_b[0] * div_b[1], (int32_t)n});
    }
    std::stable_sort(iv.begin(), iv.end(), [](const Idx& a, const Idx& b) { return a.idx < b.idx; });
    size_t index = 0;
    while (index < iv.size()) {
        size_t i = index + 1;
        while (i < iv.size() && iv[i].idx == iv[index].idx) ++i;
        if (i - index >= min_points) {
            float sx = 0, sy = 0, sz = 0, si = 0;
            for (size_t li = index; li < i; ++li) {
                const Pt4& p = in[iv[li].pt];
                sx += p.x; sy += p.y; sz += p.z; si += p.i;
            }
            const float n = (float)(i - index);
            out.push_back(Pt4{sx / n, sy / n, sz / n, si / n});
            if (counts) counts->push_back((int32_t)(i - index));
        }
        index = i;
    }
    return (int64_t)out.size();
}

// T = Translation(x, y, z) * Quaterniond(qw, qx, qy, qz) built in double, narrowed to a float 4x4 (poses.txt line:
// "x y z qw qx qy qz", tool/occupancy_mapping/src/mapping_server.cc:466-497); points moved by pcl::transformPointCloud
// (PCL 1.7/1.8 scalar form, float).
void pose_matrix(const double* p7, float M[12]) {
    const double w = p7[3], x = p7[4], y = p7[5], z = p7[6];
    const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                         2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) M[r * 4 + c] = (float)R[r * 3 + c];
        M[r * 4 + 3] = (float)p7[r];
    }
}

}  // namespace

extern "C" {

/* pcl::VoxelGrid::filter on (x, y, z, intensity) records; returns the number of output points (<= max written) */
int64_t orc_voxel_grid(const float* xyzi, int64_t n, int64_t stride, float leaf, int32_t min_points, float* out_xyzi, int32_t* out_count,
                       int64_t max) {
    std::vector<Pt4> in((size_t)n), out;
    for (int64_t k = 0; k < n; ++k) {
        const float* p = (const float*)((const char*)xyzi + k * stride);
        in[k] = Pt4{p[0], p[1], p[2], stride >= 16 ? p[3] : 0.0f};
    }
    std::vector<int32_t> cnt;
    const int64_t m = voxel_grid(in, leaf, (unsigned)min_points, out, &cnt);
    for (int64_t k = 0; k < m && k < max; ++k) {
        if (out_xyzi) memcpy(out_xyzi + k * 4, &out[k], 16);
        if (out_count) out_count[k] = cnt[k];
    }
    return m;
}

/* construct_full_map: keyframes (concatenated xyzi records, offsets[k]..offsets[k+1]) moved by poses7[k], merged, VoxelGrid(leaf) */
int64_t orc_full_map(const float* xyzi, const int64_t* offsets, int64_t n_frames, const double* poses7, float leaf, float* out_xyzi,
                     int32_t* out_count, int64_t max) {
    std::vector<Pt4> all, out;
    all.reserve((size_t)offsets[n_frames]);
    for (int64_t f = 0; f < n_frames; ++f) {
        float M[12];
        pose_matrix(poses7 + 7 * f, M);
        for (int64_t k = offsets[f]; k < offsets[f + 1]; ++k) {
            const float* p = xyzi + 4 * k;
            Pt4 q;
            q.x = ((M[0] * p[0] + M[1] * p[1]) + M[2] * p[2]) + M[3];
            q.y = ((M[4] * p[0] + M[5] * p[1]) + M[6] * p[2]) + M[7];
            q.z = ((M[8] * p[0] + M[9] * p[1]) + M[10] * p[2]) + M[11];
            q.i = p[3];
            all.push_back(q);
        }
    }
    std::vector<int32_t> cnt;
    const int64_t m = voxel_grid(all, leaf, 0, out, &cnt);
    for (int64_t k = 0; k < m && k < max; ++k) {
        if (out_xyzi) memcpy(out_xyzi + k * 4, &out[k], 16);
        if (out_count) out_count[k] = cnt[k];
    }
    return m;
}
}
