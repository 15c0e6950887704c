// construct_full_map <poses.txt> <frames_dir> <out.pcd> [leaf=0.1] [device=0]
//
// Command-line stand-in for the reference's dynamic_map/construct_full_map (scripts/construct_full_map.sh:6; the
// reference ships the launcher but not the sources): every keyframe frames/<i>.pcd is moved by the i-th
// pose of poses.txt and merged into one voxel-grid map (leaf metres, centroid of x, y, z, intensity per voxel) on the
// GPU through the b200reg C ABI; the result is written as a binary PointXYZI .pcd.
#include <chrono>
#include <cstdio>

#include "b200reg.h"
#include "pcd_io.hpp"

int main(int argc, char** argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s <poses.txt> <frames_dir> <out.pcd> [leaf=0.1] [device=0] [capacity_voxels=64000000]\n", argv[0]);
        return 64;
    }
    using namespace b200host;
    try {
        const float leaf = argc > 4 ? (float)atof(argv[4]) : 0.1f;
        const int device = argc > 5 ? atoi(argv[5]) : 0;
        const uint64_t capacity = argc > 6 ? strtoull(argv[6], nullptr, 10) : 64000000ull;
        const auto poses = load_poses(argv[1]);
        const auto frames = list_frames(argv[2]);
        if (frames.size() != poses.size())
            fprintf(stderr, "warning: %zu frames but %zu poses; using the first %zu\n", frames.size(), poses.size(), std::min(frames.size(), poses.size()));
        const size_t n = std::min(frames.size(), poses.size());
        b200_mapbuild* mb = nullptr;
        if (b200_mapbuild_create(leaf, capacity, device, &mb) != B200_OK) throw std::runtime_error(b200_last_error());
        const auto t0 = std::chrono::steady_clock::now();
        size_t total = 0;
        for (size_t i = 0; i < n; ++i) {
            const std::vector<PointXYZI> cloud = load_pcd(frames[i]);
            if (cloud.empty()) continue;
            if (b200_mapbuild_add_keyframe(mb, &cloud[0].x, (int64_t)cloud.size(), sizeof(PointXYZI), poses[i].data()) != B200_OK)
                throw std::runtime_error(b200_last_error());
            total += cloud.size();
        }
        const int64_t m = b200_mapbuild_extract(mb, nullptr, nullptr, 0);
        if (m < 0) throw std::runtime_error(b200_last_error());
        std::vector<PointXYZI> map((size_t)m);
        if (m && b200_mapbuild_extract(mb, &map[0].x, nullptr, m) != m) throw std::runtime_error(b200_last_error());
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        save_pcd_binary(argv[3], map.data(), map.size());
        b200_mapbuild_destroy(mb);
        printf("construct_full_map: %zu keyframes, %zu points -> %lld voxels (leaf %.3f) in %.3f s -> %s\n", n, total, (long long)m, leaf, sec, argv[3]);
    } catch (const std::exception& e) {
        fprintf(stderr, "construct_full_map: %s\n", e.what());
        return 1;
    }
    return 0;
}
