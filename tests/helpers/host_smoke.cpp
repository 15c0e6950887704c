// Compiles the header-only host adaptors against the C ABI and runs them end to end (used by tests/test_host_cpp.py).
// Without a GPU it must fail loudly at construction; with one it registers a small synthetic scene.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "esekf_gpu.hpp"
#include "gicp_gpu.hpp"
#include "ndt_gpu.hpp"
#include "scan2map_gpu.hpp"

struct PointXYZINormal {  // layout of pcl::PointXYZINormal: 48 bytes
    float x, y, z, pad0, normal_x, normal_y, normal_z, pad1, intensity, curvature, pad2, pad3;
};
struct Cloud { std::vector<PointXYZINormal> points; };

int main() {
    using namespace b200host;
    using IVoxType = IVox<3, IVoxNodeType::DEFAULT, PointXYZINormal>;
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> u(-20.f, 20.f), h(0.f, 6.f);
    std::normal_distribution<float> nz(0.f, 0.01f);
    // a room: floor z=0, ceiling z=6, walls x=+-20, y=+-20, plus a step so that x/y are observable
    std::vector<PointXYZINormal> map;
    auto add = [&](float x, float y, float z) { PointXYZINormal p{}; p.x = x; p.y = y; p.z = z; map.push_back(p); };
    for (int i = 0; i < 60000; ++i) { add(u(rng), u(rng), nz(rng)); add(u(rng), u(rng), 6.f + nz(rng)); }
    for (int i = 0; i < 20000; ++i) { add(20.f + nz(rng), u(rng), h(rng)); add(-20.f + nz(rng), u(rng), h(rng)); add(u(rng), 20.f + nz(rng), h(rng)); add(u(rng), -20.f + nz(rng), h(rng)); }
    try {
        IVoxType::Options opt;
        opt.resolution_ = 0.5f;
        opt.nearby_type_ = IVoxType::NearbyType::NEARBY18;
        IVoxType ivox(opt);
        ivox.AddPoints(map);
        std::vector<PointXYZINormal> near;
        PointXYZINormal q{}; q.x = 1.f; q.y = 2.f; q.z = 0.02f;
        const bool ok = ivox.GetClosestPoint(q, near);
        std::printf("ivox voxels %zu points %zu knn %d first d=%.4f\n", ivox.NumValidGrids(), ivox.NumPoints(), ok ? (int)near.size() : 0,
                    ok ? std::sqrt((near[0].x - q.x) * (near[0].x - q.x) + (near[0].y - q.y) * (near[0].y - q.y) + (near[0].z - q.z) * (near[0].z - q.z)) : -1.f);
        if (!ok || near.size() != 5) return 2;
        // scan = map points near the origin seen from a body frame shifted by (0.03, -0.02, 0.01)
        std::vector<PointXYZINormal> scan;
        for (size_t i = 0; i < map.size() && scan.size() < 4000; i += 17) { PointXYZINormal p = map[i]; p.x -= 0.03f; p.y += 0.02f; p.z -= 0.01f; scan.push_back(p); }
        Esekf<IVoxType> kf(ivox);
        StateVec x{}; x[6] = 1.0; x[10] = 1.0; x[25] = -9.809;
        CovMat P{};
        for (int i = 0; i < 23; ++i) P[i * 23 + i] = 0.01;
        kf.change_x(x);
        kf.change_P(P);
        double ms = 0;
        const bool valid = kf.update_iterated_dyn_share_modified(scan, ms);
        const StateVec& xo = kf.get_x();
        std::printf("iekf valid %d passes %d n_eff %d gpu_ms %.3f pos %.4f %.4f %.4f\n", (int)valid, kf.stats().passes, kf.stats().n_eff[0], ms, xo[0], xo[1], xo[2]);
        if (!valid || std::fabs(xo[0] - 0.03) > 0.01 || std::fabs(xo[1] + 0.02) > 0.01 || std::fabs(xo[2] - 0.01) > 0.01) return 3;
        {   // the same scan from a page-locked buffer (unpacked on the device): same posterior on a fresh filter
            PinnedScan<PointXYZINormal> pinned(scan.size());
            pinned.resize(scan.size());
            for (size_t i = 0; i < scan.size(); ++i) pinned[i] = scan[i];
            Esekf<decltype(ivox)> kf_pin(ivox);
            kf_pin.change_x(x);
            kf_pin.change_P(P);
            double ms2 = 0;
            Esekf<decltype(ivox)> kf_ref(ivox);
            kf_ref.change_x(x);
            kf_ref.change_P(P);
            if (!kf_pin.update_iterated_dyn_share_modified(pinned, ms2) || !kf_ref.update_iterated_dyn_share_modified(scan, ms2)) return 31;
            for (int i = 0; i < 26; ++i)
                if (kf_pin.get_x()[i] != kf_ref.get_x()[i]) return 32;
        }
        auto target = std::make_shared<Cloud>();
        target->points = map;
        auto source = std::make_shared<Cloud>();
        source->points = scan;
        NormalDistributionsTransform<Cloud> ndt;
        ndt.setResolution(2.0f);
        ndt.setTransformationEpsilon(0.001);
        ndt.setNeighborhoodSearchMethod(DIRECT7);
        ndt.setInputTarget(target);
        ndt.setInputSource(source);
        Cloud aligned;
        ndt.align(aligned);
        const float* T = ndt.getFinalTransformation();
        std::printf("ndt converged %d iters %d t %.4f %.4f %.4f prob %.4f\n", (int)ndt.hasConverged(), ndt.getFinalNumIteration(), T[12], T[13], T[14],
                    ndt.getTransformationProbability());
        if (!ndt.hasConverged() || aligned.points.size() != scan.size()) return 4;
        std::printf("fitness %.5f\n", ndt.getFitnessScore());
        {   // the "GICP_OMP" registration object on the same clouds
            GeneralizedIterativeClosestPoint<Cloud> gicp;
            gicp.setInputTarget(target);
            gicp.setInputSource(source);
            Cloud aligned2;
            gicp.align(aligned2);
            const float* G = gicp.getFinalTransformation();
            std::printf("gicp converged %d iters %d m %d t %.4f %.4f %.4f fitness %.6f\n", (int)gicp.hasConverged(), gicp.getFinalNumIteration(), gicp.result().last_m,
                        G[12], G[13], G[14], gicp.getFitnessScore());
            if (!gicp.hasConverged() || aligned2.points.size() != scan.size() || std::fabs(G[12] - 0.03f) > 0.02f || std::fabs(G[13] + 0.02f) > 0.02f ||
                std::fabs(G[14] - 0.01f) > 0.02f)
                return 8;
        }
        // scan pre-processing chain: downsample on the device, then the update straight from device memory
        ScanPreprocessor pre;
        double pose22[2 * 22] = {0};
        pose22[13] = pose22[17] = pose22[21] = 1.0;                       // identity rotation, no motion
        pose22[22] = 0.1; pose22[22 + 13] = pose22[22 + 17] = pose22[22 + 21] = 1.0;
        std::vector<PointXYZINormal> timed = scan;
        for (size_t i = 0; i < timed.size(); ++i) timed[i].curvature = 100.0f * (float)i / (float)timed.size();
        pre.undistort(timed, 9, 8, pose22, 2, x.data());
        const int64_t n_down = pre.filterStaged(0.5f);
        int64_t n_dev = 0;
        const void* d_scan = pre.devicePoints(&n_dev);
        b200_iekf_stats st{};
        StateVec x2 = x;
        CovMat P2 = P;
        if (n_down < 100 || n_dev != n_down || b200_iekf_update_device(kf.handle(), d_scan, n_dev, x2.data(), P2.data(), &st) != B200_OK) return 5;
        std::printf("preprocess %lld -> %lld points, update passes %d pos %.4f %.4f %.4f\n", (long long)timed.size(), (long long)n_down, st.passes, x2[0], x2[1], x2[2]);
        if (std::fabs(x2[0] - 0.03) > 0.01) return 6;
        // LOAM scan-to-map with the surface cloud as both feature maps' source
        ScanToMap<Cloud> s2m(400000);
        Cloud corner_map, corner_scan;   // no edge features in this scene: the surf features alone constrain the pose
        s2m.setInputCloud(corner_map, *target);
        float t6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const bool conv = s2m.scan2MapOptimization(corner_scan, *source, t6, 30);
        std::printf("scan2map converged %d iters %d n_sel %d t %.4f %.4f %.4f degenerate %d\n", (int)conv, s2m.stats().iters, s2m.stats().n_sel, t6[3], t6[4], t6[5],
                    (int)s2m.degenerate());
        if (!conv || std::fabs(t6[3] - 0.03f) > 0.01f || std::fabs(t6[4] + 0.02f) > 0.01f) return 7;
    } catch (const std::exception& e) {
        std::printf("FAILED LOUDLY: %s\n", e.what());
        return 10;
    }
    std::printf("host adaptors ok\n");
    return 0;
}
