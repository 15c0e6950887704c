// b200reg host adaptor — jueying_lio::IVox's interface on top of the C ABI (include/b200reg.h).
//
// Mirrors jueying_lio/include/ivox3d/ivox3d.h:40-104: same Options / NearbyType names, AddPoints,
// GetClosestPoint (single point and, new, a whole scan at once), NumValidGrids.  Header-only; compiled by the ROS
// package, links libb200reg.so.  PointType is any struct whose first three floats are x, y, z
// (pcl::PointXYZINormal is 48 bytes, pcl::PointXYZI 32/16).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200reg.h"

namespace b200host {

inline void check(int32_t rc, const char* what) {
    if (rc < 0) throw std::runtime_error(std::string(what) + ": " + b200_last_error());
}

enum class IVoxNodeType { DEFAULT, PHC };  // PHC is a compile-time option of the reference that is OFF by default

template <int dim = 3, IVoxNodeType node_type = IVoxNodeType::DEFAULT, typename PointType = void>
class IVox {
   public:
    using PointVector = std::vector<PointType>;
    enum class NearbyType { CENTER, NEARBY6, NEARBY18, NEARBY26 };  // ivox3d.h:46-51

    struct Options {  // ivox3d.h:53-58
        float resolution_ = 0.2f;
        float inv_resolution_ = 10.0f;
        NearbyType nearby_type_ = NearbyType::NEARBY6;
        std::size_t capacity_ = 1000000;
        int device_ = 0;               // new: CUDA device ordinal
        std::size_t max_points_ = 0;   // new: point-pool size hint (0 = 8M)
    };

    explicit IVox(Options options) : options_(options) {
        static_assert(dim == 3, "the GPU map is 3-D");
        b200_map_params p{};
        p.resolution = options.resolution_;
        p.nearby = options.nearby_type_ == NearbyType::CENTER ? 0 : options.nearby_type_ == NearbyType::NEARBY6 ? 6
                   : options.nearby_type_ == NearbyType::NEARBY18 ? 18 : 26;
        p.capacity_voxels = options.capacity_;
        p.max_range = 5.0f;
        p.max_points = options.max_points_;
        check(b200_map_create(&p, options.device_, &map_), "b200_map_create");
    }
    ~IVox() { b200_map_destroy(map_); }
    IVox(const IVox&) = delete;
    IVox& operator=(const IVox&) = delete;

    /// IVox::AddPoints (ivox3d.h:73)
    void AddPoints(const PointVector& points_to_add) {
        if (points_to_add.empty()) return;
        check(b200_map_insert(map_, reinterpret_cast<const float*>(points_to_add.data()), (int64_t)points_to_add.size(), sizeof(PointType)),
              "b200_map_insert");
    }

    /// IVox::GetClosestPoint(pt, closest_pt, max_num, max_range) (ivox3d.h:79).  max_num must be 5 (NUM_MATCH_POINTS) and
    /// max_range the map's (5.0): both are compile-time constants at the reference's only call site (laser_mapping.cc:618).
    bool GetClosestPoint(const PointType& pt, PointVector& closest_pt, int max_num = 5, double max_range = 5.0) {
        (void)max_num; (void)max_range;
        int32_t idx[5], cnt = 0;
        float d2[5], nb[15];
        check(b200_map_knn5_points(map_, reinterpret_cast<const float*>(&pt), 1, sizeof(PointType), idx, d2, &cnt, nb), "b200_map_knn5_points");
        if (cnt == 0) return false;  // the reference returns before touching closest_pt (ivox3d.h:151-153)
        closest_pt.clear();
        for (int k = 0; k < cnt; ++k) {  // coordinates come from the device map itself (it also grows through MapIncremental);
            PointType p{};               // the other fields of the record are not kept on the device and stay zero
            p.x = nb[3 * k]; p.y = nb[3 * k + 1]; p.z = nb[3 * k + 2];
            closest_pt.push_back(p);
        }
        return true;
    }

    /// Batched form: one launch for a whole scan.  idx = n x 5 insertion ordinals (-1 padded), ascending by distance.
    void GetClosestPoints(const PointVector& queries, std::vector<int32_t>& idx, std::vector<float>& sqdist, std::vector<int32_t>& count) {
        const int64_t n = (int64_t)queries.size();
        idx.resize(n * 5); sqdist.resize(n * 5); count.resize(n);
        if (n) check(b200_map_knn5(map_, reinterpret_cast<const float*>(queries.data()), n, sizeof(PointType), idx.data(), sqdist.data(), count.data()),
                     "b200_map_knn5");
    }

    std::size_t NumValidGrids() const { return (std::size_t)b200_map_num_voxels(map_); }  // ivox3d.h:88
    std::size_t NumPoints() const { return (std::size_t)b200_map_num_points(map_); }
    b200_map* handle() const { return map_; }

   private:
    Options options_;
    b200_map* map_ = nullptr;
};

}  // namespace b200host
