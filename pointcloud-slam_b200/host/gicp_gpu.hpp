// b200reg host adaptor — pclomp::GeneralizedIterativeClosestPoint's interface on top of the C ABI.
//
// Mirrors pointcloud_match/ndt_omp/include/pclomp/gicp_omp.h:60-283 and the pcl::Registration calls the reference makes
// on the object it selects with "GICP_OMP" (jueying_slam/src/localization.cpp:163,175-177,277,323-328): setInputTarget,
// setInputSource, align, hasConverged, getFitnessScore, getFinalTransformation, plus the class's own setters.  CloudT is
// any type with a `.points` vector of structs starting with float x, y, z - PCL itself is not needed.  4x4 transforms are
// float[16] column-major, i.e. Eigen::Matrix4f::data().
#pragma once
#include <memory>

#include "ivox_gpu.hpp"

namespace b200host {

template <typename CloudT>
class GeneralizedIterativeClosestPoint {
   public:
    explicit GeneralizedIterativeClosestPoint(int device = 0) : device_(device) {
        // constructor defaults, gicp_omp.h:115-127
        prm_.k_correspondences = 20; prm_.gicp_epsilon = 0.001; prm_.rotation_epsilon = 2e-3; prm_.transformation_epsilon = 5e-4;
        prm_.corr_dist_threshold = 5.0; prm_.max_iterations = 200; prm_.max_inner_iterations = 20;
        for (int i = 0; i < 16; ++i) final_[i] = (i % 5 == 0) ? 1.f : 0.f;
    }
    ~GeneralizedIterativeClosestPoint() { b200_gicp_destroy(gicp_); }

    void setRotationEpsilon(double e) { prm_.rotation_epsilon = e; dirty_ = true; }              // gicp_omp.h:224
    double getRotationEpsilon() const { return prm_.rotation_epsilon; }                          // :230
    void setCorrespondenceRandomness(int k) { prm_.k_correspondences = k; dirty_ = true; }       // :240
    int getCorrespondenceRandomness() const { return prm_.k_correspondences; }                   // :246
    void setMaximumOptimizerIterations(int n) { prm_.max_inner_iterations = n; dirty_ = true; }  // :252
    int getMaximumOptimizerIterations() const { return prm_.max_inner_iterations; }              // :256
    void setTransformationEpsilon(double e) { prm_.transformation_epsilon = e; dirty_ = true; }  // pcl::Registration
    void setMaxCorrespondenceDistance(double d) { prm_.corr_dist_threshold = d; dirty_ = true; } // pcl::Registration
    void setMaximumIterations(int n) { prm_.max_iterations = n; dirty_ = true; }                 // pcl::Registration
    void setInputTarget(const std::shared_ptr<const CloudT>& cloud) { target_ = cloud; target_dirty_ = true; }  // :173-178
    void setInputSource(const std::shared_ptr<const CloudT>& cloud) { source_ = cloud; source_dirty_ = true; }  // :141-157

    /// align(output, guess): output = source transformed by the final transformation (gicp_omp_impl.hpp:513-516)
    void align(CloudT& output, const float* guess16 = nullptr) {
        sync();
        check(b200_gicp_align(gicp_, guess16, final_, &result_), "b200_gicp_align");
        output = *source_;
        for (auto& p : output.points) {
            const float x = p.x, y = p.y, z = p.z;
            p.x = final_[0] * x + final_[4] * y + final_[8] * z + final_[12];
            p.y = final_[1] * x + final_[5] * y + final_[9] * z + final_[13];
            p.z = final_[2] * x + final_[6] * y + final_[10] * z + final_[14];
        }
    }
    bool hasConverged() const { return result_.converged != 0; }
    const float* getFinalTransformation() const { return final_; }  // column-major 4x4
    int getFinalNumIteration() const { return result_.iterations; }
    const b200_gicp_result& result() const { return result_; }
    double getFitnessScore(double max_range = 1.7976931348623157e308) {
        double s = 0;
        int64_t nr = 0;
        check(b200_gicp_fitness_score(gicp_, final_, max_range, &s, &nr), "b200_gicp_fitness_score");
        return s;
    }
    b200_gicp* handle() { sync(); return gicp_; }

   private:
    void sync() {
        if (!gicp_ || dirty_) {
            if (gicp_) b200_gicp_destroy(gicp_);
            gicp_ = nullptr;
            check(b200_gicp_create(&prm_, device_, &gicp_), "b200_gicp_create");
            dirty_ = false;
            target_dirty_ = target_ != nullptr;
            source_dirty_ = source_ != nullptr;
        }
        using PointT = typename std::remove_reference<decltype(target_->points[0])>::type;
        if (target_dirty_ && target_) {
            check(b200_gicp_set_target(gicp_, reinterpret_cast<const float*>(target_->points.data()), (int64_t)target_->points.size(), sizeof(PointT)),
                  "b200_gicp_set_target");
            target_dirty_ = false;
        }
        if (source_dirty_ && source_) {
            check(b200_gicp_set_source(gicp_, reinterpret_cast<const float*>(source_->points.data()), (int64_t)source_->points.size(), sizeof(PointT)),
                  "b200_gicp_set_source");
            source_dirty_ = false;
        }
    }
    int device_;
    b200_gicp_params prm_{};
    b200_gicp* gicp_ = nullptr;
    bool dirty_ = true, target_dirty_ = false, source_dirty_ = false;
    std::shared_ptr<const CloudT> target_, source_;
    float final_[16];
    b200_gicp_result result_{};
};

}  // namespace b200host
