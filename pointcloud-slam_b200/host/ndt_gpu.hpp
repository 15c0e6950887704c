// b200reg host adaptor — pclomp::NormalDistributionsTransform's interface on top of the C ABI.
//
// Mirrors pointcloud_match/ndt_omp/include/pclomp/ndt_omp.h:117-261 and the pcl::Registration calls the reference
// makes on it (jueying_slam/src/mapOptmization.cpp:683-693, localization.cpp:162-188,317-340): setInputTarget,
// setInputSource, align, hasConverged, getFinalTransformation, getTransformationProbability, getFinalNumIteration,
// setResolution / setStepSize / setOutlierRatio / setTransformationEpsilon / setMaximumIterations /
// setNeighborhoodSearchMethod / setNumThreads, calculateScore.  CloudT is any type with a `.points` vector of
// structs starting with float x, y, z (pcl::PointCloud<pcl::PointXYZI> qualifies) - PCL itself is not needed.
// 4x4 transforms are float[16] column-major, i.e. Eigen::Matrix4f::data().
#pragma once
#include <memory>

#include "ivox_gpu.hpp"

namespace b200host {

enum NeighborSearchMethod { KDTREE, DIRECT26, DIRECT7, DIRECT1 };  // ndt_omp.h:61-66

template <typename CloudT>
class NormalDistributionsTransform {
   public:
    explicit NormalDistributionsTransform(int device = 0) : device_(device) {
        // constructor defaults, ndt_omp_impl.hpp:47-65
        prm_.resolution = 1.0f; prm_.step_size = 0.1; prm_.outlier_ratio = 0.55; prm_.trans_eps = 0.1;
        prm_.max_iter = 35; prm_.search = 7; prm_.min_pts = 6; prm_.eig_ratio = 0.01;
        for (int i = 0; i < 16; ++i) final_[i] = (i % 5 == 0) ? 1.f : 0.f;
    }
    ~NormalDistributionsTransform() { b200_ndt_destroy(ndt_); }

    void setResolution(float r) { prm_.resolution = r; dirty_ = true; }               // ndt_omp.h:142
    void setStepSize(double s) { prm_.step_size = s; dirty_ = true; }                  // :166
    void setOutlierRatio(double o) { prm_.outlier_ratio = o; dirty_ = true; }          // :184
    void setTransformationEpsilon(double e) { prm_.trans_eps = e; dirty_ = true; }     // pcl::Registration
    void setMaximumIterations(int n) { prm_.max_iter = n; dirty_ = true; }             // pcl::Registration
    void setNumThreads(int) {}                                                         // :117, the device decides
    void setNeighborhoodSearchMethod(NeighborSearchMethod m) {                         // :189
        prm_.search = m == KDTREE ? 0 : m == DIRECT1 ? 1 : m == DIRECT26 ? 27 : 7;
        dirty_ = true;
    }
    void setInputTarget(const std::shared_ptr<const CloudT>& cloud) { target_ = cloud; target_dirty_ = true; }   // :125-130
    void setInputSource(const std::shared_ptr<const CloudT>& cloud) { source_ = cloud; source_dirty_ = true; }

    /// align(output, guess): output = source transformed by the final transformation (pcl::Registration::align)
    void align(CloudT& output, const float* guess16 = nullptr) {
        sync();
        float I[16];
        for (int i = 0; i < 16; ++i) I[i] = (i % 5 == 0) ? 1.f : 0.f;
        int32_t rc = b200_ndt_align(ndt_, guess16 ? guess16 : I, final_, &result_);
        check(rc, "b200_ndt_align");
        output = *source_;
        for (auto& p : output.points) {
            const float x = p.x, y = p.y, z = p.z;
            p.x = final_[0] * x + final_[4] * y + final_[8] * z + final_[12];
            p.y = final_[1] * x + final_[5] * y + final_[9] * z + final_[13];
            p.z = final_[2] * x + final_[6] * y + final_[10] * z + final_[14];
        }
    }
    bool hasConverged() const { return result_.converged != 0; }
    const float* getFinalTransformation() const { return final_; }                        // column-major 4x4
    double getTransformationProbability() const { return result_.trans_probability; }   // ndt_omp.h:200-204
    int getFinalNumIteration() const { return result_.iters; }                            // :226-230
    double getMaxEigen() const {                                                          // :209-223 (localization-lost heuristic)
        double v = 0;
        check(b200_ndt_max_eigen(result_.hessian, &v), "b200_ndt_max_eigen");
        return v;
    }
    const double* getHessian() const { return result_.hessian; }
    /// pcl::Registration::getFitnessScore(max_range): mean squared nearest-neighbour distance after the last align
    double getFitnessScore(double max_range = 1.7976931348623157e308) {
        double s = 0;
        int64_t nr = 0;
        check(b200_ndt_fitness_score(ndt_, final_, max_range, &s, &nr), "b200_ndt_fitness_score");
        return s;
    }

    /// calculateScore (ndt_omp_impl.hpp:836-880) for h candidate poses (h x 16 floats, column-major)
    void calculateScore(const float* poses16, int64_t h, double* scores) {
        sync();
        check(b200_ndt_score_batch(ndt_, poses16, h, scores), "b200_ndt_score_batch");
    }
    b200_ndt* handle() { sync(); return ndt_; }

   private:
    void sync() {
        if (!ndt_ || dirty_) {
            if (ndt_) b200_ndt_destroy(ndt_);
            ndt_ = nullptr;
            check(b200_ndt_create(&prm_, device_, &ndt_), "b200_ndt_create");
            dirty_ = false;
            target_dirty_ = target_ != nullptr;
            source_dirty_ = source_ != nullptr;
        }
        using PointT = typename std::remove_reference<decltype(target_->points[0])>::type;
        if (target_dirty_ && target_) {
            check(b200_ndt_set_target(ndt_, reinterpret_cast<const float*>(target_->points.data()), (int64_t)target_->points.size(), sizeof(PointT)),
                  "b200_ndt_set_target");
            target_dirty_ = false;
        }
        if (source_dirty_ && source_) {
            check(b200_ndt_set_source(ndt_, reinterpret_cast<const float*>(source_->points.data()), (int64_t)source_->points.size(), sizeof(PointT)),
                  "b200_ndt_set_source");
            source_dirty_ = false;
        }
    }
    int device_;
    b200_ndt_params prm_{};
    b200_ndt* ndt_ = nullptr;
    bool dirty_ = true, target_dirty_ = false, source_dirty_ = false;
    std::shared_ptr<const CloudT> target_, source_;
    float final_[16];
    b200_ndt_result result_{};
};

}  // namespace b200host
