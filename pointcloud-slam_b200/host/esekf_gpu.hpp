// b200reg host adaptor — the esekf + ObsModel pair of jueying_lio on top of the C ABI.
//
// In the reference, LaserMapping registers ObsModel with kf_.init_dyn_share(...) (laser_mapping.cc:19-23) and calls
// kf_.update_iterated_dyn_share_modified(LASER_POINT_COV, solve_time) once per scan (:347): the filter calls back
// into ObsModel on every pass.  Here the whole loop (k-NN, plane fit, residual, Jacobian, H^T H reduction, 23x23
// solve, boxplus, convergence logic) runs on the device in one call; the adaptor keeps the method names
// (esekfom.hpp:1526,1836-1848) and moves the scan in as an argument.
#pragma once
#include <array>

#include "ivox_gpu.hpp"

namespace b200host {

/// state_ikfom as 26 doubles: pos(3) rot(x,y,z,w) offset_R_L_I(x,y,z,w) offset_T_L_I(3) vel(3) bg(3) ba(3) grav(3)
/// (use-ikfom.hpp:14-15; quaternions in Eigen coefficient order).
using StateVec = std::array<double, 26>;
/// esekf::cov, 23 x 23 row-major
using CovMat = std::array<double, 23 * 23>;

struct EsekfOptions {
    int max_iter = 3;                 // NUM_MAX_ITERATIONS (config/livox.yaml max_iteration)
    float plane_thr = 0.1f;           // ESTI_PLANE_THRESHOLD
    bool extrinsic_est_en = false;    // config/livox.yaml extrinsic_est_en
    double R = 0.001;                 // LASER_POINT_COV (options.h:12)
    double limit = 0.001;             // epsi (laser_mapping.cc:19)
    double filter_size_map = 0.5;     // filter_size_map_min_
};

/// A scan buffer in page-locked memory (b200_host_alloc): update_iterated_dyn_share_modified then ships the records as they
/// are and unpacks them on the device instead of packing them on the host first.  data() / size() like std::vector, so it
/// can be handed to the update directly; fill it where the downsampled scan is produced.
template <typename PointT>
class PinnedScan {
   public:
    using value_type = PointT;
    explicit PinnedScan(size_t capacity) : cap_(capacity) {
        void* p = nullptr;
        check(b200_host_alloc(capacity * sizeof(PointT), &p), "b200_host_alloc");
        data_ = static_cast<PointT*>(p);
    }
    ~PinnedScan() { b200_host_free(data_); }
    PinnedScan(const PinnedScan&) = delete;
    PinnedScan& operator=(const PinnedScan&) = delete;
    PointT* data() { return data_; }
    const PointT* data() const { return data_; }
    size_t size() const { return n_; }
    size_t capacity() const { return cap_; }
    void resize(size_t n) { n_ = n <= cap_ ? n : cap_; }
    PointT& operator[](size_t i) { return data_[i]; }
    const PointT& operator[](size_t i) const { return data_[i]; }

   private:
    PointT* data_ = nullptr;
    size_t n_ = 0, cap_ = 0;
};

template <typename MapT>
class Esekf {
   public:
    Esekf(MapT& map, const EsekfOptions& o = EsekfOptions()) {
        b200_iekf_params p{};
        p.max_iter = o.max_iter;
        p.plane_thr = o.plane_thr;
        p.extrinsic_est_en = o.extrinsic_est_en ? 1 : 0;
        p.R = o.R;
        for (double& l : p.limit) l = o.limit;
        p.filter_size_map = o.filter_size_map;
        check(b200_iekf_create(&p, map.handle(), &kf_), "b200_iekf_create");
        P_.fill(0.0);
        x_.fill(0.0);
        x_[6] = x_[10] = 1.0;  // identity quaternions
    }
    ~Esekf() { b200_iekf_destroy(kf_); }
    Esekf(const Esekf&) = delete;
    Esekf& operator=(const Esekf&) = delete;

    const StateVec& get_x() const { return x_; }          // esekfom.hpp:1836
    const CovMat& get_P() const { return P_; }            // esekfom.hpp:1839
    void change_x(const StateVec& x) { x_ = x; }          // esekfom.hpp:1842
    void change_P(const CovMat& P) { P_ = P; }            // esekfom.hpp:1845

    /// update_iterated_dyn_share_modified(R, solve_time) for the downsampled body-frame scan (laser_mapping.cc:335-351).
    /// Returns false when no pass had an effective point (the reference's ekfom_data.valid == false, :657-661).
    template <typename PointVector>
    bool update_iterated_dyn_share_modified(const PointVector& scan_down_body, double& solve_time_ms) {
        using PointT = typename PointVector::value_type;
        int32_t rc = b200_iekf_update(kf_, reinterpret_cast<const float*>(scan_down_body.data()), (int64_t)scan_down_body.size(), sizeof(PointT),
                                      x_.data(), P_.data(), &stats_);
        check(rc, "b200_iekf_update");
        solve_time_ms = stats_.gpu_ms;
        return rc == B200_OK;
    }

    /// One IMU interval as ImuProcess::UndistortPcl hands it to kf_state.predict (imu_processing.hpp:190-241)
    struct ImuStep {
        double dt, offs_t, acc_avr[3], angvel_avr[3];
    };
    /// esekf::predict (esekfom.hpp:269-374) over all IMU intervals of a scan in one device launch.  Q12 = diagonal of Q_
    /// (cov_gyr, cov_acc, cov_bias_gyr, cov_bias_acc).  imu_poses (optional) receives IMUpose_ as 22 doubles per step
    /// {offset_time, acc, gyr, vel, pos, rot row-major} - the input of b200_scan_undistort.
    void predict(const std::vector<ImuStep>& steps, const double Q12[12], std::vector<double>* imu_poses = nullptr) {
        static_assert(sizeof(ImuStep) == 8 * sizeof(double), "ImuStep must be 8 packed doubles");
        if (steps.empty()) return;
        if (imu_poses) imu_poses->resize(steps.size() * 22);
        check(b200_iekf_predict(kf_, reinterpret_cast<const double*>(steps.data()), (int32_t)steps.size(), Q12, x_.data(), P_.data(),
                                imu_poses ? imu_poses->data() : nullptr),
              "b200_iekf_predict");
    }

    /// LaserMapping::MapIncremental (laser_mapping.cc:525-583) with the neighbours cached by the last update
    void MapIncremental(bool flg_EKF_inited, int* n_added = nullptr, int* n_no_downsample = nullptr) {
        int32_t a = 0, b = 0;
        check(b200_iekf_map_incremental(kf_, x_.data(), flg_EKF_inited ? 1 : 0, &a, &b), "b200_iekf_map_incremental");
        if (n_added) *n_added = a;
        if (n_no_downsample) *n_no_downsample = b;
    }

    const b200_iekf_stats& stats() const { return stats_; }
    b200_iekf* handle() const { return kf_; }

   private:
    b200_iekf* kf_ = nullptr;
    StateVec x_;
    CovMat P_;
    b200_iekf_stats stats_{};
};

}  // namespace b200host
