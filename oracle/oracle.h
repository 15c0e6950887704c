/* TEST INFRASTRUCTURE ONLY — C API of the CPU oracle.
 *
 * The oracle is a CPU restatement of the reference's scan-to-map hot path
 * (SURVEY.md §8a).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product library
 * (pointcloud-slam_b200/csrc) never does.
 *
 * PARITY UNPINNED: the reference ships no golden vectors, known-answer tests
 * or fixtures for the iVox / IEKF / pclomp NDT / GICP path and
 * cannot be compiled here (no PCL / Boost / TBB; the vendored Eigen lacks
 * Eigen/Core).  The Eigen decompositions the path calls are
 * restated line by line from the vendored Eigen sources (smallmat.h).  The oracle
 * is pinned by its own self-checks only (tests/test_oracle_*.py): stencil kNN
 * == brute force, Jacobian rows == finite differences, NDT gradient ==
 * numeric derivative of the score, pose recovery on noise-free data.
 */
#ifndef ORACLE_H_
#define ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- iVox local map + IEKF point-to-plane update (jueying_lio) ---------- */
typedef struct orc_lio orc_lio;

typedef struct {
    float resolution;          /* ivox_grid_resolution */
    int32_t nearby;            /* 0, 6, 18, 26 */
    uint64_t capacity_voxels;  /* IVox::Options::capacity_ */
    int32_t max_iter;          /* NUM_MAX_ITERATIONS */
    float plane_thr;           /* ESTI_PLANE_THRESHOLD */
    int32_t extrinsic_est_en;
    double R;                  /* LASER_POINT_COV */
    double limit[23];          /* epsi */
    double filter_size_map;    /* filter_size_map_min_ */
    int32_t num_threads;       /* OpenMP threads for the par_unseq loops; 0 = max */
} orc_lio_params;

#define ORC_MAX_PASSES 8
typedef struct {
    int32_t status;            /* 0 ok, 1 no effective points (every pass) */
    int32_t passes;            /* ObsModel calls made */
    int32_t knn_passes;        /* of which searched the map */
    int32_t converged;         /* t > 1 at exit */
    int32_t n_eff[ORC_MAX_PASSES];
    int32_t knn[ORC_MAX_PASSES];
    double x_in[ORC_MAX_PASSES][26];   /* state each pass was evaluated at */
    double HtH[ORC_MAX_PASSES][144];   /* h_x^T h_x, row-major 12x12 */
    double Hth[ORC_MAX_PASSES][12];    /* h_x^T h */
    double ms_match;           /* "ObsModel (Lidar Match)" total */
    double ms_jacobian;        /* "ObsModel (IEKF Build Jacobian)" total */
    double ms_solve;           /* rest of update_iterated_dyn_share_modified */
} orc_iekf_stats;

orc_lio* orc_lio_create(const orc_lio_params* p);
void orc_lio_destroy(orc_lio* h);
/* IVox::AddPoints; returns number of points inserted so far */
int64_t orc_map_insert(orc_lio* h, const float* xyz, int64_t n, int64_t stride_bytes);
int64_t orc_map_num_voxels(orc_lio* h);
int64_t orc_map_num_points(orc_lio* h);
/* IVox::GetClosestPoint(pt, out, 5, 5.0) for n queries.  idx: n*5 insertion
 * ordinals (-1 padded) ascending by (distance, enumeration rank). */
void orc_map_knn5(orc_lio* h, const float* xyz, int64_t n, int64_t stride_bytes, int32_t* idx, float* sqdist,
                  int32_t* count);
/* bytes-per-query statistics for the roofline: total occupied-stencil-cell points over the queries */
int64_t orc_map_knn_candidates(orc_lio* h, const float* xyz, int64_t n, int64_t stride_bytes);

/* esekf::update_iterated_dyn_share_modified with LaserMapping::ObsModel.
 * x: 26 doubles pos(3) rot(xyzw) offR(xyzw) offT(3) vel(3) bg(3) ba(3) grav(3);
 * P: 23x23 row-major (symmetric in practice). */
int32_t orc_iekf_update(orc_lio* h, const float* scan_body, int64_t n, int64_t stride_bytes, double* x, double* P,
                        orc_iekf_stats* st);
/* One ObsModel call at state x with dyn_share.converge = converge (uses and
 * updates the persistent per-point arrays exactly like the reference). */
int32_t orc_obs_model(orc_lio* h, const float* scan_body, int64_t n, int64_t stride_bytes, const double* x,
                      int32_t converge, double* HtH, double* Hth, int32_t* n_eff);
/* per-point arrays after the last ObsModel call (n entries) */
void orc_point_state(orc_lio* h, int64_t n, float* plane4, float* residual, uint8_t* selected, int32_t* nn_idx5,
                     int32_t* nn_count);
/* Full rows of the last pass: h_x (n_eff x 12 row-major) and h (n_eff) */
int32_t orc_last_rows(orc_lio* h, double* h_x, double* hvec, int32_t max_rows);
/* LaserMapping::MapIncremental at state x; returns points added */
int64_t orc_map_incremental(orc_lio* h, const float* scan_body, int64_t n, int64_t stride_bytes, const double* x,
                            int32_t ekf_inited, int32_t* n_add, int32_t* n_nodown);

/* stand-alone pieces for unit tests */
int32_t orc_esti_plane(const float* pts_xyz, int32_t n, float thr, float* plane4);
void orc_state_boxplus(double* x26, const double* dx23);
void orc_state_boxminus(const double* x26, const double* y26, double* dx23);
void orc_inverse(const double* A, int32_t n, double* out);
/* esekf::predict over K IMU intervals; steps K x 8 {dt, offs_t, acc_avr[3], angvel_avr[3]}, Q12 = diag(Q_), poses22 K x 22 (optional) */
void orc_predict(const double* steps, int32_t K, const double* Q12, double* x26, double* P, double* poses22);

/* ---- pclomp NDT -------------------------------------------------------- */
typedef struct orc_ndt orc_ndt;
typedef struct {
    float resolution;
    double step_size;
    double outlier_ratio;
    double trans_eps;
    int32_t max_iter;
    int32_t search;     /* 1, 7 (DIRECT7), 27 (DIRECT26) */
    int32_t min_pts;    /* 6 */
    double eig_ratio;   /* 0.01 */
    int32_t num_threads;
} orc_ndt_params;
typedef struct {
    int32_t converged;
    int32_t iters;
    int32_t evals;      /* computeDerivatives calls */
    int32_t hess_evals; /* computeHessian calls */
    double trans_probability;
    double hessian[36];
    double score;
    double p_final[6];
} orc_ndt_result;

orc_ndt* orc_ndt_create(const orc_ndt_params* p);
void orc_ndt_destroy(orc_ndt* h);
int64_t orc_ndt_set_target(orc_ndt* h, const float* xyz, int64_t n, int64_t stride_bytes); /* returns #leaves with >= min_pts */
void orc_ndt_set_source(orc_ndt* h, const float* xyz, int64_t n, int64_t stride_bytes);
int64_t orc_ndt_num_leaves(orc_ndt* h);      /* all leaves in the std::map */
/* dump valid leaves sorted by leaf id: id, n, mean(3), cov(9), icov(9) */
int64_t orc_ndt_leaves(orc_ndt* h, int64_t max, int64_t* ids, int32_t* npts, double* mean, double* cov, double* icov);
void orc_ndt_grid(orc_ndt* h, int32_t* min_b, int32_t* div_b);
/* computeDerivatives at pose vector p (source transformed by the float 4x4 built from p) */
double orc_ndt_derivatives(orc_ndt* h, const double* p6, double* g6, double* H36, int32_t compute_hessian);
/* computeHessian (double path) at p */
void orc_ndt_hessian(orc_ndt* h, const double* p6, double* H36);
int32_t orc_ndt_align(orc_ndt* h, const float* guess16_colmajor, float* final16_colmajor, orc_ndt_result* r);
/* calculateScore for h poses (col-major float 4x4 each) */
void orc_ndt_score_batch(orc_ndt* h, const float* poses16, int64_t nposes, double* scores);
int64_t orc_ndt_nbhd_total(orc_ndt* h, const double* p6);
/* pcl::Registration::getFitnessScore(max_range) with the source moved by T (col-major 4x4) */
double orc_ndt_fitness(orc_ndt* h, const float* T16_colmajor, double max_range, int64_t* n_in_range); /* sum of neighbourhood sizes (roofline bytes) */
/* test switch: 1 = round-1 substitutes (cyclic Jacobi) for SelfAdjointEigenSolver / JacobiSVD instead of the restatements of
 * the Eigen sources vendored in the reference; and stand-alone probes of those restatements */
void orc_set_legacy_eigen(int32_t on);
int32_t orc_eigen_selfadjoint3(const double* A9_rowmajor, double* w3, double* V9_rowmajor);
void orc_jacobi_svd_solve6(const double* H36_rowmajor, const double* rhs6, double* x6, double* sv6);
void orc_euler_from_matrix(const float* m16_colmajor, float* rpy);
void orc_matrix_from_pose(const double* p6, float* m16_colmajor);

/* ---- pcl::VoxelGrid (scan downsample) and the keyframe-merge map builder (construct_full_map) ---- */
/* records are (x, y, z, intensity) floats at `stride` bytes (stride 12: no intensity); out: 4 floats per voxel in
 * ascending leaf-index order, out_count the points per voxel; returns the number of voxels */
int64_t orc_voxel_grid(const float* xyzi, int64_t n, int64_t stride, float leaf, int32_t min_points, float* out_xyzi, int32_t* out_count,
                       int64_t max);
int64_t orc_full_map(const float* xyzi, const int64_t* offsets, int64_t n_frames, const double* poses7, float leaf, float* out_xyzi,
                     int32_t* out_count, int64_t max);

/* ---- per-point motion compensation (ImuProcess::UndistortPcl, backward half) ---- */
void orc_undistort(const float* pts, int64_t n, int64_t stride, int32_t time_index, int32_t intensity_index, const double* poses22, int32_t K,
                   const double* x_end26, float* out_xyzi, int32_t* out_order);

/* ---- LOAM-style scan-to-map optimisation of jueying_slam (mapOptmization.cpp:1255-1590) ---- */
typedef struct orc_loam orc_loam;
orc_loam* orc_loam_create(int32_t num_threads);
void orc_loam_destroy(orc_loam* h);
void orc_loam_set_map(orc_loam* h, const float* corner, int64_t nc, int64_t sc, const float* surf, int64_t ns, int64_t ss);
int32_t orc_loam_features(orc_loam* h, const float* corner, int64_t nc, int64_t sc, const float* surf, int64_t ns, int64_t ss, const float* t6,
                          uint8_t* flags, float* coeff4);
int32_t orc_loam_optimize(orc_loam* h, const float* corner, int64_t nc, int64_t sc, const float* surf, int64_t ns, int64_t ss, float* t6,
                          int32_t iter_num, int32_t* n_sel, int32_t* converged, int32_t* degenerate, double* AtA_first);

/* ---- pclomp::GeneralizedIterativeClosestPoint (gicp_omp.h / gicp_omp_impl.hpp) ---- */
typedef struct orc_gicp orc_gicp;
typedef struct {
    int32_t k_correspondences;      /* 20 */
    double gicp_epsilon;            /* 0.001 */
    double rotation_epsilon;        /* 2e-3 */
    double transformation_epsilon;  /* 5e-4 */
    double corr_dist_threshold;     /* 5.0 */
    int32_t max_iterations;         /* 200 */
    int32_t max_inner_iterations;   /* 20 */
    int32_t num_threads;
} orc_gicp_params;
typedef struct {
    int32_t converged;
    int32_t iterations;  /* nr_iterations_ */
    int32_t last_m;      /* correspondences of the last pass */
    int32_t last_inner;  /* BFGS steps of the last pass */
    int32_t last_status; /* BFGSSpace status the last pass ended with */
    int32_t inner_total;
    int32_t n_f, n_df, n_fdf; /* functor calls over the whole align */
} orc_gicp_result;
orc_gicp* orc_gicp_create(const orc_gicp_params* p);
void orc_gicp_destroy(orc_gicp* h);
void orc_gicp_set_target(orc_gicp* h, const float* xyz, int64_t n, int64_t stride_bytes);
void orc_gicp_set_source(orc_gicp* h, const float* xyz, int64_t n, int64_t stride_bytes);
/* computeCovariances of the source (which = 0) or the target (1): n x 9 doubles, row-major; -1 when the cloud has fewer than k points */
int32_t orc_gicp_covariances(orc_gicp* h, int32_t which, double* cov9);
/* align(output, guess): 0 converged, 2 not converged, -1 error */
int32_t orc_gicp_align(orc_gicp* h, const float* guess16_colmajor, float* final16_colmajor, orc_gicp_result* r);
/* the correspondence half of one pass of computeTransformation at (transformation_, guess): tgt_idx[i] = matched target point or -1,
 * maha9 = mahalanobis_[i].block<3,3> (row-major), d2 = squared distance to the nearest target point; returns the match count */
int64_t orc_gicp_correspondences(orc_gicp* h, const float* trans16_colmajor, const float* guess16_colmajor, int32_t* tgt_idx, float* maha9, float* d2);
/* the cost functor on the correspondences of the last orc_gicp_correspondences call: operator(), fdf's f, df's and fdf's gradients */
void orc_gicp_cost(orc_gicp* h, const double* x6, double* f_op, double* f_fdf, double* g_df6, double* g_fdf6);
/* estimateRigidTransformationBFGS on those correspondences, starting from (and returning) trans16; -1 = exception */
int32_t orc_gicp_estimate(orc_gicp* h, float* trans16_colmajor, int32_t* inner, int32_t* status, int32_t* calls3);
void orc_gicp_apply_state(const double* x6, float* t16_colmajor); /* applyState on the identity */
/* exact k nearest neighbours (ascending distance, ties to the lower index): grid search, or brute force when brute != 0 */
void orc_gicp_knn(const float* xyz, int64_t n, int64_t stride, const float* q, int64_t nq, int64_t qstride, int32_t k, int32_t brute, int32_t* idx,
                  float* d2);
/* pcl::BFGS driven like estimateRigidTransformationBFGS on a fixed smooth 6-D test function; returns the BFGSSpace status */
int32_t orc_bfgs_test(double* x6, int32_t max_inner, int32_t* inner, int32_t* calls3);

#ifdef __cplusplus
}
#endif
#endif
