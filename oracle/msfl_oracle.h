/*
 * msfl_oracle.h -- CPU ORACLE for the LOAM scan-matching hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is a plain-C restatement of the reference algorithm (kekeliu-whu/MSF_LOAM @ 96924b3)
 * for the path scanRegistration -> laserOdometry -> laserMapping.  It exists to CHECK the CUDA
 * engine and to be TIMED as the CPU baseline.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product (msf_loam_b200/)
 * never links, imports or calls anything in this directory.
 *
 * PINNED TO THE REFERENCE'S OWN CODE: msf_loam_node.cc (scan registration), odometry_scan_matcher.cc,
 * mapping_scan_matcher.cc, scan_matcher.cc, lidar_factor.cc, pose_local_parameterization.cc, scan_undistortion.cc and
 * hybrid_grid.cc are compiled UNMODIFIED into oracle/_ref/libmsfl_ref.so (Makefile target `ref`; ref_*_shim.cc) and this
 * restatement is bit-equal to them: factors, Plus, TransformPoint, GetDeltaQP, the registered clouds and feature lists, the
 * poses / correspondence counts / iteration traces of MatchScan2Map (both branches) and MatchScan2Scan, the HybridGrid
 * surround clouds (tests/test_ref_factors.py, test_ref_matchers.py, test_ref_extract.py, test_golden.py, test_stgm.py).
 * RESTATED, NOT PINNED: the third-party numerics underneath, which those sources get from stand-in headers
 * (oracle/ref_stubs/) because the libraries (PCL 1.10 / FLANN 1.9, Ceres <= 2.1,
 * Eigen 3.3, ROS) are un-vendored and absent from the image (no network); the reference's own tests hold
 * no golden vector for this path (only quaternion identities, src/slam/imu_fusion/utility_test.cc:8-34).
 * That arithmetic is restated from the libraries' published algorithms:
 *   - pcl::KdTreeFLANN::nearestKSearch  -> exact k-NN, FLANN L2_Simple<float> distance
 *     (fp32, ((dx*dx)+dy*dy)+dz*dz, no FMA), ascending, ties broken on the lower index;
 *   - pcl::VoxelGrid<PointXYZI>::filter -> centroid voxel filter (fp32 sums, output ascending
 *     voxel index, all fields averaged);
 *   - Eigen::SelfAdjointEigenSolver<Matrix3d> -> cyclic Jacobi, ascending eigenvalues;
 *   - Eigen colPivHouseholderQr().solve  -> column-pivoted Householder least squares;
 *   - ceres::Solve (TRUST_REGION / LEVENBERG_MARQUARDT, HuberLoss(0.1), jacobi scaling,
 *     max_num_iterations=6) -> trust_region_minimizer.cc / levenberg_marquardt_strategy.cc /
 *     corrector.cc / loss_function.cc semantics (see msflo_lm_solve).
 * It is validated in tests/ against an actual FLANN KDTreeSingleIndex (the copy OpenCV bundles:
 * same neighbours, order, fp32 distances and gate on the VLP-16 case), numpy / scipy (cKDTree, eigh,
 * lstsq), finite-difference Jacobians, an independent numpy LM, scipy's least_squares with Huber
 * loss at convergence, and known-transform recovery.  That pins the k-NN row against third-party
 * code; the Eigen decompositions, the Ceres loop and PCL's VoxelGrid rest on these independent restatements.
 */
#ifndef MSFL_ORACLE_H
#define MSFL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* All constants of the hot path; defaults = the reference's compile-time constants. */
typedef struct msflo_params {
  /* feature extraction -- src/msf_loam_node.cc */
  double min_range;          /* 0.3   :434  minimum_range                           */
  double scan_period;        /* 0.1   :80   kScanPeriod                             */
  double curvature_thresh;   /* 0.1   :275,:312                                     */
  double neighbor_gap_sq;    /* 0.05  :293                                          */
  int n_sectors;             /* 6     :255                                          */
  int n_sharp;               /* 2     :277                                          */
  int n_less_sharp;          /* 20    :281                                          */
  int n_flat;                /* 4     :317                                          */
  /* scan-to-scan -- odometry_scan_matcher.cc:15-18,:262 */
  double dist_sq_thresh;     /* 25                                                  */
  double nearby_scan;        /* 2.5                                                 */
  int min_correspondences;   /* 10                                                  */
  /* scan-to-map -- mapping_scan_matcher.cc:128,:147,:150,:216 */
  double knn_max_sq;         /* 1.0  d5^2 gate                                      */
  double line_eig_ratio;     /* 3.0  lambda2 > 3 lambda1                            */
  double line_half_len;      /* 0.1  point_a = c + 0.1 u                            */
  double plane_tol;          /* 0.2                                                 */
  /* solve -- call sites + Ceres defaults (a-9) */
  int num_outer;             /* 2    kOptimalNum                                    */
  int max_num_iterations;    /* 6                                                   */
  double huber_a;            /* 0.1                                                 */
  double initial_radius;     /* 1e4                                                 */
  double max_radius;         /* 1e16                                                */
  double min_radius;         /* 1e-32                                               */
  double min_relative_decrease; /* 1e-3                                             */
  double min_lm_diagonal;    /* 1e-6                                                */
  double max_lm_diagonal;    /* 1e32                                                */
  double function_tolerance; /* 1e-6                                                */
  double gradient_tolerance; /* 1e-10                                               */
  double parameter_tolerance;/* 1e-8                                                */
  int max_consecutive_invalid_steps; /* 5                                           */
  int early_exit;            /* 1 = Ceres termination tests on; 0 = fixed attempt count
                                (throughput schedule "2 x L attempts")   */
} msflo_params;

void msflo_default_params(msflo_params *p);

/* One LM step attempt, Ceres-style iteration log. */
typedef struct msflo_lm_iter {
  double cost;          /* cost at x before the attempt            */
  double cost_candidate;/* cost at x+                               */
  double model_change;  /* model_cost_change                        */
  double rho;           /* relative_decrease                        */
  double radius;        /* radius used for this attempt             */
  int valid;            /* step_is_valid                            */
  int accepted;         /* step_is_successful                       */
} msflo_lm_iter;

#define MSFLO_MAX_ATTEMPTS 64
typedef struct msflo_lm_log {
  int n_attempts;
  int termination;      /* 0 max-iter, 1 param tol, 2 function tol, 3 gradient tol, 4 radius, 5 invalid */
  double initial_cost;
  double final_cost;
  msflo_lm_iter it[MSFLO_MAX_ATTEMPTS];
} msflo_lm_log;

/* correspondence: type 0 = edge (3 residuals), 1 = plane (1 residual).
 * layout: 10 doubles: [type, p(3), a_or_c(3), n(3)]                                     */
#define MSFLO_CORR_STRIDE 10

/* ---- pose helpers (pose = t xyz, q xyzw; rigid_transform.h:59-64) ---- */
void msflo_pose_plus(const double x[7], const double delta[6], double out[7]);
void msflo_transform_point_f(const double pose[7], const float in[3], float out[3]);

/* ---- a-8 factors: residual + 3x7 / 1x7 row-major global Jacobian (lidar_factor.cc:7-44) ---- */
void msflo_edge_factor(const double pose[7], const double p[3], const double a[3], const double n[3],
                       double r[3], double J[21]);
void msflo_plane_factor(const double pose[7], const double p[3], const double c[3], const double n[3],
                        double r[1], double J[7]);

/* ---- evaluate cost, H (6x6 row-major, full) and g at pose over correspondences ---- */
void msflo_accumulate(const msflo_params *P, const double *corr, int n_corr, const double pose[7],
                      double *cost, double H[36], double g[6]);

/* ---- a-9: Ceres-semantics LM on one SE(3) block; pose in/out ---- */
int msflo_lm_solve(const msflo_params *P, const double *corr, int n_corr, double pose[7], msflo_lm_log *log);

/* the same loop over a caller-supplied evaluation: eval fills cost and, when non-NULL, H (6x6 row-major) and g at pose */
typedef void (*msflo_eval_fn)(const msflo_params *P, const void *ctx, const double pose[7], double *cost, double H[36],
                              double g[6]);
int msflo_lm_solve_cb(const msflo_params *P, msflo_eval_fn eval, const void *ctx, int n_blocks, double pose[7],
                      msflo_lm_log *log);

/* ---- exact k-NN (kd-tree, FLANN L2_Simple<float> semantics) ---- */
typedef struct msflo_kdtree msflo_kdtree;
msflo_kdtree *msflo_kdtree_build(const float *xyzi, int n);       /* xyzi: n x 4 float, not copied */
void msflo_kdtree_free(msflo_kdtree *t);
/* returns number found (<= k); idx/d2 ascending by (d2, idx) */
int msflo_kdtree_knn(const msflo_kdtree *t, const float q[3], int k, int *idx, float *d2);
/* batch helper for tests: nq queries (nq x 3 float) */
void msflo_knn_batch(const float *xyzi, int n, const float *q, int nq, int k, int *idx, float *d2);
/* brute force, same semantics (validation) */
void msflo_knn_brute(const float *xyzi, int n, const float *q, int nq, int k, int *idx, float *d2);

/* ---- 3x3 symmetric eigen (ascending) and 5x3 col-piv Householder LS ---- */
void msflo_sym_eig3(const double A[9], double evals[3], double evecs[9] /* row-major, columns = vectors */);
void msflo_lstsq_5x3(const double A[15] /* row-major 5x3 */, const double b[5], double x[3]);

/* ---- a-6/a-7 association for scan-to-map; writes corr (cap (nc+ns)*10), counts ---- */
void msflo_associate_map(const msflo_params *P,
                         const msflo_kdtree *tree_corner, const float *map_corner,
                         const msflo_kdtree *tree_surf, const float *map_surf,
                         const float *scan_corner, int n_scan_corner,
                         const float *scan_surf, int n_scan_surf,
                         const double pose[7], double *corr, int *n_edge, int *n_plane,
                         int *knn_idx_out /* optional (nc+ns) x 5, -1 if gate failed */);

/* ---- full MatchScan2Map, LiDAR-only branch (mapping_scan_matcher.cc:63-278) ----
 * logs: optional array of num_outer logs; counts: optional 2*num_outer ints (edge, plane) */
int msflo_scan2map(const msflo_params *P,
                   const float *map_corner, int n_map_corner, const float *map_surf, int n_map_surf,
                   const float *scan_corner, int n_scan_corner, const float *scan_surf, int n_scan_surf,
                   double pose[7], msflo_lm_log *logs, int *counts);

/* batch over independent scans against one submap, pthreads; clouds packed, offsets[B+1] */
int msflo_scan2map_batch(const msflo_params *P,
                         const float *map_corner, int n_map_corner, const float *map_surf, int n_map_surf,
                         int B, const float *scan_corner, const int *corner_off,
                         const float *scan_surf, const int *surf_off,
                         double *poses /* B x 7 */, int n_threads);

/* ---- 8f row 3: IMU-deskew branch of MatchScan2Map (is_initialized == true), LiDAR part only
 *      (mapping_scan_matcher.cc:107-246 with the Deskew factors, lidar_factor.cc:46-100;
 *      GetDeltaQP scan_undistortion.cc:22-42).  The IMU-only predict (:35-60) is the IMU side-car
 *      and stays with the caller: `pose` comes in as pose_j after that predict.  The speed-bias
 *      block is constant in the reference's problem (:94), so only the pose is optimised. ---- */
typedef struct msflo_deskew {
  const double *sum_dt;   /* [n] preintegration->sum_dt_buf_            */
  const double *delta_q;  /* [n][4] delta_q_buf_, x y z w               */
  const double *delta_p;  /* [n][3] delta_p_buf_                        */
  int n;
  double velocity[3];     /* Vi = bias_j.head<3>()                      */
  double gravity[3];      /* gravity_vector                             */
} msflo_deskew;
/* returns 0, or -1 when dt is outside [sum_dt.front(), sum_dt.back()] (the reference CHECK-fails) */
int msflo_get_delta_qp(const msflo_deskew *dk, double dt, double dq[4], double dp[3]);
void msflo_edge_factor_deskew(const double pose[7], const double V[3], const double p[3], const double C[3],
                              const double N[3], const double dp[3], const double dq[4], double dt, const double G[3],
                              double r[3], double J[21]);
void msflo_plane_factor_deskew(const double pose[7], const double V[3], const double p[3], const double C[3],
                               const double N[3], const double dp[3], const double dq[4], double dt, const double G[3],
                               double r[1], double J[7]);
int msflo_scan2map_deskew(const msflo_params *P,
                          const float *map_corner, int n_map_corner, const float *map_surf, int n_map_surf,
                          const float *scan_corner, int n_scan_corner, const float *scan_surf, int n_scan_surf,
                          const msflo_deskew *dk, double pose[7], msflo_lm_log *logs, int *counts,
                          int *knn_idx_out /* optional, outer 0 */);

/* ---- a-5 full MatchScan2Scan (odometry_scan_matcher.cc:43-285) ----
 * returns 0 ok, 1 too few correspondences */
int msflo_scan2scan(const msflo_params *P,
                    const float *last_corner, const uint16_t *last_corner_ring, int n_last_corner,
                    const float *last_surf, const uint16_t *last_surf_ring, int n_last_surf,
                    const float *curr_sharp, int n_curr_sharp,
                    const float *curr_flat, int n_curr_flat,
                    double pose[7], msflo_lm_log *logs, int *counts,
                    int *assoc_out /* optional: outer0 only, (n_sharp*2 + n_flat*3) ints */);

/* ---- a-1..a-4 feature extraction (msf_loam_node.cc:86-371) ----
 * in: raw cloud (n x 4 float xyzi, ring u16).  out: ring-major "full" cloud after invalid
 * removal with intensity := relative time, extrinsic applied; curvature (pre-extrinsic);
 * labels; 4 index lists into the full cloud in the reference's push order.
 * All output arrays must have capacity n.  Returns 0, or <0 on bad input. */
int msflo_extract_features(const msflo_params *P, const float *xyzi, const uint16_t *ring, int n,
                           const double T_ext[7],
                           float *full_xyzi, uint16_t *full_ring, int *n_full,
                           float *curvature, int *label,
                           int *idx_sharp, int *n_sharp, int *idx_less_sharp, int *n_less_sharp,
                           int *idx_flat, int *n_flat, int *idx_less_flat, int *n_less_flat);

/* ---- 8f row 1: STGM map (HybridGrid, hybrid_grid.cc:403-521): 3 m cells each holding a cloud that is
 *      re-voxel-filtered on every insert; surround query = cells hit by scan points +-1 m.
 *      The reference concatenates the selected cells in unordered_set<shared_ptr> order (heap-address
 *      dependent); this restatement (and the GPU producer) use ascending cell key (z, y, x). ---- */
typedef struct msflo_stgm msflo_stgm;
msflo_stgm *msflo_stgm_create(float resolution, float leaf);
void msflo_stgm_free(msflo_stgm *m);
/* HybridGridImpl::InsertScan (:503-521): scan_world already transformed (laser_mapping.cc:330-338) */
void msflo_stgm_insert(msflo_stgm *m, const float *scan_world_xyzi, int n);
/* HybridGridImpl::GetSurroundedCloud (:470-501); out capacity msflo_stgm_size(); returns n_out */
int msflo_stgm_surround(const msflo_stgm *m, const float *scan_xyzi, int n, const double pose[7], float *out_xyzi);
int msflo_stgm_size(const msflo_stgm *m, int *n_cells);
/* all points, cells in ascending key order; out capacity msflo_stgm_size() */
int msflo_stgm_dump(const msflo_stgm *m, float *out_xyzi);

/* ---- pcl::VoxelGrid<PointXYZI>::filter restatement; out capacity n; returns n_out ---- */
int msflo_voxel_grid(const float *xyzi, int n, float leaf, float *out_xyzi);

#ifdef __cplusplus
}
#endif
#endif
