/*
 * msfl.h -- C ABI of the Blackwell-native LOAM scan-matching engine (libmsfl.so).
 *
 * This is the drop-in boundary for ONE hot path of kekeliu-whu/MSF_LOAM @ 96924b3
 * (scanRegistration -> laserOdometry -> laserMapping).  Each entry point names the reference
 * interface it replaces (paths relative to the reference tree).  Plain pointers and sizes only;
 * no C++/torch types.  All functions return an int status and never throw:
 *     MSFL_OK (0), MSFL_TOO_FEW (1: fewer than min_correspondences, odometry only --
 *     odometry_scan_matcher.cc:262-267; the pose keeps the result of the completed outer
 *     iterations), < 0 error (msfl_last_error() gives the text).
 * There is no CPU fallback: every compute entry point fails with MSFL_ERR_CUDA when no sm_100
 * device / kernel image is available.
 *
 * Conventions
 *   pose      double[7] = [tx ty tz qx qy qz qw]   Rigid3d::ToVector7, rigid_transform.h:59-64
 *   cloud     host AoS view (ptr, n, stride, field byte offsets) so a pcl::PointCloud<P> is passed
 *             as (&cloud.points[0], cloud.size(), sizeof(P), offsetof(P,x), offsetof(P,intensity),
 *             offsetof(P,ring) or MSFL_NO_FIELD).  Caller keeps ownership; nothing is retained
 *             after return (laser_odometry.cc:90 shallow-copies scan_last_ = scan_curr).
 *   threading one engine per matcher thread (MatchScan2Scan runs on the LiDAR callback thread,
 *             MatchScan2Map on LaserMapping::Run, laser_mapping.cc:86); an engine owns one CUDA
 *             stream and its device buffers and is not re-entrant.
 */
#ifndef MSFL_H
#define MSFL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSFL_OK 0
#define MSFL_TOO_FEW 1
#define MSFL_ERR_ARG (-1)
#define MSFL_ERR_CUDA (-2)
#define MSFL_ERR_RING (-3)     /* ring >= 128 (kMaxScanNum, msf_loam_node.cc:79,136) or clouds not ring-sorted */
#define MSFL_ERR_EMPTY (-4)    /* no valid points (CHECK_GT(_N, 0), msf_loam_node.cc:200) */
#define MSFL_ERR_NOSUBMAP (-5)
#define MSFL_ERR_GRID (-6)     /* submap bounding box too large for the dense cell index */

#define MSFL_NO_FIELD ((size_t)-1)
#define MSFL_MAX_OUTER 4
#define MSFL_MAX_ATTEMPTS 16
#define MSFL_MAX_RINGS 128

/* Every constant of the hot path in one POD; msfl_default_params() fills the reference values. */
typedef struct msfl_params {
  /* feature extraction -- src/msf_loam_node.cc */
  double min_range;            /* 0.3   ROS param minimum_range :434                  */
  double scan_period;          /* 0.1   kScanPeriod :80                               */
  double curvature_thresh;     /* 0.1   :275, :312                                    */
  double neighbor_gap_sq;      /* 0.05  :293                                          */
  int32_t n_sectors;           /* 6     :255                                          */
  int32_t n_sharp;             /* 2     :277                                          */
  int32_t n_less_sharp;        /* 20    :281                                          */
  int32_t n_flat;              /* 4     :317                                          */
  /* scan-to-scan -- odometry_scan_matcher.cc:15-18, :262 */
  double dist_sq_thresh;       /* 25    kDistanceSqThreshold                          */
  double nearby_scan;          /* 2.5   kNearByScan                                   */
  int32_t min_correspondences; /* 10                                                  */
  int32_t _pad0;
  /* scan-to-map -- mapping_scan_matcher.cc:128, :147, :150, :216 */
  double knn_max_sq;           /* 1.0   pointSearchSqDis[4] < 1.0                     */
  double line_eig_ratio;       /* 3.0   eigenvalues[2] > 3 eigenvalues[1]             */
  double line_half_len;        /* 0.1   point_a = 0.1 u + c                           */
  double plane_tol;            /* 0.2                                                 */
  /* solve -- call sites + Ceres defaults */
  int32_t num_outer;           /* 2     kOptimalNum (<= MSFL_MAX_OUTER)               */
  int32_t max_num_iterations;  /* 6     options.max_num_iterations (<= MSFL_MAX_ATTEMPTS) */
  double huber_a;              /* 0.1   ceres::HuberLoss(0.1)                         */
  double initial_radius;       /* 1e4   initial_trust_region_radius                   */
  double max_radius;           /* 1e16  max_trust_region_radius                       */
  double min_radius;           /* 1e-32 min_trust_region_radius                       */
  double min_relative_decrease;/* 1e-3                                                */
  double min_lm_diagonal;      /* 1e-6                                                */
  double max_lm_diagonal;      /* 1e32                                                */
  double function_tolerance;   /* 1e-6                                                */
  double gradient_tolerance;   /* 1e-10                                               */
  double parameter_tolerance;  /* 1e-8                                                */
  int32_t max_consecutive_invalid_steps; /* 5                                         */
  int32_t early_exit;          /* 1 = Ceres termination tests; 0 = fixed attempt count
                                  (throughput schedule)                 */
  /* engine */
  int32_t lm_cluster;          /* CTAs (thread-block cluster size) per scan: 0 / 1 = one CTA per scan (a scan's
                                  pose is then bit-identical alone and anywhere in a batch; calls of >= 5 x SM-count
                                  scans solve with three-warp CTAs: same pose to < 1e-12 m), 2/4/8/16 = partial sums meet
                                  over distributed shared memory: small batches of large scans fill the chip, and
                                  a SINGLE scan (msfl_scan2map, the ROS call pattern) runs as one fused launch --
                                  association into shared memory + solve, scan2map_fused.cu; 16 needs a GPC with
                                  16 free SMs and falls back to 8 */
  int32_t assoc_sorted;        /* scan-to-map association order: 0 = auto (order the batch's queries
                                  by submap cell when it holds >= 65536 queries), 1 = never, 2 = always,
                                  3 = always + search against TMA-staged shared-memory tiles of the
                                  3x3x3 cell neighbourhood (k_knn5_tiled) */
} msfl_params;

/* Host AoS cloud view (see "Conventions"). */
typedef struct msfl_cloud {
  const void *data;
  size_t n;
  size_t stride;
  size_t off_xyz;        /* byte offset of float x; y, z follow                      */
  size_t off_intensity;  /* byte offset of float intensity, or MSFL_NO_FIELD         */
  size_t off_ring;       /* byte offset of uint16 ring, or MSFL_NO_FIELD             */
} msfl_cloud;

/* Ceres-style iteration log (summary.iterations / minimizer_progress_to_stdout). */
typedef struct msfl_lm_iter {
  double cost, cost_candidate, model_change, rho, radius;
  int32_t valid, accepted;
} msfl_lm_iter;

typedef struct msfl_lm_log {
  int32_t n_attempts;
  int32_t termination; /* 0 max-iter, 1 parameter tol, 2 function tol, 3 gradient tol, 4 radius, 5 invalid */
  double initial_cost, final_cost;
  msfl_lm_iter it[MSFL_MAX_ATTEMPTS];
} msfl_lm_log;

typedef struct msfl_stats {
  int32_t status;                    /* per-scan status (MSFL_OK / MSFL_TOO_FEW)     */
  int32_t n_outer;                   /* outer iterations executed                    */
  int32_t n_edge[MSFL_MAX_OUTER];    /* corner_num / corner_correspondence           */
  int32_t n_plane[MSFL_MAX_OUTER];   /* surf_num / plane_correspondence              */
  msfl_lm_log lm[MSFL_MAX_OUTER];
} msfl_stats;

/* Output of msfl_extract_features: caller-allocated arrays of capacity n_in. */
typedef struct msfl_features {
  float *full_xyzi;        /* [n_full][4] ring-major cloud_full_res, intensity := rel. time, extrinsic applied */
  uint16_t *full_ring;     /* [n_full]                                               */
  float *curvature;        /* [n_full] optional (NULL ok)                            */
  int32_t *label;          /* [n_full] optional: 0 unknown 1 sharp 2 less-sharp 3 flat */
  int32_t *idx_sharp;      /* indices into full_* in the reference's push order      */
  int32_t *idx_less_sharp;
  int32_t *idx_flat;
  int32_t *idx_less_flat;
  int32_t n_full, n_sharp, n_less_sharp, n_flat, n_less_flat;
} msfl_features;

typedef struct msfl_engine msfl_engine;

void msfl_default_params(msfl_params *p);
const char *msfl_last_error(void);
const char *msfl_version(void);
/* Binding self-check: pass the caller's sizeof() of the five structs above/below; MSFL_ERR_ARG (and the sizes this
 * library was compiled with in msfl_last_error()) when a binding generated from another msfl.h is in use. */
int msfl_abi_check(size_t sizeof_params, size_t sizeof_stats, size_t sizeof_cloud, size_t sizeof_features,
                   size_t sizeof_deskew);

/* One engine per matcher object (replaces the members of OdometryScanMatcher /
 * MappingScanMatcher, laser_odometry.cc:56, laser_mapping.cc:42).  `stream` (a cudaStream_t)
 * may be NULL: the engine then creates its own non-blocking stream. */
int msfl_create(const msfl_params *params, int device, msfl_engine **out);
int msfl_create_on_stream(const msfl_params *params, int device, void *stream, msfl_engine **out);
void msfl_destroy(msfl_engine *e);
int msfl_sync(msfl_engine *e);
void *msfl_stream(msfl_engine *e); /* cudaStream_t the engine launches on (for CUDA-event timing) */
/* number of kernel launches issued by this engine since creation (evidence for gpu_launches) */
uint64_t msfl_launch_count(const msfl_engine *e);

/* Per-stage device timing (the reference's LOG_STEP_TIME lines, tic_toc.h:29-30, as CUDA events on
 * the engine stream).  Stages: 0 "Data association" (mapping_scan_matcher.cc:248), 1 "Solver time"
 * (:264), 2 query transform + cell sort (part of data association).  msfl_get_profile synchronises,
 * returns the summed milliseconds and launch counts since the last call and resets them. */
#define MSFL_N_STAGES 4
int msfl_set_profiling(msfl_engine *e, int on);
int msfl_get_profile(msfl_engine *e, double ms[MSFL_N_STAGES], int32_t count[MSFL_N_STAGES]);

/* ---- scan-to-map: MappingScanMatcher::MatchScan2Map, LiDAR-only branch
 *      (mapping_scan_matcher.h:14-21, mapping_scan_matcher.cc:63-278) ------------------------- */

/* cloud_map.cloud_corner_less_sharp / cloud_surf_less_flat (mapping_scan_matcher.cc:66-72: the
 * two kd-tree builds).  H2D copy + cell-index build; the submap stays resident until replaced. */
int msfl_set_submap(msfl_engine *e, const msfl_cloud *map_corner, const msfl_cloud *map_surf);
/* Same, from device-resident packed float4 (x,y,z,*) arrays, e.g. after an NCCL broadcast. */
int msfl_set_submap_device(msfl_engine *e, const float *d_corner_xyzi, size_t n_corner,
                           const float *d_surf_xyzi, size_t n_surf);
/* Device pointers of the resident packed submap arrays (for broadcasting from the owner rank). */
int msfl_get_submap_device(msfl_engine *e, const float **d_corner_xyzi, size_t *n_corner,
                           const float **d_surf_xyzi, size_t *n_surf);

/* ---- multi-GPU: independent scans are sharded over ranks, one process per GPU; the ONE collective
 *      is the broadcast of the submap from the rank that owns the map, once per map version.  The payload carries
 *      the cell index (points in caller order, cell-sorted copy, cell table), so the other ranks adopt it instead of
 *      repeating the build of mapping_scan_matcher.cc:66-72.  `nccl_comm` is an ncclComm_t (any communicator whose
 *      ranks each own one engine: the host's own, or one made by msfl_nccl_comm_init); the broadcast runs on the
 *      engine stream.  Collective: every rank of the communicator must call it.  NCCL is bound at run time
 *      (libnccl.so.2), MSFL_ERR_CUDA when it is not installed. ------------------------------------------------- */
int msfl_bcast_submap(msfl_engine *e, void *nccl_comm, int root);
/* Convenience for hosts without a communicator of their own (ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy):
 * rank 0 makes the id, the host ships its 128 bytes to the other ranks by any means, every rank calls comm_init. */
#define MSFL_NCCL_UNIQUE_ID_BYTES 128
int msfl_nccl_get_unique_id(unsigned char id[MSFL_NCCL_UNIQUE_ID_BYTES]);
int msfl_nccl_comm_init(msfl_engine *e, const unsigned char id[MSFL_NCCL_UNIQUE_ID_BYTES], int nranks, int rank, void **nccl_comm);
int msfl_nccl_comm_destroy(void *nccl_comm);

/* One scan vs the resident submap.  pose_tq: in = initial guess (*pose_estimate_map_scan2world),
 * out = estimate.  stats may be NULL. */
int msfl_scan2map(msfl_engine *e, const msfl_cloud *scan_corner, const msfl_cloud *scan_surf,
                  double pose_tq[7], msfl_stats *stats);
/* B independent scans vs the resident submap (BASELINE.json configs 4-5).  poses_tq: B x 7
 * in-out; stats: B entries or NULL.  Returns MSFL_OK or the first error. */
int msfl_scan2map_batch(msfl_engine *e, int B, const msfl_cloud *scan_corner,
                        const msfl_cloud *scan_surf, double *poses_tq, msfl_stats *stats);
/* Same with inputs already resident in HBM: packed float4 queries, per-scan offset tables
 * (B+1 int32, device), poses (B x 7 double, device, in-out).  No host<->device traffic and no
 * host synchronisation: returns after enqueueing on the engine stream. */
int msfl_scan2map_batch_device(msfl_engine *e, int B,
                               const float *d_corner_xyzi, const int32_t *d_corner_off, size_t n_corner_total,
                               const float *d_surf_xyzi, const int32_t *d_surf_off, size_t n_surf_total,
                               double *d_poses_tq, msfl_stats *d_stats /* device, B entries, or NULL */);
/* Asynchronous form of msfl_scan2map_batch for replay / multi-robot streams of batches: submit
 * enqueues the H2D copies (copy stream), the kernels and the D2H copy of the poses (engine stream)
 * and returns a ticket without waiting; wait blocks until that batch is done and writes its B x 7
 * poses (and B stats entries when want_stats was set; stats may be NULL otherwise).  Up to
 * MSFL_MAX_INFLIGHT batches may be in flight, so the upload of batch k+1 overlaps the kernels of
 * batch k (and, for clouds that need the host repack, the repack of batch k+2 overlaps both); batches complete in submission order and give bit-identical poses to the synchronous
 * call.  The clouds must stay valid and unchanged until the matching wait returns (packed,
 * page-locked clouds are DMA'd straight from the caller's memory: float4 points, stride 16, or xyz-only points, stride 12
 * with off_intensity = MSFL_NO_FIELD -- the LiDAR-only matcher never reads a query's intensity, and 12 B points put a
 * quarter less on the PCIe link; any other layout is repacked into a pinned slot by host threads).  Submitting while
 * MSFL_MAX_INFLIGHT tickets are outstanding returns MSFL_ERR_ARG. */
#define MSFL_MAX_INFLIGHT 3
int msfl_scan2map_batch_submit(msfl_engine *e, int B, const msfl_cloud *scan_corner,
                               const msfl_cloud *scan_surf, const double *poses_tq_in, int want_stats,
                               int *ticket);
int msfl_scan2map_batch_wait(msfl_engine *e, int ticket, double *poses_tq_out, msfl_stats *stats);
/* Association only (a-6/a-7) at a given pose, for parity tests: knn_idx (n_corner+n_surf) x 5
 * original submap indices (-1 where the d5^2 gate failed), corr (n_corner+n_surf) x 6 doubles
 * [a_or_c(3), n(3)] (n = 0 where no factor was created).  Either output may be NULL. */
int msfl_associate_map(msfl_engine *e, const msfl_cloud *scan_corner, const msfl_cloud *scan_surf,
                       const double pose_tq[7], int32_t *knn_idx, double *corr);

/* ---- scan-to-map, IMU-initialised branch: the LiDAR part of
 *      MappingScanMatcher::MatchScan2Map with is_initialized == true
 *      (mapping_scan_matcher.cc:107-246 with LidarEdgeFactorDeskewSE3 / LidarPlaneFactorDeskewSE3,
 *      lidar_factor.cc:46-100, and GetDeltaQP, scan_undistortion.cc:22-42).  Each point carries its
 *      relative time in `intensity`; the preintegration buffers give (delta_q, delta_p) at that time.
 *      The speed-bias block is constant in the reference's problem (:94) so only the pose is
 *      optimised and *velocity is returned unchanged.  The IMU-only predict (:35-60) stays with the
 *      caller: pose_tq comes in as pose_j after it. ------------------------------------------------ */
typedef struct msfl_deskew {
  const double *sum_dt;   /* [n]    IntegrationBase::sum_dt_buf_                */
  const double *delta_q;  /* [n][4] delta_q_buf_ coefficients x y z w           */
  const double *delta_p;  /* [n][3] delta_p_buf_                                */
  int32_t n;
  int32_t _pad;
  double velocity[3];     /* bias_j.head<3>()                                   */
  double gravity[3];      /* gravity_vector                                     */
} msfl_deskew;
/* MSFL_ERR_ARG when a point time lies outside [sum_dt[0], sum_dt[n-1]] (the reference CHECK-fails). */
int msfl_scan2map_deskew(msfl_engine *e, const msfl_cloud *scan_corner, const msfl_cloud *scan_surf,
                         const msfl_deskew *deskew, double pose_tq[7], msfl_stats *stats);
/* B scans of a replayed log through the same branch in one launch sequence (arrays of B clouds / B tables, every scan
 * with its own preintegration buffers, velocity and gravity; poses_tq 7 B doubles in-out = the poses after each scan's
 * IMU-only predict; stats B entries or NULL).  Results are those of B msfl_scan2map_deskew calls.  MSFL_ERR_ARG (no pose
 * written) when a point time of any scan lies outside that scan's window. */
int msfl_scan2map_deskew_batch(msfl_engine *e, int B, const msfl_cloud *scan_corner, const msfl_cloud *scan_surf,
                               const msfl_deskew *deskew, double *poses_tq, msfl_stats *stats);

/* ---- scan-to-scan: OdometryScanMatcher::MatchScan2Scan
 *      (odometry_scan_matcher.h:10-12, odometry_scan_matcher.cc:43-285) ----------------------
 * last_* must carry rings and be ring-sorted (Appendix B of SURVEY.md; the extraction emits
 * ring by ring).  pose_tq in-out = *pose_estimate_curr2last. */
int msfl_scan2scan(msfl_engine *e, const msfl_cloud *last_corner_less_sharp,
                   const msfl_cloud *last_surf_less_flat, const msfl_cloud *curr_corner_sharp,
                   const msfl_cloud *curr_surf_flat, double pose_tq[7], msfl_stats *stats);

/* B independent MatchScan2Scan problems in one call (replay of a log: pair b = scans b and b + 1, laser_odometry.cc:75).
 * Arrays of B clouds each; poses_tq 7 B doubles in-out (the B initial guesses / estimates of pose_curr2last);
 * status: B entries (MSFL_OK / MSFL_TOO_FEW per pair, pose of a MSFL_TOO_FEW pair as in the single call) or NULL;
 * stats: B entries or NULL.  Every pair gets its own 1 m cell index over its last-scan clouds (built for all pairs at
 * once); results are those of B msfl_scan2scan calls.  Returns MSFL_OK or the first error (no pose written). */
int msfl_scan2scan_batch(msfl_engine *e, int B, const msfl_cloud *last_corner_less_sharp,
                         const msfl_cloud *last_surf_less_flat, const msfl_cloud *curr_corner_sharp,
                         const msfl_cloud *curr_surf_flat, double *poses_tq, int32_t *status, msfl_stats *stats);
/* outer-iteration-0 association only (parity tests): assoc = n_sharp x 2 then n_flat x 3 ints */
int msfl_associate_scan(msfl_engine *e, const msfl_cloud *last_corner_less_sharp,
                        const msfl_cloud *last_surf_less_flat, const msfl_cloud *curr_corner_sharp,
                        const msfl_cloud *curr_surf_flat, const double pose_tq[7], int32_t *assoc);

/* ---- feature extraction: the block of RealHandleLaserCloudMessage between laser_cloud_in
 *      (msf_loam_node.cc:166) and scan.* (msf_loam_node.cc:360-371) --------------------------
 * raw needs xyz + ring (+ intensity, overwritten by relative time as the reference does).
 * T_lidar2imu: extrinsic applied to all outputs (msf_loam_node.cc:367-371), NULL = identity. */
int msfl_extract_features(msfl_engine *e, const msfl_cloud *raw, const double T_lidar2imu[7],
                          msfl_features *out);

/* The same for B independent scans in one launch sequence (replay of a log: every kernel runs with grid.y = scan, one
 * stable sort orders the whole batch by (scan, ring)); outs[b] receives scan b's results, bit-identical to B single calls. */
int msfl_extract_features_batch(msfl_engine *e, int B, const msfl_cloud *raw, const double T_lidar2imu[7],
                                msfl_features *outs);

/* ---- raw clouds -> poses for B independent scans against the resident submap (replay of a log, re-localisation):
 *      scan registration (msf_loam_node.cc:160-371), the caller's VoxelGrid of the less-sharp / less-flat features
 *      (laser_mapping.cc:264-270; leaf 0.2 / 0.4 = mapping_line_resolution / mapping_plane_resolution) and
 *      MatchScan2Map (mapping_scan_matcher.cc:63-278), every stage once for the whole batch; intermediate clouds stay in
 *      HBM.  poses_tq: B x 7, in = initial guesses, out = estimates; counts / stats: B entries or NULL.  Results are
 *      bit-identical to msfl_extract_features -> msfl_voxel_grid x 2 -> msfl_scan2map_batch on the same scans. */
typedef struct msfl_chain_counts {
  int32_t n_full, n_sharp, n_less_sharp, n_flat, n_less_flat;  /* registration output sizes            */
  int32_t n_corner_queries, n_surf_queries;                     /* after the VoxelGrid: the matcher's queries */
} msfl_chain_counts;
int msfl_register_and_match_batch(msfl_engine *e, int B, const msfl_cloud *raw, const double T_lidar2imu[7],
                                  float leaf_corner, float leaf_surf, double *poses_tq, msfl_chain_counts *counts,
                                  msfl_stats *stats);

/* Replay of B CONSECUTIVE raw scans: msfl_register_and_match_batch plus the odometry between them, all from one
 * registration pass whose features stay in HBM.  Pair b (b >= 1) is OdometryScanMatcher::MatchScan2Scan(last = scan
 * b - 1, curr = scan b) (laser_odometry.cc:75), all B - 1 pairs in one launch sequence (msfl_scan2scan_batch's):
 *   odom_tq      7 B doubles; entry b >= 1 in-out = pose_curr2last of scan b (initial guess in, estimate out; the
 *                reference feeds the previous frame's estimate, a batch takes the caller's); entry 0 is not read
 *   odom_status  B entries or NULL (MSFL_OK / MSFL_TOO_FEW; [0] = MSFL_OK)
 *   compose      != 0: the map matcher's initial guesses are dead-reckoned from poses_tq[0]:
 *                poses_tq[b] = poses_tq[b - 1] * odom_tq[b] (pose_scan2world_ * pose_curr2last_, laser_odometry.cc:79);
 *                == 0: poses_tq holds the B guesses as in msfl_register_and_match_batch
 * then VoxelGrid + scan-to-map of all B scans against the current submap; poses_tq returns the B estimates. */
int msfl_replay_batch(msfl_engine *e, int B, const msfl_cloud *raw, const double T_lidar2imu[7], float leaf_corner,
                      float leaf_surf, double *odom_tq, int32_t *odom_status, int compose, double *poses_tq,
                      msfl_chain_counts *counts, msfl_stats *stats);

/* ---- GPU-resident STGM submap producer: HybridGrid of hybrid_grid.h:32-35,
 *      hybrid_grid.cc:403-521, one map per feature class (laser_mapping.h hybrid_grid_map_corner_ /
 *      _surf_, resolution 3 m; leaf = mapping_line_resolution 0.2 / mapping_plane_resolution 0.4).
 *      The surround cloud stays in HBM and becomes the submap without an H2D copy.  Cells are
 *      concatenated in ascending (z, y, x) cell order (the reference's order is heap-address dependent). */
typedef struct msfl_map msfl_map;
int msfl_map_create(msfl_engine *e, float resolution, float leaf, msfl_map **out);
void msfl_map_destroy(msfl_map *m);
/* HybridGrid::InsertScan.  pose_tq != NULL: the scan is first moved to the world frame with
 * TransformPointCloud (laser_mapping.cc:330-338); NULL: the cloud is already in the world frame. */
int msfl_map_insert(msfl_map *m, const msfl_cloud *scan, const double pose_tq[7]);
/* HybridGrid::GetSurroundedCloud(scan, pose) (laser_mapping.cc:273-278); result kept on the device. */
int msfl_map_surround(msfl_map *m, const msfl_cloud *scan, const double pose_tq[7], size_t *n_out);
int msfl_map_size(const msfl_map *m, size_t *n_points, size_t *n_cells);
/* which = 0: last surround result, 1: the whole map (cells in ascending order). */
int msfl_map_download(msfl_map *m, int which, float *out_xyzi, size_t capacity, size_t *n_out);
/* cloud_map.cloud_corner_less_sharp / cloud_surf_less_flat := the two last surround results. */
int msfl_set_submap_from_maps(msfl_engine *e, msfl_map *corner, msfl_map *surf);
/* One frame of the caller, LaserMapping::MatchScan2Map + InsertScan2Map (laser_mapping.cc:258-340), in one call with every
 * intermediate on the device: VoxelGrid of the scan's two feature clouds with the maps' leaf sizes (:264-270),
 * GetSurroundedCloud of both maps at the incoming pose (:273-278), the "> 10 corner and > 50 surf map points" gate
 * (:284-285), MappingScanMatcher::MatchScan2Map (LiDAR-only branch) against the two surround clouds (:304-311), then
 * InsertScan of the un-down-sampled clouds at the refined pose (:330-338) -- which happens whether or not the gate passed.
 * pose_tq in-out = pose_map_scan2world_; *matched = 1 when the gate passed (may be NULL); stats may be NULL.  The maps
 * must belong to the engine.  Results are those of the separate calls (msfl_voxel_grid, msfl_map_surround,
 * msfl_set_submap_from_maps, msfl_scan2map, msfl_map_insert). */
int msfl_mapping_frame(msfl_engine *e, msfl_map *map_corner, msfl_map *map_surf, const msfl_cloud *corner_less_sharp,
                       const msfl_cloud *surf_less_flat, double pose_tq[7], int32_t *matched, msfl_stats *stats);

/* ---- wire format: view a sensor_msgs/PointCloud2 data buffer as an msfl_cloud
 *      (what pcl::fromROSMsg does at msf_loam_node.cc:166-167 with the field list of common.h:53-62).
 *      x, y, z must be FLOAT32 (datatype 7), intensity FLOAT32 (optional), ring UINT16 (datatype 4,
 *      optional); little-endian, rows dense (row_step == width * point_step).  No copy is made: the
 *      unpack to float4 happens on the GPU inside msfl_extract_features. -------------------------- */
typedef struct msfl_pc2_field {
  const char *name;
  uint32_t offset;
  uint8_t datatype;
  uint32_t count;
} msfl_pc2_field;
int msfl_cloud_from_pointcloud2(const uint8_t *data, uint32_t width, uint32_t height, uint32_t point_step,
                                uint32_t row_step, int is_bigendian, const msfl_pc2_field *fields, int n_fields,
                                msfl_cloud *out);

/* ---- caller-side pcl::VoxelGrid<PointXYZI> (laser_mapping.cc:264-270) --
 * out_xyzi capacity in->n x 4 floats; *n_out receives the number of centroids. */
int msfl_voxel_grid(msfl_engine *e, const msfl_cloud *in, float leaf, float *out_xyzi, size_t *n_out);

/* ---- LM building block, exposed for parity tests (a-8 + reduction): cost, H (6x6 row-major),
 *      g at pose over explicit correspondences: p n x 3 float, corr n x 6 double, first
 *      n_edge rows are edge factors, the rest plane factors. ---------------------------------- */
int msfl_accumulate(msfl_engine *e, const float *p_xyz, const double *corr, int n_edge, int n_plane,
                    const double pose_tq[7], double *cost, double H[36], double g[6]);

#ifdef __cplusplus
}
#endif
#endif /* MSFL_H */
