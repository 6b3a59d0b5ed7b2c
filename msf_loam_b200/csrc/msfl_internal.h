// msfl_internal.h -- engine state shared by the translation units of libmsfl.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/msfl.h"

namespace msfl {

void set_error(const char *fmt, ...);

#define MSFL_CUDA_OK(expr)                                                                       \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      msfl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MSFL_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

// grow-only device / pinned-host buffers
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  void release();
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};
struct PinBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  void release();
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Dense cell index over one submap class.  Cells are `edge` metres wide; the grid covers the
// occupied cell range padded by 2 cells on every side, so any query whose cell lies in
// [1, n-2] on every axis has its whole 3x3x3 neighbourhood in range and any other query has no
// occupied neighbour.  Points are sorted by linear cell id ((z*ny + y)*nx + x), so the three
// x-adjacent cells of a neighbourhood row are one contiguous range.
struct GridView {
  const float4 *pts_sorted;   // xyz + original index (int bits in w)
  const float4 *pts_orig;     // original order (xyz, *), for gathering the k neighbours
  const uint32_t *cell_start; // ncell + 1
  const uint16_t *row_mask;   // per cell: bit r set <=> row r (nearest-first order of knn5_grid) of its 3x3x3
                              // neighbourhood holds a point; nullptr = not built (every row is visited)
  int nx, ny, nz;
  int ox, oy, oz;             // cell coordinate (floor(x * inv_edge)) of grid cell (0,0,0)
  float inv_edge;
  uint32_t n;
};

struct Submap {
  DevBuf orig, sorted, cell_start, keys, keys_alt, vals, vals_alt, cub_tmp, bounds, row_mask;
  GridView view{};
  size_t n = 0;
};

// packed constants handed to the kernels
struct KParams {
  float knn_max_sq_f;     // smallest float >= knn_max_sq (so that (double)d2 < T  <=>  d2 < T_f)
  float dist_sq_thresh_f;
  double dist_sq_thresh;
  double nearby_scan;
  double line_eig_ratio, line_half_len, plane_tol;
  int max_it, early_exit, max_invalid, min_corr;
  double huber_a, huber_sqrt_a, initial_radius, max_radius, min_radius, min_rel_decrease;
  double min_diag, max_diag, ftol, gtol, ptol;
};

}  // namespace msfl

struct msfl_engine {
  msfl_params params;
  msfl::KParams kp;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;         // H2D of chunk c+1 overlaps the kernels of chunk c
  std::vector<cudaEvent_t> chunk_events;
  uint64_t launches = 0;
  bool lm_attr_set[8] = {false, false, false, false, false, false, false, false};
  int lm_shape_scans = 0;  // > 0: the LM launches of a chunked call pick their CTA shape as a launch of this many scans would
  bool fused_attr_set = false;  // k_scan2map_fused: dynamic smem + non-portable cluster size attributes
  int fused_max16 = -1;         // cudaOccupancyMaxActiveClusters of the 16-CTA configuration (-1: not asked yet)
  int pick_attr_mask = 0;      // k_feat_pick shapes whose dynamic shared-memory attribute is set (per device, so per engine)
  int pick_seen_ring = -1, pick_seen_sector = -1;  // longest ring / sector of the previous extraction (-1: none yet)
  // the cell keys of a batch are counting-sorted while the bin table (64 sub-cell bins per submap cell) stays small
  // enough to live in L2; larger (sparse, far-spread) submaps fall back to a radix sort of the keys
  long long count_sort_max_bins = 16ll << 20;
  long long pair_cell_budget = 1ll << 27;  // batched scan-to-scan: cells (uint32) of the pair grids indexed at once
  int sm_count = 148;
  int pack_threads = 1;  // host threads that repack strided AoS clouds of a large batch into the pinned staging slot

  // per-stage CUDA-event timing (msfl_set_profiling)
  bool profiling = false;
  struct StageEv { cudaEvent_t a, b; int stage; };
  std::vector<StageEv> stage_events;
  std::vector<cudaEvent_t> event_pool;

  msfl::Submap map_corner, map_surf;
  msfl::Submap last_corner_grid, last_surf_grid;  // scan-to-scan: cell index over the last scan's features
  bool has_submap = false;

  // asynchronous batches (msfl_scan2map_batch_submit / _wait): per-slot input / output buffers; the
  // association / LM scratch below is shared because the kernels of all batches run in order on `stream`
  struct BatchSlot {
    msfl::DevBuf d_in, d_in3, d_stats;
    msfl::PinBuf h_stage, h_out, h_stats;
    cudaEvent_t uploaded = nullptr, done = nullptr;
    bool busy = false, want_stats = false;
    int B = 0, ticket = -1;
  };
  BatchSlot slots[MSFL_MAX_INFLIGHT];
  int next_ticket = 0;

  // batch scratch
  msfl::DevBuf d_queries, d_corr, d_poses, d_status, d_stats, d_knn, d_off, d_misc;
  msfl::PinBuf h_stage, h_poses, h_stats, h_misc, h_map_stage;
  cudaEvent_t map_uploaded = nullptr;  // the last msfl_set_submap upload has left h_map_stage
  // sorted association scratch: transformed queries, cell keys / permutation (double-buffered), cub temp
  msfl::DevBuf a_xq, a_keys, a_keys_alt, a_vals, a_vals_alt, a_tmp, a_hist;
  bool fuse_fit = true;  // batch path: plane fit inside the search kernel (MSFL_FUSE_FIT=0: separate k_fit launch)
  msfl::DevBuf a_fb;  // plane queries handed to the Householder fallback kernel: [count | slots]
  msfl::DevBuf fr_qc, fr_qs, fr_misc;  // msfl_mapping_frame: VoxelGrid-ed scan clouds, offsets + pose + stats
  msfl::DevBuf k_table, k_dsk, k_pprime, k_o4;  // deskew branch: preintegration tables, per-query (dq, dp, dt), p', offset o
  const uint32_t *a_perm = nullptr;  // cell-order permutation of the batch being associated

  // odometry scratch
  msfl::DevBuf d_last_corner, d_last_surf, d_last_corner_ring, d_last_surf_ring, d_ring_tab, d_assoc;
  // batched odometry scratch: inputs, per-grid bounds / headers, cell-ordered points + rings, cell table, sort scratch
  msfl::DevBuf ob_in, ob_bounds, ob_hdr, ob_sorted, ob_ring_sorted, ob_cells, ob_keys, ob_rank, ob_tmp;

  // feature extraction scratch
  msfl::DevBuf f_raw, f_keys, f_keys_alt, f_vals, f_vals_alt, f_tmp, f_full, f_ring, f_curv, f_label,
      f_idx, f_cnt, f_angle, f_misc, f_soff;
  // voxel grid scratch
  msfl::DevBuf v_in, v_keys, v_keys_alt, v_vals, v_vals_alt, v_tmp, v_out, v_misc;
  msfl::DevBuf vb_keys, vb_keys_alt, vb_vals, vb_vals_alt, vb_tmp, vb_misc;  // batched VoxelGrid scratch
  msfl::DevBuf c_in, c_off, c_q;                                              // chain: gathered feature clouds, offset tables, queries
};

namespace msfl {

// stage timing helpers (msfl_api.cu): no-ops unless profiling is on
void stage_begin(msfl_engine *e, int stage);
void stage_end(msfl_engine *e);

// ---- submap_index.cu
int submap_build(msfl_engine *e, Submap &m, const float4 *d_pts, size_t n, float edge, bool want_row_mask = true);
int submap_row_mask(msfl_engine *e, Submap &m);  // (re)builds m.view.row_mask from the cell table; no-op for huge grids
int submap_build_host_bounds(msfl_engine *e, Submap &m, size_t n, float edge, const int lo[3], const int hi[3]);
void submap_release(Submap &m);

// ---- associate_map.cu
// batch layout: queries = [all corner | all surf] as two packed float4 arrays with per-scan
// offset tables (B+1 int32, device); corr = (n_corner_total + n_surf_total) x 6 doubles in the
// same flat order.
int launch_associate_map(msfl_engine *e, int B, const float4 *d_qc, const int32_t *d_c_off, uint32_t n_corner_total,
                         const float4 *d_qs, const int32_t *d_s_off, uint32_t n_surf_total, const double *d_poses,
                         double *d_corr, int32_t *d_knn, bool compact = false);

// Deskew branch (mapping_scan_matcher.cc with is_initialized == true), B >= 1 scans.  The preintegration tables of all
// scans are concatenated on the device (sum_dt [N], delta_q [N][4], delta_p [N][3]); scan b's table is rows
// [row0, row0 + n) and it has its own velocity / gravity.  prepare: per query (dq, dp, dt) -> dsk[8], p' -> pprime[4],
// o = V dt - g dt^2 / 2 -> o4[4]; flags[b] |= 1 when a point time of scan b is outside its table.
struct DeskewScan {
  uint32_t row0;
  int32_t n;
  double V[3], G[3];
};
int launch_deskew_prepare(msfl_engine *e, int B, const DeskewScan *d_scans, const double *d_sum_dt, const double *d_dq,
                          const double *d_dp, const float4 *d_q, const int32_t *d_c_off, const int32_t *d_s_off, uint32_t nc,
                          uint32_t n, double *d_dsk, double *d_pprime, double *d_o4, int *d_flags);
int launch_associate_map_deskew(msfl_engine *e, int B, const float4 *d_qc, const int32_t *d_c_off, uint32_t nc,
                                const float4 *d_qs, const int32_t *d_s_off, uint32_t ns, const double *d_pose,
                                const double *d_o4, const double *d_dsk, double *d_corr, int32_t *d_knn);

// ---- lm_solve.cu
// outer: outer-iteration index (stats slot); min_corr: 0 for mapping, params.min_correspondences for odometry
int launch_lm_solve(msfl_engine *e, int B, const float4 *d_qe, const int32_t *d_e_off, uint32_t n_edge_total,
                    const float4 *d_qp, const int32_t *d_p_off, const double *d_corr, double *d_poses, int32_t *d_status,
                    msfl_stats *d_stats, int outer, int min_corr, int plane_bytes = 48);
int launch_lm_solve_pd(msfl_engine *e, int B, const double *d_qe, const int32_t *d_e_off, uint32_t n_edge_total,
                       const double *d_qp, const int32_t *d_p_off, const double *d_corr, double *d_poses, int32_t *d_status,
                       msfl_stats *d_stats, int outer, int min_corr);
int launch_accumulate(msfl_engine *e, const float4 *d_p, const double *d_corr, int n_edge, int n_plane,
                      const double *d_pose, double *d_out28);

// ---- scan2map_fused.cu: one scan, one launch (MSFL_OK = enqueued, 1 = does not qualify, < 0 error)
int launch_scan2map_fused(msfl_engine *e, const float4 *d_qc, uint32_t nc, const float4 *d_qs, uint32_t ns, double *d_pose,
                          int32_t *d_status, msfl_stats *d_stats);

// ---- scan2scan.cu
int launch_associate_scan(msfl_engine *e, const float4 *d_last_corner, const uint16_t *d_last_corner_ring, uint32_t n_lc,
                          const float4 *d_last_surf, const uint16_t *d_last_surf_ring, uint32_t n_ls,
                          const uint32_t *d_ring_start_corner, const uint32_t *d_ring_start_surf,
                          const float4 *d_queries, uint32_t n_sharp, uint32_t n_flat, const double *d_pose,
                          double *d_corr, int32_t *d_assoc);

int scan2scan_batch_device(msfl_engine *e, int B, const float4 *last_pts, const uint16_t *last_ring, const uint32_t *d_goff,
                           const uint32_t *h_goff, const float4 *q_sharp, const int32_t *d_e_off, const int32_t *h_e_off,
                           const float4 *q_flat, const int32_t *d_p_off, const int32_t *h_p_off, double *d_poses,
                           int32_t *d_status, msfl_stats *d_stats);

// ---- features.cu
int run_extract_features(msfl_engine *e, const msfl_cloud *raw, const double T[7], msfl_features *out);
// device-side results of one extraction batch (views into the engine's scratch; valid until the next extraction)
struct FeatMeta;
struct FeatDevice {
  float4 *full_post;   // [N] extrinsic applied, intensity = relative time; scan b at soff[b], n_valid[b] points
  uint16_t *ring;
  float *curv;
  int32_t *label;
  int32_t *o_sharp, *o_less, *o_flat, *o_lf;  // per scan at soff[b]: indices into the scan's own full cloud
  FeatMeta *metas;       // [B]
  const uint32_t *soff;  // [B + 1] device
};
int extract_batch_to_device(msfl_engine *e, int B, const msfl_cloud *raw, const double T[7], FeatDevice *fd,
                            std::vector<uint32_t> &h_off, std::vector<int32_t> &h_counts);
// ---- msfl_api.cu: the launch sequence of a scan-to-map batch whose inputs are in HBM
int scan2map_enqueue(msfl_engine *e, int B, const float4 *d_qc, const int32_t *d_c_off, uint32_t nct, const float4 *d_qs,
                     const int32_t *d_s_off, uint32_t nst, double *d_poses, msfl_stats *d_stats);

// ---- voxel_grid.cu
int run_voxel_grid(msfl_engine *e, const float4 *d_in, size_t n, float leaf, float4 *d_out, size_t *n_out);
// B clouds back to back in d_in (scan b at d_in_off[b], device offsets): centroids back to back in d_out, d_out_off
// (B + 1 int32, device) receives where each scan's centroids start.  No synchronisation.  vb = 0 / 1 selects one of two
// scratch sets so that two batches (corner, surf) can be in flight.
int run_voxel_grid_batch(msfl_engine *e, int B, const float4 *d_in, const uint32_t *d_in_off, size_t n_total, uint32_t max_n,
                         float leaf, float4 *d_out, int32_t *d_out_off, int vb);

// ---- cloud-view validation and repack (msfl_api.cu)
int check_cloud(const msfl_cloud *c, bool need_ring, const char *what);
// B strided AoS clouds -> packed float4 (dst4 + 4 off[b]) and, when ring_dst is given, uint16 rings (ring_dst + off[b]),
// split over the engine's pack threads
void pack_clouds_parallel(msfl_engine *e, int B, const msfl_cloud *clouds, float *dst4, uint16_t *ring_dst, const uint32_t *off);
int upload_cloud_packed(msfl_engine *e, const msfl_cloud *c, float4 *d_dst, uint16_t *d_ring_dst);

}  // namespace msfl
