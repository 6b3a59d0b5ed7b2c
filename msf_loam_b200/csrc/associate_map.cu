// associate_map.cu -- scan-to-map data association:
//   per corner query: 5-NN in the corner submap, d5^2 < 1 gate, centroid + 3x3 covariance +
//   eigen line test  -> edge factor constants        (mapping_scan_matcher.cc:109-176)
//   per surf query:   5-NN in the surf submap, gate, 5x3 least-squares plane, 0.2 m validity
//                                                    -> plane factor constants (:178-246)
// replacing pcl::KdTreeFLANN::nearestKSearch + Eigen.  One thread per query over a flat 1-D grid
// covering every query of the batch (scan id by binary search in the offset table).  kNN distances are fp32 ((dx*dx)+dy*dy)+dz*dz without FMA
// contraction (FLANN L2_Simple<float>), candidates are ordered by (d2, original index), the fit
// is fp64.  Output per query: 6 doubles [a_or_c(3), n(3)]; n = 0 marks "no factor".
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "assoc_device.cuh"
#include "msfl_internal.h"
#include "msfl_math.cuh"

namespace msfl {

// Per-query streams of a batch (queries, transformed queries, keys, ranks, factor constants: 0.3-0.7 GB per launch) are
// read or written exactly once per kernel, while the submap, its cell index and the counting-sort bin table (a few MB)
// are hit by every query: the streams carry the evict-first hint (ld.global.cs / st.global.cs) so that they do not push
// the hot tables out of L1 / L2.  -DMSFL_STREAM_HINTS=0 builds the plain loads / stores for A/B runs.
#ifndef MSFL_STREAM_HINTS
#define MSFL_STREAM_HINTS 1
#endif
template <typename T>
__device__ __forceinline__ T ld_stream(const T *p) {
#if MSFL_STREAM_HINTS
  return __ldcs(p);
#else
  return __ldg(p);
#endif
}
template <typename T>
__device__ __forceinline__ void st_stream(T *p, const T &v) {
#if MSFL_STREAM_HINTS
  __stcs(p, v);
#else
  *p = v;
#endif
}

__device__ __forceinline__ void store_corr(double *corr, size_t q, const double a[3], const double n[3]) {
  double2 *o = reinterpret_cast<double2 *>(corr + q * 6);
  st_stream(o, make_double2(a[0], a[1]));
  st_stream(o + 1, make_double2(a[2], n[0]));
  st_stream(o + 2, make_double2(n[1], n[2]));
}

// scan id of flat query k: largest b with off[b] <= k (off has B+1 ascending entries)
__device__ __forceinline__ int find_scan(const int32_t *__restrict__ off, int B, uint32_t k) {
  int lo = 0, hi = B;  // invariant: off[lo] <= k < off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((uint32_t)__ldg(off + mid) <= k) lo = mid;
    else hi = mid;
  }
  return lo;
}

// cell id of x in grid g, or -1 when x has no occupied neighbour cell
__device__ __forceinline__ int cell_of(const GridView &g, float x, float y, float z) {
  const int cx = (int)floorf(x * g.inv_edge) - g.ox, cy = (int)floorf(y * g.inv_edge) - g.oy,
            cz = (int)floorf(z * g.inv_edge) - g.oz;
  if (cx < 1 || cy < 1 || cz < 1 || cx > g.nx - 2 || cy > g.ny - 2 || cz > g.nz - 2) return -1;
  return (cz * g.ny + cy) * g.nx + cx;
}

// Pass 1 of the sorted path: transform every query by its scan's pose (mapping_scan_matcher.cc:123 /
// :193) and emit a sort key = cell id of the transformed point (corner grid first, then surf grid; one
// sentinel cell per class for queries with no occupied neighbourhood).
// COUNT: counting-sort form -- the key's bin counter is bumped and the returned rank (this query's position inside
// its bin) is stored instead of the query index; k_scatter_perm turns (key, rank) into the cell-order permutation
// after an exclusive scan of the bins.  The order inside a bin is whatever order the atomics happened in: the
// permutation is only a locality hint (every result is written at the query's own index), so poses do not depend on it.
template <bool COUNT>
__global__ void __launch_bounds__(256)
k_transform_keys(GridView gc, GridView gs, int B, const float4 *__restrict__ qc, const int32_t *__restrict__ c_off,
                 uint32_t n_corner_total, const float4 *__restrict__ qs, const int32_t *__restrict__ s_off,
                 uint32_t n_surf_total, const double *__restrict__ poses, float4 *__restrict__ xq,
                 uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t *__restrict__ hist, int sub_log2) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_total = n_corner_total + n_surf_total;
  const bool is_corner = k < n_corner_total;
  const uint32_t kk = is_corner ? k : k - n_corner_total;
  const int32_t *off = is_corner ? c_off : s_off;
  // scan id: the 32 queries of a warp are consecutive, so ONE binary search (lane 0) and a forward walk that almost
  // never takes a step replace 32 chains of log2(B) dependent loads; a warp that straddles the corner / surf boundary
  // searches per lane
  const bool class0 = __shfl_sync(0xffffffffu, is_corner, 0);
  int scan = 0;
  if ((threadIdx.x & 31) == 0 && k < n_total) scan = find_scan(off, B, kk);
  scan = __shfl_sync(0xffffffffu, scan, 0);
  if (k >= n_total) return;
  if (is_corner == class0) {
    while ((uint32_t)__ldg(off + scan + 1) <= kk) ++scan;
  } else {
    scan = find_scan(off, B, kk);
  }
  double pose[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) pose[i] = __ldg(poses + (size_t)scan * 7 + i);
  const float4 p = ld_stream((is_corner ? qc : qs) + kk);
  const float3 x = transform_point_f(pose, p.x, p.y, p.z);
  const uint32_t ncell_c = (uint32_t)(gc.nx * gc.ny * gc.nz), ncell_s = (uint32_t)(gs.nx * gs.ny * gs.nz);
  const int c = cell_of(is_corner ? gc : gs, x.x, x.y, x.z);
  uint32_t key;
  if (is_corner) key = c < 0 ? ncell_c : (uint32_t)c;
  else key = ncell_c + 1u + (c < 0 ? ncell_s : (uint32_t)c);
  // refine the order inside a cell by a sub-cell index (4x4x4 when the key has room: sub_log2 = 2) so the lanes of a
  // warp are spatial neighbours (<= 0.25 m apart): their row / cell pruning decisions in knn5_grid then agree
  if (sub_log2 > 0) {
    const GridView &g = is_corner ? gc : gs;
    const int ns = 1 << sub_log2;
    const float fs = (float)ns;
    const float ux = x.x * g.inv_edge, uy = x.y * g.inv_edge, uz = x.z * g.inv_edge;
    const int sx = min(ns - 1, max(0, (int)((ux - floorf(ux)) * fs))), sy = min(ns - 1, max(0, (int)((uy - floorf(uy)) * fs))),
              sz = min(ns - 1, max(0, (int)((uz - floorf(uz)) * fs)));
    key = (key << (3 * sub_log2)) | (uint32_t)((((sz << sub_log2) + sy) << sub_log2) + sx);
  }
  st_stream(xq + k, make_float4(x.x, x.y, x.z, 0.f));
  st_stream(keys + k, key);
  st_stream(vals + k, COUNT ? atomicAdd(hist + key, 1u) : k);
}

// counting sort, last pass: slot = first slot of the query's bin + its rank inside the bin
__global__ void __launch_bounds__(256)
k_scatter_perm(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ rank, const uint32_t *__restrict__ bin_start,
               uint32_t n, uint32_t *__restrict__ perm) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  perm[__ldg(bin_start + ld_stream(keys + k)) + ld_stream(rank + k)] = k;
}

// ---------------------------------------------------------------------------------------------
// IMU-deskew branch (mapping_scan_matcher.cc:112-121, :182-191 with
// is_initialized == true).  Per query: (dq, dp) = GetDeltaQP(preintegration, dt)
// (scan_undistortion.cc:22-42: upper_bound + Eigen slerp + lerp), dt = intensity; the kNN query is
//   TransformPoint(pose * Rigid3d{q^-1 (V dt - g dt^2 / 2) + dp, dq}, p)
// and the factor sees p' = dq p + dp (fp64) and the constant offset o = V dt - g dt^2 / 2
// (lidar_factor.cc:53,81), which is folded into the line point / plane centre (C' = C - o).
// ---------------------------------------------------------------------------------------------
struct DeskewTable {  // one scan's preintegration table as the prepare kernel sees it
  const double *sum_dt, *delta_q, *delta_p;  // device arrays: [n], [n][4] xyzw, [n][3]
  int n;
};

// what the search and fit kernels of the deskew branch need besides dsk: the per-query offset o = V dt - g dt^2 / 2
// (4 doubles per query), written by k_deskew_prepare with the velocity / gravity of the query's own scan
struct DeskewOffsets {
  const double *o4;
};

// o = Vi dt - 0.5 g dt^2 as the reference evaluates it (mapping_scan_matcher.cc:120, lidar_factor.cc:53)
__device__ __forceinline__ void deskew_offset_vg(const double V[3], const double G[3], double dt, double o[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = __dsub_rn(__dmul_rn(V[i], dt), __dmul_rn(__dmul_rn(__dmul_rn(0.5, G[i]), dt), dt));
}
__device__ __forceinline__ void deskew_offset(const DeskewOffsets &tb, size_t k, double o[3]) {
  const double2 a = *reinterpret_cast<const double2 *>(tb.o4 + k * 4);
  o[0] = a.x; o[1] = a.y; o[2] = tb.o4[k * 4 + 2];
}

// per query: dq(4) dp(3) dt(1) -> dsk[8]; p' -> pprime (double4).  Returns false when dt is out of range.
__device__ __forceinline__ bool deskew_prepare_one(const DeskewTable &tb, const float4 p, double *__restrict__ o,
                                                   double *__restrict__ pp) {
  const double dt = (double)p.w;  // auto dt = pointOri.intensity  (:114)
  if (tb.n < 2 || !(dt <= tb.sum_dt[tb.n - 1] && dt >= tb.sum_dt[0])) return false;  // CHECK scan_undistortion.cc:26
  int lo = 0, hi = tb.n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (dt < tb.sum_dt[mid]) hi = mid; else lo = mid + 1;
  }
  int idx = lo - 1;
  if (idx > tb.n - 2) idx = tb.n - 2;
  const double s = __ddiv_rn(__dsub_rn(dt, tb.sum_dt[idx]), __dsub_rn(tb.sum_dt[idx + 1], tb.sum_dt[idx]));
  const double *qa = tb.delta_q + 4 * idx, *qb = qa + 4;
  const double d = qa[0] * qb[0] + qa[1] * qb[1] + qa[2] * qb[2] + qa[3] * qb[3];
  double s0, s1;
  if (fabs(d) >= 1.0 - 2.220446049250313e-16) {
    s0 = 1.0 - s; s1 = s;
  } else {
    const double th = acos(fabs(d)), sn = sin(th);
    s0 = sin((1.0 - s) * th) / sn;
    s1 = sin(s * th) / sn;
  }
  if (d < 0) s1 = -s1;
  double dq[4], dp[3];
#pragma unroll
  for (int i = 0; i < 4; ++i) dq[i] = __dadd_rn(__dmul_rn(s0, qa[i]), __dmul_rn(s1, qb[i]));
  const double *pa = tb.delta_p + 3 * idx, *pb = pa + 3;
#pragma unroll
  for (int i = 0; i < 3; ++i) dp[i] = __dadd_rn(__dmul_rn(1 - s, pa[i]), __dmul_rn(s, pb[i]));
  o[0] = dq[0]; o[1] = dq[1]; o[2] = dq[2]; o[3] = dq[3]; o[4] = dp[0]; o[5] = dp[1]; o[6] = dp[2]; o[7] = dt;
  double r0, r1, r2;
  quat_rotate_exact(dq, (double)p.x, (double)p.y, (double)p.z, r0, r1, r2);
  pp[0] = __dadd_rn(r0, dp[0]); pp[1] = __dadd_rn(r1, dp[1]); pp[2] = __dadd_rn(r2, dp[2]); pp[3] = 0.0;
  return true;
}

// B scans, every scan with its own table, velocity and gravity: the query's scan comes from the offset tables of the
// [all corner | all surf] layout; also stores the per-query offset o (DeskewOffsets).  flags[scan] |= 1 when a dt of that
// scan is out of range.
__global__ void k_deskew_prepare(const DeskewScan *__restrict__ scans, const double *__restrict__ sum_dt,
                                       const double *__restrict__ delta_q, const double *__restrict__ delta_p, int B,
                                       const int32_t *__restrict__ c_off, const int32_t *__restrict__ s_off,
                                       uint32_t n_corner_total, const float4 *__restrict__ q, uint32_t n_total,
                                       double *__restrict__ dsk, double *__restrict__ pprime, double *__restrict__ o4,
                                       int *__restrict__ flags) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_total) return;
  const bool is_corner = k < n_corner_total;
  const int scan = find_scan(is_corner ? c_off : s_off, B, is_corner ? k : k - n_corner_total);
  const DeskewScan sc = scans[scan];
  const DeskewTable tb{sum_dt + sc.row0, delta_q + (size_t)sc.row0 * 4, delta_p + (size_t)sc.row0 * 3, sc.n};
  const float4 p = q[k];
  if (!deskew_prepare_one(tb, p, dsk + (size_t)k * 8, pprime + (size_t)k * 4)) {
    atomicOr(flags + scan, 1);
    return;
  }
  double o[3];
  deskew_offset_vg(sc.V, sc.G, (double)p.w, o);
  o4[(size_t)k * 4] = o[0]; o4[(size_t)k * 4 + 1] = o[1]; o4[(size_t)k * 4 + 2] = o[2]; o4[(size_t)k * 4 + 3] = 0.0;
}

// the total transform of the deskew branch applied to p, rounded to fp32 like TransformPoint
__device__ __forceinline__ float3 deskew_transform(const double pose[7], const DeskewOffsets &tb, size_t k, const double *dsk8,
                                                   float px, float py, float pz, double o[3]) {
  deskew_offset(tb, k, o);
  const double qc[4] = {-pose[3], -pose[4], -pose[5], pose[6]};
  double t0, t1, t2, u0, u1, u2;
  quat_rotate_exact(qc, o[0], o[1], o[2], t0, t1, t2);
  t0 = __dadd_rn(t0, dsk8[4]); t1 = __dadd_rn(t1, dsk8[5]); t2 = __dadd_rn(t2, dsk8[6]);
  quat_rotate_exact(pose + 3, t0, t1, t2, u0, u1, u2);
  const double *a = pose + 3, *b = dsk8;  // (q * dq).normalized(), Eigen product order
  double qt[4];
  qt[3] = __dsub_rn(__dsub_rn(__dsub_rn(__dmul_rn(a[3], b[3]), __dmul_rn(a[0], b[0])), __dmul_rn(a[1], b[1])), __dmul_rn(a[2], b[2]));
  qt[0] = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(a[3], b[0]), __dmul_rn(a[0], b[3])), __dmul_rn(a[1], b[2])), __dmul_rn(a[2], b[1]));
  qt[1] = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(a[3], b[1]), __dmul_rn(a[1], b[3])), __dmul_rn(a[2], b[0])), __dmul_rn(a[0], b[2]));
  qt[2] = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(a[3], b[2]), __dmul_rn(a[2], b[3])), __dmul_rn(a[0], b[1])), __dmul_rn(a[1], b[0]));
  const double nq = sqrt(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(qt[0], qt[0]), __dmul_rn(qt[1], qt[1])), __dmul_rn(qt[2], qt[2])), __dmul_rn(qt[3], qt[3])));
  double T[7] = {__dadd_rn(u0, pose[0]), __dadd_rn(u1, pose[1]), __dadd_rn(u2, pose[2]),
                 __ddiv_rn(qt[0], nq), __ddiv_rn(qt[1], nq), __ddiv_rn(qt[2], nq), __ddiv_rn(qt[3], nq)};
  return transform_point_f(T, px, py, pz);
}

// Search kernel.  SORTED = false: thread k handles flat query k and transforms it itself.
// SORTED = true: thread s handles query perm[s] (queries ordered by the cell of their transformed
// point, so the lanes of a warp walk the same candidate ranges: uniform trip counts and broadcast
// loads; the order is only a locality hint, the result does not depend on it).  STORED_X: read the
// transformed point stored by k_transform_keys instead of transforming here.
// BY_SLOT: the five indices are stored at the thread's slot (coalesced; k_fit<.., true> then walks the
// same cell order) instead of at the query's flat index k.
template <bool SORTED, bool STORED_X, bool DESKEW, bool BY_SLOT>
__global__ void __launch_bounds__(128, 10)
k_knn5(GridView gc, GridView gs, KParams kp, int B, const float4 *__restrict__ qc,
       const int32_t *__restrict__ c_off, uint32_t n_corner_total, const float4 *__restrict__ qs,
       const int32_t *__restrict__ s_off, uint32_t n_surf_total, const double *__restrict__ poses,
       const float4 *__restrict__ xq, const uint32_t *__restrict__ perm, int32_t *__restrict__ knn_out, DeskewOffsets tb,
       const double *__restrict__ dsk) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_corner_total + n_surf_total) return;
  const uint32_t k = SORTED ? __ldg(perm + slot) : slot;
  const bool is_corner = k < n_corner_total;
  float3 x;
  double dsk_o[3] = {0, 0, 0};  // deskew branch: per-point offset V dt - g dt^2 / 2
  if (STORED_X) {
    const float4 xs = __ldg(xq + k);
    x = make_float3(xs.x, xs.y, xs.z);
  } else {
    const uint32_t kk = is_corner ? k : k - n_corner_total;
    const int scan = find_scan(is_corner ? c_off : s_off, B, kk);
    double pose[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) pose[i] = __ldg(poses + (size_t)scan * 7 + i);
    const float4 p = __ldg((is_corner ? qc : qs) + kk);
    if (DESKEW) x = deskew_transform(pose, tb, k, dsk + (size_t)k * 8, p.x, p.y, p.z, dsk_o);  // :120 / :190
    else x = transform_point_f(pose, p.x, p.y, p.z);  // mapping_scan_matcher.cc:123 / :193
  }
  const GridView &g = is_corner ? gc : gs;
  Top5 t;
  const bool gate = knn5_grid(g, x.x, x.y, x.z, kp.knn_max_sq_f, t);  // :125-128 / :195-198
  // 5 neighbour indices per query (-1 when the d5^2 gate fails), consumed by k_fit
  int32_t *o = knn_out + (size_t)(BY_SLOT ? slot : k) * 5;
#pragma unroll
  for (int s = 0; s < 5; ++s) o[s] = gate ? t.i[s] : -1;
}

// ---------------------------------------------------------------------------------------------
// Batch search kernel with TMA-staged submap tiles.  The slots of a batch are ordered by submap cell,
// so the 128 queries of a CTA almost always sit in ONE cell (thousands of queries per cell at batch
// sizes): the CTA takes the cell of its first query as anchor, warp 0 reads the 9 row ranges of the
// anchor's 3x3x3 neighbourhood from the cell index (the three x-adjacent cells of a row are one
// contiguous range of pts_sorted) and copies them into shared memory with cp.async.bulk (one bulk
// copy per row, completing on an mbarrier), together with a 9-entry row table.  Every query of the
// anchor cell then runs the same exact pruned search as knn5_grid against shared memory -- no
// cell-index arithmetic, no L2 round trips; queries of another cell (CTAs that straddle a cell
// boundary) or oversized neighbourhoods take the global-memory path.  Results are identical.
// ---------------------------------------------------------------------------------------------
constexpr int kTileCap = 1024;  // staged points per CTA (16 KB)

__device__ __forceinline__ uint32_t a_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// same visiting order, bounds and (d2, index) candidate order as knn5_grid; rows[r] = tile indices
// {left cell start, centre cell start, right cell start, end} of row r (nearest-first order)
__device__ __forceinline__ bool knn5_tile(const float4 *__restrict__ tile, const uint4 *__restrict__ rows, bool exact_cells, float qx,
                                          float qy, float qz, float fxq, float fyq, float fzq, float thresh, Top5 &t) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    t.d[s] = thresh;
    t.i[s] = 0;
  }
  const float lox = exact_cells ? __fsub_rn(qx, fxq) : 0.f, hix = exact_cells ? __fsub_rn(fxq + 1.0f, qx) : 0.f;
  const float loy = exact_cells ? __fsub_rn(qy, fyq) : 0.f, hiy = exact_cells ? __fsub_rn(fyq + 1.0f, qy) : 0.f;
  const float loz = exact_cells ? __fsub_rn(qz, fzq) : 0.f, hiz = exact_cells ? __fsub_rn(fzq + 1.0f, qz) : 0.f;
  const float bxl = __fmul_rn(lox, lox), bxr = __fmul_rn(hix, hix);
  constexpr uint32_t kDyPacked = 1u | (0u << 2) | (2u << 4) | (1u << 6) | (1u << 8) | (0u << 10) | (0u << 12) | (2u << 14) | (2u << 16);
  constexpr uint32_t kDzPacked = 1u | (1u << 2) | (1u << 4) | (0u << 6) | (2u << 8) | (0u << 10) | (2u << 12) | (0u << 14) | (2u << 16);
#pragma unroll 1
  for (int r = 0; r < 9; ++r) {
    const int dy = (int)((kDyPacked >> (2 * r)) & 3u) - 1, dz = (int)((kDzPacked >> (2 * r)) & 3u) - 1;
    const float by = dy == 0 ? 0.f : (dy < 0 ? loy : hiy), bz = dz == 0 ? 0.f : (dz < 0 ? loz : hiz);
    const float by2 = __fmul_rn(by, by), bz2 = __fmul_rn(bz, bz);
    const float row_lb = __fadd_rn(by2, bz2);
    if (row_lb > t.d[4] || row_lb >= thresh) continue;
    const uint4 rr = rows[r];
    const float lbl = __fadd_rn(__fadd_rn(bxl, by2), bz2), lbr = __fadd_rn(__fadd_rn(bxr, by2), bz2);
    const uint32_t js = (lbl > t.d[4] || lbl >= thresh) ? rr.y : rr.x;
    const uint32_t je = (lbr > t.d[4] || lbr >= thresh) ? rr.z : rr.w;
#pragma unroll 2
    for (uint32_t j = js; j < je; ++j) {
      const float4 m = tile[j];
      const float dx = __fsub_rn(qx, m.x), dy2 = __fsub_rn(qy, m.y), dz2 = __fsub_rn(qz, m.z);
      float d = __fmul_rn(dx, dx);
      d = __fadd_rn(d, __fmul_rn(dy2, dy2));
      d = __fadd_rn(d, __fmul_rn(dz2, dz2));
      if (d <= t.d[4]) {
        const int id = __float_as_int(m.w);
        if (cand_less(d, id, t.d[4], t.i[4])) top5_insert(t, d, id);
      }
    }
  }
  return t.d[4] < thresh;
}

__global__ void __launch_bounds__(128, 10)
k_knn5_tiled(GridView gc, GridView gs, KParams kp, uint32_t n_corner_total, uint32_t n_total, const float4 *__restrict__ xq,
             const uint32_t *__restrict__ perm, int32_t *__restrict__ knn_out) {
  __shared__ __align__(128) float4 tile[kTileCap];
  __shared__ __align__(16) uint4 rows[9];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int anchor[5];  // is_corner, cx, cy, cz (grid-relative), staged
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = slot < n_total;
  uint32_t k = 0;
  float3 x = make_float3(0.f, 0.f, 0.f);
  bool is_corner = false;
  if (live) {
    k = __ldg(perm + slot);
    is_corner = k < n_corner_total;
    const float4 xs = __ldg(xq + k);
    x = make_float3(xs.x, xs.y, xs.z);
  }
  const GridView &g = is_corner ? gc : gs;
  const float fxq = floorf(x.x * g.inv_edge), fyq = floorf(x.y * g.inv_edge), fzq = floorf(x.z * g.inv_edge);
  const int cx = (int)fxq - g.ox, cy = (int)fyq - g.oy, cz = (int)fzq - g.oz;
  if (threadIdx.x == 0) {  // slot of thread 0 is always live
    const bool inside = !(cx < 1 || cy < 1 || cz < 1 || cx > g.nx - 2 || cy > g.ny - 2 || cz > g.nz - 2);
    anchor[0] = is_corner;
    anchor[1] = cx; anchor[2] = cy; anchor[3] = cz;
    anchor[4] = inside;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a_smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32 && anchor[4]) {  // warp 0: row table + one bulk copy per row
    const uint32_t lane = threadIdx.x;
    const GridView &ga = anchor[0] ? gc : gs;
    constexpr uint32_t kDyPacked = 1u | (0u << 2) | (2u << 4) | (1u << 6) | (1u << 8) | (0u << 10) | (0u << 12) | (2u << 14) | (2u << 16);
    constexpr uint32_t kDzPacked = 1u | (1u << 2) | (1u << 4) | (0u << 6) | (2u << 8) | (0u << 10) | (2u << 12) | (0u << 14) | (2u << 16);
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    if (lane < 9) {
      const int dy = (int)((kDyPacked >> (2 * lane)) & 3u) - 1, dz = (int)((kDzPacked >> (2 * lane)) & 3u) - 1;
      const int row = ((anchor[3] + dz) * ga.ny + (anchor[2] + dy)) * ga.nx + anchor[1];
      s0 = __ldg(ga.cell_start + row - 1); s1 = __ldg(ga.cell_start + row);
      s2 = __ldg(ga.cell_start + row + 1); s3 = __ldg(ga.cell_start + row + 2);
    }
    const uint32_t cnt = s3 - s0;
    uint32_t incl = cnt;  // inclusive prefix over the lanes (rows)
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 8);
    const bool staged = total <= (uint32_t)kTileCap;
    const uint32_t base = incl - cnt;
    if (lane == 0) {
      anchor[4] = staged ? 2 : 1;  // 2 = tile staged, 1 = anchor valid but neighbourhood too large
      if (staged)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a_smem_u32(&bar)), "r"(total * 16u) : "memory");
    }
    __syncwarp();
    if (staged && lane < 9) {
      rows[lane] = make_uint4(base, base + (s1 - s0), base + (s2 - s0), base + cnt);
      if (cnt)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         a_smem_u32(tile + base)),
                     "l"(ga.pts_sorted + s0), "r"(cnt * 16u), "r"(a_smem_u32(&bar))
                     : "memory");
    }
  }
  __syncthreads();
  const bool staged = anchor[4] == 2;
  if (staged) {  // every thread waits for the tile (phase 0 of the one-shot barrier)
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(a_smem_u32(&bar)),
        "r"(0u)
        : "memory");
  }
  if (!live) return;
  Top5 t;
  bool gate;
  if (staged && (int)is_corner == anchor[0] && cx == anchor[1] && cy == anchor[2] && cz == anchor[3])
    gate = knn5_tile(tile, rows, g.inv_edge == 1.0f, x.x, x.y, x.z, fxq, fyq, fzq, kp.knn_max_sq_f, t);
  else
    gate = knn5_grid(g, x.x, x.y, x.z, kp.knn_max_sq_f, t);
  int32_t *o = knn_out + (size_t)slot * 5;
#pragma unroll
  for (int s = 0; s < 5; ++s) o[s] = gate ? t.i[s] : -1;
}

// Line / plane fit of one gated query (fp64): reads its five neighbours and writes the factor constants
// [a_or_c(3), n(3)]; n = 0 marks "no factor".
// COMPACT: plane entries are written as 32 B {n, n.c} at corr + 48 n_corner_total + 32 i (the batch path: the LM
// kernel only ever needs the plane's offset along its normal); otherwise 48 B {c, n} like the edge entries.
// BY_SLOT: slot s holds query perm[s] and the neighbour indices k_knn5 stored at slot s.
// QR: planes take the pivoted-Householder solve (the fallback kernel); otherwise plane_fit_fast, and a query that
// needs the Householder path is appended to fb_list instead of being written.
// fit_from_idx: the neighbour indices are already in registers (the fused search + fit kernel); k = the query's flat
// index, slot = what a declined plane query is listed under.
template <bool DESKEW, bool COMPACT, bool QR>
__device__ __forceinline__ void fit_from_idx(const GridView &g, const KParams &kp, bool is_corner, uint32_t k, uint32_t slot,
                                             uint32_t n_corner_total, const int (&idx)[5], double *__restrict__ corr,
                                             const DeskewOffsets &tb, const double *__restrict__ dsk, uint32_t *__restrict__ fb_list,
                                             uint32_t *__restrict__ fb_count) {
  double a[3] = {0, 0, 0}, n[3] = {0, 0, 0};
  if (idx[4] >= 0) {
    // the five neighbours stay in registers as the fp32 values they are (15 registers, not 30) and are widened
    // where they are used -- the conversion is exact
    float mf[5][3];
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      const float4 mp = __ldg(g.pts_orig + idx[s]);
      mf[s][0] = mp.x; mf[s][1] = mp.y; mf[s][2] = mp.z;
    }
    double c[3];
    centroid5(mf, c);
    if (is_corner) {
      line_fit(mf, c, kp, a, n);
    } else if (QR) {
      plane_fit_qr(mf, c, kp, n);
      if (n[0] != 0.0 || n[1] != 0.0 || n[2] != 0.0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) a[d] = c[d];
      }
    } else {
      if (plane_fit_fast(mf, c, kp.plane_tol, n)) {
        fb_list[atomicAdd(fb_count, 1u)] = slot;  // the Householder kernel writes this entry
        return;
      }
      if (n[0] != 0.0 || n[1] != 0.0 || n[2] != 0.0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) a[d] = c[d];
      }
    }
  }
  if (DESKEW && (n[0] != 0.0 || n[1] != 0.0 || n[2] != 0.0)) {  // fold the constant offset: C' = C - o, o = V dt - g dt^2 / 2
    double off[3];
    deskew_offset(tb, k, off);
#pragma unroll
    for (int d = 0; d < 3; ++d) a[d] -= off[d];
  }
  if (COMPACT && !is_corner) {
    double2 *o = reinterpret_cast<double2 *>(reinterpret_cast<unsigned char *>(corr + (size_t)n_corner_total * 6) +
                                             (size_t)(k - n_corner_total) * 32);
    o[0] = make_double2(n[0], n[1]);
    o[1] = make_double2(n[2], __fma_rn(n[2], a[2], __fma_rn(n[1], a[1], __dmul_rn(n[0], a[0]))));
  } else {
    store_corr(corr, k, a, n);
  }
}

template <bool DESKEW, bool COMPACT, bool BY_SLOT, bool QR>
__device__ __forceinline__ void fit_query(const GridView &g, const KParams &kp, bool is_corner, uint32_t slot, uint32_t n_corner_total,
                                          const int32_t *__restrict__ knn, double *__restrict__ corr, const DeskewOffsets &tb,
                                          const double *__restrict__ dsk, const uint32_t *__restrict__ perm,
                                          uint32_t *__restrict__ fb_list, uint32_t *__restrict__ fb_count) {
  const uint32_t k = BY_SLOT ? __ldg(perm + slot) : slot;
  int idx[5];
#pragma unroll
  for (int s = 0; s < 5; ++s) idx[s] = __ldg(knn + (size_t)slot * 5 + s);
  fit_from_idx<DESKEW, COMPACT, QR>(g, kp, is_corner, k, slot, n_corner_total, idx, corr, tb, dsk, fb_list, fb_count);
}

// Batch path, fused: the search of k_knn5<true, true, false, true> followed by the fit on the indices still in
// registers -- closed-form plane fit for surf slots, closed-form eigen line test for corner slots -- so the neighbour
// lists (20 B per query written and read back) never touch memory and the whole association of an outer iteration is
// ONE kernel after the sort; only the plane queries the closed form declines store their list and go to the
// Householder kernel.
#ifndef MSFL_KNNFIT_MINB
#define MSFL_KNNFIT_MINB 10
#endif
__global__ void __launch_bounds__(128, MSFL_KNNFIT_MINB)
k_knn5_fit(GridView gc, GridView gs, KParams kp, uint32_t n_corner_total, uint32_t n_total, const float4 *__restrict__ xq,
           const uint32_t *__restrict__ perm, int32_t *__restrict__ knn_out, double *__restrict__ corr,
           uint32_t *__restrict__ fb_list, uint32_t *__restrict__ fb_count) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_total) return;
  const uint32_t k = ld_stream(perm + slot);
  const bool is_corner = slot < n_corner_total;
  const float4 xs = ld_stream(xq + k);
  const GridView &g = is_corner ? gc : gs;
  Top5 t;
  const bool gate = knn5_grid(g, xs.x, xs.y, xs.z, kp.knn_max_sq_f, t);
  int idx[5];
#pragma unroll
  for (int s = 0; s < 5; ++s) idx[s] = gate ? t.i[s] : -1;
  float mf[5][3];
  if (gate) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      const float4 mp = __ldg(g.pts_orig + idx[s]);
      mf[s][0] = mp.x; mf[s][1] = mp.y; mf[s][2] = mp.z;
    }
  }
  double c[3] = {0, 0, 0}, n[3] = {0, 0, 0};
  if (gate) centroid5(mf, c);
  if (is_corner) {  // line test + direction (closed-form eigen-solver): the edge entry {a, n}, 48 B
    double a[3] = {0, 0, 0};
    if (gate) line_fit(mf, c, kp, a, n);
    store_corr(corr, k, a, n);
    return;
  }
  // a declined plane query is re-fitted by k_fit_qr_list from its stored list
  bool declined = false;
  if (gate) declined = plane_fit_fast(mf, c, kp.plane_tol, n);
  if (declined) {
#pragma unroll
    for (int s = 0; s < 5; ++s) knn_out[(size_t)slot * 5 + s] = idx[s];
    fb_list[atomicAdd(fb_count, 1u)] = slot;
    return;
  }
  double2 *o = reinterpret_cast<double2 *>(reinterpret_cast<unsigned char *>(corr + (size_t)n_corner_total * 6) +
                                           (size_t)(k - n_corner_total) * 32);
  st_stream(o, make_double2(n[0], n[1]));
  st_stream(o + 1, make_double2(n[2], __fma_rn(n[2], c[2], __fma_rn(n[1], c[1], __dmul_rn(n[0], c[0])))));
}

// One thread per query.  Kept apart from the search kernel so that the search runs at 40 registers / 75 %
// occupancy while the fits do not throttle it.
// CLS: -1 = one launch over all queries (class decided per thread); 0 / 1 = a launch over the corner / surf queries
// only (the batch path: in cell order as in flat order all corner queries come first), so the Jacobi eigen-solver
// and the plane fit each get their own register allocation and a warp never holds both classes.
template <bool DESKEW, bool COMPACT, bool BY_SLOT, int CLS = -1>
__global__ void __launch_bounds__(128)
k_fit(GridView gc, GridView gs, KParams kp, uint32_t n_corner_total, uint32_t n_total, const int32_t *__restrict__ knn,
      double *__restrict__ corr, DeskewOffsets tb, const double *__restrict__ dsk, const uint32_t *__restrict__ perm,
      uint32_t *__restrict__ fb_list, uint32_t *__restrict__ fb_count) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x + (CLS == 1 ? n_corner_total : 0u);
  if (slot >= (CLS == 0 ? n_corner_total : n_total)) return;
  // corner slots come first in cell order as in flat order, so the class follows from the slot in both
  const bool is_corner = CLS < 0 ? slot < n_corner_total : CLS == 0;
  fit_query<DESKEW, COMPACT, BY_SLOT, false>(is_corner ? gc : gs, kp, is_corner, slot, n_corner_total, knn, corr, tb, dsk, perm,
                                             fb_list, fb_count);
}

// The Householder path for the plane queries plane_fit_fast declined (grid-stride over the list; its length is only
// known on the device).  The list order is arbitrary, but every entry is written at its own query's place.
template <bool DESKEW, bool COMPACT, bool BY_SLOT>
__global__ void __launch_bounds__(128)
k_fit_qr_list(GridView gs, KParams kp, uint32_t n_corner_total, const int32_t *__restrict__ knn, double *__restrict__ corr,
              DeskewOffsets tb, const double *__restrict__ dsk, const uint32_t *__restrict__ perm,
              const uint32_t *__restrict__ fb_list, const uint32_t *__restrict__ fb_count) {
  const uint32_t cnt = *fb_count;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x)
    fit_query<DESKEW, COMPACT, BY_SLOT, true>(gs, kp, false, fb_list[i], n_corner_total, knn, corr, tb, dsk, perm, nullptr, nullptr);
}

// grid of the Householder fallback kernel: it strides over a list whose length the host does not know
static unsigned qr_list_grid(const msfl_engine *e, uint32_t n_plane_queries) {
  const unsigned want = (n_plane_queries / 64u + 127u) / 128u + 1u;  // ~1.5 % of the queries land on the list
  return std::min(want, (unsigned)e->sm_count * 4u);
}

int launch_associate_map(msfl_engine *e, int B, const float4 *d_qc, const int32_t *d_c_off, uint32_t n_corner_total,
                         const float4 *d_qs, const int32_t *d_s_off, uint32_t n_surf_total, const double *d_poses,
                         double *d_corr, int32_t *d_knn, bool compact) {
  const uint32_t total = n_corner_total + n_surf_total;
  if (B <= 0 || total == 0) return MSFL_OK;
  const int tb = 128;
  const GridView &gc = e->map_corner.view, &gs = e->map_surf.view;
  const bool own_knn = d_knn == nullptr;
  int rc;
  if (own_knn) {  // neighbour indices travel from the search kernel to the fit kernel through this scratch
    if ((rc = e->d_knn.reserve((size_t)total * 5 * 4))) return rc;
    d_knn = e->d_knn.as<int32_t>();
  }
  if ((rc = e->a_fb.reserve(((size_t)n_surf_total + 2) * 4))) return rc;
  uint32_t *fb_count = e->a_fb.as<uint32_t>(), *fb_list = fb_count + 1;
  MSFL_CUDA_OK(cudaMemsetAsync(fb_count, 0, 4, e->stream));
  const DeskewOffsets nt{};
  const int mode = e->params.assoc_sorted;  // 0 auto, 1 never, 2 always
  const bool sorted = mode == 2 || mode == 3 || (mode == 0 && total >= 65536u);
  if (!sorted) {
    stage_begin(e, 0);
    const unsigned grid = (total + tb - 1) / tb;
    k_knn5<false, false, false, false><<<grid, tb, 0, e->stream>>>(gc, gs, e->kp, B, d_qc, d_c_off, n_corner_total, d_qs, d_s_off,
                                                                  n_surf_total, d_poses, nullptr, nullptr, d_knn, nt, nullptr);
    if (compact) {
      k_fit<false, true, false><<<grid, tb, 0, e->stream>>>(gc, gs, e->kp, n_corner_total, total, d_knn, d_corr, nt, nullptr, nullptr, fb_list, fb_count);
      k_fit_qr_list<false, true, false><<<qr_list_grid(e, n_surf_total), tb, 0, e->stream>>>(gs, e->kp, n_corner_total, d_knn, d_corr, nt, nullptr, nullptr, fb_list, fb_count);
    } else {
      k_fit<false, false, false><<<grid, tb, 0, e->stream>>>(gc, gs, e->kp, n_corner_total, total, d_knn, d_corr, nt, nullptr, nullptr, fb_list, fb_count);
      k_fit_qr_list<false, false, false><<<qr_list_grid(e, n_surf_total), tb, 0, e->stream>>>(gs, e->kp, n_corner_total, d_knn, d_corr, nt, nullptr, nullptr, fb_list, fb_count);
    }
    stage_end(e);
    e->launches += 3;
    MSFL_CUDA_OK(cudaGetLastError());
    return MSFL_OK;
  }
  // sorted path: transform + cell keys -> counting / radix sort -> association in cell order
  if ((rc = e->a_xq.reserve((size_t)total * 16))) return rc;
  if ((rc = e->a_keys.reserve((size_t)total * 4))) return rc;
  if ((rc = e->a_keys_alt.reserve((size_t)total * 4))) return rc;
  if ((rc = e->a_vals.reserve((size_t)total * 4))) return rc;
  if ((rc = e->a_vals_alt.reserve((size_t)total * 4))) return rc;
  // The cell order is rebuilt for every outer iteration.  Measured on B200 (VLP-16, 2048 scans, 0.10 m / 1 deg initial
  // error: the first solve moves the queries by 0.2 m median, 0.47 m max): keeping the first order costs the second
  // search 40 % (its lanes stop agreeing on rows and cells), more than the 0.27 ms re-sort; seeding the second search
  // with the distance bound of the first iteration's neighbours (9 % of the lists survive unchanged, 28 % as sets) gains
  // nothing either -- the centre row establishes the same bound after ~10 candidates.
  {
    const long long ncell = (long long)gc.nx * gc.ny * gc.nz + (long long)gs.nx * gs.ny * gs.nz + 2;
    int cell_bits = 1;
    while ((1ll << cell_bits) < ncell) ++cell_bits;
    // sub-cell refinement of the key: 4x4x4 when the 32-bit key has room, 2x2x2 or none for far-spread submaps
    const int sub_log2 = cell_bits + 6 <= 32 ? 2 : (cell_bits + 3 <= 32 ? 1 : 0);
    const int end_bit = cell_bits + 3 * sub_log2;
    if (end_bit > 32) { set_error("submap grid too large for the sorted association path"); return MSFL_ERR_GRID; }  // ncell <= 2^27 + 2: unreachable
    const long long nbins = ncell << (3 * sub_log2);
    if (nbins <= e->count_sort_max_bins) {
      // counting sort: one atomic per query into a bin table that lives in L2, a scan of the bins, one scatter
      size_t tmp = 0;
      if ((rc = e->a_hist.reserve((size_t)nbins * 4))) return rc;
      uint32_t *hist = e->a_hist.as<uint32_t>();
      MSFL_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, hist, hist, (int)nbins, e->stream));
      if ((rc = e->a_tmp.reserve(tmp))) return rc;
      stage_begin(e, 2);
      MSFL_CUDA_OK(cudaMemsetAsync(hist, 0, (size_t)nbins * 4, e->stream));
      k_transform_keys<true><<<(total + 255) / 256, 256, 0, e->stream>>>(gc, gs, B, d_qc, d_c_off, n_corner_total, d_qs, d_s_off,
                                                                         n_surf_total, d_poses, e->a_xq.as<float4>(),
                                                                         e->a_keys.as<uint32_t>(), e->a_vals.as<uint32_t>(), hist, sub_log2);
      MSFL_CUDA_OK(cub::DeviceScan::ExclusiveSum(e->a_tmp.p, tmp, hist, hist, (int)nbins, e->stream));
      k_scatter_perm<<<(total + 255) / 256, 256, 0, e->stream>>>(e->a_keys.as<uint32_t>(), e->a_vals.as<uint32_t>(), hist, total,
                                                                 e->a_vals_alt.as<uint32_t>());
      stage_end(e);
      e->a_perm = e->a_vals_alt.as<uint32_t>();
      e->launches += 3;  // + the cub scan kernels
    } else {
      cub::DoubleBuffer<uint32_t> dk(e->a_keys.as<uint32_t>(), e->a_keys_alt.as<uint32_t>()),
          dv(e->a_vals.as<uint32_t>(), e->a_vals_alt.as<uint32_t>());
      size_t tmp = 0;
      MSFL_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int)total, 0, end_bit, e->stream));
      if ((rc = e->a_tmp.reserve(tmp))) return rc;
      stage_begin(e, 2);
      k_transform_keys<false><<<(total + 255) / 256, 256, 0, e->stream>>>(gc, gs, B, d_qc, d_c_off, n_corner_total, d_qs, d_s_off,
                                                                          n_surf_total, d_poses, e->a_xq.as<float4>(), dk.Current(),
                                                                          dv.Current(), nullptr, sub_log2);
      MSFL_CUDA_OK(cub::DeviceRadixSort::SortPairs(e->a_tmp.p, tmp, dk, dv, (int)total, 0, end_bit, e->stream));
      stage_end(e);
      e->a_perm = dv.Current();
      e->launches += 3;
    }
  }
  // neighbour indices stay in cell order between the two kernels unless the caller wants them back (test hook)
  const bool by_slot = own_knn;
  const unsigned grid = (total + tb - 1) / tb;
  stage_begin(e, 0);
  const bool fused = by_slot && compact && e->fuse_fit && mode != 3;
  if (fused)
    k_knn5_fit<<<grid, tb, 0, e->stream>>>(gc, gs, e->kp, n_corner_total, total, e->a_xq.as<float4>(), e->a_perm, d_knn, d_corr, fb_list, fb_count);
  else if (by_slot && mode == 3)  // measured on B200 (VLP-16, 2048 scans): 0.523 ms staged vs 0.507 ms direct -- the search is
                                  // instruction-issue-bound (78 % of issue slots), not latency-bound, so staging is opt-in
    k_knn5_tiled<<<grid, tb, 0, e->stream>>>(gc, gs, e->kp, n_corner_total, total, e->a_xq.as<float4>(), e->a_perm, d_knn);
  else if (by_slot)
    k_knn5<true, true, false, true><<<grid, tb, 0, e->stream>>>(gc, gs, e->kp, B, d_qc, d_c_off, n_corner_total, d_qs, d_s_off,
                                                               n_surf_total, d_poses, e->a_xq.as<float4>(), e->a_perm, d_knn, nt, nullptr);
  else
    k_knn5<true, true, false, false><<<grid, tb, 0, e->stream>>>(gc, gs, e->kp, B, d_qc, d_c_off, n_corner_total, d_qs, d_s_off,
                                                                n_surf_total, d_poses, e->a_xq.as<float4>(), e->a_perm, d_knn, nt, nullptr);
  stage_end(e);
  stage_begin(e, 3);
  {
    if (by_slot && compact) {
      const unsigned grid_c = (n_corner_total + tb - 1) / tb, grid_s = (n_surf_total + tb - 1) / tb;
      if (grid_c && !fused) k_fit<false, true, true, 0><<<grid_c, tb, 0, e->stream>>>(gc, gs, e->kp, n_corner_total, total, d_knn, d_corr, nt, nullptr, e->a_perm, fb_list, fb_count);
      if (grid_s) {
        if (!fused) k_fit<false, true, true, 1><<<grid_s, tb, 0, e->stream>>>(gc, gs, e->kp, n_corner_total, total, d_knn, d_corr, nt, nullptr, e->a_perm, fb_list, fb_count);
        k_fit_qr_list<false, true, true><<<qr_list_grid(e, n_surf_total), tb, 0, e->stream>>>(gs, e->kp, n_corner_total, d_knn, d_corr, nt, nullptr, e->a_perm, fb_list, fb_count);
      }
      e->launches += 2;
    } else if (by_slot) {
      k_fit<false, false, true><<<grid, tb, 0, e->stream>>>(gc, gs, e->kp, n_corner_total, total, d_knn, d_corr, nt, nullptr, e->a_perm, fb_list, fb_count);
      k_fit_qr_list<false, false, true><<<qr_list_grid(e, n_surf_total), tb, 0, e->stream>>>(gs, e->kp, n_corner_total, d_knn, d_corr, nt, nullptr, e->a_perm, fb_list, fb_count);
      e->launches += 1;
    } else {
      if (compact) {
        k_fit<false, true, false><<<grid, tb, 0, e->stream>>>(gc, gs, e->kp, n_corner_total, total, d_knn, d_corr, nt, nullptr, nullptr, fb_list, fb_count);
        k_fit_qr_list<false, true, false><<<qr_list_grid(e, n_surf_total), tb, 0, e->stream>>>(gs, e->kp, n_corner_total, d_knn, d_corr, nt, nullptr, nullptr, fb_list, fb_count);
      } else {
        k_fit<false, false, false><<<grid, tb, 0, e->stream>>>(gc, gs, e->kp, n_corner_total, total, d_knn, d_corr, nt, nullptr, nullptr, fb_list, fb_count);
        k_fit_qr_list<false, false, false><<<qr_list_grid(e, n_surf_total), tb, 0, e->stream>>>(gs, e->kp, n_corner_total, d_knn, d_corr, nt, nullptr, nullptr, fb_list, fb_count);
      }
      e->launches += 1;
    }
  }
  stage_end(e);
  e->launches += 3;
  MSFL_CUDA_OK(cudaGetLastError());
  return MSFL_OK;
}

// ---- deskew branch launchers -----------------------------------------------------------------------
int launch_deskew_prepare(msfl_engine *e, int B, const DeskewScan *d_scans, const double *d_sum_dt, const double *d_dq,
                          const double *d_dp, const float4 *d_q, const int32_t *d_c_off, const int32_t *d_s_off, uint32_t nc,
                          uint32_t n, double *d_dsk, double *d_pprime, double *d_o4, int *d_flags) {
  if (n == 0) return MSFL_OK;
  k_deskew_prepare<<<(n + 127) / 128, 128, 0, e->stream>>>(d_scans, d_sum_dt, d_dq, d_dp, B, d_c_off, d_s_off, nc, d_q, n, d_dsk,
                                                           d_pprime, d_o4, d_flags);
  e->launches += 1;
  MSFL_CUDA_OK(cudaGetLastError());
  return MSFL_OK;
}

int launch_associate_map_deskew(msfl_engine *e, int B, const float4 *d_qc, const int32_t *d_c_off, uint32_t nc,
                                const float4 *d_qs, const int32_t *d_s_off, uint32_t ns, const double *d_pose,
                                const double *d_o4, const double *d_dsk, double *d_corr, int32_t *d_knn) {
  const uint32_t total = nc + ns;
  if (total == 0) return MSFL_OK;
  if (!d_knn) {
    int rck;
    if ((rck = e->d_knn.reserve((size_t)total * 5 * 4))) return rck;
    d_knn = e->d_knn.as<int32_t>();
  }
  const DeskewOffsets tb{d_o4};
  stage_begin(e, 0);
  k_knn5<false, false, true, false><<<(total + 127) / 128, 128, 0, e->stream>>>(
      e->map_corner.view, e->map_surf.view, e->kp, B, d_qc, d_c_off, nc, d_qs, d_s_off, ns, d_pose, nullptr, nullptr, d_knn, tb,
      d_dsk);
  int rcf;
  if ((rcf = e->a_fb.reserve(((size_t)ns + 2) * 4))) return rcf;
  uint32_t *fb_count = e->a_fb.as<uint32_t>(), *fb_list = fb_count + 1;
  MSFL_CUDA_OK(cudaMemsetAsync(fb_count, 0, 4, e->stream));
  k_fit<true, false, false><<<(total + 127) / 128, 128, 0, e->stream>>>(e->map_corner.view, e->map_surf.view, e->kp, nc, total, d_knn, d_corr, tb, d_dsk, nullptr, fb_list, fb_count);
  k_fit_qr_list<true, false, false><<<qr_list_grid(e, ns), 128, 0, e->stream>>>(e->map_surf.view, e->kp, nc, d_knn, d_corr, tb, d_dsk, nullptr, fb_list, fb_count);
  stage_end(e);
  e->launches += 3;
  MSFL_CUDA_OK(cudaGetLastError());
  return MSFL_OK;
}

}  // namespace msfl
