// scan2scan.cu -- scan-to-scan data association, OdometryScanMatcher::MatchScan2Scan
// (odometry_scan_matcher.cc:43-285):
//   per current SHARP point: 1-NN in last less-sharp (d2 < 25), second point = nearest (d2 < 25)
//   in a different ring within +-2.5 rings                     -> LidarEdgeFactorSE3(p, a, (a-b)/|a-b|)
//   per current FLAT point : 1-NN in last less-flat, j = nearest in the SAME ring, l = nearest in
//   another ring within +-2.5 rings                            -> LidarPlaneFactorSE3(p, a, b, c)
// The reference walks the ring-sorted arrays linearly; given ring-sorted input (which its own
// extraction guarantees and this entry point checks) those walks are "nearest point subject to a
// ring filter", evaluated here on the 1 m cell index by expanding Chebyshev shells (one warp per query) until the
// best candidate is provably nearest or the 5 m radius is exhausted.  Ties (equal fp32 distance) follow
// the reference's visiting order: forward indices ascending, then backward indices descending.
#include <cub/device/device_scan.cuh>

#include <limits.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "msfl_internal.h"
#include "msfl_math.cuh"

namespace msfl {

struct ScanGrid {
  GridView g;
  const uint16_t *ring;  // per original index
  uint32_t n;
};

__device__ __forceinline__ float sqdist_f(float qx, float qy, float qz, const float4 m) {
  // odometry_scan_matcher.cc:102-108: (a-b)*(a-b) + ... in float, left to right, no contraction.
  // (point - query) and (query - point) square to the same value.
  const float dx = __fsub_rn(m.x, qx), dy = __fsub_rn(m.y, qy), dz = __fsub_rn(m.z, qz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// visiting-order priority of candidate j relative to the closest point (lower = visited earlier)
__device__ __forceinline__ uint32_t visit_prio(int j, int closest, uint32_t n) {
  return j > closest ? (uint32_t)(j - closest) : (uint32_t)(n + closest - j);
}

struct Best {
  float d;
  int idx;
  uint32_t prio;
};

__device__ __forceinline__ void store6(double *corr, size_t q, const double a[3], const double n[3]) {
  double *o = corr + q * 6;
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = n[0]; o[4] = n[1]; o[5] = n[2];
}

// ---------------------------------------------------------------------------------------------
// One WARP per query.  A query's searches walk hundreds to thousands of candidates (5 m radius on
// a 1 m cell index), far too many for one thread: the lanes of the warp look up the row ranges of a
// shell in parallel, then stride together over every range, each lane keeping its own best under the
// exact (distance, tie-rule) order; a shuffle reduction with the same order merges the lanes after
// every shell, which is also where the termination test of shell_search is evaluated.  The two
// ring-filtered searches of a flat point share one shell walk (a search that is already decided
// cannot change any more: every later candidate is strictly farther).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool better_nn(float d, int idx, const Best &b) { return d < b.d || (d == b.d && idx < b.idx); }
__device__ __forceinline__ bool better_ring(float d, uint32_t pr, const Best &b) {
  return d < b.d || (d == b.d && b.idx >= 0 && pr < b.prio);
}

__device__ __forceinline__ void warp_merge_nn(Best &b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float d = __shfl_xor_sync(0xffffffffu, b.d, o);
    const int idx = __shfl_xor_sync(0xffffffffu, b.idx, o);
    if (idx >= 0 && (b.idx < 0 || better_nn(d, idx, b))) { b.d = d; b.idx = idx; }
  }
}
__device__ __forceinline__ void warp_merge_ring(Best &b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float d = __shfl_xor_sync(0xffffffffu, b.d, o);
    const int idx = __shfl_xor_sync(0xffffffffu, b.idx, o);
    const uint32_t pr = __shfl_xor_sync(0xffffffffu, b.prio, o);
    if (idx >= 0 && (b.idx < 0 || better_ring(d, pr, b))) { b.d = d; b.idx = idx; b.prio = pr; }
  }
}

// PHASE 0: plain 1-NN into b0.  PHASE 1: ring-filtered searches around `closest` (ring `id`): b_other = nearest point
// of another ring within +-nearby rings; b_same (only when WANT_SAME) = nearest other point of the same ring.
template <int PHASE, bool WANT_SAME>
__device__ __forceinline__ void warp_shell_search(const ScanGrid &sg, const uint16_t *__restrict__ ring_sorted, float qx, float qy,
                                                  float qz, float thresh, int closest, int id, double nearby, Best &b0,
                                                  Best &b_other, Best &b_same) {
  const GridView &g = sg.g;
  const uint32_t lane = threadIdx.x & 31;
  if (PHASE == 0) { b0.d = thresh; b0.idx = -1; b0.prio = 0xffffffffu; }
  else {
    b_other.d = thresh; b_other.idx = -1; b_other.prio = 0xffffffffu;
    b_same.d = thresh; b_same.idx = -1; b_same.prio = 0xffffffffu;
  }
  const float edge = 1.0f / g.inv_edge;
  const int cx = (int)floorf(qx * g.inv_edge) - g.ox, cy = (int)floorf(qy * g.inv_edge) - g.oy,
            cz = (int)floorf(qz * g.inv_edge) - g.oz;
  const int smax = (int)ceilf(sqrtf(thresh) * g.inv_edge) + 1;
  for (int s = 0; s <= smax; ++s) {
    const int side = 2 * s + 1, n_rows = side * side;
    for (int r0 = 0; r0 < n_rows; r0 += 32) {
      // lane-parallel lookup: the (up to two) candidate ranges of row r0 + lane of this shell
      const int r = r0 + (int)lane;
      uint32_t js0 = 0, je0 = 0, js1 = 0, je1 = 0;
      if (r < n_rows) {
        const int dz = r / side - s, dy = r % side - s;
        const int z = cz + dz, y = cy + dy;
        if (z >= 0 && z < g.nz && y >= 0 && y < g.ny) {
          const int row = (z * g.ny + y) * g.nx;
          if (dz == -s || dz == s || dy == -s || dy == s) {  // a face row of the shell: the whole x range
            const int x0 = max(cx - s, 0), x1 = min(cx + s, g.nx - 1);
            if (x0 <= x1) { js0 = __ldg(g.cell_start + row + x0); je0 = __ldg(g.cell_start + row + x1 + 1); }
          } else {                                           // interior row: only the two end cells
            const int xa = cx - s, xb = cx + s;
            if (xa >= 0 && xa < g.nx) { js0 = __ldg(g.cell_start + row + xa); je0 = __ldg(g.cell_start + row + xa + 1); }
            if (xb >= 0 && xb < g.nx) { js1 = __ldg(g.cell_start + row + xb); je1 = __ldg(g.cell_start + row + xb + 1); }
          }
        }
      }
      uint32_t live = __ballot_sync(0xffffffffu, je0 > js0 || je1 > js1);
      while (live) {
        const int l = __ffs(live) - 1;
        live &= live - 1;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          const uint32_t js = __shfl_sync(0xffffffffu, part ? js1 : js0, l), je = __shfl_sync(0xffffffffu, part ? je1 : je0, l);
          for (uint32_t j = js + lane; j < je; j += 32) {
            const float4 m = __ldg(g.pts_sorted + j);
            const int idx = __float_as_int(m.w);
            if (PHASE == 0) {
              const float d = sqdist_f(qx, qy, qz, m);
              if (better_nn(d, idx, b0)) { b0.d = d; b0.idx = idx; }
            } else {
              if (idx == closest) continue;
              const int rg = (int)__ldg(ring_sorted + j);
              const bool same = rg == id;
              if (same ? !WANT_SAME : ((double)rg > id + nearby || (double)rg < id - nearby)) continue;
              const float d = sqdist_f(qx, qy, qz, m);
              const uint32_t pr = visit_prio(idx, closest, sg.n);
              Best &b = same ? b_same : b_other;
              if (better_ring(d, pr, b)) { b.d = d; b.idx = idx; b.prio = pr; }
            }
          }
        }
      }
    }
    // merge the lanes; every unvisited point lies outside the (2s+1)^3 block, i.e. at least s*edge away on some
    // axis, and squaring / fp32 rounding are monotone, so its fp32 distance is >= fl((s*edge)^2)
    const float reach = (float)s * edge;
    const float reach2 = __fmul_rn(reach, reach);
    bool done;
    if (PHASE == 0) {
      Best t = b0;
      warp_merge_nn(t);
      done = t.idx >= 0 && t.d < reach2;
    } else {
      Best t = b_other;
      warp_merge_ring(t);
      done = t.idx >= 0 && t.d < reach2;
      if (WANT_SAME) {
        Best u = b_same;
        warp_merge_ring(u);
        done = done && u.idx >= 0 && u.d < reach2;
      }
    }
    if (done || reach2 >= thresh) break;
  }
  if (PHASE == 0) warp_merge_nn(b0);
  else {
    warp_merge_ring(b_other);
    if (WANT_SAME) warp_merge_ring(b_same);
  }
}

// One query (the whole warp): the searches of a sharp / flat point against the last scan's cloud `sg` and the factor
// constants [a(3), n(3)] (n = 0: no factor).  nn / b2 / b3 return the association indices.
__device__ __forceinline__ void associate_scan_query(const ScanGrid &sg, const uint16_t *__restrict__ ring_sorted, const KParams &kp,
                                                     const double pose[7], const float4 p, bool is_sharp, double a[3], double n[3],
                                                     Best &nn, Best &b2, Best &b3) {
  const float3 x = transform_point_f(pose, p.x, p.y, p.z);  // TransformToStart, s = 1 (:21-33)
  a[0] = a[1] = a[2] = 0; n[0] = n[1] = n[2] = 0;
  b2.idx = b3.idx = -1;
  warp_shell_search<0, false>(sg, ring_sorted, x.x, x.y, x.z, kp.dist_sq_thresh_f, -1, 0, 0.0, nn, b2, b3);  // :84-87 / :169-173
  if (nn.idx >= 0) {
    const int closest = nn.idx;
    const int id = (int)__ldg(sg.ring + closest);
    const float4 pa = __ldg(sg.g.pts_orig + closest);
    if (is_sharp) {
      Best unused;
      warp_shell_search<1, false>(sg, ring_sorted, x.x, x.y, x.z, kp.dist_sq_thresh_f, closest, id, kp.nearby_scan, nn, b2,
                                  unused);  // :93-140
      if (b2.idx >= 0) {  // :143-162
        const float4 pb = __ldg(sg.g.pts_orig + b2.idx);
        a[0] = pa.x; a[1] = pa.y; a[2] = pa.z;
        n[0] = (double)pa.x - (double)pb.x; n[1] = (double)pa.y - (double)pb.y; n[2] = (double)pa.z - (double)pb.z;
        const double nn2 = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        if (nn2 > 0) { n[0] /= nn2; n[1] /= nn2; n[2] /= nn2; }
      }
    } else {
      // b2 = nearest in the same ring, b3 = nearest in another ring (one shared shell walk)
      warp_shell_search<1, true>(sg, ring_sorted, x.x, x.y, x.z, kp.dist_sq_thresh_f, closest, id, kp.nearby_scan, nn, b3, b2);
      if (b2.idx >= 0 && b3.idx >= 0) {  // :234-256, LidarPlaneFactorSE3 4-point ctor (lidar_factor.h:70-78)
        const float4 pb = __ldg(sg.g.pts_orig + b2.idx), pc = __ldg(sg.g.pts_orig + b3.idx);
        const double A[3] = {pa.x, pa.y, pa.z}, Bv[3] = {pb.x, pb.y, pb.z}, Cv[3] = {pc.x, pc.y, pc.z};
        const double ab[3] = {A[0] - Bv[0], A[1] - Bv[1], A[2] - Bv[2]}, ac[3] = {A[0] - Cv[0], A[1] - Cv[1], A[2] - Cv[2]};
        n[0] = ab[1] * ac[2] - ab[2] * ac[1];
        n[1] = ab[2] * ac[0] - ab[0] * ac[2];
        n[2] = ab[0] * ac[1] - ab[1] * ac[0];
        const double nn2 = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        if (nn2 > 0) { n[0] /= nn2; n[1] /= nn2; n[2] /= nn2; }
        for (int d = 0; d < 3; ++d) a[d] = (A[d] + Bv[d] + Cv[d]) / 3;
      }
    }
  }
}

__global__ void __launch_bounds__(128)
k_associate_scan(ScanGrid gc, ScanGrid gs, const uint16_t *__restrict__ ring_sorted_c, const uint16_t *__restrict__ ring_sorted_s,
                 KParams kp, const float4 *__restrict__ q_sharp, uint32_t n_sharp, const float4 *__restrict__ q_flat,
                 uint32_t n_flat, const double *__restrict__ pose_g, double *__restrict__ corr, int32_t *__restrict__ assoc) {
  const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per query
  const uint32_t lane = threadIdx.x & 31;
  if (k >= n_sharp + n_flat) return;
  double pose[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) pose[i] = pose_g[i];
  const bool is_sharp = k < n_sharp;
  const float4 p = is_sharp ? q_sharp[k] : q_flat[k - n_sharp];
  double a[3], n[3];
  Best nn, b2, b3;
  associate_scan_query(is_sharp ? gc : gs, is_sharp ? ring_sorted_c : ring_sorted_s, kp, pose, p, is_sharp, a, n, nn, b2, b3);
  if (lane != 0) return;
  store6(corr, k, a, n);
  if (assoc) {
    if (is_sharp) {
      assoc[2 * k] = nn.idx;
      assoc[2 * k + 1] = b2.idx;
    } else {
      int32_t *o = assoc + 2 * n_sharp + 3 * (k - n_sharp);
      o[0] = nn.idx; o[1] = b2.idx; o[2] = b3.idx;
    }
  }
}

// ring of every point in cell order (coalesced beside pts_sorted)
__global__ void k_gather_ring(const float4 *__restrict__ pts_sorted, const uint16_t *__restrict__ ring, uint32_t n,
                              uint16_t *__restrict__ ring_sorted) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) ring_sorted[j] = ring[__float_as_int(pts_sorted[j].w)];
}

}  // namespace msfl

using namespace msfl;

// uploads one cloud (+ rings) into a device buffer as packed float4 / uint16; checks ring order
static int upload_with_rings(msfl_engine *e, const msfl_cloud *c, DevBuf &d_pts, DevBuf &d_ring, size_t stage_off,
                             bool need_ring) {
  const size_t n = c->n;
  int rc;
  if ((rc = d_pts.reserve(n * 16 + 16))) return rc;
  if ((rc = d_ring.reserve(n * 2 + 16))) return rc;
  char *h = e->h_stage.as<char>() + stage_off;
  float *h4 = (float *)h;
  uint16_t *hr = (uint16_t *)(h + n * 16);
  const char *base = (const char *)c->data;
  const bool has_i = c->off_intensity != MSFL_NO_FIELD;
  uint16_t prev = 0;
  for (size_t i = 0; i < n; ++i) {
    const char *pt = base + i * c->stride;
    memcpy(h4 + 4 * i, pt + c->off_xyz, 12);
    float w = 0.f;
    if (has_i) memcpy(&w, pt + c->off_intensity, 4);
    h4[4 * i + 3] = w;
    if (need_ring) {
      uint16_t r;
      memcpy(&r, pt + c->off_ring, 2);
      if (r < prev) { set_error("scan2scan: last-scan cloud is not ring-sorted at point %zu", i); return MSFL_ERR_RING; }
      if (r >= MSFL_MAX_RINGS) { set_error("scan2scan: ring %u >= %d", (unsigned)r, MSFL_MAX_RINGS); return MSFL_ERR_RING; }
      prev = r;
      hr[i] = r;
    }
  }
  MSFL_CUDA_OK(cudaMemcpyAsync(d_pts.p, h4, n * 16, cudaMemcpyHostToDevice, e->stream));
  if (need_ring) MSFL_CUDA_OK(cudaMemcpyAsync(d_ring.p, hr, n * 2, cudaMemcpyHostToDevice, e->stream));
  return MSFL_OK;
}

static int scan2scan_impl(msfl_engine *e, const msfl_cloud *lc, const msfl_cloud *ls, const msfl_cloud *cs,
                          const msfl_cloud *cf, double pose_tq[7], msfl_stats *stats, int32_t *assoc_out, bool assoc_only) {
  const msfl_cloud *cl[4] = {lc, ls, cs, cf};
  static const char *const names[4] = {"scan2scan last_corner_less_sharp", "scan2scan last_surf_less_flat",
                                       "scan2scan curr_corner_sharp", "scan2scan curr_surf_flat"};
  for (int i = 0; i < 4; ++i) {
    if (!cl[i]) { set_error("scan2scan: null cloud"); return MSFL_ERR_ARG; }
    int rck;
    if ((rck = check_cloud(cl[i], /*need_ring=*/i < 2, names[i]))) return rck;
  }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  const uint32_t n_sharp = (uint32_t)cs->n, n_flat = (uint32_t)cf->n, nq = n_sharp + n_flat;
  int rc;
  if (stats) memset(stats, 0, sizeof *stats);
  // Too few points to ever reach min_correspondences, or nothing to search in: same outcome as the
  // reference (return false before the first solve; pose untouched).
  if (lc->n == 0 || ls->n == 0 || nq == 0) {
    if (stats) { stats->status = MSFL_TOO_FEW; stats->n_outer = 1; }
    if (assoc_out) for (size_t i = 0; i < 2 * (size_t)n_sharp + 3 * (size_t)n_flat; ++i) assoc_out[i] = -1;
    return assoc_only ? MSFL_OK : MSFL_TOO_FEW;
  }
  const size_t stage = (lc->n + ls->n) * 18 + (size_t)nq * 16 + 256;
  if ((rc = e->h_stage.reserve(stage))) return rc;
  size_t off = 0;
  if ((rc = upload_with_rings(e, lc, e->d_last_corner, e->d_last_corner_ring, off, true))) return rc;
  off += ((lc->n * 18 + 15) & ~(size_t)15);
  if ((rc = upload_with_rings(e, ls, e->d_last_surf, e->d_last_surf_ring, off, true))) return rc;
  off += ((ls->n * 18 + 15) & ~(size_t)15);
  // queries: [sharp | flat] float4, offsets tables, pose
  const size_t q_bytes = (size_t)nq * 16;
  if ((rc = e->d_queries.reserve(q_bytes + 64 + 64))) return rc;
  {
    char *h = e->h_stage.as<char>() + off;
    float *hq = (float *)h;
    const msfl_cloud *qc[2] = {cs, cf};
    size_t w = 0;
    for (int c = 0; c < 2; ++c) {
      const char *base = (const char *)qc[c]->data;
      for (size_t i = 0; i < qc[c]->n; ++i, ++w) {
        memcpy(hq + 4 * w, base + i * qc[c]->stride + qc[c]->off_xyz, 12);
        hq[4 * w + 3] = 0.f;
      }
    }
    int32_t *hoff = (int32_t *)(h + q_bytes);
    hoff[0] = 0; hoff[1] = (int32_t)n_sharp; hoff[2] = 0; hoff[3] = (int32_t)n_flat;
    double *hp = (double *)(h + q_bytes + 64);
    memcpy(hp, pose_tq, 56);
    MSFL_CUDA_OK(cudaMemcpyAsync(e->d_queries.p, h, q_bytes + 64 + 56, cudaMemcpyHostToDevice, st));
  }
  // cell indices over the last scan's feature clouds (the two kd-tree builds, :57-61)
  Submap &g_corner = e->last_corner_grid, &g_surf = e->last_surf_grid;  // rebuilt every call, like the reference
  if ((rc = submap_build(e, g_corner, e->d_last_corner.as<float4>(), lc->n, 1.0f, false))) return rc;
  if ((rc = submap_build(e, g_surf, e->d_last_surf.as<float4>(), ls->n, 1.0f, false))) return rc;
  ScanGrid gc{g_corner.view, e->d_last_corner_ring.as<uint16_t>(), (uint32_t)lc->n};
  ScanGrid gs{g_surf.view, e->d_last_surf_ring.as<uint16_t>(), (uint32_t)ls->n};
  if ((rc = e->d_ring_tab.reserve((lc->n + ls->n) * 2 + 64))) return rc;
  uint16_t *ring_sorted_c = e->d_ring_tab.as<uint16_t>(), *ring_sorted_s = ring_sorted_c + ((lc->n + 7) & ~(size_t)7);
  k_gather_ring<<<((unsigned)lc->n + 255) / 256, 256, 0, st>>>(g_corner.view.pts_sorted, gc.ring, (uint32_t)lc->n, ring_sorted_c);
  k_gather_ring<<<((unsigned)ls->n + 255) / 256, 256, 0, st>>>(g_surf.view.pts_sorted, gs.ring, (uint32_t)ls->n, ring_sorted_s);
  e->launches += 2;
  char *d = e->d_queries.as<char>();
  const float4 *d_sharp = (const float4 *)d, *d_flat = d_sharp + n_sharp;
  const int32_t *d_e_off = (const int32_t *)(d + q_bytes), *d_p_off = d_e_off + 2;
  double *d_pose = (double *)(d + q_bytes + 64);
  if ((rc = e->d_corr.reserve(((size_t)nq + 1) * 48))) return rc;
  if ((rc = e->d_status.reserve(16))) return rc;
  if ((rc = e->d_assoc.reserve((2 * (size_t)n_sharp + 3 * (size_t)n_flat + 1) * 4))) return rc;
  msfl_stats *d_stats = nullptr;
  if (stats) {
    if ((rc = e->d_stats.reserve(sizeof(msfl_stats)))) return rc;
    d_stats = e->d_stats.as<msfl_stats>();
    MSFL_CUDA_OK(cudaMemsetAsync(d_stats, 0, sizeof(msfl_stats), st));
  }
  const int tb = 128;
  const int n_outer = assoc_only ? 1 : e->params.num_outer;
  for (int outer = 0; outer < n_outer; ++outer) {  // :64
    k_associate_scan<<<(nq * 32 + tb - 1) / tb, tb, 0, st>>>(gc, gs, ring_sorted_c, ring_sorted_s, e->kp, d_sharp, n_sharp, d_flat, n_flat, d_pose,
                                                        e->d_corr.as<double>(),
                                                        (assoc_out && outer == 0) ? e->d_assoc.as<int32_t>() : nullptr);
    e->launches += 1;
    MSFL_CUDA_OK(cudaGetLastError());
    if (assoc_only) break;
    if ((rc = launch_lm_solve(e, 1, d_sharp, d_e_off, n_sharp, d_flat, d_p_off, e->d_corr.as<double>(), d_pose,
                              e->d_status.as<int32_t>(), d_stats, outer, e->params.min_correspondences)))
      return rc;
  }
  int32_t h_status = MSFL_OK;
  if (!assoc_only) {
    MSFL_CUDA_OK(cudaMemcpyAsync(&h_status, e->d_status.p, 4, cudaMemcpyDeviceToHost, st));
    MSFL_CUDA_OK(cudaMemcpyAsync(pose_tq, d_pose, 56, cudaMemcpyDeviceToHost, st));
    if (stats) MSFL_CUDA_OK(cudaMemcpyAsync(stats, d_stats, sizeof(msfl_stats), cudaMemcpyDeviceToHost, st));
  }
  if (assoc_out)
    MSFL_CUDA_OK(cudaMemcpyAsync(assoc_out, e->d_assoc.p, (2 * (size_t)n_sharp + 3 * (size_t)n_flat) * 4, cudaMemcpyDeviceToHost, st));
  MSFL_CUDA_OK(cudaStreamSynchronize(st));
  return h_status;
}

// =============================================================================================================
// Batched form: B independent MatchScan2Scan problems (replay of a log: pair b = scans b, b + 1) in one launch sequence.
// The 2B last-scan clouds (grid g = 2 b + class, class 0 = less-sharp, 1 = less-flat) lie back to back in HBM; every
// grid gets its own dense 1 m cell table inside ONE cell array (header table: origin, dims, first cell, first point),
// built for all grids at once by a counting sort: one atomic per point (rank inside its cell), ONE exclusive scan over
// the cells of all grids, one scatter.  The search then runs one warp per query of the whole batch and the LM kernel
// one CTA (or cluster) per pair, exactly as in the batched scan-to-map.  Pairs are processed in chunks whose cell
// tables fit the engine's pair_cell_budget (2^27 cells = 0.5 GB).
// =============================================================================================================
namespace msfl {

struct PairGridHdr {
  int ox, oy, oz, nx, ny, nz;
  uint32_t cell_base;  // first entry of this grid in the chunk's cell array (ncell + 1 entries)
  uint32_t pt_base;    // first point of this grid in the concatenated last-scan arrays
  uint32_t n, chunk_pt_base;  // points in this grid; first point of the grid's chunk
};

__global__ void k_pair_bounds_init(int *bounds, uint32_t G) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * 8) return;
  const uint32_t f = i & 7u;
  bounds[i] = f < 3 ? INT_MAX : (f < 6 ? INT_MIN : 0);
}

// occupied cell range of every grid (blockIdx.y = grid); bounds[g] = {lo[3], hi[3], bad, -}
__global__ void k_pair_bounds(const float4 *__restrict__ pts, const uint32_t *__restrict__ goff, float inv_edge, int *bounds) {
  const uint32_t g = blockIdx.y, base = goff[g], n = goff[g + 1] - base;
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  int bad = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[base + i];
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) || fabsf(p.x) > 1e6f || fabsf(p.y) > 1e6f || fabsf(p.z) > 1e6f) {
      bad = 1;
      continue;
    }
    const int c[3] = {(int)floorf(p.x * inv_edge), (int)floorf(p.y * inv_edge), (int)floorf(p.z * inv_edge)};
#pragma unroll
    for (int d = 0; d < 3; ++d) { lo[d] = min(lo[d], c[d]); hi[d] = max(hi[d], c[d]); }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d)
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  bad = __any_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    int *b = bounds + 8 * g;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (lo[d] != INT_MAX) atomicMin(&b[d], lo[d]);
      if (hi[d] != INT_MIN) atomicMax(&b[3 + d], hi[d]);
    }
    if (bad) atomicOr(&b[6], 1);
  }
}

// counting sort, pass 1 (blockIdx.y = grid of the chunk): cell key + rank inside the cell
__global__ void k_pair_cell_count(const float4 *__restrict__ pts, const PairGridHdr *__restrict__ hdr, uint32_t g0, float inv_edge,
                                  uint32_t *__restrict__ cells, uint32_t *__restrict__ keys, uint32_t *__restrict__ rank) {
  const PairGridHdr h = hdr[g0 + blockIdx.y];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h.n) return;
  const float4 p = pts[h.pt_base + i];
  const int cx = (int)floorf(p.x * inv_edge) - h.ox, cy = (int)floorf(p.y * inv_edge) - h.oy, cz = (int)floorf(p.z * inv_edge) - h.oz;
  const uint32_t key = h.cell_base + (uint32_t)((cz * h.ny + cy) * h.nx + cx);
  keys[h.pt_base + i] = key;
  rank[h.pt_base + i] = atomicAdd(cells + key, 1u);
}

// pass 2: the point (index inside its own cloud in .w) and its ring go to cell order
__global__ void k_pair_cell_scatter(const float4 *__restrict__ pts, const uint16_t *__restrict__ ring,
                                    const PairGridHdr *__restrict__ hdr, uint32_t g0, const uint32_t *__restrict__ cells,
                                    const uint32_t *__restrict__ keys, const uint32_t *__restrict__ rank,
                                    float4 *__restrict__ sorted, uint16_t *__restrict__ ring_sorted) {
  const PairGridHdr h = hdr[g0 + blockIdx.y];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h.n) return;
  const uint32_t src = h.pt_base + i;
  float4 p = pts[src];
  p.w = __int_as_float((int)i);
  const uint32_t dst = h.chunk_pt_base + __ldg(cells + __ldg(keys + src)) + __ldg(rank + src);
  sorted[dst] = p;
  ring_sorted[dst] = ring[src];
}

struct PairArrays {
  const float4 *pts;         // last-scan clouds, caller order
  const uint16_t *ring;
  const float4 *sorted;      // cell order (chunk-relative positions in `cells`)
  const uint16_t *ring_sorted;
  const uint32_t *cells;
  const PairGridHdr *hdr;
  float inv_edge;
};

// one warp per query of pairs [b0, b1): the chunk's sharp queries first, then its flat queries
__global__ void __launch_bounds__(128)
k_associate_scan_batch(PairArrays pa, KParams kp, int B, int b0, int b1, const float4 *__restrict__ q_sharp,
                       const int32_t *__restrict__ e_off, uint32_t n_sharp_total, const float4 *__restrict__ q_flat,
                       const int32_t *__restrict__ p_off, const double *__restrict__ poses, const int32_t *__restrict__ status,
                       int outer, double *__restrict__ corr) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t s0 = (uint32_t)e_off[b0], ns = (uint32_t)e_off[b1] - s0, f0 = (uint32_t)p_off[b0], nf = (uint32_t)p_off[b1] - f0;
  if (w >= ns + nf) return;
  const bool is_sharp = w < ns;
  const uint32_t kk = is_sharp ? s0 + w : f0 + (w - ns);  // index inside the class
  int lo = b0, hi = b1;                                   // pair: largest b with off[b] <= kk
  const int32_t *off = is_sharp ? e_off : p_off;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((uint32_t)__ldg(off + mid) <= kk) lo = mid;
    else hi = mid;
  }
  const int b = lo;
  if (outer > 0 && status[b] != MSFL_OK) return;  // this pair bailed out in an earlier outer iteration (:262-267)
  const PairGridHdr h = pa.hdr[2 * b + (is_sharp ? 0 : 1)];
  ScanGrid sg;
  sg.g.pts_sorted = pa.sorted + h.chunk_pt_base;
  sg.g.pts_orig = pa.pts + h.pt_base;
  sg.g.cell_start = pa.cells + h.cell_base;
  sg.g.row_mask = nullptr;
  sg.g.nx = h.nx; sg.g.ny = h.ny; sg.g.nz = h.nz;
  sg.g.ox = h.ox; sg.g.oy = h.oy; sg.g.oz = h.oz;
  sg.g.inv_edge = pa.inv_edge;
  sg.g.n = h.n;
  sg.ring = pa.ring + h.pt_base;
  sg.n = h.n;
  double pose[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) pose[i] = poses[(size_t)b * 7 + i];
  const float4 p = is_sharp ? q_sharp[kk] : q_flat[kk];
  double a[3], n[3];
  Best nn, b2, b3;
  associate_scan_query(sg, pa.ring_sorted + h.chunk_pt_base, kp, pose, p, is_sharp, a, n, nn, b2, b3);
  if (lane == 0) store6(corr, is_sharp ? kk : n_sharp_total + kk, a, n);
}

// Inputs in HBM: last_pts / last_ring = the 2B last-scan clouds back to back (h_goff[2B + 1] host copy of d_goff);
// q_sharp / q_flat = the queries of all pairs with offset tables d_e_off / d_p_off [B + 1] (host copies h_*); d_poses 7B
// in-out; d_status B; d_stats B or null.  One host synchronisation (the grid bounds).
int scan2scan_batch_device(msfl_engine *e, int B, const float4 *last_pts, const uint16_t *last_ring, const uint32_t *d_goff,
                           const uint32_t *h_goff, const float4 *q_sharp, const int32_t *d_e_off, const int32_t *h_e_off,
                           const float4 *q_flat, const int32_t *d_p_off, const int32_t *h_p_off, double *d_poses,
                           int32_t *d_status, msfl_stats *d_stats) {
  cudaStream_t st = e->stream;
  const uint32_t G = 2u * (uint32_t)B;
  const float inv_edge = 1.0f;
  const size_t n_last = h_goff[G];
  const uint32_t n_sharp_total = (uint32_t)h_e_off[B], n_flat_total = (uint32_t)h_p_off[B];
  int rc;
  uint32_t max_n = 0;
  for (uint32_t g = 0; g < G; ++g) max_n = std::max(max_n, h_goff[g + 1] - h_goff[g]);
  // 1. occupied cell range of every grid
  if ((rc = e->ob_bounds.reserve((size_t)G * 32))) return rc;
  int *d_bounds = e->ob_bounds.as<int>();
  k_pair_bounds_init<<<(G * 8 + 255) / 256, 256, 0, st>>>(d_bounds, G);
  if (max_n > 0) k_pair_bounds<<<dim3(std::min((max_n + 255u) / 256u, 32u), G), 256, 0, st>>>(last_pts, d_goff, inv_edge, d_bounds);
  e->launches += 2;
  std::vector<int> hb((size_t)G * 8);
  MSFL_CUDA_OK(cudaMemcpyAsync(hb.data(), d_bounds, (size_t)G * 32, cudaMemcpyDeviceToHost, st));
  MSFL_CUDA_OK(cudaStreamSynchronize(st));
  // 2. headers and chunks of pairs
  std::vector<PairGridHdr> hdr(G);
  std::vector<int> chunk_first;  // first pair of every chunk (+ B at the end)
  std::vector<long long> chunk_cells;
  long long cells_here = 0, max_chunk_cells = 0;
  for (int b = 0; b < B; ++b) {
    long long pair_cells = 0;
    for (uint32_t g = 2u * b; g < 2u * b + 2; ++g) {
      PairGridHdr &h = hdr[g];
      const int *bb = &hb[(size_t)g * 8];
      h.pt_base = h_goff[g];
      h.n = h_goff[g + 1] - h_goff[g];
      if (bb[6]) { set_error("scan2scan batch: pair %d holds non-finite or out-of-range (>1e6 m) points", b); return MSFL_ERR_ARG; }
      if (h.n == 0) { h.ox = h.oy = h.oz = 0; h.nx = h.ny = h.nz = 1; }
      else {
        const long long nx = (long long)bb[3] - bb[0] + 5, ny = (long long)bb[4] - bb[1] + 5, nz = (long long)bb[5] - bb[2] + 5;
        if (nx * ny * nz > (1ll << 26)) {
          set_error("scan2scan batch: pair %d needs %lld cells (> 2^26); dense cell index refused", b, nx * ny * nz);
          return MSFL_ERR_GRID;
        }
        h.ox = bb[0] - 2; h.oy = bb[1] - 2; h.oz = bb[2] - 2;
        h.nx = (int)nx; h.ny = (int)ny; h.nz = (int)nz;
      }
      pair_cells += (long long)h.nx * h.ny * h.nz + 1;
    }
    if (chunk_first.empty() || cells_here + pair_cells > e->pair_cell_budget) {
      if (!chunk_first.empty()) chunk_cells.push_back(cells_here);
      chunk_first.push_back(b);
      cells_here = 0;
    }
    for (uint32_t g = 2u * b; g < 2u * b + 2; ++g) {
      hdr[g].cell_base = (uint32_t)cells_here;
      hdr[g].chunk_pt_base = h_goff[2 * chunk_first.back()];
      cells_here += (long long)hdr[g].nx * hdr[g].ny * hdr[g].nz + 1;
    }
  }
  chunk_cells.push_back(cells_here);
  chunk_first.push_back(B);
  for (long long c : chunk_cells) max_chunk_cells = std::max(max_chunk_cells, c);
  // 3. scratch
  if ((rc = e->ob_hdr.reserve((size_t)G * sizeof(PairGridHdr)))) return rc;
  if ((rc = e->ob_sorted.reserve(n_last * 16 + 16))) return rc;
  if ((rc = e->ob_ring_sorted.reserve(n_last * 2 + 16))) return rc;
  if ((rc = e->ob_keys.reserve(n_last * 4 + 16))) return rc;
  if ((rc = e->ob_rank.reserve(n_last * 4 + 16))) return rc;
  if ((rc = e->ob_cells.reserve((size_t)max_chunk_cells * 4 + 16))) return rc;
  if ((rc = e->d_corr.reserve(((size_t)n_sharp_total + n_flat_total + 1) * 48))) return rc;
  size_t tmp_bytes = 0;
  uint32_t *cells = e->ob_cells.as<uint32_t>();
  MSFL_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cells, cells, (int)max_chunk_cells, st));
  if ((rc = e->ob_tmp.reserve(tmp_bytes))) return rc;
  MSFL_CUDA_OK(cudaMemcpyAsync(e->ob_hdr.p, hdr.data(), (size_t)G * sizeof(PairGridHdr), cudaMemcpyHostToDevice, st));
  PairArrays pa{last_pts, last_ring, e->ob_sorted.as<float4>(), e->ob_ring_sorted.as<uint16_t>(), cells,
                e->ob_hdr.as<PairGridHdr>(), inv_edge};
  // 4. per chunk: index build, then the reference's outer loop (:64) for every pair of the chunk at once
  for (size_t c = 0; c + 1 < chunk_first.size(); ++c) {
    const int b0 = chunk_first[c], b1 = chunk_first[c + 1];
    const uint32_t g0 = 2u * b0, ng = 2u * (uint32_t)(b1 - b0);
    uint32_t cmax = 0;
    for (uint32_t g = g0; g < g0 + ng; ++g) cmax = std::max(cmax, hdr[g].n);
    const long long ncell = chunk_cells[c];
    MSFL_CUDA_OK(cudaMemsetAsync(cells, 0, (size_t)ncell * 4, st));
    if (cmax > 0) {
      k_pair_cell_count<<<dim3((cmax + 255) / 256, ng), 256, 0, st>>>(last_pts, pa.hdr, g0, inv_edge, cells, e->ob_keys.as<uint32_t>(),
                                                                     e->ob_rank.as<uint32_t>());
      MSFL_CUDA_OK(cub::DeviceScan::ExclusiveSum(e->ob_tmp.p, tmp_bytes, cells, cells, (int)ncell, st));
      k_pair_cell_scatter<<<dim3((cmax + 255) / 256, ng), 256, 0, st>>>(last_pts, last_ring, pa.hdr, g0, cells, e->ob_keys.as<uint32_t>(),
                                                                       e->ob_rank.as<uint32_t>(), e->ob_sorted.as<float4>(),
                                                                       e->ob_ring_sorted.as<uint16_t>());
      e->launches += 2 + 2;
    }
    const uint32_t nq = (uint32_t)(h_e_off[b1] - h_e_off[b0]) + (uint32_t)(h_p_off[b1] - h_p_off[b0]);
    for (int outer = 0; outer < e->params.num_outer; ++outer) {
      if (nq > 0) {
        k_associate_scan_batch<<<(unsigned)(((size_t)nq * 32 + 127) / 128), 128, 0, st>>>(
            pa, e->kp, B, b0, b1, q_sharp, d_e_off, n_sharp_total, q_flat, d_p_off, d_poses, d_status, outer, e->d_corr.as<double>());
        e->launches += 1;
      }
      MSFL_CUDA_OK(cudaGetLastError());
      if ((rc = launch_lm_solve(e, b1 - b0, q_sharp, d_e_off + b0, n_sharp_total, q_flat, d_p_off + b0, e->d_corr.as<double>(),
                                d_poses + (size_t)7 * b0, d_status + b0, d_stats ? d_stats + b0 : nullptr, outer,
                                e->params.min_correspondences)))
        return rc;
    }
  }
  return MSFL_OK;
}

}  // namespace msfl

extern "C" int msfl_scan2scan_batch(msfl_engine *e, int B, const msfl_cloud *last_corner_less_sharp,
                                    const msfl_cloud *last_surf_less_flat, const msfl_cloud *curr_corner_sharp,
                                    const msfl_cloud *curr_surf_flat, double *poses_tq, int32_t *status, msfl_stats *stats) {
  if (!e || B <= 0 || !last_corner_less_sharp || !last_surf_less_flat || !curr_corner_sharp || !curr_surf_flat || !poses_tq) {
    set_error("msfl_scan2scan_batch: bad argument");
    return MSFL_ERR_ARG;
  }
  int rc;
  size_t n_last = 0, n_sharp = 0, n_flat = 0;
  for (int b = 0; b < B; ++b) {
    if ((rc = check_cloud(&last_corner_less_sharp[b], true, "scan2scan_batch last_corner_less_sharp"))) return rc;
    if ((rc = check_cloud(&last_surf_less_flat[b], true, "scan2scan_batch last_surf_less_flat"))) return rc;
    if ((rc = check_cloud(&curr_corner_sharp[b], false, "scan2scan_batch curr_corner_sharp"))) return rc;
    if ((rc = check_cloud(&curr_surf_flat[b], false, "scan2scan_batch curr_surf_flat"))) return rc;
    n_last += last_corner_less_sharp[b].n + last_surf_less_flat[b].n;
    n_sharp += curr_corner_sharp[b].n;
    n_flat += curr_surf_flat[b].n;
  }
  if (n_last > 0x7fffffffull || n_sharp + n_flat > 0x7fffffffull) { set_error("msfl_scan2scan_batch: batch too large"); return MSFL_ERR_ARG; }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  // pinned staging: [last pts float4 | sharp float4 | flat float4 | rings u16 | goff | e_off | p_off | poses]
  const size_t nq = n_sharp + n_flat;
  auto al16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  const size_t at_ring = (n_last + nq) * 16, at_goff = al16(at_ring + n_last * 2);
  const size_t at_eoff = at_goff + al16(((size_t)2 * B + 1) * 4), at_poff = at_eoff + al16(((size_t)B + 1) * 4);
  const size_t at_pose = at_poff + al16(((size_t)B + 1) * 4), total = at_pose + (size_t)B * 56;
  if ((rc = e->h_stage.reserve(total))) return rc;
  if ((rc = e->ob_in.reserve(total))) return rc;
  char *h = e->h_stage.as<char>();
  float *hl = (float *)h, *hs = hl + 4 * n_last, *hf = hs + 4 * n_sharp;
  uint16_t *hr = (uint16_t *)(h + at_ring);
  uint32_t *goff = (uint32_t *)(h + at_goff);
  int32_t *e_off = (int32_t *)(h + at_eoff), *p_off = (int32_t *)(h + at_poff);
  size_t wl = 0, ws = 0, wf = 0;
  for (int b = 0; b < B; ++b) {
    const msfl_cloud *lc[2] = {&last_corner_less_sharp[b], &last_surf_less_flat[b]};
    for (int c = 0; c < 2; ++c) {
      goff[2 * b + c] = (uint32_t)wl;
      const char *base = (const char *)lc[c]->data;
      const bool has_i = lc[c]->off_intensity != MSFL_NO_FIELD;
      uint16_t prev = 0;
      for (size_t i = 0; i < lc[c]->n; ++i, ++wl) {
        const char *pt = base + i * lc[c]->stride;
        memcpy(hl + 4 * wl, pt + lc[c]->off_xyz, 12);
        float w = 0.f;
        if (has_i) memcpy(&w, pt + lc[c]->off_intensity, 4);
        hl[4 * wl + 3] = w;
        uint16_t r;
        memcpy(&r, pt + lc[c]->off_ring, 2);
        if (r < prev) { set_error("scan2scan batch: last-scan cloud of pair %d is not ring-sorted at point %zu", b, i); return MSFL_ERR_RING; }
        if (r >= MSFL_MAX_RINGS) { set_error("scan2scan batch: ring %u >= %d", (unsigned)r, MSFL_MAX_RINGS); return MSFL_ERR_RING; }
        prev = r;
        hr[wl] = r;
      }
    }
    e_off[b] = (int32_t)ws;
    p_off[b] = (int32_t)wf;
    const msfl_cloud *qc[2] = {&curr_corner_sharp[b], &curr_surf_flat[b]};
    // nothing to search in (or with): the pair keeps no queries and comes back MSFL_TOO_FEW with its pose untouched,
    // like the single call
    const bool dead = lc[0]->n == 0 || lc[1]->n == 0 || qc[0]->n + qc[1]->n == 0;
    for (int c = 0; c < 2 && !dead; ++c) {
      const char *base = (const char *)qc[c]->data;
      float *dst = c ? hf : hs;
      size_t &w = c ? wf : ws;
      for (size_t i = 0; i < qc[c]->n; ++i, ++w) {
        memcpy(dst + 4 * w, base + i * qc[c]->stride + qc[c]->off_xyz, 12);
        dst[4 * w + 3] = 0.f;
      }
    }
  }
  goff[2 * B] = (uint32_t)wl;
  e_off[B] = (int32_t)ws;
  p_off[B] = (int32_t)wf;
  memcpy(h + at_pose, poses_tq, (size_t)B * 56);
  char *d = e->ob_in.as<char>();
  MSFL_CUDA_OK(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, st));
  if ((rc = e->d_status.reserve((size_t)B * 4 + 16))) return rc;
  MSFL_CUDA_OK(cudaMemsetAsync(e->d_status.p, 0, (size_t)B * 4, st));
  msfl_stats *d_stats = nullptr;
  if (stats) {
    if ((rc = e->d_stats.reserve((size_t)B * sizeof(msfl_stats)))) return rc;
    d_stats = e->d_stats.as<msfl_stats>();
    MSFL_CUDA_OK(cudaMemsetAsync(d_stats, 0, (size_t)B * sizeof(msfl_stats), st));
  }
  const float4 *d_last = (const float4 *)d, *d_sharp = d_last + n_last, *d_flat = d_sharp + n_sharp;
  double *d_poses = (double *)(d + at_pose);
  // the host copies of the offset tables stay valid: h_stage is not touched again before the final synchronisation
  if ((rc = scan2scan_batch_device(e, B, d_last, (const uint16_t *)(d + at_ring), (const uint32_t *)(d + at_goff), goff, d_sharp,
                                   (const int32_t *)(d + at_eoff), e_off, d_flat, (const int32_t *)(d + at_poff), p_off, d_poses,
                                   e->d_status.as<int32_t>(), d_stats)))
    return rc;
  if ((rc = e->h_poses.reserve((size_t)B * 60))) return rc;
  char *ho = e->h_poses.as<char>();
  MSFL_CUDA_OK(cudaMemcpyAsync(ho, d_poses, (size_t)B * 56, cudaMemcpyDeviceToHost, st));
  MSFL_CUDA_OK(cudaMemcpyAsync(ho + (size_t)B * 56, e->d_status.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  if (stats) {
    if ((rc = e->h_stats.reserve((size_t)B * sizeof(msfl_stats)))) return rc;
    MSFL_CUDA_OK(cudaMemcpyAsync(e->h_stats.p, d_stats, (size_t)B * sizeof(msfl_stats), cudaMemcpyDeviceToHost, st));
  }
  MSFL_CUDA_OK(cudaStreamSynchronize(st));
  memcpy(poses_tq, ho, (size_t)B * 56);
  if (status) memcpy(status, ho + (size_t)B * 56, (size_t)B * 4);
  if (stats) memcpy(stats, e->h_stats.p, (size_t)B * sizeof(msfl_stats));
  return MSFL_OK;
}

extern "C" int msfl_scan2scan(msfl_engine *e, const msfl_cloud *last_corner_less_sharp, const msfl_cloud *last_surf_less_flat,
                              const msfl_cloud *curr_corner_sharp, const msfl_cloud *curr_surf_flat, double pose_tq[7],
                              msfl_stats *stats) {
  if (!e || !pose_tq) { set_error("msfl_scan2scan: bad argument"); return MSFL_ERR_ARG; }
  return scan2scan_impl(e, last_corner_less_sharp, last_surf_less_flat, curr_corner_sharp, curr_surf_flat, pose_tq, stats,
                        nullptr, false);
}

extern "C" int msfl_associate_scan(msfl_engine *e, const msfl_cloud *last_corner_less_sharp,
                                   const msfl_cloud *last_surf_less_flat, const msfl_cloud *curr_corner_sharp,
                                   const msfl_cloud *curr_surf_flat, const double pose_tq[7], int32_t *assoc) {
  if (!e || !pose_tq || !assoc) { set_error("msfl_associate_scan: bad argument"); return MSFL_ERR_ARG; }
  double pose[7];
  memcpy(pose, pose_tq, sizeof pose);
  return scan2scan_impl(e, last_corner_less_sharp, last_surf_less_flat, curr_corner_sharp, curr_surf_flat, pose, nullptr,
                        assoc, true);
}
