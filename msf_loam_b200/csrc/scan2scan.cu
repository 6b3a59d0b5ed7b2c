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
  const float3 x = transform_point_f(pose, p.x, p.y, p.z);  // TransformToStart, s = 1 (:21-33)
  const ScanGrid &sg = is_sharp ? gc : gs;
  const uint16_t *ring_sorted = is_sharp ? ring_sorted_c : ring_sorted_s;
  double a[3] = {0, 0, 0}, n[3] = {0, 0, 0};
  Best nn, b2, b3;
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
