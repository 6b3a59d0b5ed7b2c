// lm_solve.cu -- the Gauss-Newton / Levenberg-Marquardt solve on SE(3):
//   LidarEdgeFactorSE3 / LidarPlaneFactorSE3 residual + analytic Jacobian (lidar_factor.cc:7-44),
//   Huber(0.1) corrector, warp-shuffle + fixed-order block reduction of the 6x6 J^T J, J^T r and
//   cost, and a Ceres-compatible trust-region loop (replacing ceres::Solve at
//   odometry_scan_matcher.cc:270-280 and mapping_scan_matcher.cc:250-272) -- all inside ONE kernel
//   per outer iteration, one thread-block cluster per scan, no host round trip between attempts.
//
// Every attempt is ONE sweep over the scan's correspondences: cost, H and g are evaluated together
// at the candidate x+, so an accepted step needs no second pass (the (1+L) factor of the byte
// formula in SURVEY.md 8d).  The sweep streams the scan's points and factor constants from HBM with
// warp-private, double-buffered TMA bulk copies (cp.async.bulk + mbarrier), see sweep_warp.  All
// factor math is fp64 (the reference's is); sums are combined in a fixed order, so results are
// bit-reproducible run to run and independent of the position in the batch (and of its size within a CTA-shape class,
// see kLmThreadsSmall).
#include <cooperative_groups.h>

#include "lm_device.cuh"
#include <algorithm>

#include "msfl_internal.h"
#include "msfl_math.cuh"

namespace cg = cooperative_groups;

namespace msfl {

// Two CTA shapes, picked by launch size (launch_lm_solve_t).  96 threads = 3 warps per CTA, 5 CTAs per SM (128 registers):
// 15 warps per SM like 4 x 4, but 740 resident scans instead of 592 (2048 scans = 2.8 waves instead of 3.5) and one warp
// fewer at every CTA barrier -- measured on B200 (VLP-16 x 2048): 64 / 96 / 128 / 160 threads -> 0.684 / 0.666 / 0.698 /
// 0.725 ms per launch.  A launch that does not fill those 740 slots (HDL-64 x 512, OS1-128 x 64 with clusters of 4, any
// single scan) gains nothing from a fifth CTA per SM and loses a quarter of its warps: it keeps 128 threads (measured:
// HDL-64 x 512 222.9 k vs 230.9 k scans/s, OS1-128 x 64 28.4 k vs 31.3 k).  The summation order depends on the warp count,
// so a scan's pose is bit-identical across launches of the same shape class, and within 1e-14 m across the two.
constexpr int kLmThreadsBig = 128;  // also the test hook's block size
constexpr int kLmThreadsSmall = 96;
constexpr int kLmThreads = kLmThreadsBig;
using LmShared = LmSharedT<kLmThreadsBig / 32>;

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk, 1-D) staging of correspondence tiles.  The per-scan arrays are contiguous, so a
// tile is two bulk copies (points, constants) that land in shared memory and complete on an mbarrier.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

struct TileSrc {
  const unsigned char *pe, *pp;  // point arrays (PB bytes per entry)
  const double *ce, *cp;         // constants: 48 B per edge entry, PC bytes per plane entry
  uint32_t n_e, n_p;
};

// ---------------------------------------------------------------------------------------------
// Warp-private streaming (the throughput configuration).  Every warp owns a double-buffered ring of
// tiles and its own mbarriers, issues its own bulk copies (lane 0) and consumes them after a
// __syncwarp: there is no block-wide barrier inside a sweep, so a warp never waits for the slowest
// warp of its CTA.  Warp w of W streams tiles w, w + W, w + 2W, ... of each class (edge entries,
// then plane entries), i.e. the CTA as a whole still walks the scan's arrays front to back.  While a
// warp works on the last tile of a sweep it already fetches the first tile of the next sweep (the
// data does not depend on the pose), so the copy latency also hides behind the block reduction and
// the thread-0 LM step.  All shared-memory traffic uses 32-bit shared-space addresses (ld.shared /
// mbarrier on precomputed offsets): the per-tile bookkeeping is ~40 instructions.
//
// PC = bytes of plane constants per entry: 32 = {n, n.c} written by k_fit for the batch path,
// 48 = {c, n} (odometry, deskew, test hooks).  A stage holds TE edge entries or TP plane entries.
// ---------------------------------------------------------------------------------------------
#ifndef MSFL_LM_TILE_BYTES
#define MSFL_LM_TILE_BYTES 6144u
#endif
#ifndef MSFL_LM_STAGES
#define MSFL_LM_STAGES 2
#endif
constexpr int kWarpStages = MSFL_LM_STAGES;  // streaming: double buffer per warp
constexpr int kMaxWarpStages = 9;  // resident configuration (small launches): a warp's tiles stay in smem across sweeps
template <int PB, int PC>
struct WarpTile {
  static constexpr uint32_t SB = PB == 16 ? MSFL_LM_TILE_BYTES : 10240u;  // bytes per stage
  static constexpr uint32_t TE = SB / (PB + 48), TP = SB / (PB + PC);
  static_assert((TE * PB) % 16 == 0 && (TP * PB) % 16 == 0, "constant arrays must stay 16 B aligned");
};

__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ double2 lds_d2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d_s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// per-warp pipeline state, kept in registers across the sweeps of one solve
struct WarpPipe {
  uint32_t ring;     // shared-space address of this warp's ring (n_stages x SB bytes)
  uint32_t bars;     // shared-space address of this warp's mbarriers (8 B each)
  uint32_t it;       // streaming: tiles consumed so far (stage = it & 1, phase parity = (it >> 1) & 1)
  uint32_t my_tiles; // tiles of this warp per sweep
  bool resident;     // all of this warp's tiles fit its ring: loaded once, re-read by every sweep
  bool loaded;       // resident: tiles are in smem / streaming: the first tile of the next sweep is in flight
};

// class 0 = edge entries, 1 = plane entries, 2 = end of the sweep
struct TileIt { uint32_t cls, base; };

template <int PB, int PC>
__device__ __forceinline__ TileIt tile_first(const TileSrc &ts, uint32_t warp) {
  TileIt t{0u, warp * WarpTile<PB, PC>::TE};
  if (t.base >= ts.n_e) {
    t.cls = 1u;
    t.base = warp * WarpTile<PB, PC>::TP;
    if (t.base >= ts.n_p) t.cls = 2u;
  }
  return t;
}
template <int PB, int PC>
__device__ __forceinline__ TileIt tile_next(const TileSrc &ts, uint32_t warp, TileIt t) {
  if (t.cls == 0u) {
    t.base += (blockDim.x >> 5) * WarpTile<PB, PC>::TE;
    if (t.base >= ts.n_e) {
      t.cls = 1u;
      t.base = warp * WarpTile<PB, PC>::TP;
      if (t.base >= ts.n_p) t.cls = 2u;
    }
  } else {
    t.base += (blockDim.x >> 5) * WarpTile<PB, PC>::TP;
    if (t.base >= ts.n_p) t.cls = 2u;
  }
  return t;
}

// lane 0: two bulk copies (points, constants) of tile t into the stage at shared address dst
template <int PB, int PC>
__device__ __forceinline__ void issue_warp_tile(const TileSrc &ts, TileIt t, uint32_t dst, uint32_t bar) {
  using WT = WarpTile<PB, PC>;
  if (t.cls == 0u) {
    const uint32_t cnt = min(WT::TE, ts.n_e - t.base);
    mbar_expect_tx_s(bar, cnt * (uint32_t)(PB + 48));
    tma_load_1d_s(dst, ts.pe + (size_t)t.base * PB, cnt * (uint32_t)PB, bar);
    tma_load_1d_s(dst + WT::TE * PB, (const unsigned char *)ts.ce + (size_t)t.base * 48, cnt * 48u, bar);
  } else {
    const uint32_t cnt = min(WT::TP, ts.n_p - t.base);
    mbar_expect_tx_s(bar, cnt * (uint32_t)(PB + PC));
    tma_load_1d_s(dst, ts.pp + (size_t)t.base * PB, cnt * (uint32_t)PB, bar);
    tma_load_1d_s(dst + WT::TP * PB, (const unsigned char *)ts.cp + (size_t)t.base * PC, cnt * (uint32_t)PC, bar);
  }
}

struct SweepCtx {
  double R[9], t0, t1, t2;
  Huber huber_a;
};

template <int PB>
__device__ __forceinline__ void load_point_s(uint32_t a, double &p0, double &p1, double &p2) {
  if (PB == 16) {
    const float4 pf = lds_f4(a);
    p0 = pf.x; p1 = pf.y; p2 = pf.z;
  } else {
    const double2 pa = lds_d2(a), pb = lds_d2(a + 16);
    p0 = pa.x; p1 = pa.y; p2 = pb.x;
  }
}

// all entries of one tile that sits in shared memory at `buf`
template <int PB, int PC>
__device__ __forceinline__ void consume_tile(double (&acc)[kAcc], const SweepCtx &cx, const TileSrc &ts, TileIt t, uint32_t buf,
                                             uint32_t lane, int &cnt_edge, int &cnt_plane) {
  using WT = WarpTile<PB, PC>;
  if (t.cls == 0u) {
    const uint32_t cnt = min(WT::TE, ts.n_e - t.base);
    uint32_t pa = buf + lane * PB, ca = buf + WT::TE * PB + lane * 48;
#pragma unroll 1
    for (uint32_t ent = lane; ent < cnt; ent += 32, pa += 32 * PB, ca += 32 * 48) {
      const double2 c1 = lds_d2(ca + 16), c2 = lds_d2(ca + 32);
      const double n0 = c1.y, n1 = c2.x, n2 = c2.y;
      if (!(n0 == 0.0 && n1 == 0.0 && n2 == 0.0)) {
        const double2 c0 = lds_d2(ca);
        double p0, p1, p2;
        load_point_s<PB>(pa, p0, p1, p2);
        ++cnt_edge;
        eval_edge(acc, cx.R, cx.t0, cx.t1, cx.t2, p0, p1, p2, c0.x, c0.y, c1.x, n0, n1, n2, cx.huber_a);
      }
    }
  } else {
    const uint32_t cnt = min(WT::TP, ts.n_p - t.base);
    uint32_t pa = buf + lane * PB, ca = buf + WT::TP * PB + lane * PC;
    // 32 B entries: lanes 4..7 (mod 8) fetch the upper half first so that a quarter-warp's eight 16 B reads
    // cover all 32 banks
    const uint32_t hi = PC == 32 ? ((lane >> 2) & 1u) * 16u : 0u;
#pragma unroll 1
    for (uint32_t ent = lane; ent < cnt; ent += 32, pa += 32 * PB, ca += 32 * PC) {
      double n0, n1, n2, nc;
      if (PC == 32) {
        const double2 f = lds_d2(ca + hi), g2 = lds_d2(ca + (hi ^ 16u));
        const double2 c0 = hi ? g2 : f, c1 = hi ? f : g2;
        n0 = c0.x; n1 = c0.y; n2 = c1.x; nc = c1.y;
      } else {
        const double2 c0 = lds_d2(ca), c1 = lds_d2(ca + 16), c2 = lds_d2(ca + 32);
        n0 = c1.y; n1 = c2.x; n2 = c2.y;
        nc = n0 * c0.x + n1 * c0.y + n2 * c1.x;
      }
      if (!(n0 == 0.0 && n1 == 0.0 && n2 == 0.0)) {
        double p0, p1, p2;
        load_point_s<PB>(pa, p0, p1, p2);
        ++cnt_plane;
        eval_plane(acc, cx.R, cx.t0, cx.t1, cx.t2, p0, p1, p2, n0, n1, n2, nc, cx.huber_a);
      }
    }
  }
}

template <int PB, int PC>
__device__ __forceinline__ void sweep_warp(double (&acc)[kAcc], const TileSrc &ts, WarpPipe &wp, const double *pose, double huber_a,
                                           double sqrt_huber_a, int &cnt_edge, int &cnt_plane) {
  using WT = WarpTile<PB, PC>;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < kAcc; ++k) acc[k] = 0.0;
  SweepCtx cx;
  quat_to_R(pose + 3, cx.R);
  cx.t0 = pose[0]; cx.t1 = pose[1]; cx.t2 = pose[2];
  cx.huber_a = Huber{huber_a, huber_a * huber_a, sqrt_huber_a};
  cnt_edge = 0;
  cnt_plane = 0;
  TileIt cur = tile_first<PB, PC>(ts, warp);
  if (cur.cls == 2u) return;  // fewer tiles than warps: nothing for this warp
  if (wp.resident) {
    if (!wp.loaded && lane == 0) {
      uint32_t k = 0;
      for (TileIt t = cur; t.cls != 2u; t = tile_next<PB, PC>(ts, warp, t), ++k)
        issue_warp_tile<PB, PC>(ts, t, wp.ring + k * WT::SB, wp.bars + k * 8u);
    }
    uint32_t k = 0;
    for (; cur.cls != 2u; cur = tile_next<PB, PC>(ts, warp, cur), ++k) {
      if (!wp.loaded) mbar_wait_s(wp.bars + k * 8u, 0u);
      consume_tile<PB, PC>(acc, cx, ts, cur, wp.ring + k * WT::SB, lane, cnt_edge, cnt_plane);
    }
    wp.loaded = true;
    return;
  }
  if (!wp.loaded && lane == 0) issue_warp_tile<PB, PC>(ts, cur, wp.ring + (wp.it & 1u) * WT::SB, wp.bars + (wp.it & 1u) * 8u);
  for (;;) {
    TileIt nxt = tile_next<PB, PC>(ts, warp, cur);
    const bool last = nxt.cls == 2u;
    if (last) nxt = tile_first<PB, PC>(ts, warp);  // prefetch across the sweep boundary
    const uint32_t s = wp.it & 1u;
    // the other stage held the previous tile: every lane left it at the __syncwarp below
    if (lane == 0) issue_warp_tile<PB, PC>(ts, nxt, wp.ring + (s ^ 1u) * WT::SB, wp.bars + (s ^ 1u) * 8u);
    mbar_wait_s(wp.bars + s * 8u, (wp.it >> 1) & 1u);
    consume_tile<PB, PC>(acc, cx, ts, cur, wp.ring + s * WT::SB, lane, cnt_edge, cnt_plane);
    __syncwarp();
    ++wp.it;
    if (last) break;
    cur = nxt;
  }
  wp.loaded = true;
}

// before the CTA exits: a streaming warp still has the prefetched first tile of a sweep that will never run in flight
__device__ __forceinline__ void warp_pipe_drain(const WarpPipe &wp) {
  if (!wp.resident && wp.loaded) mbar_wait_s(wp.bars + (wp.it & 1u) * 8u, (wp.it >> 1) & 1u);
}

template <int PB, int PC>
__device__ __forceinline__ WarpPipe warp_pipe_init(const TileSrc &ts, unsigned char *ring_cta, uint64_t *bars_cta, uint32_t n_stages) {
  using WT = WarpTile<PB, PC>;
  const uint32_t warp = threadIdx.x >> 5;
  WarpPipe wp;
  wp.ring = smem_u32(ring_cta) + warp * n_stages * WT::SB;
  wp.bars = smem_u32(bars_cta) + warp * kMaxWarpStages * 8u;
  wp.it = 0;
  const uint32_t tt_e = (ts.n_e + WT::TE - 1) / WT::TE, tt_p = (ts.n_p + WT::TP - 1) / WT::TP;
  const uint32_t nw = blockDim.x >> 5;
  wp.my_tiles = (tt_e > warp ? (tt_e - warp + nw - 1) / nw : 0u) + (tt_p > warp ? (tt_p - warp + nw - 1) / nw : 0u);
  wp.resident = n_stages > (uint32_t)kWarpStages && wp.my_tiles <= n_stages;
  wp.loaded = false;
  return wp;
}

// ---- cluster-level combine over distributed shared memory (G CTAs per scan) ---------------------
// rank 0 adds the other CTAs' block sums in rank order (deterministic), then owns the LM step.
template <class SH>
__device__ __forceinline__ void cluster_reduce(cg::cluster_group &cluster, SH &sh, uint32_t G, uint32_t rank,
                                               bool with_counts) {
  cluster.sync();  // every CTA's sh.cand (and counts) are written
  if (rank == 0) {
    if (threadIdx.x < kAcc) {
      double v = sh.cand[threadIdx.x];
      for (uint32_t r = 1; r < G; ++r) v += *cluster.map_shared_rank(&sh.cand[threadIdx.x], r);
      sh.cand[threadIdx.x] = v;
    } else if (with_counts && threadIdx.x == 32) {
      int ne = sh.n_edge, np = sh.n_plane;
      for (uint32_t r = 1; r < G; ++r) {
        ne += *cluster.map_shared_rank(&sh.n_edge, r);
        np += *cluster.map_shared_rank(&sh.n_plane, r);
      }
      sh.n_edge = ne;
      sh.n_plane = np;
    }
    __syncthreads();
  }
}

// rank 0 publishes the candidate pose / done flag; the other CTAs copy them into their own smem
template <class SH>
__device__ __forceinline__ void cluster_broadcast(cg::cluster_group &cluster, SH &sh, uint32_t rank) {
  cluster.sync();  // rank 0's xc / done are final
  if (rank != 0) {
    if (threadIdx.x < 7) sh.xc[threadIdx.x] = *cluster.map_shared_rank(&sh.xc[threadIdx.x], 0);
    if (threadIdx.x == 7) sh.done = *cluster.map_shared_rank(&sh.done, 0);
    if (threadIdx.x == 8) sh.too_few = *cluster.map_shared_rank(&sh.too_few, 0);
  }
  __syncthreads();
}

// PB: bytes per point (16 float4 / 32 double4), PC: bytes of plane constants per entry (32 / 48)
template <int PB, int PC, int NT>
__global__ void __launch_bounds__(NT, 512 / NT)
k_lm_solve(KParams kp, const void *__restrict__ qe, const int32_t *__restrict__ e_off, uint32_t n_edge_total,
           const void *__restrict__ qp, const int32_t *__restrict__ p_off, const double *__restrict__ corr,
           double *__restrict__ poses, int32_t *__restrict__ status, msfl_stats *__restrict__ stats, int outer,
           int min_corr, uint32_t n_stages) {
  constexpr uint32_t kLmWarps = NT / 32;
  __shared__ LmSharedT<NT / 32> sh;
  __shared__ __align__(8) uint64_t bars[kLmWarps * kMaxWarpStages];
  extern __shared__ __align__(128) unsigned char ring[];
  // One thread-block cluster per scan: G CTAs (G = 1, 2, 4 or 8) each sweep 1/G of the scan's
  // correspondences; partial sums meet in the rank-0 CTA through distributed shared memory.
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t G = cluster.num_blocks(), rank = cluster.block_rank();
  const int b = blockIdx.x / G;
  uint32_t eo = (uint32_t)e_off[b], n_e = (uint32_t)e_off[b + 1] - eo;
  uint32_t po = (uint32_t)p_off[b], n_p = (uint32_t)p_off[b + 1] - po;
  if (G > 1) {  // this CTA's contiguous share of each class
    const uint32_t ce_chunk = (n_e + G - 1) / G, cp_chunk = (n_p + G - 1) / G;
    const uint32_t e0 = min(n_e, rank * ce_chunk), e1 = min(n_e, e0 + ce_chunk);
    const uint32_t p0 = min(n_p, rank * cp_chunk), p1 = min(n_p, p0 + cp_chunk);
    eo += e0; n_e = e1 - e0;
    po += p0; n_p = p1 - p0;
  }
  const unsigned char *pe = (const unsigned char *)qe + (size_t)eo * PB, *pp = (const unsigned char *)qp + (size_t)po * PB;
  const double *ce_ = corr + (size_t)eo * 6;
  const double *cp_ = (const double *)((const unsigned char *)(corr + (size_t)n_edge_total * 6) + (size_t)po * PC);
  msfl_stats *st = (stats && rank == 0) ? stats + b : nullptr;
  msfl_lm_log *log = st ? &st->lm[outer] : nullptr;
  const uint32_t tid = threadIdx.x;

  if (outer > 0 && status[b] != MSFL_OK) return;  // an earlier outer iteration bailed out (odometry :266)

  if (tid < 7) sh.x[tid] = poses[(size_t)b * 7 + tid];
  if (tid == 0) {
    sh.done = 0;
    sh.too_few = 0;
    for (int i = 0; i < (int)(kLmWarps * kMaxWarpStages); ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  TileSrc ts;
  ts.pe = pe; ts.pp = pp; ts.ce = ce_; ts.cp = cp_;
  ts.n_e = n_e; ts.n_p = n_p;
  WarpPipe wp = warp_pipe_init<PB, PC>(ts, ring, bars, n_stages);

  double acc[kAcc];
  int ce, cpl;
  sweep_warp<PB, PC>(acc, ts, wp, sh.x, kp.huber_a, kp.huber_sqrt_a, ce, cpl);
  // correspondence counts (corner_num / surf_num, mapping_scan_matcher.cc:173,243)
  for (int o = 16; o > 0; o >>= 1) {
    ce += __shfl_down_sync(0xffffffffu, ce, o);
    cpl += __shfl_down_sync(0xffffffffu, cpl, o);
  }
  if ((tid & 31) == 0) { sh.cnt[tid >> 5][0] = ce; sh.cnt[tid >> 5][1] = cpl; }
  block_reduce(acc, sh);
  if (tid == 0) {
    int ne = 0, np = 0;
    for (int w = 0; w < NT / 32; ++w) { ne += sh.cnt[w][0]; np += sh.cnt[w][1]; }
    sh.n_edge = ne;
    sh.n_plane = np;
  }
  if (G > 1) cluster_reduce(cluster, sh, G, rank, /*with_counts=*/true);
  if (tid == 0 && rank == 0) {
    const int ne = sh.n_edge, np = sh.n_plane;
    if (st) {
      st->n_edge[outer] = ne;
      st->n_plane[outer] = np;
      st->n_outer = outer + 1;
      if (outer == 0) st->status = MSFL_OK;
    }
    if (outer == 0) status[b] = MSFL_OK;
    if (log) { log->n_attempts = 0; log->termination = 0; log->initial_cost = sh.cand[27]; log->final_cost = sh.cand[27]; }
    if (ne + np < min_corr) {  // odometry_scan_matcher.cc:262-267: return false, pose untouched
      sh.done = 1;
      sh.too_few = 1;
      status[b] = MSFL_TOO_FEW;
      if (st) st->status = MSFL_TOO_FEW;
    } else if (ne + np == 0) {  // no residual blocks: nothing for the solver to do
      sh.done = 1;
      if (log) log->termination = 2;
    } else {
      for (int k = 0; k < 21; ++k) sh.H[k] = sh.cand[k];
      for (int k = 0; k < 6; ++k) sh.g[k] = sh.cand[21 + k];
      sh.cost = sh.cand[27];
      // jacobi scaling, fixed at iteration 0 (trust_region_minimizer.cc)
      for (int k = 0; k < 6; ++k) sh.S[k] = 1.0 / (1.0 + sqrt(sh.H[tri6(k, k)]));
      sh.x_norm = norm7(sh.x);
      sh.radius = kp.initial_radius;
      sh.nu = 2.0;
      sh.reuse = 0;
      sh.n_invalid = 0;
      sh.iteration = 0;
      sh.step_successful = 1;
      sh.termination = 0;
      lm_prepare_step(sh, kp, log);
    }
  }
  if (G > 1) cluster_broadcast(cluster, sh, rank);
  else __syncthreads();
#ifdef MSFL_LM_TIMING  // development (-DMSFL_LM_TIMING): cycles per phase of CTA 0, printed at exit
  long long t_sweep = 0, t_red = 0, t_step = 0, t_sync = 0;
#define MSFL_TICK(v) const long long v = clock64()
#else
#define MSFL_TICK(v)
#endif
  while (!sh.done) {
    MSFL_TICK(c0);
    sweep_warp<PB, PC>(acc, ts, wp, sh.xc, kp.huber_a, kp.huber_sqrt_a, ce, cpl);
    MSFL_TICK(c1);
    block_reduce(acc, sh);
    if (G > 1) cluster_reduce(cluster, sh, G, rank, false);
    MSFL_TICK(c2);
    if (tid == 0 && rank == 0) {
      lm_finish_step(sh, kp, log);
      if (!sh.done) lm_prepare_step(sh, kp, log);
    }
    MSFL_TICK(c3);
    if (G > 1) cluster_broadcast(cluster, sh, rank);
    else __syncthreads();
#ifdef MSFL_LM_TIMING
    t_sweep += c1 - c0; t_red += c2 - c1; t_step += c3 - c2; t_sync += clock64() - c3;
#endif
  }
#ifdef MSFL_LM_TIMING
  if (blockIdx.x == 0 && (tid == 0 || tid == 64))
    printf("LM timing CTA 0 tid %d: sweep %lld reduce %lld step %lld sync %lld cycles (n_e %u n_p %u)\n", (int)tid, t_sweep, t_red,
           t_step, t_sync, n_e, n_p);
#endif
#undef MSFL_TICK
  if (rank == 0) {
    if (tid == 0 && log && !sh.too_few && sh.n_edge + sh.n_plane > 0) {
      log->termination = sh.termination;
      log->final_cost = sh.cost;
    }
    if (tid < 7 && !sh.too_few) poses[(size_t)b * 7 + tid] = sh.x[tid];
  }
  warp_pipe_drain(wp);
  if (G > 1) cluster.sync();  // no CTA may exit while a peer can still read its shared memory
}

template <int PB, int PC, int NT>
static int launch_lm_solve_nt(msfl_engine *e, int B, int G, const void *d_qe, const int32_t *d_e_off, uint32_t n_edge_total,
                              const void *d_qp, const int32_t *d_p_off, const double *d_corr, double *d_poses,
                              int32_t *d_status, msfl_stats *d_stats, int outer, int min_corr) {
  constexpr int sb = (int)((NT / 32) * WarpTile<PB, PC>::SB);  // smem per ring stage (whole CTA)
  constexpr int max_stages = kMaxWarpStages * sb <= 227 * 1024 ? kMaxWarpStages : (227 * 1024) / sb;
  bool &attr_set = e->lm_attr_set[(PB == 16 ? 0 : 1) + (PC == 32 ? 2 : 0) + (NT == kLmThreadsBig ? 0 : 4)];  // per engine: attributes are per device
  if (!attr_set) {
    MSFL_CUDA_OK(cudaFuncSetAttribute(k_lm_solve<PB, PC, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_stages * sb));
    attr_set = true;
  }
  // double-buffered streaming (4 or 5 CTAs/SM) for throughput batches; for small launches (fewer CTAs than SMs:
  // occupancy is irrelevant) a deep ring so that a warp's tiles stay resident in smem across the sweeps
  const uint32_t n_stages = ((long long)B * G <= (long long)e->sm_count) ? (uint32_t)max_stages : (uint32_t)kWarpStages;
  // (capping the resident CTAs so that the factor arrays in flight fit the 126 MB L2 -- 3 per SM: 444 x 222 KB = 99 MB
  // -- was measured on B200: 0.706 -> 0.772 ms per launch at 3 CTAs/SM, 0.976 ms at 2; the sweep needs the warps more
  // than it needs the L2 hits)
  const int smem = (int)n_stages * sb;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)B * G);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = e->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = G;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MSFL_CUDA_OK(cudaLaunchKernelEx(&cfg, k_lm_solve<PB, PC, NT>, e->kp, d_qe, d_e_off, n_edge_total, d_qp, d_p_off, d_corr, d_poses,
                                  d_status, d_stats, outer, min_corr, n_stages));
  e->launches += 1;
  MSFL_CUDA_OK(cudaGetLastError());
  return MSFL_OK;
}

template <int PB, int PC>
static int launch_lm_solve_t(msfl_engine *e, int B, const void *d_qe, const int32_t *d_e_off, uint32_t n_edge_total,
                             const void *d_qp, const int32_t *d_p_off, const double *d_corr, double *d_poses,
                             int32_t *d_status, msfl_stats *d_stats, int outer, int min_corr) {
  int G = e->params.lm_cluster;
  if (G == 16) G = 8;  // 16 is the fused single-scan kernel's non-portable size; this kernel stays portable
  if (G != 2 && G != 4 && G != 8) G = 1;  // 0 / 1: one CTA per scan
  // three-warp CTAs once the launch fills five CTAs on every SM, four-warp CTAs below that
  if ((long long)std::max(B, e->lm_shape_scans) * G >= 5ll * e->sm_count)
    return launch_lm_solve_nt<PB, PC, kLmThreadsSmall>(e, B, G, d_qe, d_e_off, n_edge_total, d_qp, d_p_off, d_corr, d_poses, d_status,
                                                       d_stats, outer, min_corr);
  return launch_lm_solve_nt<PB, PC, kLmThreadsBig>(e, B, G, d_qe, d_e_off, n_edge_total, d_qp, d_p_off, d_corr, d_poses, d_status,
                                                   d_stats, outer, min_corr);
}

// plane_bytes: 48 = plane constants {c, n} (6 doubles, same layout as the edge entries), 32 = {n, n.c} (k_fit compact)
int launch_lm_solve(msfl_engine *e, int B, const float4 *d_qe, const int32_t *d_e_off, uint32_t n_edge_total,
                    const float4 *d_qp, const int32_t *d_p_off, const double *d_corr, double *d_poses, int32_t *d_status,
                    msfl_stats *d_stats, int outer, int min_corr, int plane_bytes) {
  if (B <= 0) return MSFL_OK;
  if (plane_bytes == 32)
    return launch_lm_solve_t<16, 32>(e, B, d_qe, d_e_off, n_edge_total, d_qp, d_p_off, d_corr, d_poses, d_status, d_stats, outer,
                                     min_corr);
  return launch_lm_solve_t<16, 48>(e, B, d_qe, d_e_off, n_edge_total, d_qp, d_p_off, d_corr, d_poses, d_status, d_stats, outer,
                                   min_corr);
}

// deskewed points: double4 (p' = dq p + dp in fp64, lidar_factor.cc:53), 32 B per entry
int launch_lm_solve_pd(msfl_engine *e, int B, const double *d_qe, const int32_t *d_e_off, uint32_t n_edge_total,
                       const double *d_qp, const int32_t *d_p_off, const double *d_corr, double *d_poses, int32_t *d_status,
                       msfl_stats *d_stats, int outer, int min_corr) {
  if (B <= 0) return MSFL_OK;
  return launch_lm_solve_t<32, 48>(e, B, d_qe, d_e_off, n_edge_total, d_qp, d_p_off, d_corr, d_poses, d_status, d_stats, outer,
                                   min_corr);
}

// ---- test hook: plain accumulate at a pose (cost, H, g), one block ---------------------------
__global__ void __launch_bounds__(kLmThreads)
k_accumulate(KParams kp, const float4 *__restrict__ p, const double *__restrict__ corr, uint32_t n_edge, uint32_t n_total,
             const double *__restrict__ pose, double *__restrict__ out28) {
  __shared__ LmShared sh;
  if (threadIdx.x < 7) sh.x[threadIdx.x] = pose[threadIdx.x];
  __syncthreads();
  double acc[kAcc];
  int ce, cpl;
  sweep(acc, p, corr, n_edge, p + n_edge, corr + (size_t)n_edge * 6, n_total - n_edge, sh.x, kp.huber_a, threadIdx.x,
        kLmThreads, ce, cpl);
  block_reduce(acc, sh);
  if (threadIdx.x < kAcc) out28[threadIdx.x] = sh.cand[threadIdx.x];
}

int launch_accumulate(msfl_engine *e, const float4 *d_p, const double *d_corr, int n_edge, int n_plane,
                      const double *d_pose, double *d_out28) {
  k_accumulate<<<1, kLmThreads, 0, e->stream>>>(e->kp, d_p, d_corr, (uint32_t)n_edge, (uint32_t)(n_edge + n_plane), d_pose,
                                                d_out28);
  e->launches += 1;
  MSFL_CUDA_OK(cudaGetLastError());
  return MSFL_OK;
}

}  // namespace msfl
