// scan2map_fused.cu -- the latency path: ONE scan against the resident submap, the way the ROS node calls
// MappingScanMatcher::MatchScan2Map (laser_mapping.cc:304-311), as ONE kernel launch.
//
// The batch path runs a scan-to-map as 2 x (search, fit, fallback, solve) launches; for a single VLP-16 scan every one
// of them is a few microseconds of work behind a few microseconds of launch and dependency latency, and the solve sits
// on one SM.  Here a thread-block cluster of G CTAs (8 portable, 16 with the non-portable attribute) owns the scan for
// the whole call:
//   outer iteration:  every CTA associates its contiguous share of the queries (transform, exact 5-NN over the cell
//                     index, line / plane fit -- the device functions of the batch kernels) and keeps the raw points
//                     and factor constants IN SHARED MEMORY (64 / 48 B per query; nothing is written to HBM);
//   LM attempt:       every thread sweeps its entries out of shared memory, the CTA reduces them in a fixed order, the
//                     28 partial sums are published in a double-buffered slot of the CTA's shared memory, ONE
//                     cluster barrier, then every CTA adds all G partials in rank order over distributed shared memory
//                     and runs the same trust-region step redundantly -- no second barrier to broadcast the candidate.
// Same arithmetic as the batch kernels per query and per factor; only the order in which the per-thread sums meet
// differs, so poses agree with the batch path to round-off (~1e-15) and with the oracle like the batch path does.
#include <cooperative_groups.h>

#include "assoc_device.cuh"
#include "lm_device.cuh"

namespace cg = cooperative_groups;

namespace msfl {

#ifndef MSFL_FUSED_THREADS
#define MSFL_FUSED_THREADS 512
#endif
constexpr int kFusedThreads = MSFL_FUSED_THREADS;
using FusedShared = LmSharedT<kFusedThreads / 32>;

// this thread's entries out of shared memory: edge entries {a, n} (6 doubles), plane entries {n, n.c} (4 doubles)
__device__ __forceinline__ void sweep_smem(double (&acc)[kAcc], const float4 *pe, const double *ce, uint32_t n_e, const float4 *pp,
                                           const double *cp, uint32_t n_p, const double *pose, const Huber &hub, int &cnt_e,
                                           int &cnt_p) {
#pragma unroll
  for (int k = 0; k < kAcc; ++k) acc[k] = 0.0;
  double R[9];
  quat_to_R(pose + 3, R);
  const double t0 = pose[0], t1 = pose[1], t2 = pose[2];
  cnt_e = 0;
  cnt_p = 0;
  for (uint32_t i = threadIdx.x; i < n_e; i += kFusedThreads) {
    const double *c = ce + (size_t)i * 6;
    if (c[3] == 0.0 && c[4] == 0.0 && c[5] == 0.0) continue;  // no factor for this query
    const float4 p = pe[i];
    ++cnt_e;
    eval_edge(acc, R, t0, t1, t2, p.x, p.y, p.z, c[0], c[1], c[2], c[3], c[4], c[5], hub);
  }
  for (uint32_t i = threadIdx.x; i < n_p; i += kFusedThreads) {
    const double *c = cp + (size_t)i * 4;
    if (c[0] == 0.0 && c[1] == 0.0 && c[2] == 0.0) continue;
    const float4 p = pp[i];
    ++cnt_p;
    eval_plane(acc, R, t0, t1, t2, p.x, p.y, p.z, c[0], c[1], c[2], c[3], hub);
  }
}

__global__ void __launch_bounds__(kFusedThreads, 1)
k_scan2map_fused(GridView gc, GridView gs, KParams kp, const float4 *__restrict__ qc, uint32_t nc, const float4 *__restrict__ qs,
                 uint32_t ns, double *__restrict__ pose_io, int32_t *__restrict__ status, msfl_stats *__restrict__ stats,
                 int num_outer, uint32_t cap_e, uint32_t cap_p) {
  __shared__ FusedShared sh;
  __shared__ double part[2][kAcc];  // this CTA's block sums, double-buffered: peers read slot g while slot g^1 is rewritten
  __shared__ int part_cnt[2];       // this CTA's factor counts (edge, plane) of the current outer iteration
  extern __shared__ __align__(16) unsigned char dyn[];
  float4 *pe = reinterpret_cast<float4 *>(dyn);
  float4 *pp = pe + cap_e;
  double *ce = reinterpret_cast<double *>(pp + cap_p);
  double *cp = ce + (size_t)cap_e * 6;

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t G = cluster.num_blocks(), rank = cluster.block_rank();
  const uint32_t tid = threadIdx.x;
  // this CTA's contiguous share of each class
  const uint32_t e0 = min(nc, rank * cap_e), n_e = min(nc, e0 + cap_e) - e0;
  const uint32_t p0 = min(ns, rank * cap_p), n_p = min(ns, p0 + cap_p) - p0;
  msfl_stats *st = (stats && rank == 0) ? stats : nullptr;
  const Huber hub{kp.huber_a, kp.huber_a * kp.huber_a, kp.huber_sqrt_a};
  if (tid < 7) sh.x[tid] = pose_io[tid];
  __syncthreads();
  uint32_t gen = 0;

  // publish this CTA's block sums, meet the cluster, add all partials in rank order (every CTA gets the same bits)
  auto combine = [&](bool with_counts) {
    if (tid < kAcc) part[gen & 1u][tid] = sh.cand[tid];
    if (G > 1) cluster.sync();
    else __syncthreads();
    if (tid < kAcc) {
      double v = 0.0;
      for (uint32_t r = 0; r < G; ++r) v += *cluster.map_shared_rank(&part[gen & 1u][tid], r);
      sh.cand[tid] = v;
    } else if (with_counts && tid == 32) {
      int ne = 0, np = 0;
      for (uint32_t r = 0; r < G; ++r) {
        const int *pc = cluster.map_shared_rank(&part_cnt[0], r);
        ne += pc[0];
        np += pc[1];
      }
      sh.n_edge = ne;
      sh.n_plane = np;
    }
    __syncthreads();
    ++gen;
  };

  double acc[kAcc];
  int cnt_e, cnt_p;
#ifdef MSFL_FUSED_TIMING  // development (-DMSFL_FUSED_TIMING): cycles per phase of this thread, printed by CTA 0 at exit
  long long t_assoc = 0, t_sweep = 0, t_red = 0, t_comb = 0, t_step = 0;
  int n_attempts = 0;
#define MSFL_FTICK(v) const long long v = clock64()
#define MSFL_FADD(acc_, a, b) acc_ += (b) - (a)
#else
#define MSFL_FTICK(v)
#define MSFL_FADD(acc_, a, b)
#endif
  for (int outer = 0; outer < num_outer; ++outer) {  // mapping_scan_matcher.cc:75
    msfl_lm_log *log = st ? &st->lm[outer] : nullptr;
    // ---- data association at the current pose (:109-246) into shared memory
    MSFL_FTICK(f0);
    {
      double pose[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) pose[i] = sh.x[i];
      for (uint32_t i = tid; i < n_e + n_p; i += kFusedThreads) {
        const bool is_corner = i < n_e;
        const uint32_t j = is_corner ? i : i - n_e;
        const float4 p = is_corner ? __ldg(qc + e0 + j) : __ldg(qs + p0 + j);
        const float3 x = transform_point_f(pose, p.x, p.y, p.z);  // :123 / :193
        const GridView &g = is_corner ? gc : gs;
        Top5 t;
        const bool gate = knn5_grid(g, x.x, x.y, x.z, kp.knn_max_sq_f, t);  // :125-128 / :195-198
        double a[3] = {0, 0, 0}, n[3] = {0, 0, 0};
        if (gate) {
          float mf[5][3];
#pragma unroll
          for (int s = 0; s < 5; ++s) {
            const float4 mp = __ldg(g.pts_orig + t.i[s]);
            mf[s][0] = mp.x; mf[s][1] = mp.y; mf[s][2] = mp.z;
          }
          double c[3];
          centroid5(mf, c);
          if (is_corner) {
            line_fit(mf, c, kp, a, n);
          } else {
            if (plane_fit_fast(mf, c, kp.plane_tol, n)) plane_fit_qr(mf, c, kp, n);  // declined: the reference's own route
#pragma unroll
            for (int d = 0; d < 3; ++d) a[d] = c[d];
          }
        }
        if (is_corner) {
          pe[j] = p;
          double *o = ce + (size_t)j * 6;
          o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = n[0]; o[4] = n[1]; o[5] = n[2];
        } else {
          pp[j] = p;
          double *o = cp + (size_t)j * 4;
          o[0] = n[0]; o[1] = n[1]; o[2] = n[2];
          o[3] = __fma_rn(n[2], a[2], __fma_rn(n[1], a[1], __dmul_rn(n[0], a[0])));
        }
      }
    }
    __syncthreads();
    MSFL_FTICK(f1);
    MSFL_FADD(t_assoc, f0, f1);
    // ---- ceres::Solve (:250-272): evaluation at x, then the trust-region loop
    sweep_smem(acc, pe, ce, n_e, pp, cp, n_p, sh.x, hub, cnt_e, cnt_p);
    MSFL_FTICK(f2);
    MSFL_FADD(t_sweep, f1, f2);
    for (int o = 16; o > 0; o >>= 1) {
      cnt_e += __shfl_down_sync(0xffffffffu, cnt_e, o);
      cnt_p += __shfl_down_sync(0xffffffffu, cnt_p, o);
    }
    if ((tid & 31) == 0) { sh.cnt[tid >> 5][0] = cnt_e; sh.cnt[tid >> 5][1] = cnt_p; }
    block_reduce(acc, sh);
    if (tid == 0) {
      int ne = 0, np = 0;
      for (int w = 0; w < kFusedThreads / 32; ++w) { ne += sh.cnt[w][0]; np += sh.cnt[w][1]; }
      part_cnt[0] = ne;
      part_cnt[1] = np;
    }
    __syncthreads();
    MSFL_FTICK(f3);
    MSFL_FADD(t_red, f2, f3);
    combine(true);
    MSFL_FTICK(f4);
    MSFL_FADD(t_comb, f3, f4);
    if (tid == 0) {  // every CTA runs the same control code on the same sums
      const int ne = sh.n_edge, np = sh.n_plane;
      if (st) {
        st->n_edge[outer] = ne;
        st->n_plane[outer] = np;
        st->n_outer = outer + 1;
        st->status = MSFL_OK;
      }
      if (log) { log->n_attempts = 0; log->termination = 0; log->initial_cost = sh.cand[27]; log->final_cost = sh.cand[27]; }
      sh.done = 0;
      sh.too_few = 0;
      if (ne + np == 0) {  // no residual blocks: nothing for the solver to do
        sh.done = 1;
        if (log) log->termination = 2;
      } else {
        for (int k = 0; k < 21; ++k) sh.H[k] = sh.cand[k];
        for (int k = 0; k < 6; ++k) sh.g[k] = sh.cand[21 + k];
        sh.cost = sh.cand[27];
        for (int k = 0; k < 6; ++k) sh.S[k] = 1.0 / (1.0 + sqrt(sh.H[tri6(k, k)]));  // jacobi scaling, fixed at iteration 0
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
    __syncthreads();
    MSFL_FTICK(f5);
    MSFL_FADD(t_step, f4, f5);
    while (!sh.done) {
      MSFL_FTICK(g0);
      sweep_smem(acc, pe, ce, n_e, pp, cp, n_p, sh.xc, hub, cnt_e, cnt_p);
      MSFL_FTICK(g1);
      block_reduce(acc, sh);
      MSFL_FTICK(g2);
      combine(false);
      MSFL_FTICK(g3);
      if (tid == 0) {
        lm_finish_step(sh, kp, log);
        if (!sh.done) lm_prepare_step(sh, kp, log);
      }
      __syncthreads();
#ifdef MSFL_FUSED_TIMING
      t_sweep += g1 - g0; t_red += g2 - g1; t_comb += g3 - g2; t_step += clock64() - g3; ++n_attempts;
#endif
    }
    if (tid == 0 && log && sh.n_edge + sh.n_plane > 0) {
      log->termination = sh.termination;
      log->final_cost = sh.cost;
    }
    __syncthreads();  // sh.x is final for this outer iteration in every CTA
  }
#ifdef MSFL_FUSED_TIMING
  if (rank == 0 && (tid == 0 || tid == 33))
    printf("fused timing rank 0 tid %d: assoc %lld sweep %lld reduce %lld combine %lld step %lld cycles, %d attempts (n_e %u n_p %u)\n",
           (int)tid, t_assoc, t_sweep, t_red, t_comb, t_step, n_attempts, n_e, n_p);
#endif
#undef MSFL_FTICK
#undef MSFL_FADD
  if (rank == 0) {
    if (tid < 7) pose_io[tid] = sh.x[tid];
    if (tid == 7 && status) status[0] = MSFL_OK;
  }
  if (G > 1) cluster.sync();  // no CTA may exit while a peer can still read its partial sums
}

// Returns MSFL_OK when the fused kernel was enqueued, 1 when this scan does not qualify (the caller takes the batch
// path), < 0 on error.  G = params.lm_cluster (2, 4, 8 or 16).
int launch_scan2map_fused(msfl_engine *e, const float4 *d_qc, uint32_t nc, const float4 *d_qs, uint32_t ns, double *d_pose,
                          int32_t *d_status, msfl_stats *d_stats) {
  int G = e->params.lm_cluster;
  if (G != 2 && G != 4 && G != 8 && G != 16) return 1;
  if (nc + ns == 0) return 1;
  const uint32_t cap_e = (nc + G - 1) / G, cap_p = (ns + G - 1) / G;
  const size_t smem = (size_t)cap_e * (16 + 48) + (size_t)cap_p * (16 + 32) + 16;
  if (smem > 200 * 1024) return 1;  // a scan this large fills the chip through the batch kernels anyway
  if (!e->fused_attr_set) {
    MSFL_CUDA_OK(cudaFuncSetAttribute(k_scan2map_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MSFL_CUDA_OK(cudaFuncSetAttribute(k_scan2map_fused, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    e->fused_attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)G);
  cfg.blockDim = dim3(kFusedThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = e->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = G;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (G == 16 && e->fused_max16 < 0) {  // can this device co-schedule 16 CTAs of this size in one GPC?
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k_scan2map_fused, &cfg) != cudaSuccess) { n = 0; (void)cudaGetLastError(); }
    e->fused_max16 = n;
  }
  if (G == 16 && e->fused_max16 <= 0) {  // fall back to the portable cluster size
    G = 8;
    cfg.gridDim = dim3(8);
    attr[0].val.clusterDim.x = 8;
    const uint32_t ce8 = (nc + 7) / 8, cp8 = (ns + 7) / 8;
    cfg.dynamicSmemBytes = (size_t)ce8 * 64 + (size_t)cp8 * 48 + 16;
    if (cfg.dynamicSmemBytes > 200 * 1024) return 1;
    MSFL_CUDA_OK(cudaLaunchKernelEx(&cfg, k_scan2map_fused, e->map_corner.view, e->map_surf.view, e->kp, d_qc, nc, d_qs, ns, d_pose,
                                    d_status, d_stats, (int)e->params.num_outer, ce8, cp8));
  } else {
    MSFL_CUDA_OK(cudaLaunchKernelEx(&cfg, k_scan2map_fused, e->map_corner.view, e->map_surf.view, e->kp, d_qc, nc, d_qs, ns, d_pose,
                                    d_status, d_stats, (int)e->params.num_outer, cap_e, cap_p));
  }
  e->launches += 1;
  MSFL_CUDA_OK(cudaGetLastError());
  return MSFL_OK;
}

}  // namespace msfl
