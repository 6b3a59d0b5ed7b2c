// lm_device.cuh -- device code of the Levenberg-Marquardt solve shared by the batch kernel (lm_solve.cu) and the
// single-scan fused kernel (scan2map_fused.cu): the two factors (lidar_factor.cc:7-44), the Huber corrector, the
// deterministic block reduction and the Ceres-compatible trust-region control.
#pragma once
#include "msfl_internal.h"
#include "msfl_math.cuh"

namespace msfl {

constexpr int kAcc = 28;  // 21 upper-tri H + 6 g + 1 cost

template <int WARPS>
struct LmSharedT {
  double x[7], xc[7];
  double H[21], g[6], cost;
  double red[WARPS][kAcc];
  double cand[kAcc];
  double S[6], diag[6];
  double radius, nu, x_norm, model;
  int reuse, n_invalid, iteration, done, step_successful, termination, n_edge, n_plane, too_few;
  int cnt[WARPS][2];
  static constexpr int kWarps = WARPS;
};

// One residual row: H += J J^T, g += J r.
__device__ __forceinline__ void acc_row(double (&acc)[kAcc], const double J[6], double r) {
  int k = 0;
#pragma unroll
  for (int u = 0; u < 6; ++u)
#pragma unroll
    for (int v = u; v < 6; ++v) acc[k++] += J[u] * J[v];
#pragma unroll
  for (int u = 0; u < 6; ++u) acc[21 + u] += J[u] * r;
}

// s^(-1/4) for the Huber outlier path: two MUFU.RSQ give a 22-bit fp32 seed, two Newton steps of
// y <- y (1 + (1 - s y^4) / 4) (error e -> 2.5 e^2) bring it to fp64 round-off.  12 fp64 instructions instead of the
// ~45 of rsqrt(double) followed by sqrt(double); only a few lanes of a warp take this path, so its length is what
// the whole warp pays.
__device__ __forceinline__ double inv_fourth_root(double s) {
  float r, q;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)s));  // s^(-1/2), one MUFU.RSQ
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(r));         // r^(-1/2)
  double y = (double)(r * q);                                      // sqrt(r) = s^(-1/4)
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double y2 = y * y;
    const double e = fma(-s, y2 * y2, 1.0);
    y = fma(y, 0.25 * e, y);
  }
  return y;
}

// Huber loss + Ceres corrector (loss_function.cc HuberLoss::Evaluate; corrector.cc: rho'' <= 0 so
// residual and Jacobian are both scaled by sqrt(rho')).  Returns the scale, adds 0.5 rho to cost.
// hub = {a, a^2, sqrt(a)}.
struct Huber { double a, b, sqrt_a; };
__device__ __forceinline__ double huber_scale(double s, const Huber &hub, double &cost) {
  if (s > hub.b) {
    // rho = 2 a sqrt(s) - a^2, rho' = a / sqrt(s); with y = s^(-1/4): sqrt(s) = s y^2, sqrt(rho') = sqrt(a) y
    // (beyond the fp32 range of the seed -- |r| > 1e15 m -- the library routines take over)
    const double y = s < 1e30 ? inv_fourth_root(s) : sqrt(rsqrt(s));
    cost += 0.5 * (2.0 * hub.a * (s * (y * y)) - hub.b);
    return hub.sqrt_a * y;
  }
  cost += 0.5 * s;
  return 1.0;
}

// ---------------------------------------------------------------------------------------------
// The two factors (same residuals / Jacobians as lidar_factor.cc:7-44, regrouped so that R [p]x is
// never formed).  With w = R^T n:
//   plane:  r = n.(R p + t - c) = w.p + (n.t - n.c)            J = [ n^T | (p x w)^T ]
//   edge :  r = n x (R p + t - a)                               J = [ [n]x | (w.p) R - (R p) w^T ]
// ( -[n]x R [p]x = -R [w]x [p]x = -R (p w^T - (w.p) I) ).  Residual and Jacobian rows are scaled by
// the Huber corrector before they enter H = J^T J, g = J^T r.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void eval_edge(double (&acc)[kAcc], const double (&R)[9], double t0, double t1, double t2, double p0,
                                          double p1, double p2, double a0, double a1, double a2, double n0, double n1,
                                          double n2, const Huber &huber_a) {
  const double q0 = R[0] * p0 + R[1] * p1 + R[2] * p2, q1 = R[3] * p0 + R[4] * p1 + R[5] * p2,
               q2 = R[6] * p0 + R[7] * p1 + R[8] * p2;  // R p
  const double d0 = q0 + (t0 - a0), d1 = q1 + (t1 - a1), d2 = q2 + (t2 - a2);
  const double w0 = R[0] * n0 + R[3] * n1 + R[6] * n2, w1 = R[1] * n0 + R[4] * n1 + R[7] * n2,
               w2 = R[2] * n0 + R[5] * n1 + R[8] * n2;  // R^T n
  const double r0 = n1 * d2 - n2 * d1, r1 = n2 * d0 - n0 * d2, r2 = n0 * d1 - n1 * d0;  // n x d  (lidar_factor.cc:12)
  const double sc = huber_scale(r0 * r0 + r1 * r1 + r2 * r2, huber_a, acc[27]);
  // J_theta = (w.p) R - (R p) w^T, scaled by sc
  const double s = (w0 * p0 + w1 * p1 + w2 * p2) * sc;
  const double u0 = q0 * sc, u1 = q1 * sc, u2 = q2 * sc;
  const double m0 = n0 * sc, m1 = n1 * sc, m2 = n2 * sc;
  double J[6];
  // rows of J = [ [n]x | J_theta ] follow r0, r1, r2  (lidar_factor.cc:18-19)
  J[0] = 0.0; J[1] = -m2; J[2] = m1;
  J[3] = s * R[0] - u0 * w0; J[4] = s * R[1] - u0 * w1; J[5] = s * R[2] - u0 * w2;
  acc_row(acc, J, r0 * sc);
  J[0] = m2; J[1] = 0.0; J[2] = -m0;
  J[3] = s * R[3] - u1 * w0; J[4] = s * R[4] - u1 * w1; J[5] = s * R[5] - u1 * w2;
  acc_row(acc, J, r1 * sc);
  J[0] = -m1; J[1] = m0; J[2] = 0.0;
  J[3] = s * R[6] - u2 * w0; J[4] = s * R[7] - u2 * w1; J[5] = s * R[8] - u2 * w2;
  acc_row(acc, J, r2 * sc);
}

// nc = n . c (the plane offset along its normal)
__device__ __forceinline__ void eval_plane(double (&acc)[kAcc], const double (&R)[9], double t0, double t1, double t2, double p0,
                                           double p1, double p2, double n0, double n1, double n2, double nc, const Huber &huber_a) {
  const double w0 = R[0] * n0 + R[3] * n1 + R[6] * n2, w1 = R[1] * n0 + R[4] * n1 + R[7] * n2,
               w2 = R[2] * n0 + R[5] * n1 + R[8] * n2;  // R^T n
  const double r = (w0 * p0 + w1 * p1 + w2 * p2) + ((n0 * t0 + n1 * t1 + n2 * t2) - nc);  // lidar_factor.cc:32
  const double sc = huber_scale(r * r, huber_a, acc[27]);
  double J[6];
  J[0] = n0 * sc; J[1] = n1 * sc; J[2] = n2 * sc;  // lidar_factor.cc:38-39
  J[3] = (p1 * w2 - p2 * w1) * sc; J[4] = (p2 * w0 - p0 * w2) * sc; J[5] = (p0 * w1 - p1 * w0) * sc;
  acc_row(acc, J, r * sc);
}

// Sweep this thread's share of the correspondences at `pose`: edge factors over the corner
// queries, then plane factors over the surf queries.
// p*: query points (raw, fp32 -- quirk Q4: factors use the untransformed point); corr*: 6 doubles
// per query [a_or_c(3), n(3)], n = 0 where no factor exists.
__device__ __forceinline__ void sweep(double (&acc)[kAcc], const float4 *__restrict__ pe, const double *__restrict__ corr_e,
                                      uint32_t n_e, const float4 *__restrict__ pp, const double *__restrict__ corr_p,
                                      uint32_t n_p, const double *pose, double huber_a_, uint32_t tid, uint32_t nthreads,
                                      int &cnt_edge, int &cnt_plane) {
  const Huber huber_a{huber_a_, huber_a_ * huber_a_, sqrt(huber_a_)};
#pragma unroll
  for (int k = 0; k < kAcc; ++k) acc[k] = 0.0;
  double R[9];
  quat_to_R(pose + 3, R);
  const double t0 = pose[0], t1 = pose[1], t2 = pose[2];
  cnt_edge = 0;
  cnt_plane = 0;
  for (uint32_t i = tid; i < n_e; i += nthreads) {
    const double2 *cp = reinterpret_cast<const double2 *>(corr_e + (size_t)i * 6);
    const double2 c0 = cp[0], c1 = cp[1], c2 = cp[2];
    if (c1.y == 0.0 && c2.x == 0.0 && c2.y == 0.0) continue;  // no factor for this query
    const float4 pf = pe[i];
    ++cnt_edge;
    eval_edge(acc, R, t0, t1, t2, pf.x, pf.y, pf.z, c0.x, c0.y, c1.x, c1.y, c2.x, c2.y, huber_a);
  }
  for (uint32_t i = tid; i < n_p; i += nthreads) {
    const double2 *cp = reinterpret_cast<const double2 *>(corr_p + (size_t)i * 6);
    const double2 c0 = cp[0], c1 = cp[1], c2 = cp[2];
    if (c1.y == 0.0 && c2.x == 0.0 && c2.y == 0.0) continue;
    const float4 pf = pp[i];
    ++cnt_plane;
    eval_plane(acc, R, t0, t1, t2, pf.x, pf.y, pf.z, c1.y, c2.x, c2.y, c1.y * c0.x + c2.x * c0.y + c2.y * c1.x, huber_a);
  }
}

// Block reduction: warp shuffle tree, then warps combined in index order (deterministic).
// one butterfly level of the warp reduce-scatter: lanes whose `bit` is clear keep the lower half of v[0 .. 2 HALF)
// and receive the partner's lower half, the others the upper half; HALF sums remain
template <int HALF>
__device__ __forceinline__ void reduce_scatter_level(double (&v)[32], uint32_t lane, uint32_t bit) {
  const bool upper = (lane & bit) != 0;
#pragma unroll
  for (int k = 0; k < HALF; ++k) {
    const double keep = upper ? v[HALF + k] : v[k], send = upper ? v[k] : v[HALF + k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
  }
}

template <class SH>
__device__ __forceinline__ void block_reduce(double (&acc)[kAcc], SH &sh) {
  // warp level: a reduce-scatter (31 shuffles of a double instead of 28 x 5) that leaves the warp total of
  // accumulator l in lane l; every level adds in a fixed order, so the sums are reproducible
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = k < kAcc ? acc[k] : 0.0;
  reduce_scatter_level<16>(v, lane, 16);
  reduce_scatter_level<8>(v, lane, 8);
  reduce_scatter_level<4>(v, lane, 4);
  reduce_scatter_level<2>(v, lane, 2);
  reduce_scatter_level<1>(v, lane, 1);
  if (lane < kAcc) sh.red[warp][lane] = v[0];
  __syncthreads();
  if (threadIdx.x < kAcc) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < SH::kWarps; ++w) v += sh.red[w][threadIdx.x];
    sh.cand[threadIdx.x] = v;
  }
  __syncthreads();
}


__device__ inline double norm7(const double *x) {
  double s = 0;
  for (int i = 0; i < 7; ++i) s += x[i] * x[i];
  return sqrt(s);
}

// trust_region_minimizer.cc EvaluateGradientAndJacobian: max-norm of x - Plus(x, -g)
__device__ inline double gradient_max_norm(const double *x, const double *g) {
  double ng[6], xp[7], m = 0;
  for (int i = 0; i < 6; ++i) ng[i] = -g[i];
  pose_plus(x, ng, xp);
  for (int i = 0; i < 7; ++i) m = fmax(m, fabs(x[i] - xp[i]));
  return m;
}

// Thread-0 control: top of the Ceres loop up to the candidate point.  Sets sh.done or sh.xc.
template <class SH>
__device__ void lm_prepare_step(SH &sh, const KParams &kp, msfl_lm_log *log) {
  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (sh.iteration >= kp.max_it) { sh.done = 1; sh.termination = 0; return; }
    if (kp.early_exit && sh.step_successful && gradient_max_norm(sh.x, sh.g) <= kp.gtol) { sh.done = 1; sh.termination = 3; return; }
    if (sh.radius <= kp.min_radius) { sh.done = 1; sh.termination = 4; return; }
    sh.iteration++;
    msfl_lm_iter *L = log ? &log->it[sh.iteration - 1] : nullptr;
    if (log) log->n_attempts = sh.iteration;
    if (L) { L->cost = sh.cost; L->cost_candidate = sh.cost; L->model_change = 0; L->rho = 0; L->radius = sh.radius; L->valid = 0; L->accepted = 0; }
    // everything below lives in registers: one packed upper triangle (tri6) factored in place, every loop unrolled
    double A[21], nb[6], y[6], S[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) S[u] = sh.S[u];
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      nb[u] = -(S[u] * sh.g[u]);
#pragma unroll
      for (int v = u; v < 6; ++v) A[tri6(u, v)] = S[u] * sh.H[tri6(u, v)] * S[v];  // H' = S H S
    }
    if (!sh.reuse) {
#pragma unroll
      for (int k = 0; k < 6; ++k) sh.diag[k] = fmin(fmax(A[tri6(k, k)], kp.min_diag), kp.max_diag);
    }
    {
      const double radius = sh.radius;
#pragma unroll
      for (int k = 0; k < 6; ++k) A[tri6(k, k)] += sh.diag[k] / radius;
    }
    const bool ok = chol_solve6_packed(A, nb, y);
    sh.reuse = 1;  // LevenbergMarquardtStrategy::ComputeStep
    double model = 0;
    double delta[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) delta[k] = y[k] * S[k];
    if (ok) {
      // model_cost_change = -(y.g' + y^T H' y / 2) with g' = S g, H' = S H S and delta = S y  ==  -(delta.g + delta^T H delta / 2):
      // evaluated on the unscaled H, g still in shared memory, so no second copy of the matrix is kept in registers
      double dg = 0, dHd = 0;
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        dg += delta[u] * sh.g[u];
        double t = 0;
#pragma unroll
        for (int v = 0; v < 6; ++v) t += sh.H[u <= v ? tri6(u, v) : tri6(v, u)] * delta[v];
        dHd += delta[u] * t;
      }
      model = -(dg + 0.5 * dHd);
    }
    sh.step_successful = 0;
    if (!ok || !(model > 0.0)) {  // HandleInvalidStep
      if (L) L->model_change = model;
      if (++sh.n_invalid >= kp.max_invalid) { sh.done = 1; sh.termination = 5; return; }
      sh.radius *= 0.5;  // StepIsInvalid
      sh.reuse = 0;
      continue;
    }
    sh.n_invalid = 0;
    sh.model = model;
    double x0[7], xc[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) x0[k] = sh.x[k];
    pose_plus(x0, delta, xc);
#pragma unroll
    for (int k = 0; k < 7; ++k) sh.xc[k] = xc[k];
    if (L) { L->valid = 1; L->model_change = model; }
    return;
  }
}

// Thread-0 control after the candidate sweep: tolerance tests, step quality, accept / reject.
template <class SH>
__device__ void lm_finish_step(SH &sh, const KParams &kp, msfl_lm_log *log) {
  msfl_lm_iter *L = log ? &log->it[sh.iteration - 1] : nullptr;
  const double cost_c = sh.cand[27];
  if (L) L->cost_candidate = cost_c;
  if (kp.early_exit) {
    double sn = 0;
    for (int i = 0; i < 7; ++i) sn += (sh.x[i] - sh.xc[i]) * (sh.x[i] - sh.xc[i]);
    sn = sqrt(sn);
    if (sn <= kp.ptol * (sh.x_norm + kp.ptol)) { sh.done = 1; sh.termination = 1; return; }
    if (fabs(sh.cost - cost_c) <= kp.ftol * sh.cost) { sh.done = 1; sh.termination = 2; return; }
  }
  const double rho = (sh.cost - cost_c) / sh.model;
  if (L) L->rho = rho;
  if (rho > kp.min_rel_decrease) {
    for (int i = 0; i < 7; ++i) sh.x[i] = sh.xc[i];
    sh.x_norm = norm7(sh.x);
    for (int k = 0; k < 21; ++k) sh.H[k] = sh.cand[k];
    for (int k = 0; k < 6; ++k) sh.g[k] = sh.cand[21 + k];
    sh.cost = cost_c;
    sh.step_successful = 1;
    if (L) L->accepted = 1;
    const double t = 2.0 * rho - 1.0;
    sh.radius = fmin(kp.max_radius, sh.radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
    sh.nu = 2.0;
    sh.reuse = 0;
  } else {
    sh.radius = sh.radius / sh.nu;
    sh.nu *= 2.0;
    sh.reuse = 1;
  }
}


}  // namespace msfl
