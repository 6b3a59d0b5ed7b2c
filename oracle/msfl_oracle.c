/*
 * msfl_oracle.c -- CPU ORACLE (test infrastructure; see msfl_oracle.h header comment).
 * Checked bit for bit against the reference's own compiled sources (oracle/_ref, ref_shim.cc); the third-party arithmetic
 * underneath (FLANN, Eigen decompositions, the Ceres loop, PCL VoxelGrid) is restated, see msfl_oracle.h.
 *
 * Compile with -ffp-contract=off so that fp32 distance arithmetic matches a generic x86-64
 * build of FLANN/PCL (no FMA contraction).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#include "msfl_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------ */
/* parameters                                                                                 */
/* ------------------------------------------------------------------------------------------ */
void msflo_default_params(msflo_params *p) {
  p->min_range = 0.3;
  p->scan_period = 0.1;
  p->curvature_thresh = 0.1;
  p->neighbor_gap_sq = 0.05;
  p->n_sectors = 6;
  p->n_sharp = 2;
  p->n_less_sharp = 20;
  p->n_flat = 4;
  p->dist_sq_thresh = 25.0;
  p->nearby_scan = 2.5;
  p->min_correspondences = 10;
  p->knn_max_sq = 1.0;
  p->line_eig_ratio = 3.0;
  p->line_half_len = 0.1;
  p->plane_tol = 0.2;
  p->num_outer = 2;
  p->max_num_iterations = 6;
  p->huber_a = 0.1;
  p->initial_radius = 1e4;
  p->max_radius = 1e16;
  p->min_radius = 1e-32;
  p->min_relative_decrease = 1e-3;
  p->min_lm_diagonal = 1e-6;
  p->max_lm_diagonal = 1e32;
  p->function_tolerance = 1e-6;
  p->gradient_tolerance = 1e-10;
  p->parameter_tolerance = 1e-8;
  p->max_consecutive_invalid_steps = 5;
  p->early_exit = 1;
}

/* ------------------------------------------------------------------------------------------ */
/* quaternion / pose helpers.  pose = [tx ty tz qx qy qz qw]  (rigid_transform.h:59-64)        */
/* ------------------------------------------------------------------------------------------ */

/* Eigen::Quaternion * Vector3 (_transformVector): uv = 2 (qv x v); v + w uv + qv x uv. */
static void quat_rotate(const double q[4] /* x y z w */, const double v[3], double out[3]) {
  double uv0 = q[1] * v[2] - q[2] * v[1];
  double uv1 = q[2] * v[0] - q[0] * v[2];
  double uv2 = q[0] * v[1] - q[1] * v[0];
  uv0 += uv0; uv1 += uv1; uv2 += uv2;
  double c0 = q[1] * uv2 - q[2] * uv1;
  double c1 = q[2] * uv0 - q[0] * uv2;
  double c2 = q[0] * uv1 - q[1] * uv0;
  out[0] = v[0] + q[3] * uv0 + c0;
  out[1] = v[1] + q[3] * uv1 + c1;
  out[2] = v[2] + q[3] * uv2 + c2;
}

/* Eigen::Quaternion::toRotationMatrix, row-major R[9]. */
static void quat_to_R(const double q[4], double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

/* Utility::deltaQ (utility.h:8-31) then PoseLocalParameterization::Plus
 * (pose_local_parameterization.cc:6-21): p += dp; q = (q * dq).normalized(). */
void msflo_pose_plus(const double x[7], const double delta[6], double out[7]) {
  const double kAngleEpisode = 1e-6;
  const double vx = delta[3], vy = delta[4], vz = delta[5];
  double theta = sqrt(vx * vx + vy * vy + vz * vz);
  double half_theta = 0.5 * theta;
  double imag, real = cos(half_theta);
  if (theta < kAngleEpisode) {
    double t2 = theta * theta, t4 = t2 * t2;
    imag = 0.5 - (1 / 48.) * t2 + (1 / 3840.) * t4;
  } else {
    imag = sin(half_theta) / theta;
  }
  const double bx = imag * vx, by = imag * vy, bz = imag * vz, bw = real;
  const double ax = x[3], ay = x[4], az = x[5], aw = x[6];
  double w = aw * bw - ax * bx - ay * by - az * bz;
  double qx = aw * bx + ax * bw + ay * bz - az * by;
  double qy = aw * by + ay * bw + az * bx - ax * bz;
  double qz = aw * bz + az * bw + ax * by - ay * bx;
  double n = sqrt(qx * qx + qy * qy + qz * qz + w * w);
  out[0] = x[0] + delta[0];
  out[1] = x[1] + delta[1];
  out[2] = x[2] + delta[2];
  if (n > 0) { qx /= n; qy /= n; qz /= n; w /= n; }
  out[3] = qx; out[4] = qy; out[5] = qz; out[6] = w;
}

/* TransformPoint (rigid_transform.h:132-138): float -> double, q*p+t, -> float.
 * Also TransformToStart with s = 1 (odometry_scan_matcher.cc:21-33; slerp(1, q) == +-q). */
void msflo_transform_point_f(const double pose[7], const float in[3], float out[3]) {
  double v[3] = {(double)in[0], (double)in[1], (double)in[2]}, r[3];
  quat_rotate(pose + 3, v, r);
  out[0] = (float)(r[0] + pose[0]);
  out[1] = (float)(r[1] + pose[1]);
  out[2] = (float)(r[2] + pose[2]);
}

/* ------------------------------------------------------------------------------------------ */
/* a-8 factors (lidar_factor.cc:7-44)                                                          */
/* ------------------------------------------------------------------------------------------ */
static void skew(const double v[3], double S[9]) { /* Utility::skewSymmetric utility.h:34-41 */
  S[0] = 0;     S[1] = -v[2]; S[2] = v[1];
  S[3] = v[2];  S[4] = 0;     S[5] = -v[0];
  S[6] = -v[1]; S[7] = v[0];  S[8] = 0;
}
static void mat3_mul(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      C[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
}

/* LidarEdgeFactorSE3::Evaluate (lidar_factor.cc:7-24): r = N x (Q p + P - C);
 * J[:,0:3] = [N]x ; J[:,3:6] = -[N]x (R [p]x) ; J[:,6] = 0. */
void msflo_edge_factor(const double pose[7], const double p[3], const double a[3], const double n[3],
                       double r[3], double J[21]) {
  double x[3];
  quat_rotate(pose + 3, p, x);
  double d[3] = {x[0] + pose[0] - a[0], x[1] + pose[1] - a[1], x[2] + pose[2] - a[2]};
  r[0] = n[1] * d[2] - n[2] * d[1];
  r[1] = n[2] * d[0] - n[0] * d[2];
  r[2] = n[0] * d[1] - n[1] * d[0];
  if (J) {
    double R[9], Sp[9], Sn[9], RSp[9], M[9];
    quat_to_R(pose + 3, R);
    skew(p, Sp);
    skew(n, Sn);
    mat3_mul(R, Sp, RSp);
    mat3_mul(Sn, RSp, M);
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) {
        J[i * 7 + j] = Sn[i * 3 + j];
        J[i * 7 + 3 + j] = -M[i * 3 + j];
      }
      J[i * 7 + 6] = 0;
    }
  }
}

/* LidarPlaneFactorSE3::Evaluate (lidar_factor.cc:26-44): r = N . (Q p + P - C);
 * J[0:3] = N^T ; J[3:6] = -N^T (R [p]x). */
void msflo_plane_factor(const double pose[7], const double p[3], const double c[3], const double n[3],
                        double r[1], double J[7]) {
  double x[3];
  quat_rotate(pose + 3, p, x);
  double d[3] = {x[0] + pose[0] - c[0], x[1] + pose[1] - c[1], x[2] + pose[2] - c[2]};
  r[0] = n[0] * d[0] + n[1] * d[1] + n[2] * d[2];
  if (J) {
    double R[9], Sp[9], RSp[9];
    quat_to_R(pose + 3, R);
    skew(p, Sp);
    mat3_mul(R, Sp, RSp);
    for (int j = 0; j < 3; j++) {
      J[j] = n[j];
      J[3 + j] = -(n[0] * RSp[0 * 3 + j] + n[1] * RSp[1 * 3 + j] + n[2] * RSp[2 * 3 + j]);
    }
    J[6] = 0;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Ceres residual-block evaluation with HuberLoss + Corrector, reduced to tangent H, g.        */
/* loss_function.cc HuberLoss::Evaluate; corrector.cc (rho'' <= 0 => plain sqrt(rho') scale);  */
/* residual_block.cc: cost = 0.5 rho(s).  Tangent J = global J[:,0:6]                          */
/* (pose_local_parameterization.cc:23-27, identity-top 7x6).                                   */
/* ------------------------------------------------------------------------------------------ */
void msflo_accumulate(const msflo_params *P, const double *corr, int n_corr, const double pose[7],
                      double *cost_out, double H[36], double g[6]) {
  const double a = P->huber_a, b = a * a;
  double cost = 0;
  if (H) memset(H, 0, 36 * sizeof(double));
  if (g) memset(g, 0, 6 * sizeof(double));
  for (int i = 0; i < n_corr; i++) {
    const double *c = corr + (size_t)i * MSFLO_CORR_STRIDE;
    const int type = (int)c[0];
    double r[3], J[21];
    int nres;
    if (type == 0) {
      msflo_edge_factor(pose, c + 1, c + 4, c + 7, r, (H || g) ? J : 0);
      nres = 3;
    } else {
      msflo_plane_factor(pose, c + 1, c + 4, c + 7, r, (H || g) ? J : 0);
      nres = 1;
    }
    double s = 0;
    for (int k = 0; k < nres; k++) s += r[k] * r[k];
    double rho0, rho1;
    if (s > b) {
      const double rr = sqrt(s);
      rho0 = 2.0 * a * rr - b;
      rho1 = fmax(DBL_MIN, a / rr);
    } else {
      rho0 = s;
      rho1 = 1.0;
    }
    cost += 0.5 * rho0;
    if (H || g) {
      const double sc = sqrt(rho1);
      for (int k = 0; k < nres; k++) {
        const double rk = r[k] * sc;
        double Jk[6];
        for (int j = 0; j < 6; j++) Jk[j] = J[k * 7 + j] * sc;
        if (g)
          for (int j = 0; j < 6; j++) g[j] += Jk[j] * rk;
        if (H)
          for (int u = 0; u < 6; u++)
            for (int v = 0; v < 6; v++) H[u * 6 + v] += Jk[u] * Jk[v];
      }
    }
  }
  *cost_out = cost;
}

/* 6x6 Cholesky solve A y = b; returns 0 ok, -1 not positive definite / non-finite. */
static int chol_solve6(const double A[36], const double b[6], double y[6]) {
  double L[36];
  memset(L, 0, sizeof L);
  for (int j = 0; j < 6; j++) {
    double d = A[j * 6 + j];
    for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k];
    if (!(d > 0) || !isfinite(d)) return -1;
    d = sqrt(d);
    L[j * 6 + j] = d;
    for (int i = j + 1; i < 6; i++) {
      double s = A[i * 6 + j];
      for (int k = 0; k < j; k++) s -= L[i * 6 + k] * L[j * 6 + k];
      L[i * 6 + j] = s / d;
    }
  }
  double z[6];
  for (int i = 0; i < 6; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * 6 + k] * z[k];
    z[i] = s / L[i * 6 + i];
  }
  for (int i = 5; i >= 0; i--) {
    double s = z[i];
    for (int k = i + 1; k < 6; k++) s -= L[k * 6 + i] * y[k];
    y[i] = s / L[i * 6 + i];
  }
  for (int i = 0; i < 6; i++)
    if (!isfinite(y[i])) return -1;
  return 0;
}

static double norm7(const double x[7]) {
  double s = 0;
  for (int i = 0; i < 7; i++) s += x[i] * x[i];
  return sqrt(s);
}

/* max-norm of (x - Plus(x, -g)) : trust_region_minimizer.cc EvaluateGradientAndJacobian */
static double gradient_max_norm(const double x[7], const double g[6]) {
  double ng[6], xp[7], m = 0;
  for (int i = 0; i < 6; i++) ng[i] = -g[i];
  msflo_pose_plus(x, ng, xp);
  for (int i = 0; i < 7; i++) {
    double d = fabs(x[i] - xp[i]);
    if (d > m) m = d;
  }
  return m;
}

/*
 * a-9  ceres::Solve restatement (call sites odometry_scan_matcher.cc:270-280,
 * mapping_scan_matcher.cc:250-272).  Ceres <= 2.1 TrustRegionMinimizer::Minimize with
 * LevenbergMarquardtStrategy, one 7/6 parameter block, jacobi scaling, monotonic steps:
 *   iteration 0: evaluate cost, g, J; scale_k = 1/(1+sqrt(H_kk)) (fixed for the solve);
 *   attempt:  [if !reuse] diag_k = clamp(H'_kk, min, max);  (H' + diag/radius) y = -g';
 *             model = -(y.g' + 0.5 y'H'y); invalid if solve failed or model <= 0
 *                 -> radius /= 2, reuse = false (StepIsInvalid), counts as an iteration;
 *             delta = S y; x+ = Plus(x, delta); cost+;
 *             ||x - x+|| <= ptol (||x|| + ptol)  -> stop (x+ NOT adopted);
 *             |cost - cost+| <= ftol cost        -> stop (x+ NOT adopted);
 *             rho = (cost - cost+)/model; rho > 1e-3 -> accept: x = x+, re-evaluate,
 *                 radius = min(max_radius, radius / max(1/3, 1 - (2 rho - 1)^3)), nu = 2, reuse = false
 *             else reject: radius /= nu, nu *= 2, reuse = true;
 *   loop guard (FinalizeIterationAndCheckIfMinimizerCanContinue): iteration >= max_num_iterations,
 *             gradient max-norm <= gtol (after successful steps / iteration 0), radius <= min_radius.
 */
typedef void (*lm_eval_fn)(const msflo_params *P, const void *ctx, const double pose[7], double *cost, double H[36],
                           double g[6]);

typedef struct { const double *corr; int n; } plain_ctx;
static void eval_plain(const msflo_params *P, const void *vctx, const double pose[7], double *cost, double H[36],
                       double g[6]) {
  const plain_ctx *c = (const plain_ctx *)vctx;
  msflo_accumulate(P, c->corr, c->n, pose, cost, H, g);
}

static int lm_solve_generic(const msflo_params *P, lm_eval_fn eval, const void *ctx, int n_corr, double pose[7],
                            msflo_lm_log *log);

int msflo_lm_solve(const msflo_params *P, const double *corr, int n_corr, double pose[7], msflo_lm_log *log) {
  plain_ctx c = {corr, n_corr};
  return lm_solve_generic(P, eval_plain, &c, n_corr, pose, log);
}

/* the same loop over a caller-supplied evaluation (oracle/ref_shim.cc: the stand-in ceres::Solve evaluates the REFERENCE's
 * own cost functions through this) */
int msflo_lm_solve_cb(const msflo_params *P, msflo_eval_fn eval, const void *ctx, int n_blocks, double pose[7],
                      msflo_lm_log *log) {
  return lm_solve_generic(P, eval, ctx, n_blocks, pose, log);
}

static int lm_solve_generic(const msflo_params *P, lm_eval_fn eval, const void *ctx, int n_corr, double pose[7],
                            msflo_lm_log *log) {
  double x[7], H[36], g[6], S[6], diag[6] = {0, 0, 0, 0, 0, 0};
  double cost, radius = P->initial_radius, nu = 2.0;
  int reuse = 0, n_invalid = 0, termination = 0;
  int max_it = P->max_num_iterations;
  if (max_it > MSFLO_MAX_ATTEMPTS) max_it = MSFLO_MAX_ATTEMPTS;
  memcpy(x, pose, sizeof x);
  if (log) { memset(log, 0, sizeof *log); }
  if (n_corr <= 0) { /* no residual blocks: Ceres removes the block, nothing to do */
    if (log) log->termination = 2;
    return 0;
  }
  eval(P, ctx, x, &cost, H, g);
  for (int k = 0; k < 6; k++) S[k] = 1.0 / (1.0 + sqrt(H[k * 6 + k]));
  double x_norm = norm7(x);
  if (log) log->initial_cost = cost;
  int step_successful = 1; /* IterationZero ends with step_is_successful = true */
  int iteration = 0;
  for (;;) {
    /* FinalizeIterationAndCheckIfMinimizerCanContinue */
    if (iteration >= max_it) { termination = 0; break; }
    if (P->early_exit && step_successful && gradient_max_norm(x, g) <= P->gradient_tolerance) { termination = 3; break; }
    if (radius <= P->min_radius) { termination = 4; break; }
    iteration++;
    msflo_lm_iter *L = log ? &log->it[iteration - 1] : 0;
    if (L) { L->cost = cost; L->radius = radius; L->valid = 0; L->accepted = 0; L->rho = 0; L->model_change = 0; L->cost_candidate = cost; }
    if (log) log->n_attempts = iteration;
    /* scaled system */
    double Hs[36], gs[6], A[36], nb[6], y[6];
    for (int u = 0; u < 6; u++) {
      gs[u] = S[u] * g[u];
      for (int v = 0; v < 6; v++) Hs[u * 6 + v] = S[u] * H[u * 6 + v] * S[v];
    }
    if (!reuse)
      for (int k = 0; k < 6; k++) diag[k] = fmin(fmax(Hs[k * 6 + k], P->min_lm_diagonal), P->max_lm_diagonal);
    memcpy(A, Hs, sizeof A);
    for (int k = 0; k < 6; k++) { A[k * 6 + k] += diag[k] / radius; nb[k] = -gs[k]; }
    int ok = chol_solve6(A, nb, y) == 0;
    reuse = 1; /* LevenbergMarquardtStrategy::ComputeStep sets reuse_diagonal_ = true */
    double model = 0;
    if (ok) {
      double yg = 0, yHy = 0;
      for (int u = 0; u < 6; u++) {
        yg += y[u] * gs[u];
        double t = 0;
        for (int v = 0; v < 6; v++) t += Hs[u * 6 + v] * y[v];
        yHy += y[u] * t;
      }
      model = -(yg + 0.5 * yHy);
    }
    step_successful = 0;
    if (!ok || !(model > 0.0)) {
      /* HandleInvalidStep */
      if (L) L->model_change = model;
      if (++n_invalid >= P->max_consecutive_invalid_steps) { termination = 5; break; }
      radius *= 0.5; /* StepIsInvalid */
      reuse = 0;
      continue;
    }
    n_invalid = 0;
    double delta[6], xc[7], cost_c;
    for (int k = 0; k < 6; k++) delta[k] = y[k] * S[k];
    msflo_pose_plus(x, delta, xc);
    eval(P, ctx, xc, &cost_c, 0, 0);
    if (L) { L->valid = 1; L->model_change = model; L->cost_candidate = cost_c; }
    if (P->early_exit) {
      double sn = 0;
      for (int i = 0; i < 7; i++) sn += (x[i] - xc[i]) * (x[i] - xc[i]);
      sn = sqrt(sn);
      if (sn <= P->parameter_tolerance * (x_norm + P->parameter_tolerance)) { termination = 1; break; }
      if (fabs(cost - cost_c) <= P->function_tolerance * cost) { termination = 2; break; }
    }
    const double rho = (cost - cost_c) / model;
    if (L) L->rho = rho;
    if (rho > P->min_relative_decrease) {
      memcpy(x, xc, sizeof x);
      x_norm = norm7(x);
      eval(P, ctx, x, &cost, H, g);
      step_successful = 1;
      if (L) L->accepted = 1;
      const double t = 2.0 * rho - 1.0;
      radius = radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
      radius = fmin(P->max_radius, radius);
      nu = 2.0;
      reuse = 0;
    } else {
      radius = radius / nu;
      nu *= 2.0;
      reuse = 1;
    }
  }
  memcpy(pose, x, sizeof x);
  if (log) { log->termination = termination; log->final_cost = cost; }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* exact k-NN: kd-tree (median split, leaf 15) with FLANN L2_Simple<float> distances.          */
/* replaces pcl::KdTreeFLANN (odometry_scan_matcher.cc:57-61,84,169;                           */
/* mapping_scan_matcher.cc:66-72,125,195).  Result ascending by (d2, index).                   */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int left, right; /* children (node ids) or -1 */
  int lo, hi;      /* leaf: range in perm */
  int dim;
  float split;
} kd_node;

struct msflo_kdtree {
  const float *pts; /* n x 4 */
  int n;
  int *perm;
  kd_node *nodes;
  int n_nodes, cap_nodes;
};

static inline float l2_simple(const float *a, const float *b) {
  float r = 0.f, d;
  d = a[0] - b[0]; r += d * d;
  d = a[1] - b[1]; r += d * d;
  d = a[2] - b[2]; r += d * d;
  return r;
}

static void kd_select(const float *pts, int *perm, int lo, int hi, int k, int dim) {
  /* quickselect so that perm[k] holds the k-th smallest coord in [lo,hi) */
  while (hi - lo > 1) {
    int mid = lo + (hi - lo) / 2;
    float a = pts[4 * perm[lo] + dim], b = pts[4 * perm[mid] + dim], c = pts[4 * perm[hi - 1] + dim];
    float pv = (a < b) ? ((b < c) ? b : (a < c ? c : a)) : ((a < c) ? a : (b < c ? c : b));
    int i = lo, j = hi - 1;
    while (i <= j) {
      while (pts[4 * perm[i] + dim] < pv) i++;
      while (pts[4 * perm[j] + dim] > pv) j--;
      if (i <= j) { int t = perm[i]; perm[i] = perm[j]; perm[j] = t; i++; j--; }
    }
    if (k <= j) hi = j + 1;
    else if (k >= i) lo = i;
    else return;
  }
}

static int kd_build_rec(msflo_kdtree *t, int lo, int hi) {
  if (t->n_nodes == t->cap_nodes) {
    t->cap_nodes *= 2;
    t->nodes = (kd_node *)realloc(t->nodes, sizeof(kd_node) * t->cap_nodes);
  }
  int id = t->n_nodes++;
  kd_node nd;
  nd.left = nd.right = -1; nd.lo = lo; nd.hi = hi; nd.dim = 0; nd.split = 0;
  if (hi - lo > 15) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = lo; i < hi; i++)
      for (int d = 0; d < 3; d++) {
        float v = t->pts[4 * t->perm[i] + d];
        if (v < mn[d]) mn[d] = v;
        if (v > mx[d]) mx[d] = v;
      }
    int dim = 0;
    float ext = mx[0] - mn[0];
    if (mx[1] - mn[1] > ext) { ext = mx[1] - mn[1]; dim = 1; }
    if (mx[2] - mn[2] > ext) { ext = mx[2] - mn[2]; dim = 2; }
    int mid = lo + (hi - lo) / 2;
    kd_select(t->pts, t->perm, lo, hi, mid, dim);
    nd.dim = dim;
    nd.split = t->pts[4 * t->perm[mid] + dim];
    t->nodes[id] = nd;
    int l = kd_build_rec(t, lo, mid);
    int r = kd_build_rec(t, mid, hi);
    t->nodes[id].left = l;
    t->nodes[id].right = r;
  } else {
    t->nodes[id] = nd;
  }
  return id;
}

msflo_kdtree *msflo_kdtree_build(const float *xyzi, int n) {
  msflo_kdtree *t = (msflo_kdtree *)calloc(1, sizeof *t);
  t->pts = xyzi;
  t->n = n;
  t->perm = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
  for (int i = 0; i < n; i++) t->perm[i] = i;
  t->cap_nodes = 64;
  t->nodes = (kd_node *)malloc(sizeof(kd_node) * t->cap_nodes);
  t->n_nodes = 0;
  if (n > 0) kd_build_rec(t, 0, n);
  return t;
}

void msflo_kdtree_free(msflo_kdtree *t) {
  if (!t) return;
  free(t->perm);
  free(t->nodes);
  free(t);
}

typedef struct {
  int k, cnt;
  int *idx;
  float *d2;
} knn_set;

static inline void knn_insert(knn_set *s, float d, int id) {
  if (s->cnt == s->k) {
    float wd = s->d2[s->k - 1];
    int wi = s->idx[s->k - 1];
    if (!(d < wd || (d == wd && id < wi))) return;
  }
  int pos = (s->cnt < s->k) ? s->cnt : s->k - 1;
  while (pos > 0 && (s->d2[pos - 1] > d || (s->d2[pos - 1] == d && s->idx[pos - 1] > id))) {
    s->d2[pos] = s->d2[pos - 1];
    s->idx[pos] = s->idx[pos - 1];
    pos--;
  }
  s->d2[pos] = d;
  s->idx[pos] = id;
  if (s->cnt < s->k) s->cnt++;
}

static void kd_search_rec(const msflo_kdtree *t, int node, const float q[3], knn_set *s) {
  const kd_node *nd = &t->nodes[node];
  if (nd->left < 0) {
    for (int i = nd->lo; i < nd->hi; i++) {
      int id = t->perm[i];
      knn_insert(s, l2_simple(q, t->pts + 4 * id), id);
    }
    return;
  }
  float diff = q[nd->dim] - nd->split;
  int near = diff < 0 ? nd->left : nd->right;
  int far = diff < 0 ? nd->right : nd->left;
  kd_search_rec(t, near, q, s);
  float bound = diff * diff;
  /* fp32 rounding is monotone, so every point on the far side has l2_simple >= bound */
  if (s->cnt < s->k || bound <= s->d2[s->k - 1]) kd_search_rec(t, far, q, s);
}

int msflo_kdtree_knn(const msflo_kdtree *t, const float q[3], int k, int *idx, float *d2) {
  knn_set s;
  s.k = k; s.cnt = 0; s.idx = idx; s.d2 = d2;
  if (t->n > 0) kd_search_rec(t, 0, q, &s);
  for (int i = s.cnt; i < k; i++) { idx[i] = -1; d2[i] = INFINITY; }
  return s.cnt;
}

void msflo_knn_batch(const float *xyzi, int n, const float *q, int nq, int k, int *idx, float *d2) {
  msflo_kdtree *t = msflo_kdtree_build(xyzi, n);
  for (int i = 0; i < nq; i++) msflo_kdtree_knn(t, q + 3 * (size_t)i, k, idx + (size_t)i * k, d2 + (size_t)i * k);
  msflo_kdtree_free(t);
}

void msflo_knn_brute(const float *xyzi, int n, const float *q, int nq, int k, int *idx, float *d2) {
  for (int i = 0; i < nq; i++) {
    knn_set s;
    s.k = k; s.cnt = 0; s.idx = idx + (size_t)i * k; s.d2 = d2 + (size_t)i * k;
    for (int j = 0; j < n; j++) knn_insert(&s, l2_simple(q + 3 * (size_t)i, xyzi + 4 * (size_t)j), j);
    for (int j = s.cnt; j < k; j++) { s.idx[j] = -1; s.d2[j] = INFINITY; }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* small dense kernels standing in for Eigen                                                   */
/* ------------------------------------------------------------------------------------------ */

/* Eigen::SelfAdjointEigenSolver<Matrix3d> (mapping_scan_matcher.cc:141): cyclic Jacobi,
 * eigenvalues ascending, eigenvectors in columns (row-major storage V[r*3+c]). */
void msflo_sym_eig3(const double Ain[9], double evals[3], double V[9]) {
  double A[9];
  memcpy(A, Ain, sizeof A);
  for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    double dg = A[0] * A[0] + A[4] * A[4] + A[8] * A[8];
    if (off <= 1e-32 * dg || off == 0.0) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        double apq = A[p * 3 + q];
        if (apq == 0.0) continue;
        double app = A[p * 3 + p], aqq = A[q * 3 + q];
        double tau = (aqq - app) / (2.0 * apq);
        double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < 3; k++) { /* A <- A J */
          double akp = A[k * 3 + p], akq = A[k * 3 + q];
          A[k * 3 + p] = c * akp - s * akq;
          A[k * 3 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) { /* A <- J^T A */
          double apk = A[p * 3 + k], aqk = A[q * 3 + k];
          A[p * 3 + k] = c * apk - s * aqk;
          A[q * 3 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
          V[k * 3 + p] = c * vkp - s * vkq;
          V[k * 3 + q] = s * vkp + c * vkq;
        }
      }
  }
  double e[3] = {A[0], A[4], A[8]};
  int ord[3] = {0, 1, 2};
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2 - i; j++)
      if (e[ord[j]] > e[ord[j + 1]]) { int t = ord[j]; ord[j] = ord[j + 1]; ord[j + 1] = t; }
  double Vs[9];
  for (int c = 0; c < 3; c++) {
    evals[c] = e[ord[c]];
    for (int r = 0; r < 3; r++) Vs[r * 3 + c] = V[r * 3 + ord[c]];
    /* The SIGN of an eigenvector is implementation-defined upstream (it falls out of Eigen's tridiagonal QL iteration)
     * and the edge factor is invariant to it (n and a = c + 0.1 n flip together: r, J^T J and J^T r keep their values,
     * SURVEY.md Appendix A).  Resolved by a rule, like the std::sort ties: the largest-magnitude component is positive
     * (the first one on ties).  The CUDA eigen-solver follows the same rule. */
    double ax = fabs(Vs[0 * 3 + c]), ay = fabs(Vs[1 * 3 + c]), az = fabs(Vs[2 * 3 + c]);
    double lead = (ax >= ay && ax >= az) ? Vs[0 * 3 + c] : (ay >= az ? Vs[1 * 3 + c] : Vs[2 * 3 + c]);
    if (lead < 0.0)
      for (int r = 0; r < 3; r++) Vs[r * 3 + c] = -Vs[r * 3 + c];
  }
  memcpy(V, Vs, sizeof Vs);
}

/* Eigen colPivHouseholderQr().solve for a 5x3 system (mapping_scan_matcher.cc:210). */
void msflo_lstsq_5x3(const double Ain[15], const double bin[5], double x[3]) {
  double A[15], b[5];
  int perm[3] = {0, 1, 2};
  memcpy(A, Ain, sizeof A);
  memcpy(b, bin, sizeof b);
  double Rdiag[3] = {0, 0, 0};
  int rank = 3;
  double maxpiv = 0;
  for (int k = 0; k < 3; k++) {
    /* pivot: remaining column with the largest norm over rows k..4 */
    int best = k;
    double bn = -1;
    for (int j = k; j < 3; j++) {
      double s = 0;
      for (int i = k; i < 5; i++) s += A[i * 3 + j] * A[i * 3 + j];
      if (s > bn) { bn = s; best = j; }
    }
    if (best != k) {
      for (int i = 0; i < 5; i++) { double t = A[i * 3 + k]; A[i * 3 + k] = A[i * 3 + best]; A[i * 3 + best] = t; }
      int t = perm[k]; perm[k] = perm[best]; perm[best] = t;
    }
    double nrm = sqrt(bn);
    if (k == 0) maxpiv = nrm;
    if (!(nrm > maxpiv * 1e-14) || nrm == 0.0) { rank = k; break; }
    double alpha = (A[k * 3 + k] > 0) ? -nrm : nrm;
    double v[5] = {0, 0, 0, 0, 0};
    for (int i = k; i < 5; i++) v[i] = A[i * 3 + k];
    v[k] -= alpha;
    double vv = 0;
    for (int i = k; i < 5; i++) vv += v[i] * v[i];
    if (vv > 0) {
      for (int j = k; j < 3; j++) {
        double s = 0;
        for (int i = k; i < 5; i++) s += v[i] * A[i * 3 + j];
        s = 2.0 * s / vv;
        for (int i = k; i < 5; i++) A[i * 3 + j] -= s * v[i];
      }
      double s = 0;
      for (int i = k; i < 5; i++) s += v[i] * b[i];
      s = 2.0 * s / vv;
      for (int i = k; i < 5; i++) b[i] -= s * v[i];
    }
    Rdiag[k] = A[k * 3 + k];
  }
  double y[3] = {0, 0, 0};
  for (int i = rank - 1; i >= 0; i--) {
    double s = b[i];
    for (int j = i + 1; j < rank; j++) s -= A[i * 3 + j] * y[j];
    y[i] = s / Rdiag[i];
  }
  for (int i = 0; i < 3; i++) x[perm[i]] = y[i];
}

/* ------------------------------------------------------------------------------------------ */
/* a-6 / a-7 association against the submap (mapping_scan_matcher.cc:109-246, LiDAR-only)      */
/* ------------------------------------------------------------------------------------------ */
void msflo_associate_map(const msflo_params *P,
                         const msflo_kdtree *tree_corner, const float *map_corner,
                         const msflo_kdtree *tree_surf, const float *map_surf,
                         const float *scan_corner, int n_scan_corner,
                         const float *scan_surf, int n_scan_surf,
                         const double pose[7], double *corr, int *n_edge, int *n_plane,
                         int *knn_idx_out) {
  int ne = 0, np = 0, nc = 0;
  int idx[5];
  float d2[5];
  for (int i = 0; i < n_scan_corner; i++) { /* :109-176 */
    const float *po = scan_corner + 4 * (size_t)i;
    float sel[3];
    msflo_transform_point_f(pose, po, sel); /* :123 */
    int found = msflo_kdtree_knn(tree_corner, sel, 5, idx, d2);
    int gate = (found == 5) && ((double)d2[4] < P->knn_max_sq); /* :128 */
    if (knn_idx_out)
      for (int j = 0; j < 5; j++) knn_idx_out[(size_t)i * 5 + j] = gate ? idx[j] : -1;
    if (!gate) continue;
    double m[5][3], c[3] = {0, 0, 0};
    for (int j = 0; j < 5; j++)
      for (int d = 0; d < 3; d++) {
        m[j][d] = (double)map_corner[4 * (size_t)idx[j] + d];
        c[d] += m[j][d];
      }
    for (int d = 0; d < 3; d++) c[d] /= 5.0; /* :137 */
    double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < 5; j++) {
      double e[3] = {m[j][0] - c[0], m[j][1] - c[1], m[j][2] - c[2]};
      for (int u = 0; u < 3; u++)
        for (int v = 0; v < 3; v++) cov[u * 3 + v] += e[u] * e[v]; /* :139 */
    }
    double ev[3], V[9];
    msflo_sym_eig3(cov, ev, V); /* :141 */
    if (ev[2] > P->line_eig_ratio * ev[1]) { /* :147 */
      double u[3] = {V[0 * 3 + 2], V[1 * 3 + 2], V[2 * 3 + 2]};
      double a[3], b[3], n[3];
      for (int d = 0; d < 3; d++) {
        a[d] = P->line_half_len * u[d] + c[d];  /* :150 */
        b[d] = -P->line_half_len * u[d] + c[d]; /* :151 */
        n[d] = a[d] - b[d];
      }
      double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      if (nn > 0) for (int d = 0; d < 3; d++) n[d] /= nn; /* :168 normalized() */
      double *o = corr + (size_t)nc * MSFLO_CORR_STRIDE;
      o[0] = 0;
      for (int d = 0; d < 3; d++) { o[1 + d] = (double)po[d]; o[4 + d] = a[d]; o[7 + d] = n[d]; }
      nc++; ne++;
    }
  }
  for (int i = 0; i < n_scan_surf; i++) { /* :178-246 */
    const float *po = scan_surf + 4 * (size_t)i;
    float sel[3];
    msflo_transform_point_f(pose, po, sel); /* :193 */
    int found = msflo_kdtree_knn(tree_surf, sel, 5, idx, d2);
    int gate = (found == 5) && ((double)d2[4] < P->knn_max_sq); /* :198 */
    if (knn_idx_out)
      for (int j = 0; j < 5; j++) knn_idx_out[((size_t)n_scan_corner + i) * 5 + j] = gate ? idx[j] : -1;
    if (!gate) continue;
    double A[15], bb[5] = {-1, -1, -1, -1, -1}, nrm[3], c[3] = {0, 0, 0};
    for (int j = 0; j < 5; j++)
      for (int d = 0; d < 3; d++) {
        A[j * 3 + d] = (double)map_surf[4 * (size_t)idx[j] + d];
        c[d] += A[j * 3 + d];
      }
    msflo_lstsq_5x3(A, bb, nrm); /* :210 */
    double nn = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
    for (int d = 0; d < 3; d++) { nrm[d] /= nn; c[d] /= 5.0; } /* :211-212 */
    int valid = 1;
    for (int j = 0; j < 5; j++) { /* :214-220 */
      double dd = nrm[0] * (A[j * 3 + 0] - c[0]) + nrm[1] * (A[j * 3 + 1] - c[1]) + nrm[2] * (A[j * 3 + 2] - c[2]);
      if (!(fabs(dd) <= P->plane_tol)) { valid = 0; break; }
    }
    if (valid) {
      double *o = corr + (size_t)nc * MSFLO_CORR_STRIDE;
      o[0] = 1;
      for (int d = 0; d < 3; d++) { o[1 + d] = (double)po[d]; o[4 + d] = c[d]; o[7 + d] = nrm[d]; }
      nc++; np++;
    }
  }
  *n_edge = ne;
  *n_plane = np;
}

static int scan2map_with_trees(const msflo_params *P, const msflo_kdtree *tc, const float *map_corner,
                               const msflo_kdtree *ts, const float *map_surf,
                               const float *scan_corner, int n_scan_corner, const float *scan_surf, int n_scan_surf,
                               double pose[7], msflo_lm_log *logs, int *counts) {
  size_t cap = (size_t)(n_scan_corner + n_scan_surf) + 1;
  double *corr = (double *)malloc(sizeof(double) * MSFLO_CORR_STRIDE * cap);
  for (int it = 0; it < P->num_outer; it++) { /* :75 */
    int ne, np;
    msflo_associate_map(P, tc, map_corner, ts, map_surf, scan_corner, n_scan_corner, scan_surf, n_scan_surf,
                        pose, corr, &ne, &np, 0);
    if (counts) { counts[2 * it] = ne; counts[2 * it + 1] = np; }
    msflo_lm_solve(P, corr, ne + np, pose, logs ? &logs[it] : 0); /* :259, :271 */
  }
  free(corr);
  return 0;
}

int msflo_scan2map(const msflo_params *P,
                   const float *map_corner, int n_map_corner, const float *map_surf, int n_map_surf,
                   const float *scan_corner, int n_scan_corner, const float *scan_surf, int n_scan_surf,
                   double pose[7], msflo_lm_log *logs, int *counts) {
  /* kd-trees are rebuilt on every call, as the reference does (mapping_scan_matcher.cc:66-72) */
  msflo_kdtree *tc = msflo_kdtree_build(map_corner, n_map_corner);
  msflo_kdtree *ts = msflo_kdtree_build(map_surf, n_map_surf);
  int rc = scan2map_with_trees(P, tc, map_corner, ts, map_surf, scan_corner, n_scan_corner, scan_surf, n_scan_surf,
                               pose, logs, counts);
  msflo_kdtree_free(tc);
  msflo_kdtree_free(ts);
  return rc;
}

typedef struct {
  const msflo_params *P;
  const msflo_kdtree *tc, *ts;
  const float *map_corner, *map_surf;
  int B, tid, nthreads;
  const float *scan_corner;
  const int *corner_off;
  const float *scan_surf;
  const int *surf_off;
  double *poses;
} batch_arg;

static void *batch_worker(void *vp) {
  batch_arg *a = (batch_arg *)vp;
  for (int i = a->tid; i < a->B; i += a->nthreads) {
    scan2map_with_trees(a->P, a->tc, a->map_corner, a->ts, a->map_surf,
                        a->scan_corner + 4 * (size_t)a->corner_off[i], a->corner_off[i + 1] - a->corner_off[i],
                        a->scan_surf + 4 * (size_t)a->surf_off[i], a->surf_off[i + 1] - a->surf_off[i],
                        a->poses + 7 * (size_t)i, 0, 0);
  }
  return 0;
}

/* Batch of independent scans vs one shared submap.  The two kd-trees are built ONCE and shared
 * read-only by the workers (generous to the CPU: the reference rebuilds them every frame). */
int msflo_scan2map_batch(const msflo_params *P,
                         const float *map_corner, int n_map_corner, const float *map_surf, int n_map_surf,
                         int B, const float *scan_corner, const int *corner_off,
                         const float *scan_surf, const int *surf_off,
                         double *poses, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  msflo_kdtree *tc = msflo_kdtree_build(map_corner, n_map_corner);
  msflo_kdtree *ts = msflo_kdtree_build(map_surf, n_map_surf);
  pthread_t th[256];
  batch_arg args[256];
  for (int t = 0; t < n_threads; t++) {
    batch_arg a = {P, tc, ts, map_corner, map_surf, B, t, n_threads, scan_corner, corner_off, scan_surf, surf_off, poses};
    args[t] = a;
    if (n_threads == 1) batch_worker(&args[t]);
    else pthread_create(&th[t], 0, batch_worker, &args[t]);
  }
  if (n_threads > 1)
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], 0);
  msflo_kdtree_free(tc);
  msflo_kdtree_free(ts);
  return 0;
}


/* ------------------------------------------------------------------------------------------ */
/* 8f row 3: IMU-deskew branch (mapping_scan_matcher.cc:107-246, is_initialized == true)       */
/* ------------------------------------------------------------------------------------------ */
static void quat_mul(const double a[4], const double b[4], double o[4]) { /* Eigen product, x y z w */
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}

/* GetDeltaQP (scan_undistortion.cc:22-42): upper_bound, slerp (Eigen semantics), lerp */
int msflo_get_delta_qp(const msflo_deskew *dk, double dt, double dq[4], double dp[3]) {
  const int n = dk->n;
  if (n < 2 || !(dt <= dk->sum_dt[n - 1] && dt >= dk->sum_dt[0])) return -1; /* CHECK :26 */
  int lo = 0, hi = n; /* first index with dt < sum_dt[idx] */
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (dt < dk->sum_dt[mid]) hi = mid; else lo = mid + 1;
  }
  int idx = lo - 1;
  if (idx > n - 2) idx = n - 2; /* dt == back(): the reference reads one past the end; clamp */
  const double s = (dt - dk->sum_dt[idx]) / (dk->sum_dt[idx + 1] - dk->sum_dt[idx]);
  const double *qa = dk->delta_q + 4 * idx, *qb = qa + 4;
  const double d = qa[0] * qb[0] + qa[1] * qb[1] + qa[2] * qb[2] + qa[3] * qb[3];
  const double absd = fabs(d);
  double s0, s1;
  if (absd >= 1.0 - DBL_EPSILON) {
    s0 = 1.0 - s; s1 = s;
  } else {
    const double th = acos(absd), sn = sin(th);
    s0 = sin((1.0 - s) * th) / sn;
    s1 = sin(s * th) / sn;
  }
  if (d < 0) s1 = -s1;
  for (int k = 0; k < 4; k++) dq[k] = s0 * qa[k] + s1 * qb[k];
  const double *pa = dk->delta_p + 3 * idx, *pb = pa + 3;
  for (int k = 0; k < 3; k++) dp[k] = (1 - s) * pa[k] + s * pb[k];
  return 0;
}

/* LidarEdgeFactorDeskewSE3::Evaluate (lidar_factor.cc:46-72), pose block only */
void msflo_edge_factor_deskew(const double pose[7], const double V[3], const double p[3], const double C[3],
                              const double N[3], const double dp[3], const double dq[4], double dt, const double G[3],
                              double r[3], double J[21]) {
  double pp[3], x[3], d[3];
  quat_rotate(dq, p, pp);
  for (int k = 0; k < 3; k++) pp[k] += dp[k];
  quat_rotate(pose + 3, pp, x);
  for (int k = 0; k < 3; k++) d[k] = x[k] + V[k] * dt - 0.5 * G[k] * dt * dt + pose[k] - C[k];
  r[0] = N[1] * d[2] - N[2] * d[1];
  r[1] = N[2] * d[0] - N[0] * d[2];
  r[2] = N[0] * d[1] - N[1] * d[0];
  if (J) {
    double R[9], Sp[9], Sn[9], RSp[9], M[9];
    quat_to_R(pose + 3, R);
    skew(pp, Sp);
    skew(N, Sn);
    mat3_mul(R, Sp, RSp);
    mat3_mul(Sn, RSp, M);
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) { J[i * 7 + j] = Sn[i * 3 + j]; J[i * 7 + 3 + j] = -M[i * 3 + j]; }
      J[i * 7 + 6] = 0;
    }
  }
}

/* LidarPlaneFactorDeskewSE3::Evaluate (lidar_factor.cc:74-100), pose block only */
void msflo_plane_factor_deskew(const double pose[7], const double V[3], const double p[3], const double C[3],
                               const double N[3], const double dp[3], const double dq[4], double dt, const double G[3],
                               double r[1], double J[7]) {
  double pp[3], x[3], d[3];
  quat_rotate(dq, p, pp);
  for (int k = 0; k < 3; k++) pp[k] += dp[k];
  quat_rotate(pose + 3, pp, x);
  for (int k = 0; k < 3; k++) d[k] = x[k] + V[k] * dt - 0.5 * G[k] * dt * dt + pose[k] - C[k];
  r[0] = N[0] * d[0] + N[1] * d[1] + N[2] * d[2];
  if (J) {
    double R[9], Sp[9], RSp[9];
    quat_to_R(pose + 3, R);
    skew(pp, Sp);
    mat3_mul(R, Sp, RSp);
    for (int j = 0; j < 3; j++) {
      J[j] = N[j];
      J[3 + j] = -(N[0] * RSp[0 * 3 + j] + N[1] * RSp[1 * 3 + j] + N[2] * RSp[2 * 3 + j]);
    }
    J[6] = 0;
  }
}

/* deskew correspondence row: [type, p(3), C(3), N(3), dp(3), dq(4), dt] = 18 doubles */
#define DSK_STRIDE 18
typedef struct { const double *corr; int n; const msflo_deskew *dk; } deskew_ctx;

static void eval_deskew(const msflo_params *P, const void *vctx, const double pose[7], double *cost_out, double H[36],
                        double g[6]) {
  const deskew_ctx *c = (const deskew_ctx *)vctx;
  const double a = P->huber_a, b = a * a;
  double cost = 0;
  if (H) memset(H, 0, 36 * sizeof(double));
  if (g) memset(g, 0, 6 * sizeof(double));
  for (int i = 0; i < c->n; i++) {
    const double *e = c->corr + (size_t)i * DSK_STRIDE;
    double r[3], J[21];
    int nres;
    if ((int)e[0] == 0) {
      msflo_edge_factor_deskew(pose, c->dk->velocity, e + 1, e + 4, e + 7, e + 10, e + 13, e[17], c->dk->gravity, r, (H || g) ? J : 0);
      nres = 3;
    } else {
      msflo_plane_factor_deskew(pose, c->dk->velocity, e + 1, e + 4, e + 7, e + 10, e + 13, e[17], c->dk->gravity, r, (H || g) ? J : 0);
      nres = 1;
    }
    double s = 0;
    for (int k = 0; k < nres; k++) s += r[k] * r[k];
    double rho0, rho1;
    if (s > b) { const double rr = sqrt(s); rho0 = 2.0 * a * rr - b; rho1 = fmax(DBL_MIN, a / rr); }
    else { rho0 = s; rho1 = 1.0; }
    cost += 0.5 * rho0;
    if (H || g) {
      const double sc = sqrt(rho1);
      for (int k = 0; k < nres; k++) {
        const double rk = r[k] * sc;
        double Jk[6];
        for (int j = 0; j < 6; j++) Jk[j] = J[k * 7 + j] * sc;
        if (g) for (int j = 0; j < 6; j++) g[j] += Jk[j] * rk;
        if (H) for (int u = 0; u < 6; u++) for (int v = 0; v < 6; v++) H[u * 6 + v] += Jk[u] * Jk[v];
      }
    }
  }
  *cost_out = cost;
}

/* pointSel = TransformPoint(pose * Rigid3d{q^-1 (V dt - g dt^2/2) + dp, dq}, pointOri)  (:120, :190) */
static void deskew_transform(const double pose[7], const msflo_deskew *dk, const double dq[4], const double dp[3],
                             double dt, const float in[3], float out[3]) {
  double o[3], qc[4] = {-pose[3], -pose[4], -pose[5], pose[6]}, tr[3], tt[3], qt[4], T[7];
  for (int k = 0; k < 3; k++) o[k] = dk->velocity[k] * dt - 0.5 * dk->gravity[k] * dt * dt;
  quat_rotate(qc, o, tr);
  for (int k = 0; k < 3; k++) tr[k] += dp[k];
  quat_rotate(pose + 3, tr, tt); /* Rigid3 operator* (rigid_transform.h:105-111) */
  quat_mul(pose + 3, dq, qt);
  const double nq = sqrt(qt[0] * qt[0] + qt[1] * qt[1] + qt[2] * qt[2] + qt[3] * qt[3]);
  for (int k = 0; k < 3; k++) T[k] = tt[k] + pose[k];
  for (int k = 0; k < 4; k++) T[3 + k] = qt[k] / nq;
  msflo_transform_point_f(T, in, out);
}

int msflo_scan2map_deskew(const msflo_params *P,
                          const float *map_corner, int n_map_corner, const float *map_surf, int n_map_surf,
                          const float *scan_corner, int n_scan_corner, const float *scan_surf, int n_scan_surf,
                          const msflo_deskew *dk, double pose[7], msflo_lm_log *logs, int *counts, int *knn_idx_out) {
  msflo_kdtree *tc = msflo_kdtree_build(map_corner, n_map_corner);
  msflo_kdtree *ts = msflo_kdtree_build(map_surf, n_map_surf);
  const int nq = n_scan_corner + n_scan_surf;
  double *corr = (double *)malloc(sizeof(double) * DSK_STRIDE * ((size_t)nq + 1));
  int rc = 0;
  for (int it = 0; it < P->num_outer && rc == 0; it++) {
    int ne = 0, np = 0, nc = 0;
    for (int i = 0; i < nq; i++) {
      const int is_corner = i < n_scan_corner;
      const float *po = is_corner ? scan_corner + 4 * (size_t)i : scan_surf + 4 * (size_t)(i - n_scan_corner);
      const double dt = (double)po[3]; /* auto dt = pointOri.intensity (float) :114 */
      double dq[4], dp[3];
      if (msflo_get_delta_qp(dk, dt, dq, dp)) { rc = -1; break; }
      float sel[3];
      deskew_transform(pose, dk, dq, dp, dt, po, sel);
      int idx[5];
      float d2[5];
      const float *map = is_corner ? map_corner : map_surf;
      int found = msflo_kdtree_knn(is_corner ? tc : ts, sel, 5, idx, d2);
      int gate = (found == 5) && ((double)d2[4] < P->knn_max_sq);
      if (knn_idx_out && it == 0)
        for (int j = 0; j < 5; j++) knn_idx_out[(size_t)i * 5 + j] = gate ? idx[j] : -1;
      if (!gate) continue;
      double m[5][3], c[3] = {0, 0, 0};
      for (int j = 0; j < 5; j++)
        for (int d = 0; d < 3; d++) { m[j][d] = (double)map[4 * (size_t)idx[j] + d]; c[d] += m[j][d]; }
      for (int d = 0; d < 3; d++) c[d] /= 5.0;
      double a[3], n[3];
      int ok = 0;
      if (is_corner) {
        double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ev[3], V[9];
        for (int j = 0; j < 5; j++) {
          double e[3] = {m[j][0] - c[0], m[j][1] - c[1], m[j][2] - c[2]};
          for (int u = 0; u < 3; u++) for (int v = 0; v < 3; v++) cov[u * 3 + v] += e[u] * e[v];
        }
        msflo_sym_eig3(cov, ev, V);
        if (ev[2] > P->line_eig_ratio * ev[1]) {
          double b[3];
          for (int d = 0; d < 3; d++) {
            a[d] = P->line_half_len * V[d * 3 + 2] + c[d];
            b[d] = -P->line_half_len * V[d * 3 + 2] + c[d];
            n[d] = a[d] - b[d];
          }
          double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
          if (nn > 0) for (int d = 0; d < 3; d++) n[d] /= nn;
          ok = 1;
        }
      } else {
        double A[15], bb[5] = {-1, -1, -1, -1, -1};
        for (int j = 0; j < 5; j++) for (int d = 0; d < 3; d++) A[j * 3 + d] = m[j][d];
        msflo_lstsq_5x3(A, bb, n);
        double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        for (int d = 0; d < 3; d++) { n[d] /= nn; a[d] = c[d]; }
        ok = 1;
        for (int j = 0; j < 5; j++) {
          double dd = n[0] * (m[j][0] - c[0]) + n[1] * (m[j][1] - c[1]) + n[2] * (m[j][2] - c[2]);
          if (!(fabs(dd) <= P->plane_tol)) { ok = 0; break; }
        }
      }
      if (!ok) continue;
      double *o = corr + (size_t)nc * DSK_STRIDE;
      o[0] = is_corner ? 0 : 1;
      for (int d = 0; d < 3; d++) { o[1 + d] = (double)po[d]; o[4 + d] = a[d]; o[7 + d] = n[d]; o[10 + d] = dp[d]; }
      for (int d = 0; d < 4; d++) o[13 + d] = dq[d];
      o[17] = dt;
      nc++;
      if (is_corner) ne++; else np++;
    }
    if (rc) break;
    if (counts) { counts[2 * it] = ne; counts[2 * it + 1] = np; }
    deskew_ctx ctx = {corr, nc, dk};
    lm_solve_generic(P, eval_deskew, &ctx, nc, pose, logs ? &logs[it] : 0);
  }
  free(corr);
  msflo_kdtree_free(tc);
  msflo_kdtree_free(ts);
  return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* a-5 MatchScan2Scan (odometry_scan_matcher.cc:43-285), loops restated literally              */
/* ------------------------------------------------------------------------------------------ */
static inline double sqdist_f(const float *a, const float *s) {
  /* :102-108 -- all operands float, expression evaluated in float, widened on assignment */
  float r = (a[0] - s[0]) * (a[0] - s[0]) + (a[1] - s[1]) * (a[1] - s[1]) + (a[2] - s[2]) * (a[2] - s[2]);
  return (double)r;
}

int msflo_scan2scan(const msflo_params *P,
                    const float *last_corner, const uint16_t *last_corner_ring, int n_last_corner,
                    const float *last_surf, const uint16_t *last_surf_ring, int n_last_surf,
                    const float *curr_sharp, int n_curr_sharp,
                    const float *curr_flat, int n_curr_flat,
                    double pose[7], msflo_lm_log *logs, int *counts, int *assoc_out) {
  msflo_kdtree *tc = msflo_kdtree_build(last_corner, n_last_corner); /* :57-61 */
  msflo_kdtree *ts = msflo_kdtree_build(last_surf, n_last_surf);
  size_t cap = (size_t)(n_curr_sharp + n_curr_flat) + 1;
  double *corr = (double *)malloc(sizeof(double) * MSFLO_CORR_STRIDE * cap);
  int rc = 0;
  for (int oc = 0; oc < P->num_outer; oc++) { /* :64 */
    int ne = 0, np = 0, nc = 0;
    for (int i = 0; i < n_curr_sharp; i++) { /* :81-163 */
      float sel[3];
      int ind;
      float d2;
      msflo_transform_point_f(pose, curr_sharp + 4 * (size_t)i, sel); /* TransformToStart */
      int found = msflo_kdtree_knn(tc, sel, 1, &ind, &d2);
      int closest = -1, min2 = -1;
      if (found == 1 && (double)d2 < P->dist_sq_thresh) { /* :87 */
        closest = ind;
        int id = last_corner_ring[closest];
        double min_d2 = P->dist_sq_thresh;
        for (int j = closest + 1; j < n_last_corner; ++j) { /* :93-115 */
          if ((int)last_corner_ring[j] <= id) continue;
          if ((double)last_corner_ring[j] > id + P->nearby_scan) break;
          double d = sqdist_f(last_corner + 4 * (size_t)j, sel);
          if (d < min_d2) { min_d2 = d; min2 = j; }
        }
        for (int j = closest - 1; j >= 0; --j) { /* :118-140 */
          if ((int)last_corner_ring[j] >= id) continue;
          if ((double)last_corner_ring[j] < id - P->nearby_scan) break;
          double d = sqdist_f(last_corner + 4 * (size_t)j, sel);
          if (d < min_d2) { min_d2 = d; min2 = j; }
        }
      }
      if (assoc_out && oc == 0) { assoc_out[2 * i] = closest; assoc_out[2 * i + 1] = min2; }
      if (min2 >= 0) { /* :143-162 */
        const float *pa = last_corner + 4 * (size_t)closest, *pb = last_corner + 4 * (size_t)min2;
        double *o = corr + (size_t)nc * MSFLO_CORR_STRIDE;
        double n[3];
        o[0] = 0;
        for (int d = 0; d < 3; d++) {
          o[1 + d] = (double)curr_sharp[4 * (size_t)i + d];
          o[4 + d] = (double)pa[d];
          n[d] = (double)pa[d] - (double)pb[d];
        }
        double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        for (int d = 0; d < 3; d++) o[7 + d] = (nn > 0) ? n[d] / nn : n[d];
        nc++; ne++;
      }
    }
    for (int i = 0; i < n_curr_flat; i++) { /* :166-258 */
      float sel[3];
      int ind;
      float d2;
      msflo_transform_point_f(pose, curr_flat + 4 * (size_t)i, sel);
      int found = msflo_kdtree_knn(ts, sel, 1, &ind, &d2);
      int closest = -1, min2 = -1, min3 = -1;
      if (found == 1 && (double)d2 < P->dist_sq_thresh) { /* :173 */
        closest = ind;
        int id = last_surf_ring[closest];
        double m2 = P->dist_sq_thresh, m3 = P->dist_sq_thresh;
        for (int j = closest + 1; j < n_last_surf; ++j) { /* :183-207 */
          if ((double)last_surf_ring[j] > id + P->nearby_scan) break;
          double d = sqdist_f(last_surf + 4 * (size_t)j, sel);
          if ((int)last_surf_ring[j] <= id && d < m2) { m2 = d; min2 = j; }
          else if ((int)last_surf_ring[j] > id && d < m3) { m3 = d; min3 = j; }
        }
        for (int j = closest - 1; j >= 0; --j) { /* :210-232 */
          if ((double)last_surf_ring[j] < id - P->nearby_scan) break;
          double d = sqdist_f(last_surf + 4 * (size_t)j, sel);
          if ((int)last_surf_ring[j] >= id && d < m2) { m2 = d; min2 = j; }
          else if ((int)last_surf_ring[j] < id && d < m3) { m3 = d; min3 = j; }
        }
      }
      if (assoc_out && oc == 0) {
        int *o = assoc_out + 2 * n_curr_sharp + 3 * i;
        o[0] = closest; o[1] = min2; o[2] = min3;
      }
      if (closest >= 0 && min2 >= 0 && min3 >= 0) { /* :234-256; LidarPlaneFactorSE3 4-arg ctor lidar_factor.h:70-78 */
        double a[3], b[3], c[3];
        for (int d = 0; d < 3; d++) {
          a[d] = (double)last_surf[4 * (size_t)closest + d];
          b[d] = (double)last_surf[4 * (size_t)min2 + d];
          c[d] = (double)last_surf[4 * (size_t)min3 + d];
        }
        double ab[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]}, ac[3] = {a[0] - c[0], a[1] - c[1], a[2] - c[2]};
        double n[3] = {ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0]};
        double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        double *o = corr + (size_t)nc * MSFLO_CORR_STRIDE;
        o[0] = 1;
        for (int d = 0; d < 3; d++) {
          o[1 + d] = (double)curr_flat[4 * (size_t)i + d];
          o[4 + d] = (a[d] + b[d] + c[d]) / 3;
          o[7 + d] = (nn > 0) ? n[d] / nn : n[d];
        }
        nc++; np++;
      }
    }
    if (counts) { counts[2 * oc] = ne; counts[2 * oc + 1] = np; }
    if (ne + np < P->min_correspondences) { rc = 1; break; } /* :262-267 */
    msflo_lm_solve(P, corr, nc, pose, logs ? &logs[oc] : 0);  /* :274, :280 */
  }
  free(corr);
  msflo_kdtree_free(tc);
  msflo_kdtree_free(ts);
  return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* a-1..a-4 feature extraction (msf_loam_node.cc:86-371)                                       */
/* ------------------------------------------------------------------------------------------ */
#define MSFLO_MAX_RINGS 128 /* kMaxScanNum :79 */

typedef struct { float c; int i; } curv_idx;
static int cmp_curv(const void *pa, const void *pb) {
  const curv_idx *a = (const curv_idx *)pa, *b = (const curv_idx *)pb;
  if (a->c < b->c) return -1;
  if (a->c > b->c) return 1;
  return (a->i > b->i) - (a->i < b->i); /* deterministic tie-break (std::sort is unstable: impl-defined) */
}

static inline float gap_sq(const float *a, const float *b) { /* Vector3f squaredNorm, fp32 :291-293 */
  float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  return dx * dx + dy * dy + dz * dz;
}

int msflo_extract_features(const msflo_params *P, const float *xyzi, const uint16_t *ring, int n,
                           const double T_ext[7],
                           float *full, uint16_t *full_ring, int *n_full_out,
                           float *curv, int *label,
                           int *idx_sharp, int *n_sharp_out, int *idx_less_sharp, int *n_less_sharp_out,
                           int *idx_flat, int *n_flat_out, int *idx_less_flat, int *n_less_flat_out) {
  *n_full_out = *n_sharp_out = *n_less_sharp_out = *n_flat_out = *n_less_flat_out = 0;
  if (n <= 0) return -2;
  /* RemoveInvalidPointsFromCloud :86-111 -- float norm vs double min_range; non-finite dropped */
  int *valid = (int *)malloc(sizeof(int) * n);
  int nv = 0;
  for (int i = 0; i < n; i++) {
    const float *p = xyzi + 4 * (size_t)i;
    float nr = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    if ((double)nr < P->min_range || !isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
    if (ring[i] >= MSFLO_MAX_RINGS) { free(valid); return -1; } /* CHECK_LT :136 */
    valid[nv++] = i;
  }
  if (nv == 0) { free(valid); return -2; }
  /* ComputeRelaTimeForEachPoint :128-156 */
  int cnt[MSFLO_MAX_RINGS], off[MSFLO_MAX_RINGS + 1], cur[MSFLO_MAX_RINGS];
  double last_angle[MSFLO_MAX_RINGS];
  memset(cnt, 0, sizeof cnt);
  for (int k = 0; k < nv; k++) cnt[ring[valid[k]]]++;
  off[0] = 0;
  for (int r = 0; r < MSFLO_MAX_RINGS; r++) { off[r + 1] = off[r] + cnt[r]; cur[r] = off[r]; last_angle[r] = -1; }
  const float *first = xyzi + 4 * (size_t)valid[0];
  /* `atan2(point.y, point.x)` at :131 / :139 is an UNQUALIFIED call with two float arguments.  The translation unit
   * includes ros/ros.h, tf and PCL headers, which pull in <math.h>; under libstdc++ that header is the C++ wrapper
   * that does `using std::atan2;`, so overload resolution picks float atan2(float, float) (with <cmath> alone the
   * arguments would be promoted to double -- both cases are demonstrated by oracle/ref_harness/atan2_overload.cc).
   * The angle is therefore a FLOAT, negated, then widened.  libm's atan2f is not correctly rounded (glibc 2.39: 1 ulp
   * off in 16 % of the cases), so this quantity is defined only to 1 float ulp (2.4e-7 rad = 3.8e-9 s) across libms. */
  const double start_ori = (double)(-atan2f(first[1], first[0]));
  for (int k = 0; k < nv; k++) {
    const int i = valid[k];
    const float *p = xyzi + 4 * (size_t)i;
    const int r = ring[i];
    double ori = (double)(-atan2f(p[1], p[0]));
    double rel = fmod(ori - start_ori + 2 * M_PI, 2 * M_PI);
    if (rel < last_angle[r]) rel += 2 * M_PI;
    last_angle[r] = rel;
    double rela_time = rel / (2 * M_PI) * P->scan_period;
    int dst = cur[r]++;
    full[4 * (size_t)dst + 0] = p[0];
    full[4 * (size_t)dst + 1] = p[1];
    full[4 * (size_t)dst + 2] = p[2];
    full[4 * (size_t)dst + 3] = (float)rela_time; /* intensity := time :152-153 */
    full_ring[dst] = (uint16_t)r;
  }
  free(valid);
  const int N = nv;
  *n_full_out = N;
  int valid_scan_num = 0;
  for (int r = MSFLO_MAX_RINGS; r > 0; --r)
    if (cnt[r - 1] > 0) { valid_scan_num = r; break; }
  /* margins :188-195 */
  int sstart[MSFLO_MAX_RINGS], send[MSFLO_MAX_RINGS];
  for (int r = 0; r < valid_scan_num; r++) { sstart[r] = off[r] + 5; send[r] = off[r + 1] - 6; }
  /* curvature :213-240 -- fp32 11-term sum in source order, squares in fp64, stored as float */
  unsigned char *picked = (unsigned char *)calloc(N, 1);
  for (int i = 0; i < N; i++) { curv[i] = 0.f; label[i] = 0; }
  for (int i = 5; i < N - 5; i++) {
    float d[3];
    for (int a = 0; a < 3; a++) {
#define PX(k) full[4 * (size_t)(i + (k)) + a]
      d[a] = PX(-5) + PX(-4) + PX(-3) + PX(-2) + PX(-1) - 10 * PX(0) + PX(1) + PX(2) + PX(3) + PX(4) + PX(5);
#undef PX
    }
    double dx = d[0], dy = d[1], dz = d[2];
    curv[i] = (float)(dx * dx + dy * dy + dz * dz);
  }
  curv_idx *sorted = (curv_idx *)malloc(sizeof(curv_idx) * (N > 0 ? N : 1));
  int ns = 0, nls = 0, nf = 0, nlf = 0;
  for (int r = 0; r < valid_scan_num; r++) { /* :251-351 */
    if (send[r] - sstart[r] < 6) continue;
    for (int j = 0; j < P->n_sectors; j++) {
      int sp = sstart[r] + (send[r] - sstart[r]) * j / P->n_sectors;
      int ep = sstart[r] + (send[r] - sstart[r]) * (j + 1) / P->n_sectors - 1;
      int m = ep - sp + 1;
      if (m <= 0) continue;
      for (int k = 0; k < m; k++) { sorted[k].c = curv[sp + k]; sorted[k].i = sp + k; }
      qsort(sorted, m, sizeof(curv_idx), cmp_curv); /* :263 */
      int largest = 0;
      for (int k = m - 1; k >= 0; k--) { /* :272-305 */
        int ind = sorted[k].i;
        if (!picked[ind] && (double)curv[ind] > P->curvature_thresh) {
          largest++;
          if (largest <= P->n_sharp) {
            label[ind] = 1;
            idx_sharp[ns++] = ind;
            idx_less_sharp[nls++] = ind;
          } else if (largest <= P->n_less_sharp) {
            label[ind] = 2;
            idx_less_sharp[nls++] = ind;
          } else {
            break;
          }
          picked[ind] = 1;
          for (int l = 1; l <= 5; l++) {
            if ((double)gap_sq(full + 4 * (size_t)(ind + l), full + 4 * (size_t)(ind + l - 1)) > P->neighbor_gap_sq) break;
            picked[ind + l] = 1;
            label[ind + l] = 2;
          }
          for (int l = -1; l >= -5; l--) {
            if ((double)gap_sq(full + 4 * (size_t)(ind + l), full + 4 * (size_t)(ind + l + 1)) > P->neighbor_gap_sq) break;
            picked[ind + l] = 1;
            label[ind + l] = 2;
          }
        }
      }
      int smallest = 0;
      for (int k = 0; k < m; k++) { /* :309-336 */
        int ind = sorted[k].i;
        if (!picked[ind] && (double)curv[ind] < P->curvature_thresh) {
          label[ind] = 3;
          idx_flat[nf++] = ind;
          smallest++;
          if (smallest >= P->n_flat) break;
          picked[ind] = 1;
          for (int l = 1; l <= 5; l++) {
            if ((double)gap_sq(full + 4 * (size_t)(ind + l), full + 4 * (size_t)(ind + l - 1)) > P->neighbor_gap_sq) break;
            picked[ind + l] = 1;
          }
          for (int l = -1; l >= -5; l--) {
            if ((double)gap_sq(full + 4 * (size_t)(ind + l), full + 4 * (size_t)(ind + l + 1)) > P->neighbor_gap_sq) break;
            picked[ind + l] = 1;
          }
        }
      }
      for (int k = sp; k <= ep; k++) /* :339-344 */
        if (label[k] == 3 || label[k] == 0) idx_less_flat[nlf++] = k;
    }
    /* VoxelGridWrapper :347-350 is an identity copy (quirk Q1: getIndices() returns the input indices) */
  }
  free(sorted);
  free(picked);
  *n_sharp_out = ns; *n_less_sharp_out = nls; *n_flat_out = nf; *n_less_flat_out = nlf;
  /* TransformPointCloudInPlace :367-371 (rigid_transform.h:140-145) */
  if (T_ext)
    for (int i = 0; i < N; i++) {
      float o[3];
      msflo_transform_point_f(T_ext, full + 4 * (size_t)i, o);
      full[4 * (size_t)i + 0] = o[0]; full[4 * (size_t)i + 1] = o[1]; full[4 * (size_t)i + 2] = o[2];
    }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* pcl::VoxelGrid<pcl::PointXYZI>::applyFilter restatement (laser_mapping.cc:264-270,           */
/* hybrid_grid.cc:518-519).  downsample_all_data = true: xyz and intensity averaged in fp32;    */
/* output ordered by ascending voxel index; accumulation in ascending point index.              */
/* ------------------------------------------------------------------------------------------ */
typedef struct { unsigned int idx; int pt; } vox_pair;
static int cmp_vox(const void *pa, const void *pb) {
  const vox_pair *a = (const vox_pair *)pa, *b = (const vox_pair *)pb;
  if (a->idx != b->idx) return a->idx < b->idx ? -1 : 1;
  return (a->pt > b->pt) - (a->pt < b->pt);
}

int msflo_voxel_grid(const float *xyzi, int n, float leaf, float *out) {
  if (n <= 0) return 0;
  const float inv = 1.0f / leaf;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) {
      float v = xyzi[4 * (size_t)i + d];
      if (v < mn[d]) mn[d] = v;
      if (v > mx[d]) mx[d] = v;
    }
  int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1,
          dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)INT32_MAX) { /* "Leaf size is too small": output = input */
    memcpy(out, xyzi, sizeof(float) * 4 * (size_t)n);
    return n;
  }
  int min_b[3], max_b[3], div_b[3], mul[3];
  for (int d = 0; d < 3; d++) {
    min_b[d] = (int)floorf(mn[d] * inv);
    max_b[d] = (int)floorf(mx[d] * inv);
    div_b[d] = max_b[d] - min_b[d] + 1;
  }
  mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];
  vox_pair *v = (vox_pair *)malloc(sizeof(vox_pair) * n);
  for (int i = 0; i < n; i++) {
    const float *p = xyzi + 4 * (size_t)i;
    int i0 = (int)(floorf(p[0] * inv) - (float)min_b[0]);
    int i1 = (int)(floorf(p[1] * inv) - (float)min_b[1]);
    int i2 = (int)(floorf(p[2] * inv) - (float)min_b[2]);
    v[i].idx = (unsigned int)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]);
    v[i].pt = i;
  }
  qsort(v, n, sizeof(vox_pair), cmp_vox);
  int nout = 0, i = 0;
  while (i < n) {
    int j = i;
    float s[4] = {0, 0, 0, 0};
    while (j < n && v[j].idx == v[i].idx) {
      const float *p = xyzi + 4 * (size_t)v[j].pt;
      s[0] += p[0]; s[1] += p[1]; s[2] += p[2]; s[3] += p[3];
      j++;
    }
    const float cntf = (float)(j - i);
    out[4 * (size_t)nout + 0] = s[0] / cntf;
    out[4 * (size_t)nout + 1] = s[1] / cntf;
    out[4 * (size_t)nout + 2] = s[2] / cntf;
    out[4 * (size_t)nout + 3] = s[3] / cntf;
    nout++;
    i = j;
  }
  free(v);
  return nout;
}

/* ------------------------------------------------------------------------------------------ */
/* 8f row 1: STGM map (HybridGrid) restatement -- hybrid_grid.cc:403-521                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t key; /* packed (z, y, x) cell index, ascending order = output order */
  float *pts;  /* n x 4 */
  int n, cap;
} stgm_cell;

struct msflo_stgm {
  float resolution, leaf;
  stgm_cell *cells; /* sorted by key */
  int n_cells, cap_cells;
};

#define STGM_OFF (1 << 20)
static int64_t stgm_key(long ix, long iy, long iz) {
  return ((int64_t)(iz + STGM_OFF) << 42) | ((int64_t)(iy + STGM_OFF) << 21) | (int64_t)(ix + STGM_OFF);
}
/* GetCellIndex (:424-428): Array3f index = point / resolution (float); RoundToInt(double) = lround */
static int64_t stgm_cell_of(const msflo_stgm *m, float x, float y, float z) {
  const float fx = x / m->resolution, fy = y / m->resolution, fz = z / m->resolution;
  return stgm_key(lround((double)fx), lround((double)fy), lround((double)fz));
}

msflo_stgm *msflo_stgm_create(float resolution, float leaf) {
  msflo_stgm *m = (msflo_stgm *)calloc(1, sizeof *m);
  m->resolution = resolution;
  m->leaf = leaf;
  m->cap_cells = 64;
  m->cells = (stgm_cell *)calloc(m->cap_cells, sizeof(stgm_cell));
  return m;
}

void msflo_stgm_free(msflo_stgm *m) {
  if (!m) return;
  for (int i = 0; i < m->n_cells; i++) free(m->cells[i].pts);
  free(m->cells);
  free(m);
}

static int stgm_find(const msflo_stgm *m, int64_t key) { /* lower_bound */
  int lo = 0, hi = m->n_cells;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (m->cells[mid].key < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

void msflo_stgm_insert(msflo_stgm *m, const float *scan, int n) {
  if (n <= 0) return; /* :504 */
  int64_t *touched = (int64_t *)malloc(sizeof(int64_t) * n);
  int nt = 0;
  for (int i = 0; i < n; i++) { /* :506-511 push_back into the point's cell */
    const float *p = scan + 4 * (size_t)i;
    const int64_t key = stgm_cell_of(m, p[0], p[1], p[2]);
    int pos = stgm_find(m, key);
    if (pos == m->n_cells || m->cells[pos].key != key) {
      if (m->n_cells == m->cap_cells) {
        m->cap_cells *= 2;
        m->cells = (stgm_cell *)realloc(m->cells, sizeof(stgm_cell) * m->cap_cells);
      }
      memmove(m->cells + pos + 1, m->cells + pos, sizeof(stgm_cell) * (m->n_cells - pos));
      m->cells[pos].key = key; m->cells[pos].pts = 0; m->cells[pos].n = 0; m->cells[pos].cap = 0;
      m->n_cells++;
    }
    stgm_cell *c = &m->cells[pos];
    if (c->n == c->cap) {
      c->cap = c->cap ? 2 * c->cap : 16;
      c->pts = (float *)realloc(c->pts, sizeof(float) * 4 * c->cap);
    }
    memcpy(c->pts + 4 * (size_t)c->n, p, 16);
    c->n++;
    touched[nt++] = key;
  }
  for (int i = 0; i < nt; i++) { /* :513-520 every touched cell is voxel-filtered once */
    stgm_cell *c = &m->cells[stgm_find(m, touched[i])];
    if (c->cap < 0) continue; /* already filtered in this call (marker) */
    float *tmp = (float *)malloc(sizeof(float) * 4 * (size_t)c->n);
    int no = msflo_voxel_grid(c->pts, c->n, m->leaf, tmp);
    memcpy(c->pts, tmp, sizeof(float) * 4 * (size_t)no);
    free(tmp);
    c->n = no;
    c->cap = -c->cap; /* mark */
  }
  for (int i = 0; i < m->n_cells; i++)
    if (m->cells[i].cap < 0) m->cells[i].cap = -m->cells[i].cap;
  free(touched);
}

/* Rigid3f * Vector3f with Eigen's float quaternion rotation, then + (i, j, k) (:474-481) */
static void stgm_transform_f(const double pose[7], const float *p, float out[3]) {
  const float qx = (float)pose[3], qy = (float)pose[4], qz = (float)pose[5], qw = (float)pose[6];
  const float tx = (float)pose[0], ty = (float)pose[1], tz = (float)pose[2];
  float uv0 = qy * p[2] - qz * p[1], uv1 = qz * p[0] - qx * p[2], uv2 = qx * p[1] - qy * p[0];
  uv0 += uv0; uv1 += uv1; uv2 += uv2;
  const float c0 = qy * uv2 - qz * uv1, c1 = qz * uv0 - qx * uv2, c2 = qx * uv1 - qy * uv0;
  out[0] = (p[0] + qw * uv0 + c0) + tx;
  out[1] = (p[1] + qw * uv1 + c1) + ty;
  out[2] = (p[2] + qw * uv2 + c2) + tz;
}

int msflo_stgm_surround(const msflo_stgm *m, const float *scan, int n, const double pose[7], float *out) {
  unsigned char *sel = (unsigned char *)calloc(m->n_cells > 0 ? m->n_cells : 1, 1);
  for (int a = 0; a < n; a++) {
    const float *p = scan + 4 * (size_t)a;
    const float nr = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    if ((double)nr > 60.0) continue; /* kDist :474, :532 */
    float w[3];
    stgm_transform_f(pose, p, w);
    for (int i = -1; i <= 1; ++i)
      for (int j = -1; j <= 1; ++j)
        for (int k = -1; k <= 1; ++k) {
          const int64_t key = stgm_cell_of(m, w[0] + (float)i, w[1] + (float)j, w[2] + (float)k);
          const int pos = stgm_find(m, key);
          if (pos < m->n_cells && m->cells[pos].key == key) sel[pos] = 1; /* TryInsertGrid :524-529 */
        }
  }
  int no = 0;
  for (int c = 0; c < m->n_cells; c++)
    if (sel[c]) {
      memcpy(out + 4 * (size_t)no, m->cells[c].pts, sizeof(float) * 4 * (size_t)m->cells[c].n);
      no += m->cells[c].n;
    }
  free(sel);
  return no;
}

int msflo_stgm_size(const msflo_stgm *m, int *n_cells) {
  int n = 0;
  for (int c = 0; c < m->n_cells; c++) n += m->cells[c].n;
  if (n_cells) *n_cells = m->n_cells;
  return n;
}

int msflo_stgm_dump(const msflo_stgm *m, float *out) {
  int no = 0;
  for (int c = 0; c < m->n_cells; c++) {
    memcpy(out + 4 * (size_t)no, m->cells[c].pts, sizeof(float) * 4 * (size_t)m->cells[c].n);
    no += m->cells[c].n;
  }
  return no;
}
