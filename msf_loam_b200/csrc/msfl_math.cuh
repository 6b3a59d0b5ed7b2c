// msfl_math.cuh -- small fp64 geometry / dense kernels shared by the CUDA kernels.
// Semantics follow the reference's Eigen/Ceres usage (file:line cited per function).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace msfl {

// ---------------------------------------------------------------------------------------------
// Rigid transform of an fp32 point with fp64 math, rounded back to fp32: TransformPoint
// (rigid_transform.h:132-138) and TransformToStart with s = 1 (odometry_scan_matcher.cc:21-33).
// Eigen's Quaternion * Vector3 is  v + w*uv + qv x uv  with uv = 2 (qv x v); written with
// explicit round-to-nearest intrinsics so no FMA contraction changes the fp32 rounding of the
// kNN query relative to a generic x86-64 build.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double mul_sub_exact(double a, double b, double c, double d) {
  return __dsub_rn(__dmul_rn(a, b), __dmul_rn(c, d));
}

__device__ __forceinline__ void quat_rotate_exact(const double q[4] /* x y z w */, double v0, double v1, double v2,
                                                  double &o0, double &o1, double &o2) {
  double uv0 = mul_sub_exact(q[1], v2, q[2], v1);
  double uv1 = mul_sub_exact(q[2], v0, q[0], v2);
  double uv2 = mul_sub_exact(q[0], v1, q[1], v0);
  uv0 = __dadd_rn(uv0, uv0);
  uv1 = __dadd_rn(uv1, uv1);
  uv2 = __dadd_rn(uv2, uv2);
  double c0 = mul_sub_exact(q[1], uv2, q[2], uv1);
  double c1 = mul_sub_exact(q[2], uv0, q[0], uv2);
  double c2 = mul_sub_exact(q[0], uv1, q[1], uv0);
  o0 = __dadd_rn(__dadd_rn(v0, __dmul_rn(q[3], uv0)), c0);
  o1 = __dadd_rn(__dadd_rn(v1, __dmul_rn(q[3], uv1)), c1);
  o2 = __dadd_rn(__dadd_rn(v2, __dmul_rn(q[3], uv2)), c2);
}

__device__ __forceinline__ float3 transform_point_f(const double pose[7], float x, float y, float z) {
  double r0, r1, r2;
  quat_rotate_exact(pose + 3, (double)x, (double)y, (double)z, r0, r1, r2);
  return make_float3((float)__dadd_rn(r0, pose[0]), (float)__dadd_rn(r1, pose[1]), (float)__dadd_rn(r2, pose[2]));
}

// Eigen::Quaternion::toRotationMatrix (used by the factors, lidar_factor.cc:19,39), row-major.
__host__ __device__ __forceinline__ void quat_to_R(const double q[4], double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// PoseLocalParameterization::Plus (pose_local_parameterization.cc:6-21) with Utility::deltaQ
// (utility.h:8-31): p += dp; q = (q * dQ(dtheta)).normalized().
__host__ __device__ inline void pose_plus(const double x[7], const double d[6], double out[7]) {
  const double vx = d[3], vy = d[4], vz = d[5];
  const double theta = sqrt(vx * vx + vy * vy + vz * vz);
  const double half_theta = 0.5 * theta;
  double imag, real, sh;
  sincos(half_theta, &sh, &real);
  if (theta < 1e-6) {
    const double t2 = theta * theta, t4 = t2 * t2;
    imag = 0.5 - (1 / 48.) * t2 + (1 / 3840.) * t4;
  } else {
    imag = sh / theta;
  }
  const double bx = imag * vx, by = imag * vy, bz = imag * vz, bw = real;
  const double ax = x[3], ay = x[4], az = x[5], aw = x[6];
  double w = aw * bw - ax * bx - ay * by - az * bz;
  double qx = aw * bx + ax * bw + ay * bz - az * by;
  double qy = aw * by + ay * bw + az * bx - ax * bz;
  double qz = aw * bz + az * bw + ax * by - ay * bx;
  const double n = sqrt(qx * qx + qy * qy + qz * qz + w * w);
  out[0] = x[0] + d[0];
  out[1] = x[1] + d[1];
  out[2] = x[2] + d[2];
  if (n > 0) { const double rn = 1.0 / n; qx *= rn; qy *= rn; qz *= rn; w *= rn; }
  out[3] = qx; out[4] = qy; out[5] = qz; out[6] = w;
}

// ---------------------------------------------------------------------------------------------
// 3x3 symmetric eigen-decomposition (stands in for Eigen::SelfAdjointEigenSolver<Matrix3d>,
// mapping_scan_matcher.cc:141).  Cyclic Jacobi on the upper triangle; returns the largest and
// middle eigenvalues and the unit eigenvector of the largest.
// ---------------------------------------------------------------------------------------------
__device__ inline void sym_eig3_top(double a00, double a01, double a02, double a11, double a12, double a22,
                                    double &lam_max, double &lam_mid, double u[3]) {
  double A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 60; ++sweep) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    const double dg = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-32 * dg || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double tau = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
    }
  }
  const double e0 = A[0][0], e1 = A[1][1], e2 = A[2][2];
  // ascending order with the oracle's tie rule (stable bubble: first index wins the lower slot)
  int i0 = 0, i1 = 1, i2 = 2;
  double f0 = e0, f1 = e1, f2 = e2;
  if (f0 > f1) { double t = f0; f0 = f1; f1 = t; int ti = i0; i0 = i1; i1 = ti; }
  if (f1 > f2) { double t = f1; f1 = f2; f2 = t; int ti = i1; i1 = i2; i2 = ti; }
  if (f0 > f1) { double t = f0; f0 = f1; f1 = t; int ti = i0; i0 = i1; i1 = ti; }
  lam_max = f2;
  lam_mid = f1;
  u[0] = (i2 == 0) ? V[0][0] : (i2 == 1 ? V[0][1] : V[0][2]);
  u[1] = (i2 == 0) ? V[1][0] : (i2 == 1 ? V[1][1] : V[1][2]);
  u[2] = (i2 == 0) ? V[2][0] : (i2 == 1 ? V[2][1] : V[2][2]);
}

// ---------------------------------------------------------------------------------------------
// 5x3 least squares A x = b by column-pivoted Householder QR (stands in for
// matA0.colPivHouseholderQr().solve(matB0), mapping_scan_matcher.cc:210).  Every index is a
// compile-time constant (the pivot is applied as a conditional column swap), so A, b and the
// reflector stay in registers -- the dynamically indexed version lived in local memory.
// ---------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void qr_step(double (&A)[5][3], double (&b)[5], int (&perm)[3], double (&Rdiag)[3], int &rank,
                                        double &maxpiv) {
  if (rank < 3) return;  // an earlier column was (numerically) zero: Eigen stops there too
  // pivot: first remaining column with the largest norm over rows K..4
  double cn[3] = {0, 0, 0};
#pragma unroll
  for (int j = K; j < 3; ++j)
#pragma unroll
    for (int i = K; i < 5; ++i) cn[j] += A[i][j] * A[i][j];
  int best = K;
  double bn = cn[K];
#pragma unroll
  for (int j = K + 1; j < 3; ++j)
    if (cn[j] > bn) { bn = cn[j]; best = j; }
#pragma unroll
  for (int j = K + 1; j < 3; ++j) {
    if (best == j) {
#pragma unroll
      for (int i = 0; i < 5; ++i) { const double t = A[i][K]; A[i][K] = A[i][j]; A[i][j] = t; }
      const int t = perm[K]; perm[K] = perm[j]; perm[j] = t;
    }
  }
  const double nrm = sqrt(bn);
  if (K == 0) maxpiv = nrm;
  if (!(nrm > maxpiv * 1e-14) || nrm == 0.0) { rank = K; return; }
  const double alpha = (A[K][K] > 0) ? -nrm : nrm;
  double v[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) v[i] = i >= K ? A[i][K] : 0.0;
  v[K] -= alpha;
  double vv = 0;
#pragma unroll
  for (int i = K; i < 5; ++i) vv += v[i] * v[i];
  if (vv > 0) {
    const double two_over_vv = 2.0 / vv;  // one division per reflector
#pragma unroll
    for (int j = K; j < 3; ++j) {
      double s = 0;
#pragma unroll
      for (int i = K; i < 5; ++i) s += v[i] * A[i][j];
      s = s * two_over_vv;
#pragma unroll
      for (int i = K; i < 5; ++i) A[i][j] -= s * v[i];
    }
    double s = 0;
#pragma unroll
    for (int i = K; i < 5; ++i) s += v[i] * b[i];
    s = s * two_over_vv;
#pragma unroll
    for (int i = K; i < 5; ++i) b[i] -= s * v[i];
  }
  Rdiag[K] = A[K][K];
}

__device__ __forceinline__ void lstsq_5x3(double (&A)[5][3], double (&b)[5], double (&x)[3]) {
  int perm[3] = {0, 1, 2};
  double Rdiag[3] = {0, 0, 0};
  int rank = 3;
  double maxpiv = 0;
  qr_step<0>(A, b, perm, Rdiag, rank, maxpiv);
  qr_step<1>(A, b, perm, Rdiag, rank, maxpiv);
  qr_step<2>(A, b, perm, Rdiag, rank, maxpiv);
  double y[3] = {0, 0, 0};
  if (rank > 2) y[2] = b[2] / Rdiag[2];
  if (rank > 1) y[1] = (b[1] - (rank > 2 ? A[1][2] * y[2] : 0.0)) / Rdiag[1];
  if (rank > 0) y[0] = ((b[0] - (rank > 1 ? A[0][1] * y[1] : 0.0)) - (rank > 2 ? A[0][2] * y[2] : 0.0)) / Rdiag[0];
#pragma unroll
  for (int c = 0; c < 3; ++c) x[c] = perm[0] == c ? y[0] : (perm[1] == c ? y[1] : y[2]);
}

// index of the (u,v) entry, u <= v, in the packed upper triangle of a 6x6 (row-major)
__host__ __device__ __forceinline__ constexpr int tri6(int u, int v) { return u * 6 - (u * (u - 1)) / 2 + (v - u); }

// 6x6 SPD solve by Cholesky (the LM normal equations) on the packed upper triangle, IN PLACE: A[tri6(j, i)] = A_ij for
// j <= i comes in and L_ij goes out in the same slot (column j only reads the finished columns k < j and itself).
// Returns false when not positive definite / non-finite (Ceres LINEAR_SOLVER_FAILURE -> invalid step).  Every index is
// a compile-time constant, so the 21 + 18 values live in registers -- with a separate L the serial LM step spilled --
// and one reciprocal square root per pivot replaces the sqrt + divisions of the textbook form (this runs on a single
// thread between two sweeps: its latency is exposed).
__device__ __forceinline__ bool chol_solve6_packed(double (&A)[21], const double (&b)[6], double (&y)[6]) {
  double inv[6];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = A[tri6(j, j)];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k < j) d -= A[tri6(k, j)] * A[tri6(k, j)];
    if (!(d > 0) || !isfinite(d)) ok = false;
    inv[j] = rsqrt(d);
    A[tri6(j, j)] = d * inv[j];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (i > j) {
        double s = A[tri6(j, i)];
#pragma unroll
        for (int k = 0; k < 6; ++k)
          if (k < j) s -= A[tri6(k, i)] * A[tri6(k, j)];
        A[tri6(j, i)] = s * inv[j];
      }
    }
  }
  if (!ok) return false;
  double z[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k < i) s -= A[tri6(k, i)] * z[k];
    z[i] = s * inv[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = z[i];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k > i) s -= A[tri6(i, k)] * y[k];
    y[i] = s * inv[i];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
    if (!isfinite(y[i])) ok = false;
  return ok;
}

}  // namespace msfl
