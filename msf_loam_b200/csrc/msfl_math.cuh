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
// Largest eigenpair + middle eigenvalue of a 3x3 symmetric positive semi-definite matrix (stands in for
// Eigen::SelfAdjointEigenSolver<Matrix3d> at mapping_scan_matcher.cc:141, whose only use is the line test
// lambda_max > 3 lambda_mid and the direction of the largest eigenvector, :147-151).
//
// Branch-free and iteration-free, ~200 fp64 instructions instead of ~1200 for cyclic Jacobi:
//   1. M = S / tr(S), squared five times: M^32 = sum (lambda_i / tr)^32 u_i u_i^T.  Whenever the line test can pass
//      (lambda_mid < lambda_max / 3) the other two terms are below 3^-32 = 5e-16 of the first, so any column of M^32 IS
//      the largest eigenvector to round-off; the column with the largest diagonal entry is used.
//   2. lambda_max = u^T S u (Rayleigh quotient: second-order accurate in u).
//   3. lambda_mid = the larger eigenvalue of S compressed to the plane orthogonal to u, from the cancellation-free
//      formula (m00 + m11) / 2 + sqrt(((m00 - m11) / 2)^2 + m01^2).
// When lambda_max is NOT isolated u is meaningless, but then u^T S u <= lambda_max and (Cauchy interlacing) the
// compressed eigenvalue >= lambda_mid, so the computed ratio can only be smaller than the true one: the test fails as it
// must.  The eigenvector's sign is implementation-defined upstream (it falls out of Eigen's QL iteration) and the factor
// is invariant to it; it is fixed by the rule "largest-magnitude component positive" here and in the oracle.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sym_eig3_top(double a00, double a01, double a02, double a11, double a12, double a22,
                                             double &lam_max, double &lam_mid, double u[3]) {
  const double tr = a00 + a11 + a22;
  u[0] = 1.0; u[1] = 0.0; u[2] = 0.0;
  lam_max = 0.0; lam_mid = 0.0;
  if (!(tr > 0.0) || !isfinite(tr)) return;  // zero scatter (five coincident points): 0 > 3 * 0 is false, no line
  const double it = 1.0 / tr;
  double m00 = a00 * it, m01 = a01 * it, m02 = a02 * it, m11 = a11 * it, m12 = a12 * it, m22 = a22 * it;
#pragma unroll
  for (int k = 0; k < 5; ++k) {  // M <- M M (symmetric): eigenvalues in [0, 1], the largest >= 1/3, so 3^-32 never underflows
    const double n00 = m00 * m00 + m01 * m01 + m02 * m02, n01 = m00 * m01 + m01 * m11 + m02 * m12,
                 n02 = m00 * m02 + m01 * m12 + m02 * m22, n11 = m01 * m01 + m11 * m11 + m12 * m12,
                 n12 = m01 * m02 + m11 * m12 + m12 * m22, n22 = m02 * m02 + m12 * m12 + m22 * m22;
    m00 = n00; m01 = n01; m02 = n02; m11 = n11; m12 = n12; m22 = n22;
  }
  // the column with the largest diagonal entry (= largest u_j^2)
  double v0 = m00, v1 = m01, v2 = m02, best = m00;
  if (m11 > best) { best = m11; v0 = m01; v1 = m11; v2 = m12; }
  if (m22 > best) { best = m22; v0 = m02; v1 = m12; v2 = m22; }
  const double vn = v0 * v0 + v1 * v1 + v2 * v2;
  if (!(vn > 0.0)) return;
  const double iv = rsqrt(vn);
  double x = v0 * iv, y = v1 * iv, z = v2 * iv;
  // sign rule: the largest-magnitude component is positive (first one on ties)
  {
    const double ax = fabs(x), ay = fabs(y), az = fabs(z);
    const double lead = (ax >= ay && ax >= az) ? x : (ay >= az ? y : z);
    if (lead < 0.0) { x = -x; y = -y; z = -z; }
  }
  // Rayleigh quotient on the unscaled matrix
  const double s0 = a00 * x + a01 * y + a02 * z, s1 = a01 * x + a11 * y + a12 * z, s2 = a02 * x + a12 * y + a22 * z;
  lam_max = x * s0 + y * s1 + z * s2;
  // orthonormal basis (p, q) of the plane orthogonal to u: p = e_k x u for the axis k along which u is smallest
  double p0, p1, p2;
  {
    const double ax = fabs(x), ay = fabs(y), az = fabs(z);
    if (ax <= ay && ax <= az) { p0 = 0.0; p1 = -z; p2 = y; }       // e_x x u
    else if (ay <= az) { p0 = z; p1 = 0.0; p2 = -x; }              // e_y x u
    else { p0 = -y; p1 = x; p2 = 0.0; }                            // e_z x u
    const double ip = rsqrt(p0 * p0 + p1 * p1 + p2 * p2);
    p0 *= ip; p1 *= ip; p2 *= ip;
  }
  const double q0 = y * p2 - z * p1, q1 = z * p0 - x * p2, q2 = x * p1 - y * p0;  // u x p
  const double sp0 = a00 * p0 + a01 * p1 + a02 * p2, sp1 = a01 * p0 + a11 * p1 + a12 * p2, sp2 = a02 * p0 + a12 * p1 + a22 * p2;
  const double sq0 = a00 * q0 + a01 * q1 + a02 * q2, sq1 = a01 * q0 + a11 * q1 + a12 * q2, sq2 = a02 * q0 + a12 * q1 + a22 * q2;
  const double c00 = p0 * sp0 + p1 * sp1 + p2 * sp2, c01 = q0 * sp0 + q1 * sp1 + q2 * sp2, c11 = q0 * sq0 + q1 * sq1 + q2 * sq2;
  const double h = 0.5 * (c00 - c11);
  lam_mid = 0.5 * (c00 + c11) + sqrt(h * h + c01 * c01);
  u[0] = x; u[1] = y; u[2] = z;
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
