// assoc_device.cuh -- device code of the scan-to-map association shared by the batch kernels (associate_map.cu) and
// the single-scan fused kernel (scan2map_fused.cu): exact 5-NN over the cell index (replaces
// pcl::KdTreeFLANN::nearestKSearch, mapping_scan_matcher.cc:125 / :195) and the line / plane fits (:130-151, :199-220).
#pragma once
#include "msfl_internal.h"
#include "msfl_math.cuh"

namespace msfl {

struct Top5 {
  float d[5];
  int i[5];
};

// (d, id) < (bd, bi) in the exact candidate order (distance, then original index).  Squared distances are >= +0, so
// the order of the floats is the order of their bit patterns and the pair compares as ONE 64-bit unsigned integer
// (two ISETP instead of two FSETP + ISETP + predicate logic; this comparison is 15 % of the search's instructions).
// Callers filter NaN distances with a float compare first.  Indices are non-negative.
__device__ __forceinline__ bool cand_less(float d, int id, float bd, int bi) {
  const unsigned long long a = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)id;
  const unsigned long long b = ((unsigned long long)__float_as_uint(bd) << 32) | (unsigned)bi;
  return a < b;
}

// Sorted insertion without branches and in place: l_s = "entry s stays ahead of the candidate"; every slot is
// rewritten by selects from the top down, so the hot candidate loop around it needs no register copies.
__device__ __forceinline__ void top5_insert(Top5 &t, float d, int id) {
  // precondition: (d, id) < (t.d[4], t.i[4])
  const bool l0 = cand_less(t.d[0], t.i[0], d, id), l1 = cand_less(t.d[1], t.i[1], d, id),
             l2 = cand_less(t.d[2], t.i[2], d, id), l3 = cand_less(t.d[3], t.i[3], d, id);
  t.d[4] = l3 ? d : t.d[3];                    t.i[4] = l3 ? id : t.i[3];
  t.d[3] = l3 ? t.d[3] : (l2 ? d : t.d[2]);    t.i[3] = l3 ? t.i[3] : (l2 ? id : t.i[2]);
  t.d[2] = l2 ? t.d[2] : (l1 ? d : t.d[1]);    t.i[2] = l2 ? t.i[2] : (l1 ? id : t.i[1]);
  t.d[1] = l1 ? t.d[1] : (l0 ? d : t.d[0]);    t.i[1] = l1 ? t.i[1] : (l0 ? id : t.i[0]);
  t.d[0] = l0 ? t.d[0] : d;                    t.i[0] = l0 ? t.i[0] : id;
}

// 5-NN of q among the 27 cells around it, restricted to d2 < thresh.  Returns true when five
// such neighbours exist (<=> pointSearchSqDis[4] < thresh for the exact 5-NN).
//
// Rows (dy, dz) are visited nearest-first and a row -- or its left / right cell -- is skipped when
// even the closest possible point in it cannot enter the current top-5.  The bounds are exact in
// fp32: with 1 m cells the cell boundaries are integers, float subtraction / squaring / addition
// are monotone, and the bound is accumulated in the same order as the distance itself
// ((bx^2 + by^2) + bz^2), so  bound > worst  implies  d > worst  for every point of that cell.
__device__ __forceinline__ bool knn5_grid(const GridView &g, float qx, float qy, float qz, float thresh, Top5 &t) {
  // sentinels (thresh, 0): no real candidate compares below them at d == thresh, and a real entry always has
  // d < thresh, so "five neighbours inside the gate" <=> t.d[4] < thresh
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    t.d[s] = thresh;
    t.i[s] = 0;
  }
  const float fxq = floorf(qx * g.inv_edge), fyq = floorf(qy * g.inv_edge), fzq = floorf(qz * g.inv_edge);
  const int cx = (int)fxq - g.ox, cy = (int)fyq - g.oy, cz = (int)fzq - g.oz;
  if (cx < 1 || cy < 1 || cz < 1 || cx > g.nx - 2 || cy > g.ny - 2 || cz > g.nz - 2) return false;
  const bool exact_cells = (g.inv_edge == 1.0f);
  // distance from q to the lower / upper faces of its own cell (>= 0); 0 disables pruning
  const float lox = exact_cells ? __fsub_rn(qx, fxq) : 0.f, hix = exact_cells ? __fsub_rn(fxq + 1.0f, qx) : 0.f;
  const float loy = exact_cells ? __fsub_rn(qy, fyq) : 0.f, hiy = exact_cells ? __fsub_rn(fyq + 1.0f, qy) : 0.f;
  const float loz = exact_cells ? __fsub_rn(qz, fzq) : 0.f, hiz = exact_cells ? __fsub_rn(fzq + 1.0f, qz) : 0.f;
  const float bx2[2] = {__fmul_rn(lox, lox), __fmul_rn(hix, hix)};  // left cell, right cell
  // nearest-first row order: centre, 4 face neighbours, 4 diagonal neighbours; (dy+1, dz+1) packed
  // 2 bits each so the row loop stays rolled (the unrolled form thrashed the instruction cache)
  //            r:   0      1      2      3      4      5      6      7      8
  //      (dy,dz): (0,0) (-1,0) (1,0) (0,-1) (0,1) (-1,-1) (-1,1) (1,-1) (1,1)
  constexpr uint32_t kDyPacked = 1u | (0u << 2) | (2u << 4) | (1u << 6) | (1u << 8) | (0u << 10) | (0u << 12) | (2u << 14) | (2u << 16);
  constexpr uint32_t kDzPacked = 1u | (1u << 2) | (1u << 4) | (0u << 6) | (2u << 8) | (0u << 10) | (2u << 12) | (0u << 14) | (2u << 16);
  // only the rows that hold a point at all (k_row_mask; all nine when the table was not built)
  uint32_t rows = g.row_mask ? (uint32_t)__ldg(g.row_mask + ((size_t)cz * g.ny + cy) * g.nx + cx) : 0x1ffu;
#pragma unroll 1
  for (; rows; rows &= rows - 1u) {
    const int r = __ffs(rows) - 1;
    const int dy = (int)((kDyPacked >> (2 * r)) & 3u) - 1, dz = (int)((kDzPacked >> (2 * r)) & 3u) - 1;
    const float by = dy == 0 ? 0.f : (dy < 0 ? loy : hiy), bz = dz == 0 ? 0.f : (dz < 0 ? loz : hiz);
    const float by2 = __fmul_rn(by, by), bz2 = __fmul_rn(bz, bz);
    const float row_lb = __fadd_rn(by2, bz2);  // (0 + by^2) + bz^2
    if (row_lb > t.d[4] || row_lb >= thresh) continue;
    const int row = ((cz + dz) * g.ny + (cy + dy)) * g.nx + cx;
    const uint32_t s0 = __ldg(g.cell_start + row - 1), s1 = __ldg(g.cell_start + row),
                   s2 = __ldg(g.cell_start + row + 1), s3 = __ldg(g.cell_start + row + 2);
    const float lbl = __fadd_rn(__fadd_rn(bx2[0], by2), bz2), lbr = __fadd_rn(__fadd_rn(bx2[1], by2), bz2);
    const uint32_t js = (lbl > t.d[4] || lbl >= thresh) ? s1 : s0;
    const uint32_t je = (lbr > t.d[4] || lbr >= thresh) ? s2 : s3;
#pragma unroll 2
    for (uint32_t j = js; j < je; ++j) {
      const float4 m = __ldg(g.pts_sorted + j);
      const float dx = __fsub_rn(qx, m.x), dy2 = __fsub_rn(qy, m.y), dz2 = __fsub_rn(qz, m.z);
      float d = __fmul_rn(dx, dx);
      d = __fadd_rn(d, __fmul_rn(dy2, dy2));
      d = __fadd_rn(d, __fmul_rn(dz2, dz2));
      if (d <= t.d[4]) {  // cheap filter; exact (d2, index) order inside
        const int id = __float_as_int(m.w);
        if (cand_less(d, id, t.d[4], t.i[4])) top5_insert(t, d, id);
      }
    }
  }
  return t.d[4] < thresh;
}

// ---------------------------------------------------------------------------------------------
// Plane fit without the QR.  The reference solves the 5x3 least-squares system A x = -1 (rows = the five
// neighbours) and normalises x (mapping_scan_matcher.cc:199-211).  With c = the mean row and S = sum (p-c)(p-c)^T
// the normal equations are (S + 5 c c^T) x = -5 c, so by Sherman-Morrison x is a positive multiple of -S^-1 c =
// -adj(S) c / det(S): the unit normal is  n = -adj(S) c / |adj(S) c|  -- six 2x2 minors of a CENTRED 3x3 matrix,
// one matrix-vector product and one reciprocal square root (~100 fp64 instructions, 64 registers) instead of three
// Householder reflections with their square roots and divisions (~250, 128 registers).  It is the same least-squares
// solution, not an approximation; centring removes the |c|^2 / spread^2 conditioning of the raw system, so it
// agrees with the pivoted-Householder solve to ~5e-12 on the bench clouds (tools/dev_fit_study.py).
// Round-off is amplified by tr(S)^2 |c| / |adj(S) c| (five nearly collinear neighbours): beyond 1e5, or when a
// neighbour sits within 1e-7 m of the validity bound, the query is handed to the Householder kernel instead
// (0.3 % / 1.4 % of the VLP-16 / HDL-64E queries).  Explicit round-to-nearest intrinsics: every kernel that hosts
// this function produces the same bits.
// Returns true when the query needs the Householder path; otherwise nrm (zero when the plane is invalid).
// ---------------------------------------------------------------------------------------------
constexpr double kFastFitMinRatioSq = 1e-10;  // (|adj(S) c| / (tr(S)^2 |c|))^2 below this: ill-conditioned
constexpr double kFastFitBorder = 1e-7;       // metres around plane_tol
__device__ __forceinline__ bool plane_fit_fast(const float (&mf)[5][3], const double (&c)[3], double plane_tol, double (&nrm)[3]) {
  double S00 = 0, S01 = 0, S02 = 0, S11 = 0, S12 = 0, S22 = 0;
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const double e0 = __dsub_rn((double)mf[s][0], c[0]), e1 = __dsub_rn((double)mf[s][1], c[1]), e2 = __dsub_rn((double)mf[s][2], c[2]);
    S00 = __fma_rn(e0, e0, S00); S01 = __fma_rn(e0, e1, S01); S02 = __fma_rn(e0, e2, S02);
    S11 = __fma_rn(e1, e1, S11); S12 = __fma_rn(e1, e2, S12); S22 = __fma_rn(e2, e2, S22);
  }
  // adj(S) (symmetric): the 2x2 minors
  const double A00 = __fma_rn(S11, S22, -__dmul_rn(S12, S12)), A01 = __fma_rn(S02, S12, -__dmul_rn(S01, S22)),
               A02 = __fma_rn(S01, S12, -__dmul_rn(S02, S11)), A11 = __fma_rn(S00, S22, -__dmul_rn(S02, S02)),
               A12 = __fma_rn(S01, S02, -__dmul_rn(S00, S12)), A22 = __fma_rn(S00, S11, -__dmul_rn(S01, S01));
  const double v0 = -__fma_rn(A02, c[2], __fma_rn(A01, c[1], __dmul_rn(A00, c[0]))),
               v1 = -__fma_rn(A12, c[2], __fma_rn(A11, c[1], __dmul_rn(A01, c[0]))),
               v2 = -__fma_rn(A22, c[2], __fma_rn(A12, c[1], __dmul_rn(A02, c[0])));
  const double vv = __fma_rn(v2, v2, __fma_rn(v1, v1, __dmul_rn(v0, v0)));
  const double tr = __dadd_rn(__dadd_rn(S00, S11), S22), tr2 = __dmul_rn(tr, tr);
  const double cc = __fma_rn(c[2], c[2], __fma_rn(c[1], c[1], __dmul_rn(c[0], c[0])));
  if (!(vv > __dmul_rn(__dmul_rn(kFastFitMinRatioSq, __dmul_rn(tr2, tr2)), cc))) return true;
  const double inv = rsqrt(vv);
  const double n0 = __dmul_rn(v0, inv), n1 = __dmul_rn(v1, inv), n2 = __dmul_rn(v2, inv);
  bool valid = true, border = false;
#pragma unroll
  for (int s = 0; s < 5; ++s) {  // :214-220
    const double e0 = __dsub_rn((double)mf[s][0], c[0]), e1 = __dsub_rn((double)mf[s][1], c[1]), e2 = __dsub_rn((double)mf[s][2], c[2]);
    const double dd = fabs(__fma_rn(n2, e2, __fma_rn(n1, e1, __dmul_rn(n0, e0))));
    if (!(dd <= plane_tol)) valid = false;
    if (fabs(__dsub_rn(dd, plane_tol)) < kFastFitBorder) border = true;
  }
  if (border) return true;
  nrm[0] = valid ? n0 : 0.0; nrm[1] = valid ? n1 : 0.0; nrm[2] = valid ? n2 : 0.0;
  return false;
}

// centroid of the five neighbours, summed in index order and divided by 5 (:137 / :212)
__device__ __forceinline__ void centroid5(const float (&mf)[5][3], double (&c)[3]) {
#pragma unroll
  for (int d = 0; d < 3; ++d)
    c[d] = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn((double)mf[0][d], (double)mf[1][d]), (double)mf[2][d]), (double)mf[3][d]), (double)mf[4][d]), 5.0);
}

// Line test of a corner query (:137-151, :168): S = sum (p - c)(p - c)^T (not divided by 5), eigen-decomposition,
// lambda_max > ratio * lambda_mid  ->  a = c + 0.1 u, n = unit(a - b); otherwise a = n = 0 ("no factor").
__device__ __forceinline__ void line_fit(const float (&mf)[5][3], const double (&c)[3], const KParams &kp, double (&a)[3], double (&n)[3]) {
  double cov[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const double e0 = (double)mf[s][0] - c[0], e1 = (double)mf[s][1] - c[1], e2 = (double)mf[s][2] - c[2];
    cov[0] += e0 * e0; cov[1] += e0 * e1; cov[2] += e0 * e2;
    cov[3] += e1 * e1; cov[4] += e1 * e2; cov[5] += e2 * e2;
  }
  double lmax, lmid, u[3];
  sym_eig3_top(cov[0], cov[1], cov[2], cov[3], cov[4], cov[5], lmax, lmid, u);  // :141
  if (lmax > kp.line_eig_ratio * lmid) {                                          // :147
    double b[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      a[d] = kp.line_half_len * u[d] + c[d];   // :150
      b[d] = -kp.line_half_len * u[d] + c[d];  // :151
      n[d] = a[d] - b[d];
    }
    const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);  // (point_a - point_b).normalized() :168
    if (nn > 0) { n[0] /= nn; n[1] /= nn; n[2] /= nn; }
  }
}

// Plane fit of a surf query by the reference's own route (:199-220): pivoted Householder least squares of A x = -1,
// normalised, valid iff every neighbour lies within plane_tol of the plane through the centroid.  n = 0 when invalid.
__device__ __forceinline__ void plane_fit_qr(const float (&mf)[5][3], const double (&c)[3], const KParams &kp, double (&n)[3]) {
  double A[5][3], bb[5] = {-1, -1, -1, -1, -1}, nrm[3];
#pragma unroll
  for (int s = 0; s < 5; ++s) { A[s][0] = (double)mf[s][0]; A[s][1] = (double)mf[s][1]; A[s][2] = (double)mf[s][2]; }
  lstsq_5x3(A, bb, nrm);  // :210
  const double inv_nn = 1.0 / sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
  nrm[0] *= inv_nn; nrm[1] *= inv_nn; nrm[2] *= inv_nn;  // norm.normalize() :211
  bool valid = true;
#pragma unroll
  for (int s = 0; s < 5; ++s) {  // :214-220
    const double dd = nrm[0] * ((double)mf[s][0] - c[0]) + nrm[1] * ((double)mf[s][1] - c[1]) + nrm[2] * ((double)mf[s][2] - c[2]);
    if (!(fabs(dd) <= kp.plane_tol)) valid = false;
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) n[d] = valid ? nrm[d] : 0.0;
}

}  // namespace msfl
