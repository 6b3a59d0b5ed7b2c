// associate_map.cu -- scan-to-map data association:
//   per corner query: 5-NN in the corner submap, d5^2 < 1 gate, centroid + 3x3 covariance +
//   eigen line test  -> edge factor constants        (mapping_scan_matcher.cc:109-176)
//   per surf query:   5-NN in the surf submap, gate, 5x3 least-squares plane, 0.2 m validity
//                                                    -> plane factor constants (:178-246)
// replacing pcl::KdTreeFLANN::nearestKSearch + Eigen.  One thread per query over a flat 1-D grid
// covering every query of the batch (scan id by binary search in the offset table).  kNN distances are fp32 ((dx*dx)+dy*dy)+dz*dz without FMA
// contraction (FLANN L2_Simple<float>), candidates are ordered by (d2, original index), the fit
// is fp64.  Output per query: 6 doubles [a_or_c(3), n(3)]; n = 0 marks "no factor".
#include "msfl_internal.h"
#include "msfl_math.cuh"

namespace msfl {

struct Top5 {
  float d[5];
  int i[5];
};

__device__ __forceinline__ bool cand_less(float d, int id, float bd, int bi) {
  return d < bd || (d == bd && id < bi);
}

__device__ __forceinline__ void top5_insert(Top5 &t, float d, int id) {
  // precondition: (d, id) < (t.d[4], t.i[4])
  bool placed = false;
#pragma unroll
  for (int s = 4; s > 0; --s) {
    if (!placed) {
      if (cand_less(d, id, t.d[s - 1], t.i[s - 1])) {
        t.d[s] = t.d[s - 1];
        t.i[s] = t.i[s - 1];
      } else {
        t.d[s] = d;
        t.i[s] = id;
        placed = true;
      }
    }
  }
  if (!placed) {
    t.d[0] = d;
    t.i[0] = id;
  }
}

// 5-NN of q among the 27 cells around it, restricted to d2 < thresh.  Returns true when five
// such neighbours exist (<=> pointSearchSqDis[4] < thresh for the exact 5-NN).
__device__ __forceinline__ bool knn5_grid(const GridView &g, float qx, float qy, float qz, float thresh, Top5 &t) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    t.d[s] = thresh;
    t.i[s] = -1;
  }
  const int cx = (int)floorf(qx * g.inv_edge) - g.ox;
  const int cy = (int)floorf(qy * g.inv_edge) - g.oy;
  const int cz = (int)floorf(qz * g.inv_edge) - g.oz;
  if (cx < 1 || cy < 1 || cz < 1 || cx > g.nx - 2 || cy > g.ny - 2 || cz > g.nz - 2) return false;
  for (int dz = -1; dz <= 1; ++dz) {
    for (int dy = -1; dy <= 1; ++dy) {
      const int row = ((cz + dz) * g.ny + (cy + dy)) * g.nx + cx;
      const uint32_t s = __ldg(g.cell_start + row - 1), e = __ldg(g.cell_start + row + 2);
      for (uint32_t j = s; j < e; ++j) {
        const float4 m = __ldg(g.pts_sorted + j);
        const float dx = __fsub_rn(qx, m.x), dy2 = __fsub_rn(qy, m.y), dz2 = __fsub_rn(qz, m.z);
        float d = __fmul_rn(dx, dx);
        d = __fadd_rn(d, __fmul_rn(dy2, dy2));
        d = __fadd_rn(d, __fmul_rn(dz2, dz2));
        const int id = __float_as_int(m.w);
        if (cand_less(d, id, t.d[4], t.i[4])) top5_insert(t, d, id);
      }
    }
  }
  return t.i[4] >= 0;
}

__device__ __forceinline__ void store_corr(double *corr, size_t q, const double a[3], const double n[3]) {
  double2 *o = reinterpret_cast<double2 *>(corr + q * 6);
  o[0] = make_double2(a[0], a[1]);
  o[1] = make_double2(a[2], n[0]);
  o[2] = make_double2(n[1], n[2]);
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

// Flat 1-D grid over all queries of the batch: [all corner queries | all surf queries].
__global__ void __launch_bounds__(128)
k_associate_map(GridView gc, GridView gs, KParams kp, int B, const float4 *__restrict__ qc,
                const int32_t *__restrict__ c_off, uint32_t n_corner_total, const float4 *__restrict__ qs,
                const int32_t *__restrict__ s_off, uint32_t n_surf_total, const double *__restrict__ poses,
                double *__restrict__ corr, int32_t *__restrict__ knn_out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_corner_total + n_surf_total) return;
  const bool is_corner = k < n_corner_total;
  const uint32_t kk = is_corner ? k : k - n_corner_total;
  const int scan = find_scan(is_corner ? c_off : s_off, B, kk);
  double pose[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) pose[i] = __ldg(poses + (size_t)scan * 7 + i);
  const size_t q = k;
  const float4 p = __ldg((is_corner ? qc : qs) + kk);
  const float3 x = transform_point_f(pose, p.x, p.y, p.z);  // mapping_scan_matcher.cc:123 / :193
  const GridView &g = is_corner ? gc : gs;
  Top5 t;
  const bool gate = knn5_grid(g, x.x, x.y, x.z, kp.knn_max_sq_f, t);  // :125-128 / :195-198
  if (knn_out) {
#pragma unroll
    for (int s = 0; s < 5; ++s) knn_out[q * 5 + s] = gate ? t.i[s] : -1;
  }
  double a[3] = {0, 0, 0}, n[3] = {0, 0, 0};
  if (gate) {
    double m[5][3];
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      const float4 mp = __ldg(g.pts_orig + t.i[s]);
      m[s][0] = (double)mp.x; m[s][1] = (double)mp.y; m[s][2] = (double)mp.z;
    }
    double c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) c[d] = ((((m[0][d] + m[1][d]) + m[2][d]) + m[3][d]) + m[4][d]) / 5.0;  // :137 / :212
    if (is_corner) {
      double cov[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int s = 0; s < 5; ++s) {  // :138-139 (S = sum (p-c)(p-c)^T, not divided by 5)
        const double e0 = m[s][0] - c[0], e1 = m[s][1] - c[1], e2 = m[s][2] - c[2];
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
    } else {
      double A[5][3], bb[5] = {-1, -1, -1, -1, -1}, nrm[3];
#pragma unroll
      for (int s = 0; s < 5; ++s) { A[s][0] = m[s][0]; A[s][1] = m[s][1]; A[s][2] = m[s][2]; }
      lstsq_5x3(A, bb, nrm);  // :210
      const double nn = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
      nrm[0] /= nn; nrm[1] /= nn; nrm[2] /= nn;  // :211
      bool valid = true;
#pragma unroll
      for (int s = 0; s < 5; ++s) {  // :214-220
        const double dd = nrm[0] * (m[s][0] - c[0]) + nrm[1] * (m[s][1] - c[1]) + nrm[2] * (m[s][2] - c[2]);
        if (!(fabs(dd) <= kp.plane_tol)) valid = false;
      }
      if (valid) {
#pragma unroll
        for (int d = 0; d < 3; ++d) { a[d] = c[d]; n[d] = nrm[d]; }
      }
    }
  }
  store_corr(corr, q, a, n);
}

int launch_associate_map(msfl_engine *e, int B, const float4 *d_qc, const int32_t *d_c_off, uint32_t n_corner_total,
                         const float4 *d_qs, const int32_t *d_s_off, uint32_t n_surf_total, const double *d_poses,
                         double *d_corr, int32_t *d_knn) {
  const uint32_t total = n_corner_total + n_surf_total;
  if (B <= 0 || total == 0) return MSFL_OK;
  const int tb = 128;
  k_associate_map<<<(total + tb - 1) / tb, tb, 0, e->stream>>>(e->map_corner.view, e->map_surf.view, e->kp, B, d_qc, d_c_off,
                                                              n_corner_total, d_qs, d_s_off, n_surf_total, d_poses,
                                                              d_corr, d_knn);
  e->launches += 1;
  MSFL_CUDA_OK(cudaGetLastError());
  return MSFL_OK;
}

}  // namespace msfl
