// stgm.cu -- GPU-resident STGM submap producer: the reference's HybridGrid
// (src/slam/map/hybrid_grid.cc:403-521) as the map that FEEDS scan-to-map:
//   InsertScan (:503-521)          points go to their 3 m cell (index = lround(p / resolution));
//                                  every touched cell's cloud is then re-voxel-filtered
//                                  (pcl::VoxelGrid centroid filter, leaf 0.2 / 0.4) -- old centroids and
//                                  new points together, exactly like the reference;
//   GetSurroundedCloud (:470-501)  cells hit by the scan points (range <= 60 m) moved by +-1 m on every
//                                  axis (float pose), concatenated.
// The map lives in HBM as a CSR: cell keys (sorted), cell offsets, one float4 array.  The surround
// result stays on the device and is handed to msfl_set_submap_device, so the per-frame submap needs no
// H2D copy.  Output order = ascending cell key (z, y, x); the reference's order is the iteration order
// of an unordered_set of shared_ptr (heap-address dependent) -- scan matching does not depend on it.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <utility>

#include "msfl_internal.h"
#include "msfl_math.cuh"

namespace msfl {

constexpr long long kStgmOff = 1ll << 20;

__device__ __forceinline__ unsigned long long stgm_key(float x, float y, float z, float res) {
  // GetCellIndex (:424-428): Array3f index = point / resolution; RoundToInt = lround(double)
  const long long ix = llround((double)__fdiv_rn(x, res)), iy = llround((double)__fdiv_rn(y, res)),
                  iz = llround((double)__fdiv_rn(z, res));
  return ((unsigned long long)(iz + kStgmOff) << 42) | ((unsigned long long)(iy + kStgmOff) << 21) |
         (unsigned long long)(ix + kStgmOff);
}

__device__ __forceinline__ int lower_bound64(const unsigned long long *a, int n, unsigned long long k) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < k) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

struct Pose7d { double v[7]; };

// optional world transform (TransformPointCloud, laser_mapping.cc:24-31) + cell key
__global__ void k_stgm_keys(const float4 *__restrict__ in, uint32_t n, int do_transform, Pose7d T, float res,
                            float4 *__restrict__ world, unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = in[i];
  if (do_transform) {
    const float3 x = transform_point_f(T.v, p.x, p.y, p.z);
    p.x = x.x; p.y = x.y; p.z = x.z;
  }
  world[i] = p;
  keys[i] = stgm_key(p.x, p.y, p.z, res);
  vals[i] = i;
}

__global__ void k_heads64(const unsigned long long *__restrict__ k, uint32_t n, uint32_t *__restrict__ head) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  head[j] = (j == 0 || k[j] != k[j - 1]) ? 1u : 0u;
}

// touched cells: key, first position in the sorted new points
__global__ void k_stgm_touched(const unsigned long long *__restrict__ ks, const uint32_t *__restrict__ head,
                               const uint32_t *__restrict__ pos, uint32_t n, unsigned long long *__restrict__ t_key,
                               uint32_t *__restrict__ t_start) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || !head[j]) return;
  t_key[pos[j]] = ks[j];
  t_start[pos[j]] = j;
}

// per touched cell: index of the existing cell (or -1), number of old + new points
__global__ void k_stgm_lookup(const unsigned long long *__restrict__ t_key, const uint32_t *__restrict__ t_start, uint32_t nt,
                              uint32_t n_new, const unsigned long long *__restrict__ cell_keys,
                              const uint32_t *__restrict__ cell_off, int n_cells, int *__restrict__ t_old,
                              uint32_t *__restrict__ t_total, int *__restrict__ old_touched, uint32_t *__restrict__ is_new) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const uint32_t cnt_new = (t + 1 < nt ? t_start[t + 1] : n_new) - t_start[t];
  int c = n_cells > 0 ? lower_bound64(cell_keys, n_cells, t_key[t]) : 0;
  uint32_t cnt_old = 0;
  if (c < n_cells && cell_keys[c] == t_key[t]) {
    cnt_old = cell_off[c + 1] - cell_off[c];
    old_touched[c] = (int)t;
  } else {
    c = -1;
  }
  t_old[t] = c;
  t_total[t] = cnt_old + cnt_new;
  is_new[t] = c < 0 ? 1u : 0u;
}

// one warp per touched cell: combined = old cell cloud (stored order) ++ new points (scan order);
// also the cell's VoxelGrid bounding box (min_b, div_b)
__global__ void k_stgm_fill(const int *__restrict__ t_old, const uint32_t *__restrict__ t_start, const uint32_t *__restrict__ t_off,
                            uint32_t nt, uint32_t n_new, const float4 *__restrict__ map_pts, const uint32_t *__restrict__ cell_off,
                            const float4 *__restrict__ world, const uint32_t *__restrict__ perm, float inv_leaf,
                            float4 *__restrict__ comb, uint32_t *__restrict__ comb_tc, int *__restrict__ t_minb,
                            int *__restrict__ t_divb) {
  const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= nt) return;
  const int c = t_old[t];
  const uint32_t o0 = c >= 0 ? cell_off[c] : 0, cnt_old = c >= 0 ? cell_off[c + 1] - o0 : 0;
  const uint32_t s0 = t_start[t], cnt_new = (t + 1 < nt ? t_start[t + 1] : n_new) - s0;
  const uint32_t base = t_off[t];
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (uint32_t j = lane; j < cnt_old + cnt_new; j += 32) {
    const float4 p = j < cnt_old ? map_pts[o0 + j] : world[perm[s0 + (j - cnt_old)]];
    comb[base + j] = p;
    comb_tc[base + j] = t;
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d)
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  if (lane < 3) {
    const int lo = (int)floorf(__fmul_rn(mn[lane], inv_leaf)), hi = (int)floorf(__fmul_rn(mx[lane], inv_leaf));
    t_minb[3 * t + lane] = lo;
    t_divb[3 * t + lane] = hi - lo + 1;
  }
}

// (touched cell, voxel index) key of every combined point -- pcl::VoxelGrid::applyFilter indexing
__global__ void k_stgm_voxkeys(const float4 *__restrict__ comb, const uint32_t *__restrict__ comb_tc, uint32_t n, float inv_leaf,
                               const int *__restrict__ t_minb, const int *__restrict__ t_divb,
                               unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q = comb[i];
  const uint32_t t = comb_tc[i];
  const int *mb = t_minb + 3 * t, *db = t_divb + 3 * t;
  const int i0 = (int)__fsub_rn(floorf(__fmul_rn(q.x, inv_leaf)), (float)mb[0]);
  const int i1 = (int)__fsub_rn(floorf(__fmul_rn(q.y, inv_leaf)), (float)mb[1]);
  const int i2 = (int)__fsub_rn(floorf(__fmul_rn(q.z, inv_leaf)), (float)mb[2]);
  const uint32_t idx = (uint32_t)(i0 + i1 * db[0] + i2 * db[0] * db[1]);
  keys[i] = ((unsigned long long)t << 32) | idx;
  vals[i] = i;
}

// centroid of every (cell, voxel) run, fp32 sums in stored order; counts the centroids per touched cell
__global__ void k_stgm_centroids(const float4 *__restrict__ comb, const unsigned long long *__restrict__ ks,
                                 const uint32_t *__restrict__ vals, const uint32_t *__restrict__ head,
                                 const uint32_t *__restrict__ pos, uint32_t n, float4 *__restrict__ filt,
                                 uint32_t *__restrict__ t_fcnt) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || !head[j]) return;
  const unsigned long long key = ks[j];
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  uint32_t k = j;
  for (; k < n && ks[k] == key; ++k) {
    const float4 q = comb[vals[k]];
    sx = __fadd_rn(sx, q.x); sy = __fadd_rn(sy, q.y); sz = __fadd_rn(sz, q.z); si = __fadd_rn(si, q.w);
  }
  const float c = (float)(k - j);
  filt[pos[j]] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), __fdiv_rn(si, c));
  atomicAdd(&t_fcnt[(uint32_t)(key >> 32)], 1u);
}

__global__ void k_stgm_newkeys(const unsigned long long *__restrict__ t_key, const uint32_t *__restrict__ is_new,
                               const uint32_t *__restrict__ new_rank, uint32_t nt, unsigned long long *__restrict__ nk,
                               uint32_t *__restrict__ nk_tid) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt || !is_new[t]) return;
  nk[new_rank[t]] = t_key[t];
  nk_tid[new_rank[t]] = t;
}

// merged cell list: position, key, point count and source of every cell of the new CSR
__global__ void k_stgm_merge(const unsigned long long *__restrict__ old_keys, const uint32_t *__restrict__ old_off, int n_old,
                             const int *__restrict__ old_touched, const unsigned long long *__restrict__ nk,
                             const uint32_t *__restrict__ nk_tid, int n_newcells, const uint32_t *__restrict__ t_fcnt,
                             unsigned long long *__restrict__ out_keys, uint32_t *__restrict__ out_cnt,
                             int *__restrict__ out_src /* >=0: touched id, <0: -(old cell)-1 */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_old) {
    const int pos = i + (n_newcells > 0 ? lower_bound64(nk, n_newcells, old_keys[i]) : 0);
    const int t = old_touched[i];
    out_keys[pos] = old_keys[i];
    out_cnt[pos] = t >= 0 ? t_fcnt[t] : old_off[i + 1] - old_off[i];
    out_src[pos] = t >= 0 ? t : -i - 1;
  } else if (i < n_old + n_newcells) {
    const int r = i - n_old;
    const int pos = r + (n_old > 0 ? lower_bound64(old_keys, n_old, nk[r]) : 0);
    out_keys[pos] = nk[r];
    out_cnt[pos] = t_fcnt[nk_tid[r]];
    out_src[pos] = (int)nk_tid[r];
  }
}

// one warp per cell of the new CSR: copy its points from the old storage or from the filtered buffer
__global__ void k_stgm_copy(const int *__restrict__ src, const uint32_t *__restrict__ new_off, int n_cells,
                            const float4 *__restrict__ old_pts, const uint32_t *__restrict__ old_off,
                            const float4 *__restrict__ filt, const uint32_t *__restrict__ t_foff, float4 *__restrict__ out) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= n_cells) return;
  const int s = src[c];
  const float4 *from = s >= 0 ? filt + t_foff[s] : old_pts + old_off[-s - 1];
  const uint32_t o = new_off[c], cnt = new_off[c + 1] - o;
  for (uint32_t j = lane; j < cnt; j += 32) out[o + j] = from[j];
}

// GetSurroundedCloud (:470-486): flag the cells hit by the scan points moved by (i, j, k) metres
__global__ void k_stgm_mark(const float4 *__restrict__ scan, uint32_t n, Pose7d T, float res,
                            const unsigned long long *__restrict__ cell_keys, int n_cells, uint32_t *__restrict__ flag) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const float4 p = scan[a];
  const float nr = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y)), __fmul_rn(p.z, p.z)));
  if ((double)nr > 60.0) return;  // kDist (:474, :532)
  // pose.cast<float>() * point: Eigen float quaternion rotation + translation, no FMA contraction
  const float qx = (float)T.v[3], qy = (float)T.v[4], qz = (float)T.v[5], qw = (float)T.v[6];
  const float tx = (float)T.v[0], ty = (float)T.v[1], tz = (float)T.v[2];
  float uv0 = __fsub_rn(__fmul_rn(qy, p.z), __fmul_rn(qz, p.y)), uv1 = __fsub_rn(__fmul_rn(qz, p.x), __fmul_rn(qx, p.z)),
        uv2 = __fsub_rn(__fmul_rn(qx, p.y), __fmul_rn(qy, p.x));
  uv0 = __fadd_rn(uv0, uv0); uv1 = __fadd_rn(uv1, uv1); uv2 = __fadd_rn(uv2, uv2);
  const float c0 = __fsub_rn(__fmul_rn(qy, uv2), __fmul_rn(qz, uv1)), c1 = __fsub_rn(__fmul_rn(qz, uv0), __fmul_rn(qx, uv2)),
              c2 = __fsub_rn(__fmul_rn(qx, uv1), __fmul_rn(qy, uv0));
  const float wx = __fadd_rn(__fadd_rn(__fadd_rn(p.x, __fmul_rn(qw, uv0)), c0), tx);
  const float wy = __fadd_rn(__fadd_rn(__fadd_rn(p.y, __fmul_rn(qw, uv1)), c1), ty);
  const float wz = __fadd_rn(__fadd_rn(__fadd_rn(p.z, __fmul_rn(qw, uv2)), c2), tz);
  for (int i = -1; i <= 1; ++i)
    for (int j = -1; j <= 1; ++j)
      for (int k = -1; k <= 1; ++k) {
        const unsigned long long key = stgm_key(__fadd_rn(wx, (float)i), __fadd_rn(wy, (float)j), __fadd_rn(wz, (float)k), res);
        const int c = lower_bound64(cell_keys, n_cells, key);
        if (c < n_cells && cell_keys[c] == key) flag[c] = 1u;  // TryInsertGrid (:524-529)
      }
}

__global__ void k_stgm_selcnt(const uint32_t *__restrict__ flag, const uint32_t *__restrict__ cell_off, int n_cells,
                              uint32_t *__restrict__ cnt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  cnt[c] = flag[c] ? cell_off[c + 1] - cell_off[c] : 0u;
}

__global__ void k_stgm_gather(const uint32_t *__restrict__ flag, const uint32_t *__restrict__ cell_off,
                              const uint32_t *__restrict__ out_off, int n_cells, const float4 *__restrict__ pts,
                              float4 *__restrict__ out) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= n_cells || !flag[c]) return;
  const uint32_t s = cell_off[c], cnt = cell_off[c + 1] - s, o = out_off[c];
  for (uint32_t j = lane; j < cnt; j += 32) out[o + j] = pts[s + j];
}

}  // namespace msfl

using namespace msfl;

struct msfl_map {
  msfl_engine *e = nullptr;
  float resolution = 3.0f, leaf = 0.4f;
  DevBuf cell_keys, cell_off, pts;          // CSR (current)
  DevBuf cell_keys2, cell_off2, pts2;       // CSR (next, swapped in by insert)
  size_t n_cells = 0, n_points = 0;
  DevBuf sur;                               // last surround result
  size_t sur_n = 0;
  // scratch
  DevBuf in, world, k64, k64b, v32, v32b, c64, c64b, cv32, cv32b, chead, cpos, tmp, head, pos, t_key, t_start, t_old, t_total, t_off, old_touched, is_new,
      new_rank, comb, comb_tc, t_minb, t_divb, filt, t_fcnt, t_foff, nk, nk_tid, out_cnt, out_src, flag, selcnt, seloff;
  void release_all() {
    DevBuf *b[] = {&cell_keys, &cell_off, &pts, &cell_keys2, &cell_off2, &pts2, &sur, &in, &world, &k64, &k64b, &v32, &v32b,
                   &c64, &c64b, &cv32, &cv32b, &chead, &cpos, &tmp, &head, &pos, &t_key, &t_start, &t_old, &t_total, &t_off, &old_touched, &is_new, &new_rank, &comb,
                   &comb_tc, &t_minb, &t_divb, &filt, &t_fcnt, &t_foff, &nk, &nk_tid, &out_cnt, &out_src, &flag, &selcnt,
                   &seloff};
    for (auto *x : b) x->release();
  }
};

static int upload_packed(msfl_engine *e, const msfl_cloud *c, DevBuf &dst) {
  const size_t n = c->n;
  int rc;
  if ((rc = e->h_stage.reserve(n * 16 + 16))) return rc;
  if ((rc = dst.reserve(n * 16 + 16))) return rc;
  const uint32_t off0 = 0;
  pack_clouds_parallel(e, 1, c, e->h_stage.as<float>(), nullptr, &off0);
  MSFL_CUDA_OK(cudaMemcpyAsync(dst.p, e->h_stage.p, n * 16, cudaMemcpyHostToDevice, e->stream));
  return MSFL_OK;
}

static int scan_u32(msfl_map *m, const uint32_t *in, uint32_t *out, int n) {
  size_t bytes = 0;
  MSFL_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, m->e->stream));
  int rc;
  if ((rc = m->tmp.reserve(bytes + 16))) return rc;
  MSFL_CUDA_OK(cub::DeviceScan::ExclusiveSum(m->tmp.p, bytes, in, out, n, m->e->stream));
  return MSFL_OK;
}

static int sort_pairs64(msfl_map *m, unsigned long long *k, unsigned long long *k2, uint32_t *v, uint32_t *v2, int n, int end_bit,
                        const unsigned long long **ks, const uint32_t **vs) {
  cub::DoubleBuffer<unsigned long long> dk(k, k2);
  cub::DoubleBuffer<uint32_t> dv(v, v2);
  size_t bytes = 0;
  MSFL_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, n, 0, end_bit, m->e->stream));
  int rc;
  if ((rc = m->tmp.reserve(bytes + 16))) return rc;
  MSFL_CUDA_OK(cub::DeviceRadixSort::SortPairs(m->tmp.p, bytes, dk, dv, n, 0, end_bit, m->e->stream));
  *ks = dk.Current();
  *vs = dv.Current();
  return MSFL_OK;
}

static int last_plus(msfl_map *m, const uint32_t *a, const uint32_t *b, uint32_t n, uint32_t *out) {
  // out = a[n-1] + b[n-1]  (exclusive scan total), one small D2H
  uint32_t h[2];
  MSFL_CUDA_OK(cudaMemcpyAsync(&h[0], a + (n - 1), 4, cudaMemcpyDeviceToHost, m->e->stream));
  MSFL_CUDA_OK(cudaMemcpyAsync(&h[1], b + (n - 1), 4, cudaMemcpyDeviceToHost, m->e->stream));
  MSFL_CUDA_OK(cudaStreamSynchronize(m->e->stream));
  *out = h[0] + h[1];
  return MSFL_OK;
}

static int map_insert_device(msfl_map *m, uint32_t n, int do_transform, const double pose[7]) {
  msfl_engine *e = m->e;
  cudaStream_t st = e->stream;
  const int tb = 256;
  int rc;
#define RSV(buf, bytes) do { if ((rc = m->buf.reserve(bytes))) return rc; } while (0)
  RSV(world, (size_t)n * 16); RSV(k64, (size_t)n * 8); RSV(k64b, (size_t)n * 8); RSV(v32, (size_t)n * 4); RSV(v32b, (size_t)n * 4);
  RSV(head, (size_t)n * 4); RSV(pos, (size_t)n * 4);
  Pose7d T;
  for (int i = 0; i < 7; ++i) T.v[i] = pose ? pose[i] : (i == 6 ? 1.0 : 0.0);
  // 1. (transform +) cell key of every new point; stable sort by cell keeps the scan order inside a cell
  k_stgm_keys<<<(n + tb - 1) / tb, tb, 0, st>>>(m->in.as<float4>(), n, do_transform, T, m->resolution, m->world.as<float4>(),
                                               m->k64.as<unsigned long long>(), m->v32.as<uint32_t>());
  const unsigned long long *ks;
  const uint32_t *perm;
  if ((rc = sort_pairs64(m, m->k64.as<unsigned long long>(), m->k64b.as<unsigned long long>(), m->v32.as<uint32_t>(),
                         m->v32b.as<uint32_t>(), (int)n, 63, &ks, &perm)))
    return rc;
  // 2. touched cells
  k_heads64<<<(n + tb - 1) / tb, tb, 0, st>>>(ks, n, m->head.as<uint32_t>());
  if ((rc = scan_u32(m, m->head.as<uint32_t>(), m->pos.as<uint32_t>(), (int)n))) return rc;
  uint32_t nt = 0;
  if ((rc = last_plus(m, m->head.as<uint32_t>(), m->pos.as<uint32_t>(), n, &nt))) return rc;
  const int n_old = (int)m->n_cells;
  RSV(t_key, (size_t)nt * 8); RSV(t_start, (size_t)nt * 4 + 4); RSV(t_old, (size_t)nt * 4); RSV(t_total, (size_t)nt * 4);
  RSV(t_off, (size_t)nt * 4 + 4); RSV(is_new, (size_t)nt * 4); RSV(new_rank, (size_t)nt * 4); RSV(t_minb, (size_t)nt * 12);
  RSV(t_divb, (size_t)nt * 12); RSV(t_fcnt, (size_t)nt * 4); RSV(t_foff, (size_t)nt * 4 + 4);
  RSV(nk, (size_t)nt * 8); RSV(nk_tid, (size_t)nt * 4);
  RSV(old_touched, (size_t)(n_old + 1) * 4);
  MSFL_CUDA_OK(cudaMemsetAsync(m->old_touched.p, 0xff, (size_t)(n_old + 1) * 4, st));
  MSFL_CUDA_OK(cudaMemsetAsync(m->t_fcnt.p, 0, (size_t)nt * 4, st));
  k_stgm_touched<<<(n + tb - 1) / tb, tb, 0, st>>>(ks, m->head.as<uint32_t>(), m->pos.as<uint32_t>(), n,
                                                  m->t_key.as<unsigned long long>(), m->t_start.as<uint32_t>());
  k_stgm_lookup<<<(nt + tb - 1) / tb, tb, 0, st>>>(m->t_key.as<unsigned long long>(), m->t_start.as<uint32_t>(), nt, n,
                                                  m->cell_keys.as<unsigned long long>(), m->cell_off.as<uint32_t>(), n_old,
                                                  m->t_old.as<int>(), m->t_total.as<uint32_t>(), m->old_touched.as<int>(),
                                                  m->is_new.as<uint32_t>());
  if ((rc = scan_u32(m, m->t_total.as<uint32_t>(), m->t_off.as<uint32_t>(), (int)nt))) return rc;
  if ((rc = scan_u32(m, m->is_new.as<uint32_t>(), m->new_rank.as<uint32_t>(), (int)nt))) return rc;
  uint32_t ncomb = 0, n_newcells = 0;
  if ((rc = last_plus(m, m->t_total.as<uint32_t>(), m->t_off.as<uint32_t>(), nt, &ncomb))) return rc;
  if ((rc = last_plus(m, m->is_new.as<uint32_t>(), m->new_rank.as<uint32_t>(), nt, &n_newcells))) return rc;
  // 3. per touched cell: old cloud ++ new points, VoxelGrid over the lot
  RSV(comb, (size_t)ncomb * 16); RSV(comb_tc, (size_t)ncomb * 4); RSV(filt, (size_t)ncomb * 16);
  RSV(c64, (size_t)ncomb * 8); RSV(c64b, (size_t)ncomb * 8); RSV(cv32, (size_t)ncomb * 4); RSV(cv32b, (size_t)ncomb * 4);
  RSV(chead, (size_t)ncomb * 4); RSV(cpos, (size_t)ncomb * 4);
  const float inv_leaf = 1.0f / m->leaf;
  k_stgm_fill<<<(nt * 32 + tb - 1) / tb, tb, 0, st>>>(m->t_old.as<int>(), m->t_start.as<uint32_t>(), m->t_off.as<uint32_t>(), nt, n,
                                                     m->pts.as<float4>(), m->cell_off.as<uint32_t>(), m->world.as<float4>(), perm,
                                                     inv_leaf, m->comb.as<float4>(), m->comb_tc.as<uint32_t>(),
                                                     m->t_minb.as<int>(), m->t_divb.as<int>());
  k_stgm_voxkeys<<<(ncomb + tb - 1) / tb, tb, 0, st>>>(m->comb.as<float4>(), m->comb_tc.as<uint32_t>(), ncomb, inv_leaf,
                                                      m->t_minb.as<int>(), m->t_divb.as<int>(),
                                                      m->c64.as<unsigned long long>(), m->cv32.as<uint32_t>());
  const unsigned long long *cks;
  const uint32_t *cperm;
  if ((rc = sort_pairs64(m, m->c64.as<unsigned long long>(), m->c64b.as<unsigned long long>(), m->cv32.as<uint32_t>(),
                         m->cv32b.as<uint32_t>(), (int)ncomb, 64, &cks, &cperm)))
    return rc;
  k_heads64<<<(ncomb + tb - 1) / tb, tb, 0, st>>>(cks, ncomb, m->chead.as<uint32_t>());
  if ((rc = scan_u32(m, m->chead.as<uint32_t>(), m->cpos.as<uint32_t>(), (int)ncomb))) return rc;
  k_stgm_centroids<<<(ncomb + tb - 1) / tb, tb, 0, st>>>(m->comb.as<float4>(), cks, cperm, m->chead.as<uint32_t>(),
                                                        m->cpos.as<uint32_t>(), ncomb, m->filt.as<float4>(),
                                                        m->t_fcnt.as<uint32_t>());
  if ((rc = scan_u32(m, m->t_fcnt.as<uint32_t>(), m->t_foff.as<uint32_t>(), (int)nt))) return rc;
  // 4. merged CSR: untouched cells keep their cloud, touched cells take the filtered one
  const int n_cells_new = n_old + (int)n_newcells;
  RSV(cell_keys2, (size_t)n_cells_new * 8); RSV(cell_off2, (size_t)(n_cells_new + 1) * 4);
  RSV(out_cnt, (size_t)(n_cells_new + 1) * 4); RSV(out_src, (size_t)n_cells_new * 4);
  k_stgm_newkeys<<<(nt + tb - 1) / tb, tb, 0, st>>>(m->t_key.as<unsigned long long>(), m->is_new.as<uint32_t>(),
                                                   m->new_rank.as<uint32_t>(), nt, m->nk.as<unsigned long long>(),
                                                   m->nk_tid.as<uint32_t>());
  MSFL_CUDA_OK(cudaMemsetAsync(m->out_cnt.p, 0, (size_t)(n_cells_new + 1) * 4, st));
  k_stgm_merge<<<(n_cells_new + tb - 1) / tb, tb, 0, st>>>(m->cell_keys.as<unsigned long long>(), m->cell_off.as<uint32_t>(), n_old,
                                                          m->old_touched.as<int>(), m->nk.as<unsigned long long>(),
                                                          m->nk_tid.as<uint32_t>(), (int)n_newcells, m->t_fcnt.as<uint32_t>(),
                                                          m->cell_keys2.as<unsigned long long>(), m->out_cnt.as<uint32_t>(),
                                                          m->out_src.as<int>());
  if ((rc = scan_u32(m, m->out_cnt.as<uint32_t>(), m->cell_off2.as<uint32_t>(), n_cells_new + 1))) return rc;
  uint32_t n_points_new = 0;
  MSFL_CUDA_OK(cudaMemcpyAsync(&n_points_new, m->cell_off2.as<uint32_t>() + n_cells_new, 4, cudaMemcpyDeviceToHost, st));
  MSFL_CUDA_OK(cudaStreamSynchronize(st));
  RSV(pts2, (size_t)n_points_new * 16 + 16);
  k_stgm_copy<<<(n_cells_new * 32 + tb - 1) / tb, tb, 0, st>>>(m->out_src.as<int>(), m->cell_off2.as<uint32_t>(), n_cells_new,
                                                              m->pts.as<float4>(), m->cell_off.as<uint32_t>(),
                                                              m->filt.as<float4>(), m->t_foff.as<uint32_t>(),
                                                              m->pts2.as<float4>());
  e->launches += 12 + 12;
  MSFL_CUDA_OK(cudaGetLastError());
  MSFL_CUDA_OK(cudaStreamSynchronize(st));
  std::swap(m->cell_keys, m->cell_keys2);
  std::swap(m->cell_off, m->cell_off2);
  std::swap(m->pts, m->pts2);
  m->n_cells = (size_t)n_cells_new;
  m->n_points = n_points_new;
#undef RSV
  return MSFL_OK;
}

// HybridGrid::GetSurroundedCloud with the scan already in m->in (n packed points)
static int map_surround_device(msfl_map *m, uint32_t n, const double pose_tq[7], size_t *n_out) {
  msfl_engine *e = m->e;
  cudaStream_t st = e->stream;
  m->sur_n = 0;
  if (n_out) *n_out = 0;
  if (m->n_cells == 0 || n == 0) return MSFL_OK;
  int rc;
  const int nc = (int)m->n_cells, tb = 256;
  if ((rc = m->flag.reserve((size_t)nc * 4))) return rc;
  if ((rc = m->selcnt.reserve((size_t)(nc + 1) * 4))) return rc;
  if ((rc = m->seloff.reserve((size_t)(nc + 1) * 4))) return rc;
  MSFL_CUDA_OK(cudaMemsetAsync(m->flag.p, 0, (size_t)nc * 4, st));
  MSFL_CUDA_OK(cudaMemsetAsync(m->selcnt.p, 0, (size_t)(nc + 1) * 4, st));
  Pose7d T;
  for (int i = 0; i < 7; ++i) T.v[i] = pose_tq[i];
  k_stgm_mark<<<(n + tb - 1) / tb, tb, 0, st>>>(m->in.as<float4>(), n, T, m->resolution, m->cell_keys.as<unsigned long long>(), nc,
                                               m->flag.as<uint32_t>());
  k_stgm_selcnt<<<(nc + tb - 1) / tb, tb, 0, st>>>(m->flag.as<uint32_t>(), m->cell_off.as<uint32_t>(), nc, m->selcnt.as<uint32_t>());
  if ((rc = scan_u32(m, m->selcnt.as<uint32_t>(), m->seloff.as<uint32_t>(), nc + 1))) return rc;
  uint32_t total = 0;
  MSFL_CUDA_OK(cudaMemcpyAsync(&total, m->seloff.as<uint32_t>() + nc, 4, cudaMemcpyDeviceToHost, st));
  MSFL_CUDA_OK(cudaStreamSynchronize(st));
  if ((rc = m->sur.reserve((size_t)total * 16 + 16))) return rc;
  k_stgm_gather<<<(nc * 32 + tb - 1) / tb, tb, 0, st>>>(m->flag.as<uint32_t>(), m->cell_off.as<uint32_t>(), m->seloff.as<uint32_t>(),
                                                       nc, m->pts.as<float4>(), m->sur.as<float4>());
  e->launches += 3 + 2;
  MSFL_CUDA_OK(cudaGetLastError());
  m->sur_n = total;
  if (n_out) *n_out = total;
  return MSFL_OK;
}

extern "C" {

int msfl_map_create(msfl_engine *e, float resolution, float leaf, msfl_map **out) {
  if (!e || !out || !(resolution > 0) || !(leaf > 0)) { set_error("msfl_map_create: bad argument"); return MSFL_ERR_ARG; }
  msfl_map *m = new msfl_map();
  m->e = e;
  m->resolution = resolution;
  m->leaf = leaf;
  *out = m;
  return MSFL_OK;
}

void msfl_map_destroy(msfl_map *m) {
  if (!m) return;
  cudaSetDevice(m->e->device);
  cudaStreamSynchronize(m->e->stream);
  m->release_all();
  delete m;
}

int msfl_map_insert(msfl_map *m, const msfl_cloud *scan, const double pose_tq[7]) {
  if (!m || !scan) { set_error("msfl_map_insert: bad argument"); return MSFL_ERR_ARG; }
  if (scan->n == 0) return MSFL_OK;  // hybrid_grid.cc:504
  int rc;
  if ((rc = check_cloud(scan, false, "msfl_map_insert"))) return rc;
  MSFL_CUDA_OK(cudaSetDevice(m->e->device));
  if ((rc = upload_packed(m->e, scan, m->in))) return rc;
  return map_insert_device(m, (uint32_t)scan->n, pose_tq != nullptr, pose_tq);
}

int msfl_map_size(const msfl_map *m, size_t *n_points, size_t *n_cells) {
  if (!m) return MSFL_ERR_ARG;
  if (n_points) *n_points = m->n_points;
  if (n_cells) *n_cells = m->n_cells;
  return MSFL_OK;
}

int msfl_map_surround(msfl_map *m, const msfl_cloud *scan, const double pose_tq[7], size_t *n_out) {
  if (!m || !scan || !pose_tq) { set_error("msfl_map_surround: bad argument"); return MSFL_ERR_ARG; }
  msfl_engine *e = m->e;
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  m->sur_n = 0;
  if (n_out) *n_out = 0;
  if (m->n_cells == 0 || scan->n == 0) return MSFL_OK;
  int rc;
  if ((rc = check_cloud(scan, false, "msfl_map_surround"))) return rc;
  if ((rc = upload_packed(e, scan, m->in))) return rc;
  return map_surround_device(m, (uint32_t)scan->n, pose_tq, n_out);
}

int msfl_map_download(msfl_map *m, int which, float *out_xyzi, size_t capacity, size_t *n_out) {
  if (!m || !out_xyzi) { set_error("msfl_map_download: bad argument"); return MSFL_ERR_ARG; }
  MSFL_CUDA_OK(cudaSetDevice(m->e->device));
  const size_t n = which == 0 ? m->sur_n : m->n_points;
  const void *src = which == 0 ? m->sur.p : m->pts.p;
  if (n_out) *n_out = n;
  if (n > capacity) { set_error("msfl_map_download: capacity %zu < %zu points", capacity, n); return MSFL_ERR_ARG; }
  if (n) MSFL_CUDA_OK(cudaMemcpyAsync(out_xyzi, src, n * 16, cudaMemcpyDeviceToHost, m->e->stream));
  MSFL_CUDA_OK(cudaStreamSynchronize(m->e->stream));
  return MSFL_OK;
}

// One frame of LaserMapping (laser_mapping.cc:258-340) with every intermediate on the device: the two feature clouds are
// uploaded once and serve as surround queries, VoxelGrid inputs and inserted scans.
int msfl_mapping_frame(msfl_engine *e, msfl_map *map_corner, msfl_map *map_surf, const msfl_cloud *corner_less_sharp,
                       const msfl_cloud *surf_less_flat, double pose_tq[7], int32_t *matched, msfl_stats *stats) {
  if (!e || !map_corner || !map_surf || !corner_less_sharp || !surf_less_flat || !pose_tq) { set_error("msfl_mapping_frame: bad argument"); return MSFL_ERR_ARG; }
  if (map_corner->e != e || map_surf->e != e) { set_error("msfl_mapping_frame: the maps belong to another engine"); return MSFL_ERR_ARG; }
  if (matched) *matched = 0;
  if (stats) memset(stats, 0, sizeof *stats);
  int rc;
  if ((rc = check_cloud(corner_less_sharp, false, "msfl_mapping_frame corner_less_sharp"))) return rc;
  if ((rc = check_cloud(surf_less_flat, false, "msfl_mapping_frame surf_less_flat"))) return rc;
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  const size_t nc = corner_less_sharp->n, ns = surf_less_flat->n;
  // 1. both clouds to the device, once (one staging buffer: the two copies are in flight together)
  if ((rc = e->h_stage.reserve((nc + ns) * 16 + 16))) return rc;
  {
    const msfl_cloud cl[2] = {*corner_less_sharp, *surf_less_flat};
    const uint32_t off[2] = {0u, (uint32_t)nc};
    float *h = e->h_stage.as<float>();
    pack_clouds_parallel(e, 2, cl, h, nullptr, off);
    if ((rc = map_corner->in.reserve(nc * 16 + 16))) return rc;
    if ((rc = map_surf->in.reserve(ns * 16 + 16))) return rc;
    if (nc) MSFL_CUDA_OK(cudaMemcpyAsync(map_corner->in.p, h, nc * 16, cudaMemcpyHostToDevice, st));
    if (ns) MSFL_CUDA_OK(cudaMemcpyAsync(map_surf->in.p, h + 4 * nc, ns * 16, cudaMemcpyHostToDevice, st));
  }
  // 2. GetSurroundedCloud of both maps at the incoming pose (:273-278)
  size_t m_c = 0, m_s = 0;
  if ((rc = map_surround_device(map_corner, (uint32_t)nc, pose_tq, &m_c))) return rc;
  if ((rc = map_surround_device(map_surf, (uint32_t)ns, pose_tq, &m_s))) return rc;
  // 3. the gate (:284-285), the scan's VoxelGrid (:264-270, the maps' own leaf sizes: one filter object serves both uses
  //    upstream) and MatchScan2Map (:304-311) on the device-resident clouds
  if (m_c > 10 && m_s > 50) {
    if ((rc = e->fr_qc.reserve(nc * 16 + 16))) return rc;
    if ((rc = e->fr_qs.reserve(ns * 16 + 16))) return rc;
    size_t nqc = 0, nqs = 0;
    if (nc && (rc = run_voxel_grid(e, map_corner->in.as<float4>(), nc, map_corner->leaf, e->fr_qc.as<float4>(), &nqc))) return rc;
    if (ns && (rc = run_voxel_grid(e, map_surf->in.as<float4>(), ns, map_surf->leaf, e->fr_qs.as<float4>(), &nqs))) return rc;
    if ((rc = msfl_set_submap_device(e, map_corner->sur.as<float>(), m_c, map_surf->sur.as<float>(), m_s))) return rc;
    if (nqc + nqs > 0) {
      // offsets (2 x 2 int32) + pose (7 doubles) + stats in one small device block
      if ((rc = e->fr_misc.reserve(16 + 56 + sizeof(msfl_stats) + 16))) return rc;
      if ((rc = e->h_misc.reserve(16 + 56 + sizeof(msfl_stats) + 16))) return rc;
      char *hm = e->h_misc.as<char>();
      int32_t *hoff = (int32_t *)hm;
      hoff[0] = 0; hoff[1] = (int32_t)nqc; hoff[2] = 0; hoff[3] = (int32_t)nqs;
      memcpy(hm + 16, pose_tq, 56);
      MSFL_CUDA_OK(cudaMemcpyAsync(e->fr_misc.p, hm, 16 + 56, cudaMemcpyHostToDevice, st));
      char *dm = e->fr_misc.as<char>();
      double *d_pose = (double *)(dm + 16);
      msfl_stats *d_stats = nullptr;
      if (stats) {
        d_stats = (msfl_stats *)(dm + 16 + 56 + 8);
        MSFL_CUDA_OK(cudaMemsetAsync(d_stats, 0, sizeof(msfl_stats), st));
      }
      if ((rc = scan2map_enqueue(e, 1, e->fr_qc.as<float4>(), (const int32_t *)dm, (uint32_t)nqc, e->fr_qs.as<float4>(),
                                 (const int32_t *)dm + 2, (uint32_t)nqs, d_pose, d_stats)))
        return rc;
      MSFL_CUDA_OK(cudaMemcpyAsync(hm + 16, d_pose, 56, cudaMemcpyDeviceToHost, st));
      if (stats) MSFL_CUDA_OK(cudaMemcpyAsync(hm + 16 + 56 + 8, d_stats, sizeof(msfl_stats), cudaMemcpyDeviceToHost, st));
      MSFL_CUDA_OK(cudaStreamSynchronize(st));
      memcpy(pose_tq, hm + 16, 56);
      if (stats) memcpy(stats, hm + 16 + 56 + 8, sizeof(msfl_stats));
    }
    if (matched) *matched = 1;
  }
  // 4. InsertScan2Map (:330-338): the un-down-sampled clouds at the refined pose, whether or not the gate passed
  if (nc && (rc = map_insert_device(map_corner, (uint32_t)nc, 1, pose_tq))) return rc;
  if (ns && (rc = map_insert_device(map_surf, (uint32_t)ns, 1, pose_tq))) return rc;
  return MSFL_OK;
}

int msfl_set_submap_from_maps(msfl_engine *e, msfl_map *corner, msfl_map *surf) {
  if (!e || !corner || !surf) { set_error("msfl_set_submap_from_maps: bad argument"); return MSFL_ERR_ARG; }
  if (corner->sur_n == 0 || surf->sur_n == 0) { set_error("msfl_set_submap_from_maps: empty surround cloud"); return MSFL_ERR_ARG; }
  // the surround gathers are enqueued (unsynchronised) on the maps' own engine stream: only that engine's stream is
  // ordered behind them
  if (corner->e != e || surf->e != e) { set_error("msfl_set_submap_from_maps: the maps belong to another engine"); return MSFL_ERR_ARG; }
  return msfl_set_submap_device(e, corner->sur.as<float>(), corner->sur_n, surf->sur.as<float>(), surf->sur_n);
}

}  // extern "C"
