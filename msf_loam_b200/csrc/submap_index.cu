// submap_index.cu -- dense sorted-cell index over a submap class; replaces the two
// pcl::KdTreeFLANN::setInputCloud builds of mapping_scan_matcher.cc:66-72 (rebuilt every frame in
// the reference, once per submap version here).
//
// Layout in HBM (per class):  pts_orig float4[M] (caller order), pts_sorted float4[M] (sorted by
// linear cell id, original index in .w), cell_start uint32[ncell+1].  Cell edge = search radius
// (sqrt(knn_max_sq) = 1 m), so the d5^2 < 1 gate makes a 3x3x3 neighbourhood search exact.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <limits.h>

#include "msfl_internal.h"

namespace msfl {

__global__ void k_cell_bounds(const float4 *__restrict__ pts, uint32_t n, float inv_edge, int *bounds /* 6 + flag */) {
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  int bad = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) || fabsf(p.x) > 1e6f || fabsf(p.y) > 1e6f || fabsf(p.z) > 1e6f) {
      bad = 1;
      continue;
    }
    const int c[3] = {(int)floorf(p.x * inv_edge), (int)floorf(p.y * inv_edge), (int)floorf(p.z * inv_edge)};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lo[d] = min(lo[d], c[d]);
      hi[d] = max(hi[d], c[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  }
  bad = __any_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (lo[d] != INT_MAX) atomicMin(&bounds[d], lo[d]);
      if (hi[d] != INT_MIN) atomicMax(&bounds[3 + d], hi[d]);
    }
    if (bad) atomicOr(&bounds[6], 1);
  }
}

__global__ void k_init_bounds(int *bounds) {
  if (threadIdx.x < 3) bounds[threadIdx.x] = INT_MAX;
  else if (threadIdx.x < 6) bounds[threadIdx.x] = INT_MIN;
  else if (threadIdx.x == 6) bounds[6] = 0;
}

__global__ void k_cell_keys(const float4 *__restrict__ pts, uint32_t n, float inv_edge, int ox, int oy, int oz, int nx,
                            int ny, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  const int cx = (int)floorf(p.x * inv_edge) - ox, cy = (int)floorf(p.y * inv_edge) - oy,
            cz = (int)floorf(p.z * inv_edge) - oz;
  keys[i] = (uint32_t)((cz * ny + cy) * nx + cx);
  vals[i] = i;
}

__global__ void k_gather_sorted(const float4 *__restrict__ pts, const uint32_t *__restrict__ vals, uint32_t n,
                                float4 *__restrict__ sorted) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t src = vals[j];
  float4 p = pts[src];
  p.w = __int_as_float((int)src);
  sorted[j] = p;
}

// cell_start[c] = first sorted position whose key >= c  (c in [0, ncell])
__global__ void k_cell_start(const uint32_t *__restrict__ keys_sorted, uint32_t n, uint32_t ncell,
                             uint32_t *__restrict__ cell_start) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > ncell) return;
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (keys_sorted[mid] < c) lo = mid + 1;
    else hi = mid;
  }
  cell_start[c] = lo;
}

// ---- counting-sort build (host-known grid: no device round trip) --------------------------------------------------
// one atomic per point returns its rank inside its cell; after an exclusive scan of the cell counts one scatter puts the
// point at cell_start[cell] + rank.  The order inside a cell follows the atomics, which the 5-NN does not see: its
// candidates are ordered by (distance, original index), whatever order they arrive in.
__global__ void k_cell_count(const float4 *__restrict__ pts, uint32_t n, float inv_edge, int ox, int oy, int oz, int nx, int ny,
                             uint32_t *__restrict__ count, uint32_t *__restrict__ keys, uint32_t *__restrict__ rank) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  const int cx = (int)floorf(p.x * inv_edge) - ox, cy = (int)floorf(p.y * inv_edge) - oy, cz = (int)floorf(p.z * inv_edge) - oz;
  const uint32_t key = (uint32_t)((cz * ny + cy) * nx + cx);
  keys[i] = key;
  rank[i] = atomicAdd(count + key, 1u);
}

__global__ void k_cell_scatter(const float4 *__restrict__ pts, uint32_t n, const uint32_t *__restrict__ keys,
                               const uint32_t *__restrict__ rank, const uint32_t *__restrict__ cell_start, float4 *__restrict__ sorted) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  p.w = __int_as_float((int)i);
  sorted[__ldg(cell_start + __ldg(keys + i)) + __ldg(rank + i)] = p;
}

// Cell index over points already in m.orig, with the cell bounding box known on the host (msfl_set_submap computes it
// while it repacks the caller's cloud): five stream operations, no synchronisation, no library sort.
int submap_build_host_bounds(msfl_engine *e, Submap &m, size_t n, float edge, const int lo[3], const int hi[3]) {
  cudaStream_t st = e->stream;
  if (n == 0 || n > 0x7fffffffull) { set_error("submap class is empty or too large (n=%zu)", n); return MSFL_ERR_ARG; }
  const uint32_t N = (uint32_t)n;
  const float inv_edge = 1.0f / edge;
  const long long nx = (long long)hi[0] - lo[0] + 5, ny = (long long)hi[1] - lo[1] + 5, nz = (long long)hi[2] - lo[2] + 5;
  const long long ncell = nx * ny * nz;
  if (ncell > (1ll << 26)) {
    set_error("submap bounding box needs %lld cells (> 2^26); dense cell index refused", ncell);
    return MSFL_ERR_GRID;
  }
  int rc;
  if ((rc = m.sorted.reserve(n * sizeof(float4)))) return rc;
  if ((rc = m.keys.reserve(n * 4))) return rc;
  if ((rc = m.vals.reserve(n * 4))) return rc;
  if ((rc = m.cell_start.reserve((size_t)(ncell + 1) * 4))) return rc;
  size_t tmp_bytes = 0;
  uint32_t *cs = m.cell_start.as<uint32_t>();
  MSFL_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cs, cs, (int)(ncell + 1), st));
  if ((rc = m.cub_tmp.reserve(tmp_bytes))) return rc;
  const int ox = lo[0] - 2, oy = lo[1] - 2, oz = lo[2] - 2;
  const int tb = 256;
  const float4 *pts = m.orig.as<float4>();
  MSFL_CUDA_OK(cudaMemsetAsync(cs, 0, (size_t)(ncell + 1) * 4, st));
  k_cell_count<<<(N + tb - 1) / tb, tb, 0, st>>>(pts, N, inv_edge, ox, oy, oz, (int)nx, (int)ny, cs, m.keys.as<uint32_t>(), m.vals.as<uint32_t>());
  MSFL_CUDA_OK(cub::DeviceScan::ExclusiveSum(m.cub_tmp.p, tmp_bytes, cs, cs, (int)(ncell + 1), st));
  k_cell_scatter<<<(N + tb - 1) / tb, tb, 0, st>>>(pts, N, m.keys.as<uint32_t>(), m.vals.as<uint32_t>(), cs, m.sorted.as<float4>());
  e->launches += 2 + 2;
  MSFL_CUDA_OK(cudaGetLastError());
  m.n = n;
  m.view.pts_sorted = m.sorted.as<float4>();
  m.view.pts_orig = pts;
  m.view.cell_start = cs;
  m.view.nx = (int)nx; m.view.ny = (int)ny; m.view.nz = (int)nz;
  m.view.ox = ox; m.view.oy = oy; m.view.oz = oz;
  m.view.inv_edge = inv_edge;
  m.view.n = N;
  return MSFL_OK;
}

// Row occupancy of every cell's 3x3x3 neighbourhood: most of the nine (dy, dz) rows around a query are EMPTY in a map
// of surfaces (a floor patch fills three rows of nine), and an empty row still costs the search its four cell-table
// loads and bounds.  Bit r follows knn5_grid's nearest-first row order.
__global__ void k_row_mask(const uint32_t *__restrict__ cell_start, int nx, int ny, int nz, uint16_t *__restrict__ mask) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ncell = (long long)nx * ny * nz;
  if (c >= ncell) return;
  const int cx = (int)(c % nx), cy = (int)((c / nx) % ny), cz = (int)(c / ((long long)nx * ny));
  uint32_t m = 0;
  if (cx >= 1 && cy >= 1 && cz >= 1 && cx <= nx - 2 && cy <= ny - 2 && cz <= nz - 2) {
    constexpr uint32_t kDyPacked = 1u | (0u << 2) | (2u << 4) | (1u << 6) | (1u << 8) | (0u << 10) | (0u << 12) | (2u << 14) | (2u << 16);
    constexpr uint32_t kDzPacked = 1u | (1u << 2) | (1u << 4) | (0u << 6) | (2u << 8) | (0u << 10) | (2u << 12) | (0u << 14) | (2u << 16);
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int dy = (int)((kDyPacked >> (2 * r)) & 3u) - 1, dz = (int)((kDzPacked >> (2 * r)) & 3u) - 1;
      const long long row = ((long long)(cz + dz) * ny + (cy + dy)) * nx + cx;
      if (__ldg(cell_start + row + 2) > __ldg(cell_start + row - 1)) m |= 1u << r;
    }
  }
  mask[c] = (uint16_t)m;
}

int submap_row_mask(msfl_engine *e, Submap &m) {
  m.view.row_mask = nullptr;
  const long long ncell = (long long)m.view.nx * m.view.ny * m.view.nz;
  if (ncell > (1ll << 24)) return MSFL_OK;  // far-spread grid: the table would cost more than it saves
  int rc;
  if ((rc = m.row_mask.reserve((size_t)ncell * 2))) return rc;
  k_row_mask<<<(unsigned)((ncell + 255) / 256), 256, 0, e->stream>>>(m.view.cell_start, m.view.nx, m.view.ny, m.view.nz,
                                                                    m.row_mask.as<uint16_t>());
  e->launches += 1;
  MSFL_CUDA_OK(cudaGetLastError());
  m.view.row_mask = m.row_mask.as<uint16_t>();
  return MSFL_OK;
}

void submap_release(Submap &m) {
  m.orig.release(); m.sorted.release(); m.cell_start.release(); m.keys.release(); m.keys_alt.release();
  m.vals.release(); m.vals_alt.release(); m.cub_tmp.release(); m.bounds.release(); m.row_mask.release();
  m.n = 0;
}

int submap_build(msfl_engine *e, Submap &m, const float4 *d_pts, size_t n, float edge, bool want_row_mask) {
  cudaStream_t st = e->stream;
  if (n == 0 || n > 0x7fffffffull) { set_error("submap class is empty or too large (n=%zu)", n); return MSFL_ERR_ARG; }
  const uint32_t N = (uint32_t)n;
  const float inv_edge = 1.0f / edge;
  int rc;
  if ((rc = m.orig.reserve(n * sizeof(float4)))) return rc;
  if ((rc = m.sorted.reserve(n * sizeof(float4)))) return rc;
  if ((rc = m.keys.reserve(n * 4))) return rc;
  if ((rc = m.keys_alt.reserve(n * 4))) return rc;
  if ((rc = m.vals.reserve(n * 4))) return rc;
  if ((rc = m.vals_alt.reserve(n * 4))) return rc;
  if ((rc = m.bounds.reserve(64))) return rc;
  if ((const void *)d_pts != m.orig.p)
    MSFL_CUDA_OK(cudaMemcpyAsync(m.orig.p, d_pts, n * sizeof(float4), cudaMemcpyDeviceToDevice, st));
  const float4 *pts = m.orig.as<float4>();
  int *bounds = m.bounds.as<int>();
  k_init_bounds<<<1, 32, 0, st>>>(bounds);
  const int tb = 256;
  int gb = (int)((N + tb - 1) / tb);
  if (gb > e->sm_count * 8) gb = e->sm_count * 8;
  k_cell_bounds<<<gb, tb, 0, st>>>(pts, N, inv_edge, bounds);
  e->launches += 2;
  int hb[7];
  MSFL_CUDA_OK(cudaMemcpyAsync(hb, bounds, sizeof hb, cudaMemcpyDeviceToHost, st));
  MSFL_CUDA_OK(cudaStreamSynchronize(st));
  if (hb[6]) { set_error("submap contains non-finite or out-of-range (>1e6 m) points"); return MSFL_ERR_ARG; }
  const long long nx = (long long)hb[3] - hb[0] + 5, ny = (long long)hb[4] - hb[1] + 5, nz = (long long)hb[5] - hb[2] + 5;
  const long long ncell = nx * ny * nz;
  if (ncell > (1ll << 26)) {
    set_error("submap bounding box needs %lld cells (> 2^26); dense cell index refused", ncell);
    return MSFL_ERR_GRID;
  }
  if ((rc = m.cell_start.reserve((size_t)(ncell + 1) * 4))) return rc;
  const int ox = hb[0] - 2, oy = hb[1] - 2, oz = hb[2] - 2;
  uint32_t *keys = m.keys.as<uint32_t>(), *keys_alt = m.keys_alt.as<uint32_t>();
  uint32_t *vals = m.vals.as<uint32_t>(), *vals_alt = m.vals_alt.as<uint32_t>();
  k_cell_keys<<<(N + tb - 1) / tb, tb, 0, st>>>(pts, N, inv_edge, ox, oy, oz, (int)nx, (int)ny, keys, vals);
  int end_bit = 1;
  while ((1ll << end_bit) < ncell) ++end_bit;
  cub::DoubleBuffer<uint32_t> dk(keys, keys_alt), dv(vals, vals_alt);
  size_t tmp_bytes = 0;
  MSFL_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)N, 0, end_bit, st));
  if ((rc = m.cub_tmp.reserve(tmp_bytes))) return rc;
  MSFL_CUDA_OK(cub::DeviceRadixSort::SortPairs(m.cub_tmp.p, tmp_bytes, dk, dv, (int)N, 0, end_bit, st));
  k_gather_sorted<<<(N + tb - 1) / tb, tb, 0, st>>>(pts, dv.Current(), N, m.sorted.as<float4>());
  k_cell_start<<<(unsigned)((ncell + 1 + tb - 1) / tb), tb, 0, st>>>(dk.Current(), N, (uint32_t)ncell,
                                                                    m.cell_start.as<uint32_t>());
  e->launches += 3 + 2;  // + the radix-sort passes (library)
  MSFL_CUDA_OK(cudaGetLastError());
  m.n = n;
  m.view.pts_sorted = m.sorted.as<float4>();
  m.view.pts_orig = pts;
  m.view.cell_start = m.cell_start.as<uint32_t>();
  m.view.nx = (int)nx; m.view.ny = (int)ny; m.view.nz = (int)nz;
  m.view.ox = ox; m.view.oy = oy; m.view.oz = oz;
  m.view.inv_edge = inv_edge;
  m.view.n = N;
  m.view.row_mask = nullptr;
  return want_row_mask ? submap_row_mask(e, m) : MSFL_OK;
}

}  // namespace msfl
