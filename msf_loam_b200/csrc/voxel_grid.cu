// voxel_grid.cu -- centroid voxel down-sampling with pcl::VoxelGrid<pcl::PointXYZI> semantics, as the
// caller applies it to the scan features before scan-to-map (laser_mapping.cc:264-270) and to map
// cells on insert (hybrid_grid.cc:518-519).  SURVEY.md 8f row 2.
//   min/max -> min_b = floor(min * inv_leaf), div_b; voxel id = i0 + i1*div0 + i2*div0*div1 with
//   i = (int)(floorf(x * inv_leaf) - (float)min_b); points ordered by voxel id (stable: ascending
//   point index inside a voxel); xyz and intensity averaged with fp32 sums in that order; output in
//   ascending voxel id.  All fp32 ops are explicit round-to-nearest (no FMA) => bit-exact vs the oracle.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "msfl_internal.h"

namespace msfl {

struct VoxMeta {
  uint32_t min_enc[3], max_enc[3];  // order-preserving uint encodings of the float min / max
  int min_b[3], div_b[3];
  int overflow;                     // dx*dy*dz > INT32_MAX: PCL returns the input unchanged
  uint32_t n_out;
};

__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_vox_init(VoxMeta *m) {
  if (threadIdx.x < 3) {
    m->min_enc[threadIdx.x] = 0xffffffffu;
    m->max_enc[threadIdx.x] = 0u;
  }
  if (threadIdx.x == 0) { m->overflow = 0; m->n_out = 0; }
}

__global__ void k_vox_minmax(const float4 *__restrict__ p, uint32_t n, VoxMeta *m) {
  uint32_t lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0, 0, 0};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 q = p[i];
    const uint32_t e[3] = {f2ord(q.x), f2ord(q.y), f2ord(q.z)};
#pragma unroll
    for (int d = 0; d < 3; ++d) { lo[d] = min(lo[d], e[d]); hi[d] = max(hi[d], e[d]); }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d)
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { atomicMin(&m->min_enc[d], lo[d]); atomicMax(&m->max_enc[d], hi[d]); }
  }
}

__global__ void k_vox_meta(VoxMeta *m, float inv) {
  if (threadIdx.x != 0) return;
  long long dim[3];
  for (int d = 0; d < 3; ++d) {
    const float mn = ord2f(m->min_enc[d]), mx = ord2f(m->max_enc[d]);
    dim[d] = (long long)(__fmul_rn(__fsub_rn(mx, mn), inv)) + 1;
    const int lo = (int)floorf(__fmul_rn(mn, inv)), hi = (int)floorf(__fmul_rn(mx, inv));
    m->min_b[d] = lo;
    m->div_b[d] = hi - lo + 1;
  }
  m->overflow = (dim[0] * dim[1] * dim[2] > 2147483647ll) ? 1 : 0;
}

__global__ void k_vox_keys(const float4 *__restrict__ p, uint32_t n, float inv, const VoxMeta *__restrict__ m,
                           uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q = p[i];
  const int i0 = (int)__fsub_rn(floorf(__fmul_rn(q.x, inv)), (float)m->min_b[0]);
  const int i1 = (int)__fsub_rn(floorf(__fmul_rn(q.y, inv)), (float)m->min_b[1]);
  const int i2 = (int)__fsub_rn(floorf(__fmul_rn(q.z, inv)), (float)m->min_b[2]);
  keys[i] = m->overflow ? i : (uint32_t)(i0 + i1 * m->div_b[0] + i2 * m->div_b[0] * m->div_b[1]);
  vals[i] = i;
}

__global__ void k_vox_heads(const uint32_t *__restrict__ keys, uint32_t n, uint32_t *__restrict__ head) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  head[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1u : 0u;
}

__global__ void k_vox_centroids(const float4 *__restrict__ p, const uint32_t *__restrict__ keys,
                                const uint32_t *__restrict__ vals, const uint32_t *__restrict__ head,
                                const uint32_t *__restrict__ pos, uint32_t n, float4 *__restrict__ out, VoxMeta *m) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  if (j == n - 1) m->n_out = pos[j] + head[j];
  if (!head[j]) return;
  const uint32_t key = keys[j];
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  uint32_t k = j;
  for (; k < n && keys[k] == key; ++k) {
    const float4 q = p[vals[k]];
    sx = __fadd_rn(sx, q.x); sy = __fadd_rn(sy, q.y); sz = __fadd_rn(sz, q.z); si = __fadd_rn(si, q.w);
  }
  const float c = (float)(k - j);
  out[pos[j]] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), __fdiv_rn(si, c));
}

// d_in, d_out: packed float4 device arrays (d_out capacity n).  *n_out is read back (one sync).
int run_voxel_grid(msfl_engine *e, const float4 *d_in, size_t n, float leaf, float4 *d_out, size_t *n_out) {
  *n_out = 0;
  if (n == 0) return MSFL_OK;
  if (!(leaf > 0)) { set_error("voxel_grid: leaf must be > 0"); return MSFL_ERR_ARG; }
  cudaStream_t st = e->stream;
  const uint32_t N = (uint32_t)n;
  const float inv = 1.0f / leaf;
  int rc;
  if ((rc = e->v_keys.reserve(n * 4))) return rc;
  if ((rc = e->v_keys_alt.reserve(n * 4))) return rc;
  if ((rc = e->v_vals.reserve(n * 4))) return rc;
  if ((rc = e->v_vals_alt.reserve(n * 4))) return rc;
  if ((rc = e->v_misc.reserve(n * 8 + sizeof(VoxMeta) + 64))) return rc;
  uint32_t *head = e->v_misc.as<uint32_t>();
  uint32_t *pos = head + n;
  VoxMeta *meta = reinterpret_cast<VoxMeta *>(pos + n);
  const int tb = 256;
  const unsigned gb = (N + tb - 1) / tb;
  k_vox_init<<<1, 32, 0, st>>>(meta);
  k_vox_minmax<<<min(gb, (unsigned)e->sm_count * 8), tb, 0, st>>>(d_in, N, meta);
  k_vox_meta<<<1, 32, 0, st>>>(meta, inv);
  uint32_t *keys = e->v_keys.as<uint32_t>(), *vals = e->v_vals.as<uint32_t>();
  k_vox_keys<<<gb, tb, 0, st>>>(d_in, N, inv, meta, keys, vals);
  cub::DoubleBuffer<uint32_t> dk(keys, e->v_keys_alt.as<uint32_t>()), dv(vals, e->v_vals_alt.as<uint32_t>());
  size_t tmp_sort = 0, tmp_scan = 0;
  MSFL_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, dk, dv, (int)N, 0, 32, st));
  MSFL_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, head, pos, (int)N, st));
  if ((rc = e->v_tmp.reserve(tmp_sort > tmp_scan ? tmp_sort : tmp_scan))) return rc;
  MSFL_CUDA_OK(cub::DeviceRadixSort::SortPairs(e->v_tmp.p, tmp_sort, dk, dv, (int)N, 0, 32, st));
  k_vox_heads<<<gb, tb, 0, st>>>(dk.Current(), N, head);
  MSFL_CUDA_OK(cub::DeviceScan::ExclusiveSum(e->v_tmp.p, tmp_scan, head, pos, (int)N, st));
  k_vox_centroids<<<gb, tb, 0, st>>>(d_in, dk.Current(), dv.Current(), head, pos, N, d_out, meta);
  e->launches += 6 + 6;
  MSFL_CUDA_OK(cudaGetLastError());
  uint32_t h_n = 0;
  MSFL_CUDA_OK(cudaMemcpyAsync(&h_n, &meta->n_out, 4, cudaMemcpyDeviceToHost, st));
  MSFL_CUDA_OK(cudaStreamSynchronize(st));
  *n_out = h_n;
  return MSFL_OK;
}

// ---------------------------------------------------------------------------------------------
// Batched form: B clouds back to back, one launch sequence.  Same arithmetic per cloud (own min / max, own voxel grid);
// the sort key carries the scan in its upper half, so ONE 64-bit radix sort orders every cloud of the batch by voxel
// and the centroids of scan b come out contiguous, in ascending voxel order, right after those of scan b - 1.
// ---------------------------------------------------------------------------------------------
__global__ void k_voxb_init(VoxMeta *metas) {
  VoxMeta *m = metas + blockIdx.x;
  if (threadIdx.x < 3) {
    m->min_enc[threadIdx.x] = 0xffffffffu;
    m->max_enc[threadIdx.x] = 0u;
  }
  if (threadIdx.x == 0) { m->overflow = 0; m->n_out = 0; }
}

__global__ void k_voxb_minmax(const float4 *__restrict__ p, const uint32_t *__restrict__ off, VoxMeta *metas) {
  const uint32_t b = blockIdx.y, base = off[b], n = off[b + 1] - base;
  VoxMeta *m = metas + b;
  p += base;
  uint32_t lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0, 0, 0};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 q = p[i];
    const uint32_t e[3] = {f2ord(q.x), f2ord(q.y), f2ord(q.z)};
#pragma unroll
    for (int d = 0; d < 3; ++d) { lo[d] = min(lo[d], e[d]); hi[d] = max(hi[d], e[d]); }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d)
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  if ((threadIdx.x & 31) == 0 && n > 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { atomicMin(&m->min_enc[d], lo[d]); atomicMax(&m->max_enc[d], hi[d]); }
  }
}

__global__ void k_voxb_meta(VoxMeta *metas, const uint32_t *__restrict__ off, float inv) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= gridDim.x * blockDim.x) return;
  VoxMeta *m = metas + b;
  if (off[b + 1] == off[b]) return;  // empty cloud: no voxels
  long long dim[3];
  for (int d = 0; d < 3; ++d) {
    const float mn = ord2f(m->min_enc[d]), mx = ord2f(m->max_enc[d]);
    dim[d] = (long long)(__fmul_rn(__fsub_rn(mx, mn), inv)) + 1;
    const int lo = (int)floorf(__fmul_rn(mn, inv)), hi = (int)floorf(__fmul_rn(mx, inv));
    m->min_b[d] = lo;
    m->div_b[d] = hi - lo + 1;
  }
  m->overflow = (dim[0] * dim[1] * dim[2] > 2147483647ll) ? 1 : 0;
}

__global__ void k_voxb_keys(const float4 *__restrict__ p, const uint32_t *__restrict__ off, float inv, const VoxMeta *__restrict__ metas,
                            unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals) {
  const uint32_t b = blockIdx.y, base = off[b], n = off[b + 1] - base;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const VoxMeta *m = metas + b;
  const float4 q = p[base + i];
  const int i0 = (int)__fsub_rn(floorf(__fmul_rn(q.x, inv)), (float)m->min_b[0]);
  const int i1 = (int)__fsub_rn(floorf(__fmul_rn(q.y, inv)), (float)m->min_b[1]);
  const int i2 = (int)__fsub_rn(floorf(__fmul_rn(q.z, inv)), (float)m->min_b[2]);
  const uint32_t key = m->overflow ? i : (uint32_t)(i0 + i1 * m->div_b[0] + i2 * m->div_b[0] * m->div_b[1]);
  keys[base + i] = ((unsigned long long)b << 32) | key;
  vals[base + i] = base + i;
}

__global__ void k_voxb_heads(const unsigned long long *__restrict__ keys, uint32_t n, uint32_t *__restrict__ head) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  head[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1u : 0u;
}

__global__ void k_voxb_centroids(const float4 *__restrict__ p, const unsigned long long *__restrict__ keys,
                                 const uint32_t *__restrict__ vals, const uint32_t *__restrict__ head,
                                 const uint32_t *__restrict__ pos, uint32_t n, float4 *__restrict__ out) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || !head[j]) return;
  const unsigned long long key = keys[j];
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  uint32_t k = j;
  for (; k < n && keys[k] == key; ++k) {
    const float4 q = p[vals[k]];
    sx = __fadd_rn(sx, q.x); sy = __fadd_rn(sy, q.y); sz = __fadd_rn(sz, q.z); si = __fadd_rn(si, q.w);
  }
  const float c = (float)(k - j);
  out[pos[j]] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), __fdiv_rn(si, c));
}

// where each scan's centroids start: the sorted segment of scan b is [off[b], off[b + 1]) (the key's upper half is b)
__global__ void k_voxb_out_off(const uint32_t *__restrict__ off, const uint32_t *__restrict__ head, const uint32_t *__restrict__ pos,
                               uint32_t n, int B, int32_t *__restrict__ out_off) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > B) return;
  const uint32_t total = n ? pos[n - 1] + head[n - 1] : 0u;
  const uint32_t i = off[b];
  out_off[b] = (int32_t)(i < n ? pos[i] : total);
}

int run_voxel_grid_batch(msfl_engine *e, int B, const float4 *d_in, const uint32_t *d_in_off, size_t n_total, uint32_t max_n,
                         float leaf, float4 *d_out, int32_t *d_out_off, int vb) {
  if (!(leaf > 0)) { set_error("voxel_grid: leaf must be > 0"); return MSFL_ERR_ARG; }
  cudaStream_t st = e->stream;
  const int tb = 256;
  if (n_total == 0) {
    MSFL_CUDA_OK(cudaMemsetAsync(d_out_off, 0, (size_t)(B + 1) * 4, st));
    return MSFL_OK;
  }
  const uint32_t N = (uint32_t)n_total;
  const float inv = 1.0f / leaf;
  int rc;
  // two scratch sets (vb): the corner and the surf batch of one chain call do not wait for each other's buffers
  const size_t half_keys = n_total * 8, half_vals = n_total * 4, half_misc = n_total * 8 + (size_t)B * sizeof(VoxMeta) + 64;
  if ((rc = e->vb_keys.reserve(2 * half_keys))) return rc;
  if ((rc = e->vb_keys_alt.reserve(2 * half_keys))) return rc;
  if ((rc = e->vb_vals.reserve(2 * half_vals))) return rc;
  if ((rc = e->vb_vals_alt.reserve(2 * half_vals))) return rc;
  if ((rc = e->vb_misc.reserve(2 * half_misc))) return rc;
  unsigned long long *keys = (unsigned long long *)(e->vb_keys.as<char>() + vb * half_keys);
  unsigned long long *keys_alt = (unsigned long long *)(e->vb_keys_alt.as<char>() + vb * half_keys);
  uint32_t *vals = (uint32_t *)(e->vb_vals.as<char>() + vb * half_vals), *vals_alt = (uint32_t *)(e->vb_vals_alt.as<char>() + vb * half_vals);
  uint32_t *head = (uint32_t *)(e->vb_misc.as<char>() + vb * ((half_misc + 15) & ~(size_t)15));
  uint32_t *pos = head + n_total;
  VoxMeta *metas = reinterpret_cast<VoxMeta *>(pos + n_total);
  const dim3 gp((max_n + tb - 1) / tb, (unsigned)B);
  k_voxb_init<<<B, 32, 0, st>>>(metas);
  k_voxb_minmax<<<dim3(std::min((max_n + tb - 1) / tb, 64u), (unsigned)B), tb, 0, st>>>(d_in, d_in_off, metas);
  k_voxb_meta<<<B, 1, 0, st>>>(metas, d_in_off, inv);
  k_voxb_keys<<<gp, tb, 0, st>>>(d_in, d_in_off, inv, metas, keys, vals);
  int bits = 32;
  while ((1ll << (bits - 32)) < B) ++bits;
  cub::DoubleBuffer<unsigned long long> dk(keys, keys_alt);
  cub::DoubleBuffer<uint32_t> dv(vals, vals_alt);
  size_t tmp_sort = 0, tmp_scan = 0;
  MSFL_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, dk, dv, (int)N, 0, bits, st));
  MSFL_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, head, pos, (int)N, st));
  if ((rc = e->vb_tmp.reserve(std::max(tmp_sort, tmp_scan)))) return rc;
  MSFL_CUDA_OK(cub::DeviceRadixSort::SortPairs(e->vb_tmp.p, tmp_sort, dk, dv, (int)N, 0, bits, st));
  const unsigned gb = (N + tb - 1) / tb;
  k_voxb_heads<<<gb, tb, 0, st>>>(dk.Current(), N, head);
  MSFL_CUDA_OK(cub::DeviceScan::ExclusiveSum(e->vb_tmp.p, tmp_scan, head, pos, (int)N, st));
  k_voxb_centroids<<<gb, tb, 0, st>>>(d_in, dk.Current(), dv.Current(), head, pos, N, d_out);
  k_voxb_out_off<<<(B + 1 + tb - 1) / tb, tb, 0, st>>>(d_in_off, head, pos, N, B, d_out_off);
  e->launches += 7 + 6;
  MSFL_CUDA_OK(cudaGetLastError());
  return MSFL_OK;
}

}  // namespace msfl

using namespace msfl;

extern "C" int msfl_voxel_grid(msfl_engine *e, const msfl_cloud *in, float leaf, float *out_xyzi, size_t *n_out) {
  if (!e || !in || !out_xyzi || !n_out) { set_error("msfl_voxel_grid: bad argument"); return MSFL_ERR_ARG; }
  *n_out = 0;
  if (in->n == 0) return MSFL_OK;
  int rc;
  if ((rc = check_cloud(in, false, "msfl_voxel_grid"))) return rc;
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  const size_t n = in->n;
  if ((rc = e->h_stage.reserve(n * 16))) return rc;
  if ((rc = e->v_in.reserve(n * 16))) return rc;
  if ((rc = e->v_out.reserve(n * 16))) return rc;
  float *h = e->h_stage.as<float>();
  const char *base = (const char *)in->data;
  const bool has_i = in->off_intensity != MSFL_NO_FIELD;
  for (size_t i = 0; i < n; ++i) {
    const char *pt = base + i * in->stride;
    memcpy(h + 4 * i, pt + in->off_xyz, 12);
    float w = 0.f;
    if (has_i) memcpy(&w, pt + in->off_intensity, 4);
    h[4 * i + 3] = w;
  }
  MSFL_CUDA_OK(cudaMemcpyAsync(e->v_in.p, h, n * 16, cudaMemcpyHostToDevice, e->stream));
  size_t no = 0;
  if ((rc = run_voxel_grid(e, e->v_in.as<float4>(), n, leaf, e->v_out.as<float4>(), &no))) return rc;
  MSFL_CUDA_OK(cudaMemcpyAsync(out_xyzi, e->v_out.p, no * 16, cudaMemcpyDeviceToHost, e->stream));
  MSFL_CUDA_OK(cudaStreamSynchronize(e->stream));
  *n_out = no;
  return MSFL_OK;
}
