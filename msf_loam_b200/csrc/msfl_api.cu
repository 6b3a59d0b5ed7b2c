// msfl_api.cu -- the C ABI of libmsfl.so (include/msfl.h): engine lifetime, host<->device
// marshalling of AoS cloud views, and the launch sequences of the scan-matching path.
#include <emmintrin.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "msfl_internal.h"

namespace msfl {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

int DevBuf::reserve(size_t bytes) {
  if (bytes <= cap) return MSFL_OK;
  size_t want = bytes + bytes / 4 + 256;
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
  MSFL_CUDA_OK(cudaMalloc(&p, want));
  cap = want;
  return MSFL_OK;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}
int PinBuf::reserve(size_t bytes) {
  if (bytes <= cap) return MSFL_OK;
  size_t want = bytes + bytes / 4 + 256;
  if (p) cudaFreeHost(p);
  p = nullptr;
  cap = 0;
  MSFL_CUDA_OK(cudaHostAlloc(&p, want, cudaHostAllocDefault));
  cap = want;
  return MSFL_OK;
}
void PinBuf::release() {
  if (p) cudaFreeHost(p);
  p = nullptr;
  cap = 0;
}

// smallest float f with (double)d < T  <=>  d < f for every float d
static float float_bound(double T) {
  float f = (float)T;
  if ((double)f < T) f = nextafterf(f, INFINITY);
  return f;
}

static void fill_kparams(msfl_engine *e) {
  const msfl_params &p = e->params;
  KParams &k = e->kp;
  k.knn_max_sq_f = float_bound(p.knn_max_sq);
  k.dist_sq_thresh_f = float_bound(p.dist_sq_thresh);
  k.dist_sq_thresh = p.dist_sq_thresh;
  k.nearby_scan = p.nearby_scan;
  k.line_eig_ratio = p.line_eig_ratio;
  k.line_half_len = p.line_half_len;
  k.plane_tol = p.plane_tol;
  k.max_it = p.max_num_iterations;
  k.early_exit = p.early_exit;
  k.max_invalid = p.max_consecutive_invalid_steps;
  k.min_corr = p.min_correspondences;
  k.huber_a = p.huber_a;
  k.huber_sqrt_a = sqrt(p.huber_a);
  k.initial_radius = p.initial_radius;
  k.max_radius = p.max_radius;
  k.min_radius = p.min_radius;
  k.min_rel_decrease = p.min_relative_decrease;
  k.min_diag = p.min_lm_diagonal;
  k.max_diag = p.max_lm_diagonal;
  k.ftol = p.function_tolerance;
  k.gtol = p.gradient_tolerance;
  k.ptol = p.parameter_tolerance;
}

// One validation for every entry point that takes an msfl_cloud: every field the call will read has to lie inside
// the point stride, so the last point can never be read past the caller's buffer.
int check_cloud(const msfl_cloud *c, bool need_ring, const char *what) {
  if (!c) { set_error("%s: null cloud", what); return MSFL_ERR_ARG; }
  if (c->n == 0) return MSFL_OK;
  if (!c->data) { set_error("%s: null data with n=%zu", what, c->n); return MSFL_ERR_ARG; }
  if (c->stride < 12 || c->off_xyz > c->stride - 12) { set_error("%s: xyz outside the point stride", what); return MSFL_ERR_ARG; }
  if (c->off_intensity != MSFL_NO_FIELD && (c->stride < 4 || c->off_intensity > c->stride - 4)) {
    set_error("%s: intensity outside the point stride", what);
    return MSFL_ERR_ARG;
  }
  if (c->off_ring != MSFL_NO_FIELD && c->off_ring > c->stride - 2) { set_error("%s: ring outside the point stride", what); return MSFL_ERR_ARG; }
  if (need_ring && c->off_ring == MSFL_NO_FIELD) { set_error("%s: ring field required", what); return MSFL_ERR_ARG; }
  if (c->n > 0x3fffffffull) { set_error("%s: too many points", what); return MSFL_ERR_ARG; }
  return MSFL_OK;
}

// host AoS view -> packed float4 (x, y, z, intensity) [+ uint16 ring] in (pinned) host memory
static void pack_cloud_host(const msfl_cloud *c, float *dst4, uint16_t *ring_dst) {
  const size_t n = c->n;
  const char *base = (const char *)c->data;
  const bool has_i = c->off_intensity != MSFL_NO_FIELD;
  if (c->stride == 16 && c->off_xyz == 0 && (c->off_intensity == 12 || !has_i)) {
    memcpy(dst4, base, n * 16);
  } else if (((c->stride | c->off_xyz | (has_i ? c->off_intensity : 0) | (size_t)(uintptr_t)base) & 3u) == 0 &&
             c->off_xyz + 16 <= c->stride && (((uintptr_t)dst4) & 15u) == 0) {
    // word-aligned points with 16 readable bytes at xyz (every PCL type: pcl::PointXYZI is x y z pad | intensity ...):
    // one 16-byte load, the intensity dropped into lane 3, one NON-TEMPORAL 16-byte store -- the staging slot is
    // written once and read by the DMA engine, so it should not be pulled into the cache first (write-allocate would
    // add a third of the traffic of this memory-bound loop)
    const size_t st = c->stride;
    const char *px = base + c->off_xyz, *pi = has_i ? base + c->off_intensity : nullptr;
    const __m128i keep_xyz = _mm_set_epi32(0, -1, -1, -1);
    for (size_t i = 0; i < n; ++i, px += st) {
      __m128i v = _mm_and_si128(_mm_loadu_si128((const __m128i *)px), keep_xyz);
      if (has_i) {
        v = _mm_or_si128(v, _mm_slli_si128(_mm_cvtsi32_si128(*(const int *)pi), 12));
        pi += st;
      }
      _mm_stream_si128((__m128i *)(dst4 + 4 * i), v);
    }
    _mm_sfence();
  } else if (((c->stride | c->off_xyz | (has_i ? c->off_intensity : 0) | (size_t)(uintptr_t)base) & 3u) == 0) {
    // word-aligned points (every PCL type: pcl::PointXYZI is 32 B with intensity at 16): four 32-bit moves per point
    const size_t sw = c->stride / 4, ox = c->off_xyz / 4, oi = has_i ? c->off_intensity / 4 : 0;
    const uint32_t *src = (const uint32_t *)base;
    uint32_t *dst = (uint32_t *)dst4;
    for (size_t i = 0; i < n; ++i, src += sw, dst += 4) {
      dst[0] = src[ox]; dst[1] = src[ox + 1]; dst[2] = src[ox + 2];
      dst[3] = has_i ? src[oi] : 0u;
    }
  } else {
    for (size_t i = 0; i < n; ++i) {
      const char *pt = base + i * c->stride;
      memcpy(dst4 + 4 * i, pt + c->off_xyz, 12);
      float w = 0.f;
      if (has_i) memcpy(&w, pt + c->off_intensity, 4);
      dst4[4 * i + 3] = w;
    }
  }
  if (ring_dst) {
    if (c->off_ring == MSFL_NO_FIELD) memset(ring_dst, 0, n * 2);
    else
      for (size_t i = 0; i < n; ++i) memcpy(ring_dst + i, base + i * c->stride + c->off_ring, 2);
  }
}

// Packs scans [b0, b1) of a batch into the staging layout [all corner | all surf] (hq: float4 base, corner scan b at
// c_at[b], surf scan b at nct + s_at[b]).  Large batches (the PCL-layout replay path: 32 B points in pageable memory) are
// split over host threads -- one memory-bound loop per thread; a 2048-scan VLP-16 batch is 0.3 GB read + 0.16 GB written.
static void pack_batch_host(msfl_engine *e, int b0, int b1, const msfl_cloud *corner, const msfl_cloud *surf, float *hq,
                            size_t nct, const size_t *c_at, const size_t *s_at) {
  size_t pts = 0;
  for (int b = b0; b < b1; ++b) pts += corner[b].n + surf[b].n;
  int nt = e->pack_threads;
  if (nt > b1 - b0) nt = b1 - b0;
  if (pts < 200000 || nt <= 1) {
    for (int b = b0; b < b1; ++b) {
      pack_cloud_host(&corner[b], hq + 4 * c_at[b], nullptr);
      pack_cloud_host(&surf[b], hq + 4 * (nct + s_at[b]), nullptr);
    }
    return;
  }
  // contiguous scan ranges with equal point counts
  std::vector<std::thread> th;
  th.reserve(nt);
  int b = b0;
  size_t done = 0;
  for (int t = 0; t < nt; ++t) {
    const size_t goal = pts * (size_t)(t + 1) / (size_t)nt;
    const int first = b;
    while (b < b1 && (done < goal || t == nt - 1)) { done += corner[b].n + surf[b].n; ++b; }
    const int last = b;
    if (first == last) continue;
    th.emplace_back([=]() {
      for (int i = first; i < last; ++i) {
        pack_cloud_host(&corner[i], hq + 4 * c_at[i], nullptr);
        pack_cloud_host(&surf[i], hq + 4 * (nct + s_at[i]), nullptr);
      }
    });
  }
  for (auto &t : th) t.join();
}

void pack_clouds_parallel(msfl_engine *e, int B, const msfl_cloud *clouds, float *dst4, uint16_t *ring_dst, const uint32_t *off) {
  size_t pts = 0;
  for (int b = 0; b < B; ++b) pts += clouds[b].n;
  int nt = std::min(e->pack_threads, B);
  if (pts < 200000 || nt <= 1) {
    for (int b = 0; b < B; ++b) pack_cloud_host(&clouds[b], dst4 + 4 * (size_t)off[b], ring_dst ? ring_dst + off[b] : nullptr);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(nt);
  int b = 0;
  size_t done = 0;
  for (int t = 0; t < nt; ++t) {
    const size_t goal = pts * (size_t)(t + 1) / (size_t)nt;
    const int first = b;
    while (b < B && (done < goal || t == nt - 1)) { done += clouds[b].n; ++b; }
    const int last = b;
    if (first == last) continue;
    th.emplace_back([=]() {
      for (int i = first; i < last; ++i) pack_cloud_host(&clouds[i], dst4 + 4 * (size_t)off[i], ring_dst ? ring_dst + off[i] : nullptr);
    });
  }
  for (auto &t : th) t.join();
}

// xyz-only uploads (12 B per point: the LiDAR-only matcher never reads a query's intensity) are widened to the kernels'
// float4 layout on the device; the PCIe link carries a quarter less
__global__ void k_expand_xyz(const float *__restrict__ xyz, uint32_t n, float4 *__restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = make_float4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], 0.f);
}

static cudaEvent_t take_event(msfl_engine *e) {
  if (!e->event_pool.empty()) {
    cudaEvent_t ev = e->event_pool.back();
    e->event_pool.pop_back();
    return ev;
  }
  cudaEvent_t ev = nullptr;
  cudaEventCreate(&ev);
  return ev;
}

void stage_begin(msfl_engine *e, int stage) {
  if (!e->profiling) return;
  msfl_engine::StageEv s{take_event(e), take_event(e), stage};
  cudaEventRecord(s.a, e->stream);
  e->stage_events.push_back(s);
}

void stage_end(msfl_engine *e) {
  if (!e->profiling || e->stage_events.empty()) return;
  cudaEventRecord(e->stage_events.back().b, e->stream);
}

}  // namespace msfl

using namespace msfl;

extern "C" {

int msfl_set_profiling(msfl_engine *e, int on) {
  if (!e) return MSFL_ERR_ARG;
  e->profiling = on != 0;
  return MSFL_OK;
}

int msfl_get_profile(msfl_engine *e, double ms[MSFL_N_STAGES], int32_t count[MSFL_N_STAGES]) {
  if (!e || !ms || !count) return MSFL_ERR_ARG;
  for (int i = 0; i < MSFL_N_STAGES; ++i) { ms[i] = 0; count[i] = 0; }
  MSFL_CUDA_OK(cudaStreamSynchronize(e->stream));
  for (auto &s : e->stage_events) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess && s.stage >= 0 && s.stage < MSFL_N_STAGES) {
      ms[s.stage] += t;
      count[s.stage] += 1;
    }
    e->event_pool.push_back(s.a);
    e->event_pool.push_back(s.b);
  }
  e->stage_events.clear();
  return MSFL_OK;
}

const char *msfl_last_error(void) { return g_err; }
const char *msfl_version(void) { return "msfl 0.2 (sm_100a)"; }

int msfl_abi_check(size_t sp, size_t ss, size_t sc, size_t sf, size_t sd) {
  if (sp == sizeof(msfl_params) && ss == sizeof(msfl_stats) && sc == sizeof(msfl_cloud) && sf == sizeof(msfl_features) &&
      sd == sizeof(msfl_deskew))
    return MSFL_OK;
  set_error("struct sizes differ: library has params %zu stats %zu cloud %zu features %zu deskew %zu, caller passed %zu %zu %zu %zu %zu",
            sizeof(msfl_params), sizeof(msfl_stats), sizeof(msfl_cloud), sizeof(msfl_features), sizeof(msfl_deskew), sp, ss, sc, sf, sd);
  return MSFL_ERR_ARG;
}

void msfl_default_params(msfl_params *p) {
  memset(p, 0, sizeof *p);
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
  p->lm_cluster = 0;
  p->assoc_sorted = 0;
}

int msfl_create_on_stream(const msfl_params *params, int device, void *stream, msfl_engine **out) {
  if (!out) { set_error("msfl_create: null out"); return MSFL_ERR_ARG; }
  *out = nullptr;
  msfl_params p;
  if (params) p = *params;
  else msfl_default_params(&p);
  if (p.num_outer < 1 || p.num_outer > MSFL_MAX_OUTER || p.max_num_iterations < 0 ||
      p.max_num_iterations > MSFL_MAX_ATTEMPTS || p.n_sectors < 1 || p.n_sectors > 64 || !(p.knn_max_sq > 0) ||
      !(p.dist_sq_thresh > 0)) {
    set_error("msfl_create: parameter out of range");
    return MSFL_ERR_ARG;
  }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev <= 0) {
    set_error("msfl_create: no CUDA device (%s); this engine has no CPU fallback", cudaGetErrorString(ce));
    return MSFL_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { set_error("msfl_create: device %d out of range", device); return MSFL_ERR_ARG; }
  MSFL_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  MSFL_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("msfl_create: device %d is sm_%d%d; libmsfl carries sm_100a code only", device, prop.major, prop.minor);
    return MSFL_ERR_CUDA;
  }
  msfl_engine *e = new msfl_engine();
  e->params = p;
  e->device = device;
  e->sm_count = prop.multiProcessorCount;
  fill_kparams(e);
  if (const char *v = getenv("MSFL_COUNT_SORT_MAX_BINS")) e->count_sort_max_bins = atoll(v);  // tests: force the radix path
  if (const char *v = getenv("MSFL_PAIR_CELL_BUDGET")) e->pair_cell_budget = std::max(1ll, atoll(v));  // tests: force chunks
  {
    const unsigned hc = std::thread::hardware_concurrency();
    e->pack_threads = (int)std::min(16u, std::max(1u, hc));
    if (const char *v = getenv("MSFL_PACK_THREADS")) e->pack_threads = std::max(1, atoi(v));
  }
  if (const char *v = getenv("MSFL_FUSE_FIT")) e->fuse_fit = atoi(v) != 0;  // development A/B
  if (stream) {
    e->stream = (cudaStream_t)stream;
    e->own_stream = false;
  } else {
    cudaError_t se = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) {
      set_error("cudaStreamCreate failed: %s", cudaGetErrorString(se));
      delete e;
      return MSFL_ERR_CUDA;
    }
    e->own_stream = true;
  }
  *out = e;
  return MSFL_OK;
}

int msfl_create(const msfl_params *params, int device, msfl_engine **out) {
  return msfl_create_on_stream(params, device, nullptr, out);
}

void msfl_destroy(msfl_engine *e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  submap_release(e->map_corner);
  submap_release(e->map_surf);
  submap_release(e->last_corner_grid);
  submap_release(e->last_surf_grid);
  DevBuf *dbs[] = {&e->d_queries, &e->d_corr, &e->d_poses, &e->d_status, &e->d_stats, &e->d_knn, &e->d_off, &e->d_misc,
                   &e->d_last_corner, &e->d_last_surf, &e->d_last_corner_ring, &e->d_last_surf_ring, &e->d_ring_tab,
                   &e->d_assoc, &e->f_raw, &e->f_keys, &e->f_keys_alt, &e->f_vals, &e->f_vals_alt, &e->f_tmp, &e->f_full,
                   &e->f_ring, &e->f_curv, &e->f_label, &e->f_idx, &e->f_cnt, &e->f_angle, &e->f_misc, &e->f_soff, &e->v_in,
                   &e->v_keys, &e->v_keys_alt, &e->v_vals, &e->v_vals_alt, &e->v_tmp, &e->v_out, &e->v_misc, &e->vb_keys, &e->vb_keys_alt, &e->vb_vals, &e->vb_vals_alt, &e->vb_tmp, &e->vb_misc,
                   &e->c_in, &e->c_off, &e->c_q, &e->ob_in, &e->ob_bounds, &e->ob_hdr, &e->ob_sorted, &e->ob_ring_sorted, &e->ob_cells,
                   &e->ob_keys, &e->ob_rank, &e->ob_tmp,
                   &e->a_xq, &e->a_keys, &e->a_keys_alt, &e->a_vals, &e->a_vals_alt, &e->a_tmp, &e->a_hist,
                   &e->k_table, &e->k_dsk, &e->k_pprime, &e->k_o4, &e->a_fb, &e->fr_qc, &e->fr_qs, &e->fr_misc};
  for (DevBuf *b : dbs) b->release();
  for (auto &sl : e->slots) {
    sl.d_in.release(); sl.d_in3.release(); sl.d_stats.release(); sl.h_stage.release(); sl.h_out.release(); sl.h_stats.release();
    if (sl.uploaded) cudaEventDestroy(sl.uploaded);
    if (sl.done) cudaEventDestroy(sl.done);
  }
  PinBuf *pbs[] = {&e->h_stage, &e->h_poses, &e->h_stats, &e->h_misc, &e->h_map_stage};
  if (e->map_uploaded) cudaEventDestroy(e->map_uploaded);
  for (PinBuf *b : pbs) b->release();
  for (auto &s : e->stage_events) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
  for (auto ev : e->event_pool) cudaEventDestroy(ev);
  for (auto ev : e->chunk_events) cudaEventDestroy(ev);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->own_stream) cudaStreamDestroy(e->stream);
  delete e;
}

int msfl_sync(msfl_engine *e) {
  if (!e) return MSFL_ERR_ARG;
  MSFL_CUDA_OK(cudaStreamSynchronize(e->stream));
  return MSFL_OK;
}

void *msfl_stream(msfl_engine *e) { return e ? (void *)e->stream : nullptr; }
uint64_t msfl_launch_count(const msfl_engine *e) { return e ? e->launches : 0; }

// ------------------------------------------------------------------------------------------------
// submap
// ------------------------------------------------------------------------------------------------
static float cell_edge(const msfl_engine *e) {
  const double r = sqrt(e->params.knn_max_sq);
  if (r == 1.0) return 1.0f;  // floor(x * 1.0f) is exact: the 27-cell search is exact w.r.t. the gate
  return (float)(r * 1.0001);
}

int msfl_set_submap_device(msfl_engine *e, const float *d_corner, size_t n_corner, const float *d_surf, size_t n_surf) {
  if (!e || !d_corner || !d_surf) { set_error("msfl_set_submap_device: null argument"); return MSFL_ERR_ARG; }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  e->has_submap = false;
  int rc;
  if ((rc = submap_build(e, e->map_corner, (const float4 *)d_corner, n_corner, cell_edge(e)))) return rc;
  if ((rc = submap_build(e, e->map_surf, (const float4 *)d_surf, n_surf, cell_edge(e)))) return rc;
  e->has_submap = true;
  return MSFL_OK;
}

// cell bounding box of a packed cloud for the given cell edge, computed with the kernels' own arithmetic
// (floorf(x * inv_edge) in fp32); false when a point is non-finite or beyond 1e6 m (the device build refuses those too)
static bool host_cell_bounds(const float *xyzw, size_t n, float inv_edge, int lo[3], int hi[3]) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  bool ok = true;
  for (size_t i = 0; i < n; ++i) {
    for (int d = 0; d < 3; ++d) {
      const float v = xyzw[4 * i + d];
      if (!(fabsf(v) <= 1e6f)) ok = false;  // also catches NaN
      mn[d] = v < mn[d] ? v : mn[d];
      mx[d] = v > mx[d] ? v : mx[d];
    }
  }
  if (!ok) return false;
  for (int d = 0; d < 3; ++d) {  // x -> floorf(x * inv_edge) is monotone, so the extreme points give the extreme cells
    lo[d] = (int)floorf(mn[d] * inv_edge);
    hi[d] = (int)floorf(mx[d] * inv_edge);
  }
  return true;
}

int msfl_set_submap(msfl_engine *e, const msfl_cloud *map_corner, const msfl_cloud *map_surf) {
  if (!e) { set_error("msfl_set_submap: null engine"); return MSFL_ERR_ARG; }
  int rc;
  if ((rc = check_cloud(map_corner, false, "map_corner"))) return rc;
  if ((rc = check_cloud(map_surf, false, "map_surf"))) return rc;
  if (map_corner->n == 0 || map_surf->n == 0) { set_error("msfl_set_submap: empty submap class"); return MSFL_ERR_ARG; }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  const size_t nc = map_corner->n, ns = map_surf->n;
  // own staging buffer (the batch calls repack into h_stage while this upload may still be in flight); the previous
  // map upload must have left it
  if (e->map_uploaded) MSFL_CUDA_OK(cudaEventSynchronize(e->map_uploaded));
  else MSFL_CUDA_OK(cudaEventCreateWithFlags(&e->map_uploaded, cudaEventDisableTiming));
  if ((rc = e->h_map_stage.reserve((nc + ns) * 16))) return rc;
  float *h = e->h_map_stage.as<float>();
  pack_cloud_host(map_corner, h, nullptr);
  pack_cloud_host(map_surf, h + 4 * nc, nullptr);
  if ((rc = e->map_corner.orig.reserve(nc * 16))) return rc;
  if ((rc = e->map_surf.orig.reserve(ns * 16))) return rc;
  e->has_submap = false;
  MSFL_CUDA_OK(cudaMemcpyAsync(e->map_corner.orig.p, h, nc * 16, cudaMemcpyHostToDevice, e->stream));
  MSFL_CUDA_OK(cudaMemcpyAsync(e->map_surf.orig.p, h + 4 * nc, ns * 16, cudaMemcpyHostToDevice, e->stream));
  MSFL_CUDA_OK(cudaEventRecord(e->map_uploaded, e->stream));
  // The points pass through the host anyway: take the cell bounding box here, so the index build needs no device
  // round trip (the reference rebuilds its kd-trees every frame, mapping_scan_matcher.cc:66-72 -- this call is on the
  // latency path of the drop-in).  Counting sort of the cells: no library sort, no synchronisation.
  const float edge = cell_edge(e), inv_edge = 1.0f / edge;
  int lo_c[3], hi_c[3], lo_s[3], hi_s[3];
  if (!host_cell_bounds(h, nc, inv_edge, lo_c, hi_c) || !host_cell_bounds(h + 4 * nc, ns, inv_edge, lo_s, hi_s)) {
    set_error("submap contains non-finite or out-of-range (>1e6 m) points");
    return MSFL_ERR_ARG;
  }
  if ((rc = submap_build_host_bounds(e, e->map_corner, nc, edge, lo_c, hi_c))) return rc;
  if ((rc = submap_build_host_bounds(e, e->map_surf, ns, edge, lo_s, hi_s))) return rc;
  e->has_submap = true;
  return MSFL_OK;
}

int msfl_get_submap_device(msfl_engine *e, const float **d_corner, size_t *n_corner, const float **d_surf, size_t *n_surf) {
  if (!e || !e->has_submap) { set_error("msfl_get_submap_device: no submap"); return MSFL_ERR_NOSUBMAP; }
  if (d_corner) *d_corner = e->map_corner.orig.as<float>();
  if (n_corner) *n_corner = e->map_corner.n;
  if (d_surf) *d_surf = e->map_surf.orig.as<float>();
  if (n_surf) *n_surf = e->map_surf.n;
  return MSFL_OK;
}

// ------------------------------------------------------------------------------------------------
// scan-to-map
// ------------------------------------------------------------------------------------------------
}  // extern "C"
int msfl::scan2map_enqueue(msfl_engine *e, int B, const float4 *d_qc, const int32_t *d_c_off, uint32_t nct, const float4 *d_qs,
                           const int32_t *d_s_off, uint32_t nst, double *d_poses, msfl_stats *d_stats) {
  int rc;
  if ((rc = e->d_corr.reserve(((size_t)nct + nst + 1) * 6 * sizeof(double)))) return rc;
  if ((rc = e->d_status.reserve((size_t)B * 4))) return rc;
  if (B == 1) {  // the ROS node's call pattern: one cluster owns the scan for the whole call, one launch
    stage_begin(e, 1);
    rc = launch_scan2map_fused(e, d_qc, nct, d_qs, nst, d_poses, e->d_status.as<int32_t>(), d_stats);
    stage_end(e);
    if (rc <= 0) return rc;
  }
  for (int outer = 0; outer < e->params.num_outer; ++outer) {  // mapping_scan_matcher.cc:75
    const bool compact = true;  // plane constants as 32 B {n, n.c}
    if ((rc = launch_associate_map(e, B, d_qc, d_c_off, nct, d_qs, d_s_off, nst, d_poses, e->d_corr.as<double>(), nullptr, compact)))
      return rc;
    stage_begin(e, 1);
    rc = launch_lm_solve(e, B, d_qc, d_c_off, nct, d_qs, d_s_off, e->d_corr.as<double>(), d_poses,
                         e->d_status.as<int32_t>(), d_stats, outer, /*min_corr=*/0, compact ? 32 : 48);
    stage_end(e);
    if (rc) return rc;
  }
  return MSFL_OK;
}

extern "C" {

int msfl_scan2map_batch_device(msfl_engine *e, int B, const float *d_corner, const int32_t *d_corner_off,
                               size_t n_corner_total, const float *d_surf, const int32_t *d_surf_off,
                               size_t n_surf_total, double *d_poses, msfl_stats *d_stats) {
  if (!e || B <= 0 || !d_corner_off || !d_surf_off || !d_poses) { set_error("msfl_scan2map_batch_device: bad argument"); return MSFL_ERR_ARG; }
  if (!e->has_submap) { set_error("msfl_scan2map: no submap set"); return MSFL_ERR_NOSUBMAP; }
  if (n_corner_total + n_surf_total > 0x7fffffffull) { set_error("batch too large"); return MSFL_ERR_ARG; }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  return scan2map_enqueue(e, B, (const float4 *)d_corner, d_corner_off, (uint32_t)n_corner_total, (const float4 *)d_surf,
                          d_surf_off, (uint32_t)n_surf_total, d_poses, d_stats);
}

// uploads B (corner, surf) cloud pairs into the engine's batch buffers; fills totals
static int upload_batch(msfl_engine *e, int B, const msfl_cloud *corner, const msfl_cloud *surf, const double *poses,
                        uint32_t *nct_out, uint32_t *nst_out) {
  int rc;
  size_t nct = 0, nst = 0;
  for (int b = 0; b < B; ++b) {
    if ((rc = check_cloud(&corner[b], false, "scan_corner"))) return rc;
    if ((rc = check_cloud(&surf[b], false, "scan_surf"))) return rc;
    nct += corner[b].n;
    nst += surf[b].n;
  }
  if (nct + nst > 0x7fffffffull) { set_error("batch too large"); return MSFL_ERR_ARG; }
  const size_t off_bytes = (size_t)2 * (B + 1) * 4;
  const size_t pose_bytes = (size_t)B * 7 * 8;
  // staging layout: [queries (nct+nst) float4 | offsets 2(B+1) int32 | poses B*7 double]
  const size_t q_bytes = (nct + nst) * 16;
  const size_t q_pad = (q_bytes + 15) & ~(size_t)15;
  const size_t off_pad = (off_bytes + 15) & ~(size_t)15;
  if ((rc = e->h_stage.reserve(q_pad + off_pad + pose_bytes))) return rc;
  if ((rc = e->d_queries.reserve(q_pad + off_pad + pose_bytes))) return rc;
  char *h = e->h_stage.as<char>();
  float *hq = (float *)h;
  int32_t *hoff = (int32_t *)(h + q_pad);
  double *hpose = (double *)(h + q_pad + off_pad);
  // Fast path: every cloud is packed float4 and the clouds of a class are back to back in one host
  // allocation (e.g. one pinned buffer) -> DMA straight from the caller's memory, no host repack.
  bool contiguous = true;
  for (int b = 0; b < B && contiguous; ++b) {
    const msfl_cloud *cl[2] = {&corner[b], &surf[b]};
    const msfl_cloud *nx[2] = {b + 1 < B ? &corner[b + 1] : nullptr, b + 1 < B ? &surf[b + 1] : nullptr};
    for (int c = 0; c < 2; ++c) {
      if (cl[c]->stride != 16 || cl[c]->off_xyz != 0) contiguous = false;
      if (nx[c] && (const char *)nx[c]->data != (const char *)cl[c]->data + cl[c]->n * 16) contiguous = false;
    }
  }
  size_t ci = 0, si = 0;
  std::vector<size_t> c_at(B), s_at(B);
  for (int b = 0; b < B; ++b) {
    hoff[b] = (int32_t)ci;
    hoff[B + 1 + b] = (int32_t)si;
    c_at[b] = ci; s_at[b] = si;
    ci += corner[b].n;
    si += surf[b].n;
  }
  hoff[B] = (int32_t)ci;
  hoff[2 * B + 1] = (int32_t)si;
  if (!contiguous) pack_batch_host(e, 0, B, corner, surf, hq, nct, c_at.data(), s_at.data());
  memcpy(hpose, poses, pose_bytes);
  if (contiguous) {
    char *d = e->d_queries.as<char>();
    if (nct) MSFL_CUDA_OK(cudaMemcpyAsync(d, corner[0].data, nct * 16, cudaMemcpyHostToDevice, e->stream));
    if (nst) MSFL_CUDA_OK(cudaMemcpyAsync(d + nct * 16, surf[0].data, nst * 16, cudaMemcpyHostToDevice, e->stream));
    MSFL_CUDA_OK(cudaMemcpyAsync(d + q_pad, h + q_pad, off_pad + pose_bytes, cudaMemcpyHostToDevice, e->stream));
  } else {
    // one H2D copy for the whole batch
    MSFL_CUDA_OK(cudaMemcpyAsync(e->d_queries.p, h, q_pad + off_pad + pose_bytes, cudaMemcpyHostToDevice, e->stream));
  }
  *nct_out = (uint32_t)nct;
  *nst_out = (uint32_t)nst;
  return MSFL_OK;
}

// Large host batches: split the scans into chunks and overlap the H2D copy of chunk c+1 (copy
// stream) with the kernels of chunk c (engine stream).  Scans are independent, so the poses are
// bit-identical to the single-shot path.
static int scan2map_batch_pipelined(msfl_engine *e, int B, const msfl_cloud *corner, const msfl_cloud *surf,
                                    double *poses_tq, msfl_stats *stats, int n_chunks) {
  int rc;
  size_t nct = 0, nst = 0;
  bool contiguous = true;
  for (int b = 0; b < B; ++b) {
    if ((rc = check_cloud(&corner[b], false, "scan_corner"))) return rc;
    if ((rc = check_cloud(&surf[b], false, "scan_surf"))) return rc;
    const msfl_cloud *cl[2] = {&corner[b], &surf[b]};
    const msfl_cloud *nx[2] = {b + 1 < B ? &corner[b + 1] : nullptr, b + 1 < B ? &surf[b + 1] : nullptr};
    for (int c = 0; c < 2; ++c) {
      if (cl[c]->stride != 16 || cl[c]->off_xyz != 0) contiguous = false;
      if (nx[c] && (const char *)nx[c]->data != (const char *)cl[c]->data + cl[c]->n * 16) contiguous = false;
    }
    nct += corner[b].n;
    nst += surf[b].n;
  }
  if (nct + nst > 0x7fffffffull) { set_error("batch too large"); return MSFL_ERR_ARG; }
  if (!e->copy_stream) MSFL_CUDA_OK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  while ((int)e->chunk_events.size() < n_chunks) {
    cudaEvent_t ev;
    MSFL_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    e->chunk_events.push_back(ev);
  }
  const size_t q_bytes = (nct + nst) * 16, q_pad = (q_bytes + 15) & ~(size_t)15;
  const size_t off_bytes = (size_t)2 * (B + n_chunks) * 4, off_pad = (off_bytes + 15) & ~(size_t)15;
  const size_t pose_bytes = (size_t)B * 7 * 8;
  if ((rc = e->h_stage.reserve(q_pad + off_pad + pose_bytes))) return rc;
  if ((rc = e->d_queries.reserve(q_pad + off_pad + pose_bytes))) return rc;
  char *h = e->h_stage.as<char>(), *d = e->d_queries.as<char>();
  float *hq = (float *)h;
  int32_t *hoff = (int32_t *)(h + q_pad);
  double *hpose = (double *)(h + q_pad + off_pad);
  // chunk tables: per chunk two rebased offset tables (Bc+1 each), stored back to back
  struct Chunk { int b0, b1; size_t ci0, ci1, si0, si1; size_t off_pos; };
  std::vector<Chunk> ch(n_chunks);
  std::vector<size_t> c_at(B), s_at(B);
  {
    size_t ci = 0, si = 0, pos = 0;
    for (int c = 0; c < n_chunks; ++c) {
      Chunk &k = ch[c];
      // geometric chunk sizes (1, 2, 4, ... , rest): the copy of chunk c+1 (PCIe is ~2x faster than the
      // kernels consume scans) hides behind the kernels of chunk c, and only the small first copy is exposed
      const long long denom = 1ll << n_chunks;
      k.b0 = c == 0 ? 0 : (int)((long long)B * ((1ll << c) - 1) / denom);
      k.b1 = c == n_chunks - 1 ? B : (int)((long long)B * ((1ll << (c + 1)) - 1) / denom);
      k.ci0 = ci; k.si0 = si; k.off_pos = pos;
      const int Bc = k.b1 - k.b0;
      int32_t *co = hoff + pos, *so = co + (Bc + 1);
      for (int b = k.b0; b < k.b1; ++b) {
        co[b - k.b0] = (int32_t)(ci - k.ci0);
        so[b - k.b0] = (int32_t)(si - k.si0);
        c_at[b] = ci; s_at[b] = si;
        ci += corner[b].n;
        si += surf[b].n;
      }
      co[Bc] = (int32_t)(ci - k.ci0);
      so[Bc] = (int32_t)(si - k.si0);
      k.ci1 = ci; k.si1 = si;
      pos += 2 * (size_t)(Bc + 1);
    }
  }
  memcpy(hpose, poses_tq, pose_bytes);
  // scratch sized once for the largest chunk so no buffer is re-allocated while kernels run
  size_t max_total = 0;
  for (auto &k : ch) max_total = std::max(max_total, (k.ci1 - k.ci0) + (k.si1 - k.si0));
  if ((rc = e->d_corr.reserve((max_total + 1) * 6 * sizeof(double)))) return rc;
  if ((rc = e->d_status.reserve((size_t)B * 4))) return rc;
  if ((rc = e->d_knn.reserve(max_total * 20))) return rc;
  if ((rc = e->a_fb.reserve((max_total + 2) * 4))) return rc;
  if ((rc = e->a_xq.reserve(max_total * 16))) return rc;
  if ((rc = e->a_keys.reserve(max_total * 4))) return rc;
  if ((rc = e->a_keys_alt.reserve(max_total * 4))) return rc;
  if ((rc = e->a_vals.reserve(max_total * 4))) return rc;
  if ((rc = e->a_vals_alt.reserve(max_total * 4))) return rc;
  if ((rc = e->a_tmp.reserve(max_total * 8 + (1 << 20)))) return rc;
  msfl_stats *d_stats = nullptr;
  if (stats) {
    if ((rc = e->d_stats.reserve((size_t)B * sizeof(msfl_stats)))) return rc;
    if ((rc = e->h_stats.reserve((size_t)B * sizeof(msfl_stats)))) return rc;
    d_stats = e->d_stats.as<msfl_stats>();
    MSFL_CUDA_OK(cudaMemsetAsync(d_stats, 0, (size_t)B * sizeof(msfl_stats), e->stream));
  }
  if ((rc = e->h_poses.reserve(pose_bytes))) return rc;
  cudaStream_t cs = e->copy_stream;
  MSFL_CUDA_OK(cudaMemcpyAsync(d + q_pad, h + q_pad, off_pad + pose_bytes, cudaMemcpyHostToDevice, cs));
  float4 *d_qc = (float4 *)d, *d_qs = d_qc + nct;
  const int32_t *d_off = (const int32_t *)(d + q_pad);
  double *d_poses = (double *)(d + q_pad + off_pad);
  for (int c = 0; c < n_chunks; ++c) {
    const Chunk &k = ch[c];
    const size_t ncc = k.ci1 - k.ci0, nsc = k.si1 - k.si0;
    if (contiguous) {
      if (ncc) MSFL_CUDA_OK(cudaMemcpyAsync(d_qc + k.ci0, (const char *)corner[0].data + k.ci0 * 16, ncc * 16, cudaMemcpyHostToDevice, cs));
      if (nsc) MSFL_CUDA_OK(cudaMemcpyAsync(d_qs + k.si0, (const char *)surf[0].data + k.si0 * 16, nsc * 16, cudaMemcpyHostToDevice, cs));
    } else {
      pack_batch_host(e, k.b0, k.b1, corner, surf, hq, nct, c_at.data(), s_at.data());
      if (ncc) MSFL_CUDA_OK(cudaMemcpyAsync(d_qc + k.ci0, hq + 4 * k.ci0, ncc * 16, cudaMemcpyHostToDevice, cs));
      if (nsc) MSFL_CUDA_OK(cudaMemcpyAsync(d_qs + k.si0, hq + 4 * (nct + k.si0), nsc * 16, cudaMemcpyHostToDevice, cs));
    }
    MSFL_CUDA_OK(cudaEventRecord(e->chunk_events[c], cs));
    MSFL_CUDA_OK(cudaStreamWaitEvent(e->stream, e->chunk_events[c], 0));
    const int Bc = k.b1 - k.b0;
    const int32_t *co = d_off + k.off_pos, *so = co + (Bc + 1);
    e->lm_shape_scans = B;  // every chunk solves with the CTA shape of the whole call: poses = the unchunked batch's
    rc = scan2map_enqueue(e, Bc, d_qc + k.ci0, co, (uint32_t)ncc, d_qs + k.si0, so, (uint32_t)nsc, d_poses + (size_t)k.b0 * 7,
                          d_stats ? d_stats + k.b0 : nullptr);
    e->lm_shape_scans = 0;
    if (rc) return rc;
  }
  MSFL_CUDA_OK(cudaMemcpyAsync(e->h_poses.p, d_poses, pose_bytes, cudaMemcpyDeviceToHost, e->stream));
  if (stats)
    MSFL_CUDA_OK(cudaMemcpyAsync(e->h_stats.p, d_stats, (size_t)B * sizeof(msfl_stats), cudaMemcpyDeviceToHost, e->stream));
  MSFL_CUDA_OK(cudaStreamSynchronize(e->stream));
  memcpy(poses_tq, e->h_poses.p, pose_bytes);
  if (stats) memcpy(stats, e->h_stats.p, (size_t)B * sizeof(msfl_stats));
  return MSFL_OK;
}

int msfl_scan2map_batch(msfl_engine *e, int B, const msfl_cloud *scan_corner, const msfl_cloud *scan_surf,
                        double *poses_tq, msfl_stats *stats) {
  if (!e || B <= 0 || !scan_corner || !scan_surf || !poses_tq) { set_error("msfl_scan2map_batch: bad argument"); return MSFL_ERR_ARG; }
  if (!e->has_submap) { set_error("msfl_scan2map: no submap set"); return MSFL_ERR_NOSUBMAP; }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  int rc;
  {
    // pipeline H2D against compute when the batch is big enough to split into efficient chunks
    size_t total = 0;
    for (int b = 0; b < B; ++b) total += scan_corner[b].n + scan_surf[b].n;
    const int n_chunks = (B >= 256 && total >= 2000000) ? 4 : ((B >= 64 && total >= 500000) ? 2 : 1);
    if (n_chunks >= 2) return scan2map_batch_pipelined(e, B, scan_corner, scan_surf, poses_tq, stats, n_chunks);
  }
  uint32_t nct = 0, nst = 0;
  if ((rc = upload_batch(e, B, scan_corner, scan_surf, poses_tq, &nct, &nst))) return rc;
  const size_t q_pad = (((size_t)nct + nst) * 16 + 15) & ~(size_t)15;
  const size_t off_pad = ((size_t)2 * (B + 1) * 4 + 15) & ~(size_t)15;
  char *d = e->d_queries.as<char>();
  const float4 *d_qc = (const float4 *)d;
  const float4 *d_qs = d_qc + nct;
  const int32_t *d_c_off = (const int32_t *)(d + q_pad);
  const int32_t *d_s_off = d_c_off + (B + 1);
  double *d_poses = (double *)(d + q_pad + off_pad);
  msfl_stats *d_stats = nullptr;
  if (stats) {
    if ((rc = e->d_stats.reserve((size_t)B * sizeof(msfl_stats)))) return rc;
    if ((rc = e->h_stats.reserve((size_t)B * sizeof(msfl_stats)))) return rc;
    d_stats = e->d_stats.as<msfl_stats>();
    MSFL_CUDA_OK(cudaMemsetAsync(d_stats, 0, (size_t)B * sizeof(msfl_stats), e->stream));
  }
  if ((rc = scan2map_enqueue(e, B, d_qc, d_c_off, nct, d_qs, d_s_off, nst, d_poses, d_stats))) return rc;
  if ((rc = e->h_poses.reserve((size_t)B * 7 * 8))) return rc;
  MSFL_CUDA_OK(cudaMemcpyAsync(e->h_poses.p, d_poses, (size_t)B * 7 * 8, cudaMemcpyDeviceToHost, e->stream));
  if (stats)
    MSFL_CUDA_OK(cudaMemcpyAsync(e->h_stats.p, d_stats, (size_t)B * sizeof(msfl_stats), cudaMemcpyDeviceToHost, e->stream));
  MSFL_CUDA_OK(cudaStreamSynchronize(e->stream));
  memcpy(poses_tq, e->h_poses.p, (size_t)B * 7 * 8);
  if (stats) memcpy(stats, e->h_stats.p, (size_t)B * sizeof(msfl_stats));
  return MSFL_OK;
}

// ---- asynchronous batches: H2D of batch k+1 (copy stream) overlaps the kernels of batch k (engine stream) ----
int msfl_scan2map_batch_submit(msfl_engine *e, int B, const msfl_cloud *corner, const msfl_cloud *surf, const double *poses_in,
                               int want_stats, int *ticket) {
  if (!e || B <= 0 || !corner || !surf || !poses_in || !ticket) { set_error("msfl_scan2map_batch_submit: bad argument"); return MSFL_ERR_ARG; }
  if (!e->has_submap) { set_error("msfl_scan2map: no submap set"); return MSFL_ERR_NOSUBMAP; }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  msfl_engine::BatchSlot &sl = e->slots[e->next_ticket % MSFL_MAX_INFLIGHT];
  if (sl.busy) { set_error("msfl_scan2map_batch_submit: %d batches already in flight", MSFL_MAX_INFLIGHT); return MSFL_ERR_ARG; }
  int rc;
  size_t nct = 0, nst = 0;
  bool contiguous = true;  // packed float4 clouds back to back: DMA straight from the caller's memory
  bool packed3 = true;     // packed xyz-only clouds (12 B points) back to back: DMA + widening on the device
  for (int b = 0; b < B; ++b) {
    if ((rc = check_cloud(&corner[b], false, "scan_corner"))) return rc;
    if ((rc = check_cloud(&surf[b], false, "scan_surf"))) return rc;
    const msfl_cloud *cl[2] = {&corner[b], &surf[b]};
    const msfl_cloud *nx[2] = {b + 1 < B ? &corner[b + 1] : nullptr, b + 1 < B ? &surf[b + 1] : nullptr};
    for (int c = 0; c < 2; ++c) {
      if (cl[c]->stride != 16 || cl[c]->off_xyz != 0) contiguous = false;
      if (nx[c] && (const char *)nx[c]->data != (const char *)cl[c]->data + cl[c]->n * 16) contiguous = false;
      if (cl[c]->stride != 12 || cl[c]->off_xyz != 0 || cl[c]->off_intensity != MSFL_NO_FIELD) packed3 = false;
      if (nx[c] && (const char *)nx[c]->data != (const char *)cl[c]->data + cl[c]->n * 12) packed3 = false;
    }
    nct += corner[b].n;
    nst += surf[b].n;
  }
  if (nct + nst > 0x7fffffffull) { set_error("batch too large"); return MSFL_ERR_ARG; }
  if (!e->copy_stream) MSFL_CUDA_OK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  if (!sl.uploaded) MSFL_CUDA_OK(cudaEventCreateWithFlags(&sl.uploaded, cudaEventDisableTiming));
  if (!sl.done) MSFL_CUDA_OK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
  // slot layout (host staging and device alike): [queries (nct+nst) float4 | offsets 2(B+1) int32 | poses B*7 double]
  const size_t q_bytes = (nct + nst) * 16, q_pad = (q_bytes + 15) & ~(size_t)15;
  const size_t off_bytes = (size_t)2 * (B + 1) * 4, off_pad = (off_bytes + 15) & ~(size_t)15;
  const size_t pose_bytes = (size_t)B * 7 * 8;
  const bool direct = contiguous || packed3;
  if ((rc = sl.h_stage.reserve((direct ? 0 : q_pad) + off_pad + pose_bytes))) return rc;
  if ((rc = sl.d_in.reserve(q_pad + off_pad + pose_bytes))) return rc;
  if (packed3 && (rc = sl.d_in3.reserve((nct + nst) * 12 + 16))) return rc;
  if ((rc = sl.h_out.reserve(pose_bytes))) return rc;
  char *h = sl.h_stage.as<char>(), *d = sl.d_in.as<char>();
  char *h_tab = h + (direct ? 0 : q_pad);
  int32_t *hoff = (int32_t *)h_tab;
  size_t ci = 0, si = 0;
  std::vector<size_t> c_at(B), s_at(B);
  for (int b = 0; b < B; ++b) {
    hoff[b] = (int32_t)ci;
    hoff[B + 1 + b] = (int32_t)si;
    c_at[b] = ci; s_at[b] = si;
    ci += corner[b].n;
    si += surf[b].n;
  }
  hoff[B] = (int32_t)ci;
  hoff[2 * B + 1] = (int32_t)si;
  if (!direct) pack_batch_host(e, 0, B, corner, surf, (float *)h, nct, c_at.data(), s_at.data());
  memcpy(h_tab + off_pad, poses_in, pose_bytes);
  cudaStream_t cs = e->copy_stream;
  if (contiguous) {
    if (nct) MSFL_CUDA_OK(cudaMemcpyAsync(d, corner[0].data, nct * 16, cudaMemcpyHostToDevice, cs));
    if (nst) MSFL_CUDA_OK(cudaMemcpyAsync(d + nct * 16, surf[0].data, nst * 16, cudaMemcpyHostToDevice, cs));
  } else if (packed3) {
    char *d3 = sl.d_in3.as<char>();
    if (nct) MSFL_CUDA_OK(cudaMemcpyAsync(d3, corner[0].data, nct * 12, cudaMemcpyHostToDevice, cs));
    if (nst) MSFL_CUDA_OK(cudaMemcpyAsync(d3 + nct * 12, surf[0].data, nst * 12, cudaMemcpyHostToDevice, cs));
  } else if (q_bytes) {
    MSFL_CUDA_OK(cudaMemcpyAsync(d, h, q_bytes, cudaMemcpyHostToDevice, cs));
  }
  MSFL_CUDA_OK(cudaMemcpyAsync(d + q_pad, h_tab, off_pad + pose_bytes, cudaMemcpyHostToDevice, cs));
  MSFL_CUDA_OK(cudaEventRecord(sl.uploaded, cs));
  MSFL_CUDA_OK(cudaStreamWaitEvent(e->stream, sl.uploaded, 0));
  if (packed3 && nct + nst) {
    const uint32_t n = (uint32_t)(nct + nst);
    k_expand_xyz<<<(n + 255) / 256, 256, 0, e->stream>>>(sl.d_in3.as<float>(), n, (float4 *)d);
    e->launches += 1;
  }
  msfl_stats *d_stats = nullptr;
  if (want_stats) {
    if ((rc = sl.d_stats.reserve((size_t)B * sizeof(msfl_stats)))) return rc;
    if ((rc = sl.h_stats.reserve((size_t)B * sizeof(msfl_stats)))) return rc;
    d_stats = sl.d_stats.as<msfl_stats>();
    MSFL_CUDA_OK(cudaMemsetAsync(d_stats, 0, (size_t)B * sizeof(msfl_stats), e->stream));
  }
  const float4 *d_qc = (const float4 *)d, *d_qs = d_qc + nct;
  const int32_t *d_c_off = (const int32_t *)(d + q_pad), *d_s_off = d_c_off + (B + 1);
  double *d_poses = (double *)(d + q_pad + off_pad);
  if ((rc = scan2map_enqueue(e, B, d_qc, d_c_off, (uint32_t)nct, d_qs, d_s_off, (uint32_t)nst, d_poses, d_stats))) return rc;
  MSFL_CUDA_OK(cudaMemcpyAsync(sl.h_out.p, d_poses, pose_bytes, cudaMemcpyDeviceToHost, e->stream));
  if (want_stats)
    MSFL_CUDA_OK(cudaMemcpyAsync(sl.h_stats.p, d_stats, (size_t)B * sizeof(msfl_stats), cudaMemcpyDeviceToHost, e->stream));
  MSFL_CUDA_OK(cudaEventRecord(sl.done, e->stream));
  sl.busy = true;
  sl.want_stats = want_stats != 0;
  sl.B = B;
  sl.ticket = e->next_ticket++;
  *ticket = sl.ticket;
  return MSFL_OK;
}

int msfl_scan2map_batch_wait(msfl_engine *e, int ticket, double *poses_out, msfl_stats *stats) {
  if (!e || ticket < 0 || !poses_out) { set_error("msfl_scan2map_batch_wait: bad argument"); return MSFL_ERR_ARG; }
  msfl_engine::BatchSlot &sl = e->slots[ticket % MSFL_MAX_INFLIGHT];
  if (!sl.busy || sl.ticket != ticket) { set_error("msfl_scan2map_batch_wait: ticket %d is not in flight", ticket); return MSFL_ERR_ARG; }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  MSFL_CUDA_OK(cudaEventSynchronize(sl.done));
  memcpy(poses_out, sl.h_out.p, (size_t)sl.B * 7 * 8);
  if (stats && sl.want_stats) memcpy(stats, sl.h_stats.p, (size_t)sl.B * sizeof(msfl_stats));
  sl.busy = false;
  return MSFL_OK;
}

int msfl_scan2map(msfl_engine *e, const msfl_cloud *scan_corner, const msfl_cloud *scan_surf, double pose_tq[7],
                  msfl_stats *stats) {
  return msfl_scan2map_batch(e, 1, scan_corner, scan_surf, pose_tq, stats);
}

// IMU-initialised branch for B scans: every scan with its own preintegration table, velocity and gravity.
int msfl_scan2map_deskew_batch(msfl_engine *e, int B, const msfl_cloud *scan_corner, const msfl_cloud *scan_surf,
                               const msfl_deskew *dk, double *poses_tq, msfl_stats *stats) {
  if (!e || B <= 0 || !scan_corner || !scan_surf || !dk || !poses_tq) { set_error("msfl_scan2map_deskew: bad argument"); return MSFL_ERR_ARG; }
  if (!e->has_submap) { set_error("msfl_scan2map_deskew: no submap set"); return MSFL_ERR_NOSUBMAP; }
  size_t rows = 0;
  for (int b = 0; b < B; ++b) {
    if (dk[b].n < 2 || !dk[b].sum_dt || !dk[b].delta_q || !dk[b].delta_p) { set_error("msfl_scan2map_deskew: preintegration table of scan %d needs >= 2 samples", b); return MSFL_ERR_ARG; }
    if (scan_corner[b].off_intensity == MSFL_NO_FIELD || scan_surf[b].off_intensity == MSFL_NO_FIELD) { set_error("msfl_scan2map_deskew: clouds need the intensity (relative time) field"); return MSFL_ERR_ARG; }
    rows += (size_t)dk[b].n;
  }
  if (rows > 0x7fffffffull) { set_error("msfl_scan2map_deskew: preintegration tables too large"); return MSFL_ERR_ARG; }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  int rc;
  uint32_t nc = 0, ns = 0;
  if ((rc = upload_batch(e, B, scan_corner, scan_surf, poses_tq, &nc, &ns))) return rc;
  const size_t total = (size_t)nc + ns;
  if (stats) memset(stats, 0, (size_t)B * sizeof *stats);
  if (total == 0) return MSFL_OK;
  const size_t q_pad = (total * 16 + 15) & ~(size_t)15, off_pad = ((size_t)2 * (B + 1) * 4 + 15) & ~(size_t)15;
  char *d = e->d_queries.as<char>();
  const float4 *d_qc = (const float4 *)d, *d_qs = d_qc + nc;
  const int32_t *d_c_off = (const int32_t *)(d + q_pad), *d_s_off = d_c_off + (B + 1);
  double *d_poses = (double *)(d + q_pad + off_pad);
  // preintegration tables -> device: [sum_dt N | delta_q 4N | delta_p 3N | DeskewScan B | flags B]
  const size_t tab_bytes = rows * 64, scans_bytes = (size_t)B * sizeof(msfl::DeskewScan), flag_bytes = (size_t)B * 4;
  if ((rc = e->k_table.reserve(tab_bytes + scans_bytes + flag_bytes))) return rc;
  if ((rc = e->h_misc.reserve(tab_bytes + scans_bytes + flag_bytes))) return rc;
  double *ht = e->h_misc.as<double>();
  msfl::DeskewScan *hs = (msfl::DeskewScan *)(e->h_misc.as<char>() + tab_bytes);
  size_t row = 0;
  for (int b = 0; b < B; ++b) {
    const size_t n = (size_t)dk[b].n;
    memcpy(ht + row, dk[b].sum_dt, n * 8);
    memcpy(ht + rows + 4 * row, dk[b].delta_q, n * 32);
    memcpy(ht + 5 * rows + 3 * row, dk[b].delta_p, n * 24);
    hs[b].row0 = (uint32_t)row;
    hs[b].n = dk[b].n;
    for (int i = 0; i < 3; ++i) hs[b].V[i] = dk[b].velocity[i], hs[b].G[i] = dk[b].gravity[i];
    row += n;
  }
  memset(e->h_misc.as<char>() + tab_bytes + scans_bytes, 0, flag_bytes);
  MSFL_CUDA_OK(cudaMemcpyAsync(e->k_table.p, ht, tab_bytes + scans_bytes + flag_bytes, cudaMemcpyHostToDevice, st));
  const double *t_dt = e->k_table.as<double>(), *t_dq = t_dt + rows, *t_dp = t_dt + 5 * rows;
  const msfl::DeskewScan *d_scans = (const msfl::DeskewScan *)(e->k_table.as<char>() + tab_bytes);
  int *d_flags = (int *)(e->k_table.as<char>() + tab_bytes + scans_bytes);
  if ((rc = e->k_dsk.reserve(total * 64))) return rc;
  if ((rc = e->k_pprime.reserve(total * 32))) return rc;
  if ((rc = e->k_o4.reserve(total * 32))) return rc;
  if ((rc = e->d_corr.reserve((total + 1) * 48))) return rc;
  if ((rc = e->d_status.reserve((size_t)B * 4 + 16))) return rc;
  msfl_stats *d_stats = nullptr;
  if (stats) {
    if ((rc = e->d_stats.reserve((size_t)B * sizeof(msfl_stats)))) return rc;
    if ((rc = e->h_stats.reserve((size_t)B * sizeof(msfl_stats)))) return rc;
    d_stats = e->d_stats.as<msfl_stats>();
    MSFL_CUDA_OK(cudaMemsetAsync(d_stats, 0, (size_t)B * sizeof(msfl_stats), st));
  }
  // a query whose time is outside its scan's table keeps stale (dq, dp, p', o): give them defined contents; the call
  // fails below and no pose is written, so they never reach a result
  MSFL_CUDA_OK(cudaMemsetAsync(e->k_dsk.p, 0, total * 64, st));
  MSFL_CUDA_OK(cudaMemsetAsync(e->k_pprime.p, 0, total * 32, st));
  MSFL_CUDA_OK(cudaMemsetAsync(e->k_o4.p, 0, total * 32, st));
  if ((rc = launch_deskew_prepare(e, B, d_scans, t_dt, t_dq, t_dp, d_qc, d_c_off, d_s_off, nc, (uint32_t)total,
                                  e->k_dsk.as<double>(), e->k_pprime.as<double>(), e->k_o4.as<double>(), d_flags)))
    return rc;
  const double *pp_c = e->k_pprime.as<double>(), *pp_s = pp_c + (size_t)nc * 4;
  for (int outer = 0; outer < e->params.num_outer; ++outer) {
    if ((rc = launch_associate_map_deskew(e, B, d_qc, d_c_off, nc, d_qs, d_s_off, ns, d_poses, e->k_o4.as<double>(),
                                          e->k_dsk.as<double>(), e->d_corr.as<double>(), nullptr)))
      return rc;
    stage_begin(e, 1);
    rc = launch_lm_solve_pd(e, B, pp_c, d_c_off, nc, pp_s, d_s_off, e->d_corr.as<double>(), d_poses, e->d_status.as<int32_t>(),
                            d_stats, outer, 0);
    stage_end(e);
    if (rc) return rc;
  }
  if ((rc = e->h_poses.reserve((size_t)B * 7 * 8 + flag_bytes))) return rc;
  int *h_flags = (int *)(e->h_poses.as<char>() + (size_t)B * 7 * 8);
  MSFL_CUDA_OK(cudaMemcpyAsync(h_flags, d_flags, flag_bytes, cudaMemcpyDeviceToHost, st));
  MSFL_CUDA_OK(cudaMemcpyAsync(e->h_poses.p, d_poses, (size_t)B * 7 * 8, cudaMemcpyDeviceToHost, st));
  if (stats) MSFL_CUDA_OK(cudaMemcpyAsync(e->h_stats.p, d_stats, (size_t)B * sizeof(msfl_stats), cudaMemcpyDeviceToHost, st));
  MSFL_CUDA_OK(cudaStreamSynchronize(st));
  for (int b = 0; b < B; ++b)
    if (h_flags[b]) {
      set_error("msfl_scan2map_deskew: a point time of scan %d lies outside its preintegration window [%g, %g]", b,
                dk[b].sum_dt[0], dk[b].sum_dt[dk[b].n - 1]);
      if (stats) memset(stats, 0, (size_t)B * sizeof *stats);
      return MSFL_ERR_ARG;
    }
  memcpy(poses_tq, e->h_poses.p, (size_t)B * 7 * 8);
  if (stats) memcpy(stats, e->h_stats.p, (size_t)B * sizeof(msfl_stats));
  return MSFL_OK;
}

int msfl_scan2map_deskew(msfl_engine *e, const msfl_cloud *scan_corner, const msfl_cloud *scan_surf,
                         const msfl_deskew *dk, double pose_tq[7], msfl_stats *stats) {
  return msfl_scan2map_deskew_batch(e, 1, scan_corner, scan_surf, dk, pose_tq, stats);
}

int msfl_associate_map(msfl_engine *e, const msfl_cloud *scan_corner, const msfl_cloud *scan_surf,
                       const double pose_tq[7], int32_t *knn_idx, double *corr) {
  if (!e || !scan_corner || !scan_surf || !pose_tq) { set_error("msfl_associate_map: bad argument"); return MSFL_ERR_ARG; }
  if (!e->has_submap) { set_error("msfl_associate_map: no submap set"); return MSFL_ERR_NOSUBMAP; }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  int rc;
  uint32_t nct = 0, nst = 0;
  if ((rc = upload_batch(e, 1, scan_corner, scan_surf, pose_tq, &nct, &nst))) return rc;
  const size_t total = (size_t)nct + nst;
  if (total == 0) return MSFL_OK;
  const size_t q_pad = (total * 16 + 15) & ~(size_t)15;
  const size_t off_pad = ((size_t)2 * 2 * 4 + 15) & ~(size_t)15;
  char *d = e->d_queries.as<char>();
  const float4 *d_qc = (const float4 *)d;
  const float4 *d_qs = d_qc + nct;
  const int32_t *d_c_off = (const int32_t *)(d + q_pad);
  const int32_t *d_s_off = d_c_off + 2;
  double *d_poses = (double *)(d + q_pad + off_pad);
  if ((rc = e->d_corr.reserve((total + 1) * 6 * 8))) return rc;
  if ((rc = e->d_knn.reserve(total * 5 * 4))) return rc;
  if ((rc = launch_associate_map(e, 1, d_qc, d_c_off, nct, d_qs, d_s_off, nst, d_poses, e->d_corr.as<double>(),
                                 e->d_knn.as<int32_t>())))
    return rc;
  if (knn_idx) MSFL_CUDA_OK(cudaMemcpyAsync(knn_idx, e->d_knn.p, total * 5 * 4, cudaMemcpyDeviceToHost, e->stream));
  if (corr) MSFL_CUDA_OK(cudaMemcpyAsync(corr, e->d_corr.p, total * 6 * 8, cudaMemcpyDeviceToHost, e->stream));
  MSFL_CUDA_OK(cudaStreamSynchronize(e->stream));
  return MSFL_OK;
}

int msfl_cloud_from_pointcloud2(const uint8_t *data, uint32_t width, uint32_t height, uint32_t point_step,
                                uint32_t row_step, int is_bigendian, const msfl_pc2_field *fields, int n_fields,
                                msfl_cloud *out) {
  if (!out || !fields || n_fields <= 0 || (!data && (size_t)width * height > 0)) { set_error("pointcloud2: bad argument"); return MSFL_ERR_ARG; }
  if (is_bigendian) { set_error("pointcloud2: big-endian data is not supported"); return MSFL_ERR_ARG; }
  if (row_step != width * point_step) { set_error("pointcloud2: padded rows (row_step != width * point_step)"); return MSFL_ERR_ARG; }
  size_t ox = MSFL_NO_FIELD, oy = MSFL_NO_FIELD, oz = MSFL_NO_FIELD, oi = MSFL_NO_FIELD, orr = MSFL_NO_FIELD;
  for (int i = 0; i < n_fields; ++i) {
    const msfl_pc2_field &f = fields[i];
    if (!f.name) continue;
    const bool f32 = f.datatype == 7, u16 = f.datatype == 4;
    if (!strcmp(f.name, "x") && f32) ox = f.offset;
    else if (!strcmp(f.name, "y") && f32) oy = f.offset;
    else if (!strcmp(f.name, "z") && f32) oz = f.offset;
    else if (!strcmp(f.name, "intensity") && f32) oi = f.offset;
    else if (!strcmp(f.name, "ring") && u16) orr = f.offset;
  }
  if (ox == MSFL_NO_FIELD || oy != ox + 4 || oz != ox + 8) { set_error("pointcloud2: needs consecutive FLOAT32 x, y, z fields"); return MSFL_ERR_ARG; }
  if (ox + 12 > point_step || (oi != MSFL_NO_FIELD && oi + 4 > point_step) || (orr != MSFL_NO_FIELD && orr + 2 > point_step)) {
    set_error("pointcloud2: field outside point_step");
    return MSFL_ERR_ARG;
  }
  out->data = data;
  out->n = (size_t)width * height;
  out->stride = point_step;
  out->off_xyz = ox;
  out->off_intensity = oi;
  out->off_ring = orr;
  return MSFL_OK;
}

int msfl_accumulate(msfl_engine *e, const float *p_xyz, const double *corr, int n_edge, int n_plane,
                    const double pose_tq[7], double *cost, double H[36], double g[6]) {
  if (!e || !p_xyz || !corr || !pose_tq || n_edge < 0 || n_plane < 0) { set_error("msfl_accumulate: bad argument"); return MSFL_ERR_ARG; }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  const size_t n = (size_t)n_edge + n_plane;
  int rc;
  if ((rc = e->h_stage.reserve(n * 16 + 64))) return rc;
  if ((rc = e->d_queries.reserve(n * 16 + 64))) return rc;
  if ((rc = e->d_corr.reserve((n + 1) * 48))) return rc;
  if ((rc = e->d_misc.reserve(64 * 8))) return rc;
  float *h = e->h_stage.as<float>();
  for (size_t i = 0; i < n; ++i) {
    h[4 * i] = p_xyz[3 * i]; h[4 * i + 1] = p_xyz[3 * i + 1]; h[4 * i + 2] = p_xyz[3 * i + 2]; h[4 * i + 3] = 0.f;
  }
  MSFL_CUDA_OK(cudaMemcpyAsync(e->d_queries.p, h, n * 16, cudaMemcpyHostToDevice, e->stream));
  MSFL_CUDA_OK(cudaMemcpyAsync(e->d_corr.p, corr, n * 48, cudaMemcpyHostToDevice, e->stream));
  double *d_pose = e->d_misc.as<double>();
  MSFL_CUDA_OK(cudaMemcpyAsync(d_pose, pose_tq, 56, cudaMemcpyHostToDevice, e->stream));
  if ((rc = launch_accumulate(e, e->d_queries.as<float4>(), e->d_corr.as<double>(), n_edge, n_plane, d_pose, d_pose + 8)))
    return rc;
  double out[28];
  MSFL_CUDA_OK(cudaMemcpyAsync(out, d_pose + 8, sizeof out, cudaMemcpyDeviceToHost, e->stream));
  MSFL_CUDA_OK(cudaStreamSynchronize(e->stream));
  if (cost) *cost = out[27];
  if (g) for (int i = 0; i < 6; ++i) g[i] = out[21 + i];
  if (H) {
    int k = 0;
    for (int u = 0; u < 6; ++u)
      for (int v = u; v < 6; ++v) { H[u * 6 + v] = out[k]; H[v * 6 + u] = out[k]; ++k; }
  }
  return MSFL_OK;
}

}  // extern "C"
