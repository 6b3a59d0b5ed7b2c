// features.cu -- scan registration: the block of RealHandleLaserCloudMessage
// between laser_cloud_in (msf_loam_node.cc:166) and scan.* (msf_loam_node.cc:360-371).
//   a-1  invalid-point removal (:86-111), stable ring split + relative time from azimuth (:128-156)
//   a-2  11-tap curvature, fp32 sum in source order, squares in fp64 (:213-240)
//   a-3  per ring, 6 sectors in order: sort by curvature, greedy pick of 2 sharp / 20 less-sharp /
//        4 flat with +-5 neighbour suppression, everything FLAT/UNKNOWN -> less-flat (:251-351)
//   a-4  extrinsic applied to every output cloud (:367-371)
// The greedy pick is sequential inside a ring (flags leak across sector borders), so the pick
// kernel runs one CTA per ring: the ring is staged in shared memory, its sectors are sorted by one segmented
// block-wide bitonic sort, warp 0 does the order-dependent sweeps (lane-parallel candidate scan and neighbour
// suppression) and compacts the less-flat list.
#include <cub/device/device_radix_sort.cuh>

#include <limits.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "msfl_internal.h"
#include "msfl_math.cuh"

namespace msfl {

constexpr int kMaxSectorPts = 8192;  // = the largest k_feat_pick shape's sort-key capacity
constexpr double kTwoPi = 2 * 3.14159265358979323846;

struct FeatMeta {
  int first_valid, bad_ring, n_valid, sector_overflow;
  uint32_t ring_start[MSFL_MAX_RINGS + 1];
  int first_dec[MSFL_MAX_RINGS];
  int cnt[4][MSFL_MAX_RINGS];  // per ring: sharp, less_sharp, flat, less_flat
  int tot[4];
};

// Batched form: B scans live back to back in every per-point array; soff (B + 1 entries, device) gives each scan's
// first point.  Per-point kernels run with grid.y = scan, per-ring kernels with grid = (ring, scan); inside a kernel
// every array is shifted to the scan's own base, so the bodies read like the single-scan code they were.
struct ScanView { uint32_t b, base, n; };
__device__ __forceinline__ ScanView scan_view(const uint32_t *__restrict__ soff, uint32_t b) {
  const uint32_t base = __ldg(soff + b);
  return ScanView{b, base, __ldg(soff + b + 1) - base};
}

__global__ void k_feat_init(FeatMeta *metas) {
  FeatMeta *m = metas + blockIdx.x;
  const int t = threadIdx.x;
  if (t == 0) { m->first_valid = INT_MAX; m->bad_ring = 0; m->n_valid = 0; m->sector_overflow = 0; }
  if (t < MSFL_MAX_RINGS) {
    m->first_dec[t] = INT_MAX;
    for (int k = 0; k < 4; ++k) m->cnt[k][t] = 0;
  }
  if (t < 4) m->tot[t] = 0;
}

// Wire format -> packed arrays on the GPU: the raw AoS bytes of a
// pcl::PointCloud<PointXYZIRT> / sensor_msgs::PointCloud2 data buffer (fields x, y, z, intensity, ring at
// arbitrary, possibly unaligned byte offsets; pcl::fromROSMsg at msf_loam_node.cc:166-167, field list
// common.h:53-62) are copied to the device as they are and unpacked here, so the host never loops over points.
__device__ __forceinline__ float load_f32_unaligned(const unsigned char *p) {
  const uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
  return __uint_as_float(v);
}

__global__ void k_unpack_aos(const unsigned char *__restrict__ raw, uint32_t n, uint32_t stride, uint32_t off_xyz,
                             uint32_t off_i, uint32_t off_ring, int has_i, float4 *__restrict__ pts,
                             uint16_t *__restrict__ ring) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char *p = raw + (size_t)i * stride;
  float4 o;
  if (((stride | off_xyz) & 3u) == 0) {  // aligned fast path (PCL structs are 16-byte aligned)
    const float *f = reinterpret_cast<const float *>(p + off_xyz);
    o.x = f[0]; o.y = f[1]; o.z = f[2];
  } else {
    o.x = load_f32_unaligned(p + off_xyz); o.y = load_f32_unaligned(p + off_xyz + 4); o.z = load_f32_unaligned(p + off_xyz + 8);
  }
  o.w = has_i ? load_f32_unaligned(p + off_i) : 0.f;
  pts[i] = o;
  ring[i] = (uint16_t)((uint32_t)p[off_ring] | ((uint32_t)p[off_ring + 1] << 8));
}

// RemoveInvalidPointsFromCloud (:96-103): float norm vs double min_range, non-finite dropped.
__global__ void k_feat_keys(const float4 *__restrict__ raw, const uint16_t *__restrict__ ring, const uint32_t *__restrict__ soff,
                            double min_range, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, FeatMeta *metas) {
  const ScanView sv = scan_view(soff, blockIdx.y);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sv.n) return;
  raw += sv.base; ring += sv.base; keys += sv.base; vals += sv.base;
  FeatMeta *m = metas + sv.b;
  const float4 p = raw[i];
  const float nr = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y)), __fmul_rn(p.z, p.z)));
  const bool valid = !((double)nr < min_range || !isfinite(p.x) || !isfinite(p.y) || !isfinite(p.z));
  uint32_t key = 255u;
  if (valid) {
    const uint32_t r = ring[i];
    if (r >= MSFL_MAX_RINGS) atomicOr(&m->bad_ring, 1);  // CHECK_LT(point.ring, ...) :136
    else key = r;
    atomicMin(&m->first_valid, (int)i);
  }
  keys[i] = (sv.b << 8) | key;  // scan-major, ring-minor: one stable sort orders every scan of the batch
  vals[i] = i;
}

__global__ void k_feat_ring_start(const uint32_t *__restrict__ keys_sorted, const uint32_t *__restrict__ soff, FeatMeta *metas) {
  const ScanView sv = scan_view(soff, blockIdx.x);
  keys_sorted += sv.base;
  FeatMeta *m = metas + sv.b;
  const uint32_t n = sv.n;
  const uint32_t r = threadIdx.x;
  if (r > MSFL_MAX_RINGS) return;
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if ((keys_sorted[mid] & 255u) < r) lo = mid + 1;
    else hi = mid;
  }
  m->ring_start[r] = lo;
  if (r == MSFL_MAX_RINGS) m->n_valid = (int)lo;
}

// ComputeRelaTimeForEachPoint (:131-151), first half: the raw relative angle of every point.
__global__ void k_feat_angles(const float4 *__restrict__ raw, const uint32_t *__restrict__ vals, const uint32_t *__restrict__ soff,
                              const FeatMeta *__restrict__ metas, double *__restrict__ rel) {
  const ScanView sv = scan_view(soff, blockIdx.y);
  const FeatMeta *m = metas + sv.b;
  raw += sv.base; vals += sv.base; rel += sv.base;
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= (uint32_t)m->n_valid) return;
  // :131 / :139 call atan2 unqualified on two floats with <math.h> in the include graph: the float overload is chosen
  // (see the oracle, msfl_oracle.c "atan2").  The float result is produced here by rounding the double-precision atan2
  // once -- the correctly rounded float, within 1 ulp of any libm's atan2f.
  const float4 f = raw[m->first_valid];
  const double start_ori = (double)(-(float)atan2((double)f.y, (double)f.x));  // :131
  const float4 p = raw[vals[j]];
  const double ori = (double)(-(float)atan2((double)p.y, (double)p.x));        // :139
  rel[j] = fmod(__dadd_rn(__dsub_rn(ori, start_ori), kTwoPi), kTwoPi);  // :142
}

// "if (relative_angle < last_relative_angles[ring]) += 2 pi" (:145-149): once a point of a ring is
// bumped every later point of that ring is bumped too, so the recurrence collapses to "j >= first
// position whose raw angle is below its predecessor's".
__global__ void k_feat_first_dec(const uint32_t *__restrict__ keys_sorted, const double *__restrict__ rel,
                                 const uint32_t *__restrict__ soff, FeatMeta *metas) {
  const ScanView sv = scan_view(soff, blockIdx.y);
  FeatMeta *m = metas + sv.b;
  keys_sorted += sv.base; rel += sv.base;
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0 || j >= (uint32_t)m->n_valid) return;
  const uint32_t r = keys_sorted[j] & 255u;
  if ((keys_sorted[j - 1] & 255u) == r && rel[j] < rel[j - 1]) atomicMin(&m->first_dec[r], (int)j);
}

__global__ void k_feat_full(const float4 *__restrict__ raw, const uint32_t *__restrict__ keys_sorted,
                            const uint32_t *__restrict__ vals, const double *__restrict__ rel, const uint32_t *__restrict__ soff,
                            const FeatMeta *__restrict__ metas, double scan_period, float4 *__restrict__ full,
                            uint16_t *__restrict__ ring_out) {
  const ScanView sv = scan_view(soff, blockIdx.y);
  const FeatMeta *m = metas + sv.b;
  raw += sv.base; keys_sorted += sv.base; vals += sv.base; rel += sv.base; full += sv.base; ring_out += sv.base;
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= (uint32_t)m->n_valid) return;
  const uint32_t r = keys_sorted[j] & 255u;
  double a = rel[j];
  if ((int)j >= m->first_dec[r]) a = __dadd_rn(a, kTwoPi);
  const double t = __dmul_rn(__ddiv_rn(a, kTwoPi), scan_period);  // :151
  const float4 p = raw[vals[j]];
  full[j] = make_float4(p.x, p.y, p.z, (float)t);  // intensity := time (:152-153)
  ring_out[j] = (uint16_t)r;
}

// a-2 curvature (:213-240)
__global__ void k_feat_curv(const float4 *__restrict__ full, const uint32_t *__restrict__ soff, const FeatMeta *__restrict__ metas,
                            float *__restrict__ curv, int32_t *__restrict__ label, uint8_t *__restrict__ picked) {
  const ScanView sv = scan_view(soff, blockIdx.y);
  const FeatMeta *m = metas + sv.b;
  full += sv.base; curv += sv.base; label += sv.base; picked += sv.base;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = m->n_valid;
  if (i >= N) return;
  label[i] = 0;
  picked[i] = 0;
  float c = 0.f;
  if (i >= 5 && i < N - 5) {
    float4 q[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) q[k] = full[i - 5 + k];
    double d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#define AX(k) (a == 0 ? q[k].x : (a == 1 ? q[k].y : q[k].z))
      float s = __fadd_rn(AX(0), AX(1));
      s = __fadd_rn(s, AX(2));
      s = __fadd_rn(s, AX(3));
      s = __fadd_rn(s, AX(4));
      s = __fsub_rn(s, __fmul_rn(10.0f, AX(5)));
      s = __fadd_rn(s, AX(6));
      s = __fadd_rn(s, AX(7));
      s = __fadd_rn(s, AX(8));
      s = __fadd_rn(s, AX(9));
      s = __fadd_rn(s, AX(10));
#undef AX
      d[a] = (double)s;
    }
    c = (float)__dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2]));  // :236
  }
  curv[i] = c;
}

__device__ __forceinline__ float gap_sq(const float4 a, const float4 b) {  // Vector3f squaredNorm (:291-293)
  const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// a-3: one CTA per ring.  The ring's points, labels and picked flags are staged in shared memory (rings of up to
// kRingCap points; longer rings work on the global arrays), all sectors of the ring are sorted by ONE segmented
// bitonic sort (the sort keys -- curvature bits, index -- do not depend on the picking), and warp 0 then runs the
// order-dependent sweeps sector by sector: the lanes scan 32 sorted candidates at a time for "above the threshold and
// not yet picked", the picks of a chunk are taken in order, and the +-5 neighbour suppression of a pick is evaluated
// by ten lanes at once.  Same decisions, same output order as the sequential loops of :263-350.
// Three shapes <threads, ring points staged in shared memory, sort keys in shared memory (all sectors of a ring, padded to
// powers of two)>.  The order-dependent sweeps run on ONE warp, so what a batch needs is many rings in flight per SM:
// the shape is the smallest that holds the rings the engine saw in its previous extraction (a sensor does not change
// between calls) -- VLP-16 / OS1-128 rings (<= 2048 points) run three 74 KB CTAs per SM, HDL-64E rings two 84 KB CTAs,
// anything else (and the first call) the 148 KB shape.  A ring longer than the staged capacity still works (on the
// global arrays); a sector that does not fit the keys raises sector_overflow and the host re-runs the larger shape.
template <int NT, int RING_CAP, int KEYS_CAP>
struct PickShape {
  static constexpr int kThreads = NT, kRingCap = RING_CAP, kKeysCap = KEYS_CAP;
  static constexpr size_t kSmem = (size_t)KEYS_CAP * 8 + (size_t)RING_CAP * (16 + 4 + 1);
};
using PickSmall = PickShape<512, 2048, 4096>;
using PickMid = PickShape<1024, 2560, 4096>;
using PickBig = PickShape<1024, 4096, kMaxSectorPts>;

template <class SHAPE>
__global__ void __launch_bounds__(SHAPE::kThreads)
k_feat_pick(const float4 *__restrict__ full, const float *__restrict__ curv, int32_t *__restrict__ label,
            uint8_t *__restrict__ picked, const uint32_t *__restrict__ soff, FeatMeta *metas, double curv_thr, double gap_thr,
            int n_sectors, int n_sharp, int n_less, int n_flat, int32_t *__restrict__ slot_sharp, int32_t *__restrict__ slot_less,
            int32_t *__restrict__ slot_flat, int32_t *__restrict__ lessflat_tmp) {
  const ScanView sv = scan_view(soff, blockIdx.y);
  FeatMeta *m = metas + sv.b;
  full += sv.base; curv += sv.base; label += sv.base; picked += sv.base; lessflat_tmp += sv.base;
  slot_sharp += (size_t)sv.b * MSFL_MAX_RINGS * n_sectors * n_sharp;
  slot_less += (size_t)sv.b * MSFL_MAX_RINGS * n_sectors * n_less;
  slot_flat += (size_t)sv.b * MSFL_MAX_RINGS * n_sectors * n_flat;
  constexpr int kPickThreads = SHAPE::kThreads, kRingCap = SHAPE::kRingCap, kKeysCap = SHAPE::kKeysCap;
  extern __shared__ __align__(16) unsigned char pick_smem[];
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(pick_smem);
  float4 *s_full = reinterpret_cast<float4 *>(pick_smem + (size_t)kKeysCap * 8);
  int32_t *s_label = reinterpret_cast<int32_t *>(pick_smem + (size_t)kKeysCap * 8 + (size_t)kRingCap * 16);
  uint8_t *s_picked = pick_smem + (size_t)kKeysCap * 8 + (size_t)kRingCap * 20;
  const int r = blockIdx.x;
  const int rs = (int)m->ring_start[r], re = (int)m->ring_start[r + 1];
  const int start = rs + 5, end = re - 6;  // :192-194
  if (end - start < 6) return;             // :252
  const int tid = threadIdx.x, lane = tid & 31;
  const int ring_len = re - rs;
  const bool staged = ring_len <= kRingCap;
  // F / LB / PK are indexed by (absolute point index - ob)
  const float4 *F = staged ? s_full : full;
  int32_t *LB = staged ? s_label : label;
  uint8_t *PK = staged ? s_picked : picked;
  const int ob = staged ? rs : 0;
  if (staged)
    for (int i = tid; i < ring_len; i += kPickThreads) {
      s_full[i] = full[rs + i];
      s_label[i] = 0;
      s_picked[i] = 0;
    }
  // sector geometry (:256-259) and the common padded size of the segmented sort
  int max_cnt = 0;
  for (int j = 0; j < n_sectors; ++j) {
    const int sp = start + (end - start) * j / n_sectors, ep = start + (end - start) * (j + 1) / n_sectors - 1;
    max_cnt = max(max_cnt, ep - sp + 1);
  }
  int P = 1;
  while (P < max_cnt) P <<= 1;
  if (P > kKeysCap) {
    if (tid == 0) atomicOr(&m->sector_overflow, 1);
    return;
  }
  const int batch = min(n_sectors, kKeysCap / P);  // sectors sorted together
  int ns = 0, nl = 0, nf = 0, n_lf = 0;            // list lengths (meaningful in warp 0)
  int32_t *my_sharp = slot_sharp + (size_t)r * n_sectors * n_sharp;
  int32_t *my_less = slot_less + (size_t)r * n_sectors * n_less;
  int32_t *my_flat = slot_flat + (size_t)r * n_sectors * n_flat;
  int32_t *my_lf = lessflat_tmp + rs;
  for (int j0 = 0; j0 < n_sectors; j0 += batch) {
    const int j1 = min(n_sectors, j0 + batch), nb = j1 - j0;
    __syncthreads();  // staging done / previous batch's keys no longer needed
    // std::sort by curvature (:263) -> bitonic sort on (curvature bits, index): curvature >= 0 so the uint order of the
    // bits is the float order; ties resolved by index (deterministic).  Segment q of the batch lives in keys[q P .. q P + P).
    for (int t = tid; t < nb * P; t += kPickThreads) {
      const int q = t / P, k = t - q * P, j = j0 + q;
      const int sp = start + (end - start) * j / n_sectors, ep = start + (end - start) * (j + 1) / n_sectors - 1;
      keys[t] = (k <= ep - sp) ? (((unsigned long long)__float_as_uint(curv[sp + k]) << 32) | (unsigned)(sp + k)) : ~0ull;
    }
    __syncthreads();
    const int lp = __ffs(P) - 1;  // P = 1 << lp; all index arithmetic below is shifts and masks
    for (int size = 2; size <= P; size <<= 1) {
      for (int ls = __ffs(size) - 2; ls >= 0; --ls) {  // stride = 1 << ls
        const int stride = 1 << ls;
        for (int t = tid; t < (nb << (lp - 1)); t += kPickThreads) {
          const int q = lp > 0 ? t >> (lp - 1) : t, u = t & ((P >> 1) - 1);
          const int lo = ((u >> ls) << (ls + 1)) | (u & (stride - 1));
          const int hi = lo + stride;
          const bool up = ((lo & size) == 0);
          unsigned long long *K = keys + ((size_t)q << lp);
          const unsigned long long a = K[lo], b = K[hi];
          if ((a > b) == up) { K[lo] = b; K[hi] = a; }
        }
        __syncthreads();
      }
    }
    if (tid < 32) {  // warp 0: the order-dependent part, sector after sector
      for (int j = j0; j < j1; ++j) {
        const int sp = start + (end - start) * j / n_sectors, ep = start + (end - start) * (j + 1) / n_sectors - 1;
        const int cnt = ep - sp + 1;
        if (cnt <= 0) continue;
        const unsigned long long *K = keys + (size_t)(j - j0) * P;
        // --- largest curvature first: 2 sharp, up to 20 less-sharp (:272-305)
        int largest = 0;
        bool stop = false;
        for (int k0 = cnt - 1; k0 >= 0 && !stop; k0 -= 32) {
          const int kk = k0 - lane;
          const unsigned long long key = kk >= 0 ? K[kk] : 0ull;
          const int ind = (int)(unsigned)(key & 0xffffffffull);
          const bool above = kk >= 0 && (double)__uint_as_float((unsigned)(key >> 32)) > curv_thr;
          unsigned pend = __ballot_sync(0xffffffffu, above);
          if (!pend) break;  // sorted: nothing further exceeds the threshold
          while (pend) {
            const int l = __ffs(pend) - 1;
            pend &= pend - 1;
            const int ci = __shfl_sync(0xffffffffu, ind, l);
            if (PK[ci - ob]) continue;
            ++largest;
            if (largest > n_less) { stop = true; break; }
            if (lane == 0) {
              if (largest <= n_sharp) {
                LB[ci - ob] = 1;
                my_sharp[ns++] = ci;
                my_less[nl++] = ci;
              } else {
                LB[ci - ob] = 2;
                my_less[nl++] = ci;
              }
              PK[ci - ob] = 1;
            }
            // neighbour suppression: lanes 0-4 test the forward gaps 1..5, lanes 8-12 the backward gaps 1..5 (:289-303)
            bool fail = false;
            if (lane < 5) fail = (double)gap_sq(F[ci + lane + 1 - ob], F[ci + lane - ob]) > gap_thr;
            else if (lane >= 8 && lane < 13) fail = (double)gap_sq(F[ci - (lane - 7) - ob], F[ci - (lane - 8) - ob]) > gap_thr;
            const unsigned fm = __ballot_sync(0xffffffffu, fail);
            const int nfw = (fm & 0x1fu) ? __ffs(fm & 0x1fu) - 1 : 5, nbw = ((fm >> 8) & 0x1fu) ? __ffs((fm >> 8) & 0x1fu) - 1 : 5;
            if (lane < nfw) { PK[ci + lane + 1 - ob] = 1; LB[ci + lane + 1 - ob] = 2; }
            if (lane >= 8 && lane - 8 < nbw) { PK[ci - (lane - 7) - ob] = 1; LB[ci - (lane - 7) - ob] = 2; }
            __syncwarp();
          }
        }
        __syncwarp();
        // --- smallest curvature first: 4 flat (:309-336)
        int smallest = 0;
        stop = false;
        for (int k0 = 0; k0 < cnt && !stop; k0 += 32) {
          const int kk = k0 + lane;
          const unsigned long long key = kk < cnt ? K[kk] : 0ull;
          const int ind = (int)(unsigned)(key & 0xffffffffull);
          const bool below = kk < cnt && (double)__uint_as_float((unsigned)(key >> 32)) < curv_thr;
          unsigned pend = __ballot_sync(0xffffffffu, below);
          if (!pend) break;  // sorted: everything further is at or above the threshold
          while (pend) {
            const int l = __ffs(pend) - 1;
            pend &= pend - 1;
            const int ci = __shfl_sync(0xffffffffu, ind, l);
            if (PK[ci - ob]) continue;
            if (lane == 0) {
              LB[ci - ob] = 3;
              my_flat[nf++] = ci;
            }
            if (++smallest >= n_flat) { stop = true; break; }
            bool fail = false;
            if (lane < 5) fail = (double)gap_sq(F[ci + lane + 1 - ob], F[ci + lane - ob]) > gap_thr;
            else if (lane >= 8 && lane < 13) fail = (double)gap_sq(F[ci - (lane - 7) - ob], F[ci - (lane - 8) - ob]) > gap_thr;
            const unsigned fm = __ballot_sync(0xffffffffu, fail);
            const int nfw = (fm & 0x1fu) ? __ffs(fm & 0x1fu) - 1 : 5, nbw = ((fm >> 8) & 0x1fu) ? __ffs((fm >> 8) & 0x1fu) - 1 : 5;
            if (lane == 16) PK[ci - ob] = 1;
            if (lane < nfw) PK[ci + lane + 1 - ob] = 1;
            if (lane >= 8 && lane - 8 < nbw) PK[ci - (lane - 7) - ob] = 1;
            __syncwarp();
          }
        }
        __syncwarp();
        // --- everything FLAT or UNKNOWN of this sector -> less-flat, in scan order (:339-344)
        for (int k0 = sp; k0 <= ep; k0 += 32) {
          const int k = k0 + lane;
          const bool take = (k <= ep) && (LB[k - ob] == 3 || LB[k - ob] == 0);
          const unsigned bal = __ballot_sync(0xffffffffu, take);
          if (take) my_lf[n_lf + __popc(bal & ((1u << lane) - 1u))] = k;
          n_lf += __popc(bal);
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  if (staged)
    for (int i = tid; i < ring_len; i += kPickThreads) label[rs + i] = s_label[i];
  if (tid == 0) {
    m->cnt[0][r] = ns;
    m->cnt[1][r] = nl;
    m->cnt[2][r] = nf;
    m->cnt[3][r] = n_lf;
  }
}

// ring-major concatenation of the per-ring lists (the push_back order of :279-283, :314, :350): one CTA per ring,
// its four output offsets are the exclusive prefix sums of the per-ring counts (warp w scans list w)
__global__ void __launch_bounds__(128)
k_feat_compact(const uint32_t *__restrict__ soff, FeatMeta *metas, int n_sectors, int n_sharp, int n_less, int n_flat,
               const int32_t *__restrict__ slot_sharp, const int32_t *__restrict__ slot_less,
               const int32_t *__restrict__ slot_flat, const int32_t *__restrict__ lessflat_tmp,
               int32_t *__restrict__ out_sharp, int32_t *__restrict__ out_less,
               int32_t *__restrict__ out_flat, int32_t *__restrict__ out_lf) {
  const ScanView sv = scan_view(soff, blockIdx.y);
  FeatMeta *m = metas + sv.b;
  lessflat_tmp += sv.base; out_sharp += sv.base; out_less += sv.base; out_flat += sv.base; out_lf += sv.base;
  slot_sharp += (size_t)sv.b * MSFL_MAX_RINGS * n_sectors * n_sharp;
  slot_less += (size_t)sv.b * MSFL_MAX_RINGS * n_sectors * n_less;
  slot_flat += (size_t)sv.b * MSFL_MAX_RINGS * n_sectors * n_flat;
  __shared__ int off[4], cnt[4];
  const int r = blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    int before = 0, total = 0;
    for (int q = lane; q < MSFL_MAX_RINGS; q += 32) {
      const int c = m->cnt[w][q];
      total += c;
      if (q < r) before += c;
    }
    for (int o = 16; o > 0; o >>= 1) {
      before += __shfl_xor_sync(0xffffffffu, before, o);
      total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    if (lane == 0) {
      off[w] = before;
      cnt[w] = m->cnt[w][r];
      if (r == 0) m->tot[w] = total;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < cnt[0]; k += blockDim.x) out_sharp[off[0] + k] = slot_sharp[(size_t)r * n_sectors * n_sharp + k];
  for (int k = threadIdx.x; k < cnt[1]; k += blockDim.x) out_less[off[1] + k] = slot_less[(size_t)r * n_sectors * n_less + k];
  for (int k = threadIdx.x; k < cnt[2]; k += blockDim.x) out_flat[off[2] + k] = slot_flat[(size_t)r * n_sectors * n_flat + k];
  for (int k = threadIdx.x; k < cnt[3]; k += blockDim.x) out_lf[off[3] + k] = lessflat_tmp[m->ring_start[r] + k];
}

// a-4 TransformPointCloudInPlace (:367-371; rigid_transform.h:140-145)
struct Pose7 { double v[7]; };
__global__ void k_feat_extrinsic(const float4 *__restrict__ in, const uint32_t *__restrict__ soff, const FeatMeta *__restrict__ metas,
                                 Pose7 T, float4 *__restrict__ out) {
  const ScanView sv = scan_view(soff, blockIdx.y);
  const FeatMeta *m = metas + sv.b;
  in += sv.base; out += sv.base;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m->n_valid) return;
  const float4 p = in[i];
  const float3 x = transform_point_f(T.v, p.x, p.y, p.z);
  out[i] = make_float4(x.x, x.y, x.z, p.w);
}

// scratch for a batch of N points in B scans; the packed input (float4 xyz+i, then uint16 rings) goes to f_raw
static int feat_reserve(msfl_engine *e, size_t N, int B) {
  const msfl_params &P = e->params;
  const size_t slots = (size_t)B * MSFL_MAX_RINGS * P.n_sectors * (P.n_sharp + P.n_less_sharp + P.n_flat);
  int rc;
  if ((rc = e->f_raw.reserve(N * 16 + N * 2 + 64))) return rc;
  if ((rc = e->f_keys.reserve(N * 4))) return rc;
  if ((rc = e->f_keys_alt.reserve(N * 4))) return rc;
  if ((rc = e->f_vals.reserve(N * 4))) return rc;
  if ((rc = e->f_vals_alt.reserve(N * 4))) return rc;
  if ((rc = e->f_angle.reserve(N * 8))) return rc;
  if ((rc = e->f_full.reserve(2 * N * 16))) return rc;  // pre- and post-extrinsic
  if ((rc = e->f_ring.reserve(N * 2 + 64))) return rc;
  if ((rc = e->f_curv.reserve(N * 4))) return rc;
  if ((rc = e->f_label.reserve(N * 4 + N))) return rc;  // labels + picked flags
  if ((rc = e->f_idx.reserve((5 * N + slots) * 4))) return rc;
  if ((rc = e->f_cnt.reserve((size_t)B * sizeof(FeatMeta)))) return rc;
  if ((rc = e->f_soff.reserve((size_t)(B + 1) * 4))) return rc;
  return MSFL_OK;
}

// Enqueues the whole registration block for B scans whose packed points / rings are in f_raw and whose offsets are in
// f_soff (device) / h_off (host).  No synchronisation; results stay on the device (FeatDev).
template <class SHAPE>
static int launch_feat_pick(msfl_engine *e, int which, dim3 gr, const float4 *full_pre, const float *curv, int32_t *label, uint8_t *picked,
                            const uint32_t *soff, FeatMeta *meta, int32_t *slot_sharp, int32_t *slot_less, int32_t *slot_flat,
                            int32_t *lf_tmp) {
  const msfl_params &P = e->params;
  if (!(e->pick_attr_mask & (1 << which))) {  // per engine: attributes are per device
    MSFL_CUDA_OK(cudaFuncSetAttribute(k_feat_pick<SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SHAPE::kSmem));
    e->pick_attr_mask |= 1 << which;
  }
  k_feat_pick<SHAPE><<<gr, SHAPE::kThreads, SHAPE::kSmem, e->stream>>>(full_pre, curv, label, picked, soff, meta, P.curvature_thresh,
                                                                       P.neighbor_gap_sq, P.n_sectors, P.n_sharp, P.n_less_sharp, P.n_flat,
                                                                       slot_sharp, slot_less, slot_flat, lf_tmp);
  return MSFL_OK;
}

// shape: 0 small, 1 mid, 2 big (see PickShape)
static int extract_enqueue(msfl_engine *e, int B, const uint32_t *h_off, const double T[7], FeatDevice *fd, int shape) {
  cudaStream_t st = e->stream;
  const msfl_params &P = e->params;
  const size_t n = h_off[B];
  const uint32_t N = (uint32_t)n;
  uint32_t max_n = 0;
  for (int b = 0; b < B; ++b) max_n = std::max(max_n, h_off[b + 1] - h_off[b]);
  const int S = P.n_sectors;
  int rc;
  const float4 *d_raw = e->f_raw.as<float4>();
  const uint16_t *d_ring_in = (const uint16_t *)(e->f_raw.as<char>() + n * 16);
  const uint32_t *soff = e->f_soff.as<uint32_t>();
  FeatMeta *meta = e->f_cnt.as<FeatMeta>();
  uint32_t *keys = e->f_keys.as<uint32_t>(), *vals = e->f_vals.as<uint32_t>();
  const int tb = 256;
  const dim3 gp((max_n + tb - 1) / tb, (unsigned)B);  // per-point kernels: grid.y = scan
  k_feat_init<<<B, 128, 0, st>>>(meta);
  k_feat_keys<<<gp, tb, 0, st>>>(d_raw, d_ring_in, soff, P.min_range, keys, vals, meta);
  int bits = 8;
  while ((1 << (bits - 8)) < B) ++bits;  // key = scan << 8 | ring (255 = invalid point)
  cub::DoubleBuffer<uint32_t> dk(keys, e->f_keys_alt.as<uint32_t>()), dv(vals, e->f_vals_alt.as<uint32_t>());
  size_t tmp = 0;
  MSFL_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int)N, 0, bits, st));
  if ((rc = e->f_tmp.reserve(tmp))) return rc;
  MSFL_CUDA_OK(cub::DeviceRadixSort::SortPairs(e->f_tmp.p, tmp, dk, dv, (int)N, 0, bits, st));
  const uint32_t *ks = dk.Current(), *vs = dv.Current();
  k_feat_ring_start<<<B, 160, 0, st>>>(ks, soff, meta);
  double *rel = e->f_angle.as<double>();
  float4 *full_pre = e->f_full.as<float4>(), *full_post = full_pre + n;
  uint16_t *d_ring = e->f_ring.as<uint16_t>();
  float *curv = e->f_curv.as<float>();
  int32_t *label = e->f_label.as<int32_t>();
  uint8_t *picked = (uint8_t *)(label + n);
  int32_t *o_sharp = e->f_idx.as<int32_t>(), *o_less = o_sharp + n, *o_flat = o_less + n, *o_lf = o_flat + n,
          *lf_tmp = o_lf + n, *slot_sharp = lf_tmp + n, *slot_less = slot_sharp + (size_t)B * MSFL_MAX_RINGS * S * P.n_sharp,
          *slot_flat = slot_less + (size_t)B * MSFL_MAX_RINGS * S * P.n_less_sharp;
  k_feat_angles<<<gp, tb, 0, st>>>(d_raw, vs, soff, meta, rel);
  k_feat_first_dec<<<gp, tb, 0, st>>>(ks, rel, soff, meta);
  k_feat_full<<<gp, tb, 0, st>>>(d_raw, ks, vs, rel, soff, meta, P.scan_period, full_pre, d_ring);
  k_feat_curv<<<gp, tb, 0, st>>>(full_pre, soff, meta, curv, label, picked);
  const dim3 gr(MSFL_MAX_RINGS, (unsigned)B);  // per-ring kernels: grid = (ring, scan)
  if (shape == 0) rc = launch_feat_pick<PickSmall>(e, 0, gr, full_pre, curv, label, picked, soff, meta, slot_sharp, slot_less, slot_flat, lf_tmp);
  else if (shape == 1) rc = launch_feat_pick<PickMid>(e, 1, gr, full_pre, curv, label, picked, soff, meta, slot_sharp, slot_less, slot_flat, lf_tmp);
  else rc = launch_feat_pick<PickBig>(e, 2, gr, full_pre, curv, label, picked, soff, meta, slot_sharp, slot_less, slot_flat, lf_tmp);
  if (rc) return rc;
  k_feat_compact<<<gr, 128, 0, st>>>(soff, meta, S, P.n_sharp, P.n_less_sharp, P.n_flat, slot_sharp, slot_less, slot_flat, lf_tmp, o_sharp,
                                     o_less, o_flat, o_lf);
  Pose7 T7;
  for (int i = 0; i < 7; ++i) T7.v[i] = T ? T[i] : (i == 6 ? 1.0 : 0.0);
  k_feat_extrinsic<<<gp, tb, 0, st>>>(full_pre, soff, meta, T7, full_post);
  e->launches += 11 + 3;
  MSFL_CUDA_OK(cudaGetLastError());
  fd->full_post = full_post; fd->ring = d_ring; fd->curv = curv; fd->label = label;
  fd->o_sharp = o_sharp; fd->o_less = o_less; fd->o_flat = o_flat; fd->o_lf = o_lf;
  fd->metas = meta; fd->soff = soff;
  return MSFL_OK;
}

static int check_feat_meta(const FeatMeta &hm, int b) {
  if (hm.bad_ring) { set_error("extract_features: ring >= %d (kMaxScanNum) in scan %d", MSFL_MAX_RINGS, b); return MSFL_ERR_RING; }
  if (hm.n_valid <= 0) { set_error("extract_features: no valid points in scan %d", b); return MSFL_ERR_EMPTY; }
  if (hm.sector_overflow) { set_error("extract_features: a ring sector of scan %d holds more than %d points", b, kMaxSectorPts); return MSFL_ERR_ARG; }
  return MSFL_OK;
}

// Enqueue + metas to the host + synchronise.  The pick shape follows the rings of the engine's previous extraction; a
// shape that turns out too small for a sector (sector_overflow) is re-run with the largest one -- f_raw is untouched by
// the kernels, so the block simply runs again.
static int extract_run(msfl_engine *e, int B, const uint32_t *h_off, const double T[7], FeatDevice *fd, std::vector<FeatMeta> &hm) {
  const int S = e->params.n_sectors;
  int shape = 2;
  if (e->pick_seen_ring >= 0) {
    int P = 1;
    while (P < e->pick_seen_sector) P <<= 1;
    if (e->pick_seen_ring <= PickSmall::kRingCap && (long long)P * S <= PickSmall::kKeysCap) shape = 0;
    else if (e->pick_seen_ring <= PickMid::kRingCap && (long long)P * S <= PickMid::kKeysCap) shape = 1;
  }
  if (const char *v = getenv("MSFL_PICK_SHAPE")) shape = std::min(2, std::max(0, atoi(v)));  // tests: force a shape
  hm.resize(B);
  int rc;
  for (;;) {
    if ((rc = extract_enqueue(e, B, h_off, T, fd, shape))) return rc;
    MSFL_CUDA_OK(cudaMemcpyAsync(hm.data(), fd->metas, (size_t)B * sizeof(FeatMeta), cudaMemcpyDeviceToHost, e->stream));
    MSFL_CUDA_OK(cudaStreamSynchronize(e->stream));
    bool overflow = false;
    for (int b = 0; b < B; ++b) overflow = overflow || hm[b].sector_overflow;
    if (!overflow || shape == 2) break;
    shape = 2;
  }
  int ring_max = 0, sector_max = 0;
  for (int b = 0; b < B; ++b)
    for (int r = 0; r < MSFL_MAX_RINGS; ++r) {
      const int len = (int)(hm[b].ring_start[r + 1] - hm[b].ring_start[r]);
      ring_max = std::max(ring_max, len);
      sector_max = std::max(sector_max, (len - 11 + S - 1) / S + 1);  // sectors split [start + 5, end - 6]
    }
  e->pick_seen_ring = ring_max;
  e->pick_seen_sector = std::max(sector_max, 1);
  return MSFL_OK;
}

// copies one scan's results to the caller's arrays (async; the caller synchronises)
static int download_features(msfl_engine *e, const FeatDevice &fd, uint32_t base, const FeatMeta &hm, msfl_features *out) {
  cudaStream_t st = e->stream;
  const size_t nv = (size_t)hm.n_valid;
  out->n_full = hm.n_valid;
  out->n_sharp = hm.tot[0];
  out->n_less_sharp = hm.tot[1];
  out->n_flat = hm.tot[2];
  out->n_less_flat = hm.tot[3];
  if (out->full_xyzi) MSFL_CUDA_OK(cudaMemcpyAsync(out->full_xyzi, fd.full_post + base, nv * 16, cudaMemcpyDeviceToHost, st));
  if (out->full_ring) MSFL_CUDA_OK(cudaMemcpyAsync(out->full_ring, fd.ring + base, nv * 2, cudaMemcpyDeviceToHost, st));
  if (out->curvature) MSFL_CUDA_OK(cudaMemcpyAsync(out->curvature, fd.curv + base, nv * 4, cudaMemcpyDeviceToHost, st));
  if (out->label) MSFL_CUDA_OK(cudaMemcpyAsync(out->label, fd.label + base, nv * 4, cudaMemcpyDeviceToHost, st));
  if (out->idx_sharp) MSFL_CUDA_OK(cudaMemcpyAsync(out->idx_sharp, fd.o_sharp + base, (size_t)hm.tot[0] * 4, cudaMemcpyDeviceToHost, st));
  if (out->idx_less_sharp) MSFL_CUDA_OK(cudaMemcpyAsync(out->idx_less_sharp, fd.o_less + base, (size_t)hm.tot[1] * 4, cudaMemcpyDeviceToHost, st));
  if (out->idx_flat) MSFL_CUDA_OK(cudaMemcpyAsync(out->idx_flat, fd.o_flat + base, (size_t)hm.tot[2] * 4, cudaMemcpyDeviceToHost, st));
  if (out->idx_less_flat) MSFL_CUDA_OK(cudaMemcpyAsync(out->idx_less_flat, fd.o_lf + base, (size_t)hm.tot[3] * 4, cudaMemcpyDeviceToHost, st));
  return MSFL_OK;
}

int run_extract_features(msfl_engine *e, const msfl_cloud *raw, const double T[7], msfl_features *out) {
  cudaStream_t st = e->stream;
  const size_t n = raw->n;
  const uint32_t N = (uint32_t)n;
  int rc;
  // raw AoS bytes -> device, unpacked there (no per-point host loop)
  const size_t raw_bytes = n * raw->stride;
  if ((rc = feat_reserve(e, n, 1))) return rc;
  if ((rc = e->f_misc.reserve(raw_bytes + 64))) return rc;
  const uint32_t h_off[2] = {0u, N};
  MSFL_CUDA_OK(cudaMemcpyAsync(e->f_soff.p, h_off, sizeof h_off, cudaMemcpyHostToDevice, st));
  MSFL_CUDA_OK(cudaMemcpyAsync(e->f_misc.p, raw->data, raw_bytes, cudaMemcpyHostToDevice, st));
  k_unpack_aos<<<(N + 255) / 256, 256, 0, st>>>(e->f_misc.as<unsigned char>(), N, (uint32_t)raw->stride, (uint32_t)raw->off_xyz,
                                               raw->off_intensity == MSFL_NO_FIELD ? 0u : (uint32_t)raw->off_intensity,
                                               (uint32_t)raw->off_ring, raw->off_intensity != MSFL_NO_FIELD,
                                               e->f_raw.as<float4>(), (uint16_t *)(e->f_raw.as<char>() + n * 16));
  FeatDevice fd;
  std::vector<FeatMeta> hm;
  if ((rc = extract_run(e, 1, h_off, T, &fd, hm))) return rc;
  if ((rc = check_feat_meta(hm[0], 0))) return rc;
  if ((rc = download_features(e, fd, 0, hm[0], out))) return rc;
  MSFL_CUDA_OK(cudaStreamSynchronize(st));
  return MSFL_OK;
}

// B raw clouds -> packed float4 + rings in pinned staging (host threads), one upload, one launch sequence for the batch
static int upload_raw_batch(msfl_engine *e, int B, const msfl_cloud *raw, std::vector<uint32_t> &h_off) {
  h_off.assign(B + 1, 0);
  size_t N = 0;
  for (int b = 0; b < B; ++b) {
    h_off[b] = (uint32_t)N;
    N += raw[b].n;
  }
  if (N > 0x3fffffffull) { set_error("extract_features_batch: batch too large"); return MSFL_ERR_ARG; }
  h_off[B] = (uint32_t)N;
  int rc;
  if ((rc = feat_reserve(e, N, B))) return rc;
  MSFL_CUDA_OK(cudaStreamSynchronize(e->stream));  // the staging buffer may still feed an earlier upload
  if ((rc = e->h_stage.reserve(N * 18 + (size_t)(B + 1) * 4 + 64))) return rc;
  float *h4 = e->h_stage.as<float>();
  uint16_t *hr = (uint16_t *)(e->h_stage.as<char>() + N * 16);
  uint32_t *ho = (uint32_t *)(e->h_stage.as<char>() + ((N * 18 + 15) & ~(size_t)15));
  memcpy(ho, h_off.data(), (size_t)(B + 1) * 4);
  MSFL_CUDA_OK(cudaMemcpyAsync(e->f_soff.p, ho, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, e->stream));
  // repack and upload in a few pieces: the DMA of piece k runs while the host threads repack piece k + 1
  const int n_pieces = N >= 2000000 ? 4 : 1;
  char *d_raw = e->f_raw.as<char>();
  int b0 = 0;
  for (int k = 0; k < n_pieces && b0 < B; ++k) {
    int b1 = b0;
    const size_t goal = k == n_pieces - 1 ? N : N * (size_t)(k + 1) / n_pieces;
    while (b1 < B && (h_off[b1] < goal || k == n_pieces - 1)) ++b1;
    if (b1 == b0) continue;
    pack_clouds_parallel(e, b1 - b0, raw + b0, h4, hr, h_off.data() + b0);
    const size_t p0 = h_off[b0], np = h_off[b1] - p0;
    MSFL_CUDA_OK(cudaMemcpyAsync(d_raw + p0 * 16, h4 + 4 * p0, np * 16, cudaMemcpyHostToDevice, e->stream));
    MSFL_CUDA_OK(cudaMemcpyAsync(d_raw + N * 16 + p0 * 2, hr + p0, np * 2, cudaMemcpyHostToDevice, e->stream));
    b0 = b1;
  }
  return MSFL_OK;
}

// upload + extraction of a batch, results left on the device; h_counts receives B x 5 ints (n_valid, sharp, less_sharp,
// flat, less_flat).  One synchronisation (the counts size everything downstream).
int extract_batch_to_device(msfl_engine *e, int B, const msfl_cloud *raw, const double T[7], FeatDevice *fd,
                            std::vector<uint32_t> &h_off, std::vector<int32_t> &h_counts) {
  int rc;
  if ((rc = upload_raw_batch(e, B, raw, h_off))) return rc;
  std::vector<FeatMeta> hm;
  if ((rc = extract_run(e, B, h_off.data(), T, fd, hm))) return rc;
  h_counts.resize((size_t)B * 5);
  for (int b = 0; b < B; ++b) {
    if ((rc = check_feat_meta(hm[b], b))) return rc;
    h_counts[5 * b] = hm[b].n_valid;
    for (int k = 0; k < 4; ++k) h_counts[5 * b + 1 + k] = hm[b].tot[k];
  }
  return MSFL_OK;
}

int run_extract_features_batch(msfl_engine *e, int B, const msfl_cloud *raw, const double T[7], msfl_features *outs) {
  std::vector<uint32_t> h_off;
  int rc;
  if ((rc = upload_raw_batch(e, B, raw, h_off))) return rc;
  FeatDevice fd;
  std::vector<FeatMeta> hm;
  if ((rc = extract_run(e, B, h_off.data(), T, &fd, hm))) return rc;
  for (int b = 0; b < B; ++b) {
    if ((rc = check_feat_meta(hm[b], b))) return rc;
    if ((rc = download_features(e, fd, h_off[b], hm[b], &outs[b]))) return rc;
  }
  MSFL_CUDA_OK(cudaStreamSynchronize(e->stream));
  return MSFL_OK;
}

}  // namespace msfl

using namespace msfl;

extern "C" int msfl_extract_features(msfl_engine *e, const msfl_cloud *raw, const double T_lidar2imu[7], msfl_features *out) {
  if (!e || !raw || !out) { set_error("msfl_extract_features: bad argument"); return MSFL_ERR_ARG; }
  out->n_full = out->n_sharp = out->n_less_sharp = out->n_flat = out->n_less_flat = 0;
  if (raw->n == 0 || !raw->data) { set_error("extract_features: empty cloud"); return MSFL_ERR_EMPTY; }
  if (raw->off_ring == MSFL_NO_FIELD || raw->stride < 14 || raw->n > 0x3fffffffull || raw->off_xyz + 12 > raw->stride ||
      raw->off_ring + 2 > raw->stride || (raw->off_intensity != MSFL_NO_FIELD && raw->off_intensity + 4 > raw->stride) ||
      raw->n * raw->stride > 0x7fffffffull) {
    set_error("extract_features: cloud needs xyz + ring inside the point stride");
    return MSFL_ERR_ARG;
  }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  return run_extract_features(e, raw, T_lidar2imu, out);
}

extern "C" int msfl_extract_features_batch(msfl_engine *e, int B, const msfl_cloud *raw, const double T_lidar2imu[7],
                                           msfl_features *outs) {
  if (!e || !raw || !outs || B <= 0) { set_error("msfl_extract_features_batch: bad argument"); return MSFL_ERR_ARG; }
  for (int b = 0; b < B; ++b) {
    outs[b].n_full = outs[b].n_sharp = outs[b].n_less_sharp = outs[b].n_flat = outs[b].n_less_flat = 0;
    int rc;
    if ((rc = check_cloud(&raw[b], true, "extract_features_batch"))) return rc;
    if (raw[b].n == 0) { set_error("extract_features_batch: scan %d is empty", b); return MSFL_ERR_EMPTY; }
  }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  return run_extract_features_batch(e, B, raw, T_lidar2imu, outs);
}
