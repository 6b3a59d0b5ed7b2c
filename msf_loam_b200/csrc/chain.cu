// chain.cu -- raw clouds -> poses for a batch of independent scans (replay of a log / re-localisation against a frozen
// map): scan registration (msf_loam_node.cc:160-371), the caller-side VoxelGrid of the less-sharp / less-flat features
// (laser_mapping.cc:264-270) and MappingScanMatcher::MatchScan2Map (mapping_scan_matcher.cc:63-278) back to back on the
// GPU.  Every stage runs ONCE for the whole batch (grid.y = scan); the feature clouds and the down-sampled queries never
// leave HBM -- the host sees the raw clouds going in and B poses coming out, plus two small count read-backs that size
// the next stage's launches.
#include <math.h>
#include <string.h>

#include <vector>

#include "msfl_internal.h"

namespace msfl {

// corner_in / surf_in: the less-sharp / less-flat clouds of every scan, back to back (scan b at c_off[b] / s_off[b])
__global__ void k_chain_gather(const float4 *__restrict__ full, const uint32_t *__restrict__ soff, const int32_t *__restrict__ o_less,
                               const int32_t *__restrict__ o_lf, const uint32_t *__restrict__ c_off, const uint32_t *__restrict__ s_off,
                               float4 *__restrict__ corner_in, float4 *__restrict__ surf_in) {
  const uint32_t b = blockIdx.y, base = soff[b];
  const bool surf = blockIdx.z == 1;
  const uint32_t *off = surf ? s_off : c_off;
  const uint32_t n = off[b + 1] - off[b];
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int32_t i = (surf ? o_lf : o_less)[base + k];
  (surf ? surf_in : corner_in)[off[b] + k] = full[base + i];
}

// scan-to-scan inputs from the device-resident features: item j = blockIdx.y reads index list `idx` of scan
// (j >> shift_log2) + scan_add (two lists alternate when idx_b != nullptr: even items idx, odd items idx_b) and writes
// the points (and rings) at out_off[j]
__global__ void k_odo_gather(const float4 *__restrict__ full, const uint16_t *__restrict__ ring, const uint32_t *__restrict__ soff,
                             const int32_t *__restrict__ idx, const int32_t *__restrict__ idx_b, int scan_add,
                             const uint32_t *__restrict__ out_off, float4 *__restrict__ out, uint16_t *__restrict__ out_ring) {
  const uint32_t j = blockIdx.y;
  const uint32_t scan = (idx_b ? (j >> 1) : j) + (uint32_t)scan_add, base = soff[scan];
  const uint32_t n = out_off[j + 1] - out_off[j];
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int32_t i = ((idx_b && (j & 1u)) ? idx_b : idx)[base + k];
  out[out_off[j] + k] = full[base + i];
  if (out_ring) out_ring[out_off[j] + k] = ring[base + i];
}

// OdometryScanMatcher::MatchScan2Scan for the B - 1 consecutive pairs of the batch (pair p: last = scan p, curr = scan
// p + 1; laser_odometry.cc:75), straight from the device-resident features.  odom_tq: 7 B doubles, entry b >= 1 in-out.
static int chain_odometry(msfl_engine *e, int B, const FeatDevice &fd, const std::vector<int32_t> &hc, double *odom_tq,
                          int32_t *odom_status) {
  const int NP = B - 1;
  if (odom_status) odom_status[0] = MSFL_OK;
  if (NP <= 0) return MSFL_OK;
  cudaStream_t st = e->stream;
  int rc;
  std::vector<uint32_t> tab((size_t)(2 * NP + 1) + 2 * (size_t)(NP + 1));
  uint32_t *goff = tab.data();
  int32_t *e_off = (int32_t *)(goff + 2 * NP + 1), *p_off = e_off + NP + 1;
  uint32_t n_last = 0, ns = 0, nf = 0, max_last = 0, max_s = 0, max_f = 0;
  for (int p = 0; p < NP; ++p) {
    const uint32_t nl[2] = {(uint32_t)hc[5 * p + 2], (uint32_t)hc[5 * p + 4]};
    for (int c = 0; c < 2; ++c) { goff[2 * p + c] = n_last; n_last += nl[c]; max_last = std::max(max_last, nl[c]); }
    e_off[p] = (int32_t)ns; p_off[p] = (int32_t)nf;
    // a pair with nothing to search in keeps no queries: MSFL_TOO_FEW, pose untouched (as msfl_scan2scan)
    const bool dead = nl[0] == 0 || nl[1] == 0;
    const uint32_t qs = dead ? 0u : (uint32_t)hc[5 * (p + 1) + 1], qf = dead ? 0u : (uint32_t)hc[5 * (p + 1) + 3];
    ns += qs; nf += qf;
    max_s = std::max(max_s, qs); max_f = std::max(max_f, qf);
  }
  goff[2 * NP] = n_last; e_off[NP] = (int32_t)ns; p_off[NP] = (int32_t)nf;
  // scratch: [last float4 | sharp | flat | rings u16 | tables | poses | status]
  auto al16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  const size_t at_ring = ((size_t)n_last + ns + nf) * 16, at_tab = al16(at_ring + (size_t)n_last * 2);
  const size_t at_pose = at_tab + al16(tab.size() * 4), at_status = at_pose + (size_t)NP * 56, total = at_status + (size_t)NP * 4 + 16;
  if ((rc = e->ob_in.reserve(total))) return rc;
  char *d = e->ob_in.as<char>();
  float4 *d_last = (float4 *)d, *d_sharp = d_last + n_last, *d_flat = d_sharp + ns;
  uint16_t *d_ring = (uint16_t *)(d + at_ring);
  uint32_t *d_goff = (uint32_t *)(d + at_tab);
  int32_t *d_e_off = (int32_t *)(d_goff + 2 * NP + 1), *d_p_off = d_e_off + NP + 1;
  double *d_poses = (double *)(d + at_pose);
  int32_t *d_status = (int32_t *)(d + at_status);
  MSFL_CUDA_OK(cudaMemcpyAsync(d_goff, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, st));
  MSFL_CUDA_OK(cudaMemcpyAsync(d_poses, odom_tq + 7, (size_t)NP * 56, cudaMemcpyHostToDevice, st));
  MSFL_CUDA_OK(cudaMemsetAsync(d_status, 0, (size_t)NP * 4, st));
  if (max_last > 0)
    k_odo_gather<<<dim3((max_last + 255) / 256, 2u * NP), 256, 0, st>>>(fd.full_post, fd.ring, fd.soff, fd.o_less, fd.o_lf, 0, d_goff,
                                                                       d_last, d_ring);
  if (max_s > 0)
    k_odo_gather<<<dim3((max_s + 255) / 256, (unsigned)NP), 256, 0, st>>>(fd.full_post, fd.ring, fd.soff, fd.o_sharp, nullptr, 1,
                                                                         (const uint32_t *)d_e_off, d_sharp, nullptr);
  if (max_f > 0)
    k_odo_gather<<<dim3((max_f + 255) / 256, (unsigned)NP), 256, 0, st>>>(fd.full_post, fd.ring, fd.soff, fd.o_flat, nullptr, 1,
                                                                         (const uint32_t *)d_p_off, d_flat, nullptr);
  e->launches += 3;
  MSFL_CUDA_OK(cudaGetLastError());
  if ((rc = scan2scan_batch_device(e, NP, d_last, d_ring, d_goff, goff, d_sharp, d_e_off, e_off, d_flat, d_p_off, p_off, d_poses,
                                   d_status, nullptr)))
    return rc;
  if ((rc = e->h_poses.reserve((size_t)NP * 60))) return rc;
  char *ho = e->h_poses.as<char>();
  MSFL_CUDA_OK(cudaMemcpyAsync(ho, d_poses, (size_t)NP * 56, cudaMemcpyDeviceToHost, st));
  MSFL_CUDA_OK(cudaMemcpyAsync(ho + (size_t)NP * 56, d_status, (size_t)NP * 4, cudaMemcpyDeviceToHost, st));
  MSFL_CUDA_OK(cudaStreamSynchronize(st));
  memcpy(odom_tq + 7, ho, (size_t)NP * 56);
  if (odom_status) memcpy(odom_status + 1, ho + (size_t)NP * 56, (size_t)NP * 4);
  return MSFL_OK;
}

// Rigid3d product a * b (rigid_transform.h:105-111): t = qa tb + ta, q = (qa qb).normalized(); pose = [t(3), q xyzw]
static void pose_compose(const double a[7], const double b[7], double out[7]) {
  const double ax = a[3], ay = a[4], az = a[5], aw = a[6];
  const double bx = b[3], by = b[4], bz = b[5], bw = b[6];
  // qa * tb  (Eigen: v + w * uv + qv x uv, uv = 2 qv x v)
  const double uvx = 2 * (ay * b[2] - az * b[1]), uvy = 2 * (az * b[0] - ax * b[2]), uvz = 2 * (ax * b[1] - ay * b[0]);
  out[0] = (b[0] + aw * uvx + (ay * uvz - az * uvy)) + a[0];
  out[1] = (b[1] + aw * uvy + (az * uvx - ax * uvz)) + a[1];
  out[2] = (b[2] + aw * uvz + (ax * uvy - ay * uvx)) + a[2];
  double q[4] = {aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                 aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz};
  const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) out[3 + i] = nq > 0 ? q[i] / nq : q[i];
}

}  // namespace msfl

using namespace msfl;

static int chain_impl(msfl_engine *e, int B, const msfl_cloud *raw, const double T_lidar2imu[7], float leaf_corner,
                      float leaf_surf, double *odom_tq, int32_t *odom_status, int compose, double *poses_tq,
                      msfl_chain_counts *counts, msfl_stats *stats, const char *what) {
  if (!e || !raw || !poses_tq || B <= 0) { set_error("%s: bad argument", what); return MSFL_ERR_ARG; }
  if (!e->has_submap) { set_error("%s: no submap set", what); return MSFL_ERR_NOSUBMAP; }
  int rc;
  for (int b = 0; b < B; ++b) {
    if ((rc = check_cloud(&raw[b], true, what))) return rc;
    if (raw[b].n == 0) { set_error("%s: scan %d is empty", what, b); return MSFL_ERR_EMPTY; }
  }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  // 1. registration of the whole batch; the per-scan feature counts come back (sync 1) and size the gather
  FeatDevice fd;
  std::vector<uint32_t> h_off;
  std::vector<int32_t> hc;
  if ((rc = extract_batch_to_device(e, B, raw, T_lidar2imu, &fd, h_off, hc))) return rc;
  // 1b. odometry of the consecutive pairs, and (compose) the dead-reckoned initial guesses of the map matcher:
  //     pose_scan2world = pose_scan2world * pose_curr2last (laser_odometry.cc:79), starting from poses_tq[0]
  if (odom_tq) {
    if ((rc = chain_odometry(e, B, fd, hc, odom_tq, odom_status))) return rc;
    if (compose)
      for (int b = 1; b < B; ++b) pose_compose(poses_tq + 7 * (b - 1), odom_tq + 7 * b, poses_tq + 7 * b);
  }
  std::vector<uint32_t> tab(2 * (size_t)(B + 1));
  uint32_t *c_off_h = tab.data(), *s_off_h = c_off_h + (B + 1);
  uint32_t nc_in = 0, ns_in = 0, max_c = 0, max_s = 0;
  for (int b = 0; b < B; ++b) {
    c_off_h[b] = nc_in; s_off_h[b] = ns_in;
    nc_in += (uint32_t)hc[5 * b + 2]; ns_in += (uint32_t)hc[5 * b + 4];
    max_c = std::max(max_c, (uint32_t)hc[5 * b + 2]); max_s = std::max(max_s, (uint32_t)hc[5 * b + 4]);
  }
  c_off_h[B] = nc_in; s_off_h[B] = ns_in;
  // scratch: gathered clouds | queries (capacity = input size) | tables: 2 x (B+1) uint32 in, 2 x (B+1) int32 out, poses
  const size_t n_in = (size_t)nc_in + ns_in;
  if ((rc = e->c_in.reserve(n_in * 16 + 64))) return rc;
  if ((rc = e->c_q.reserve(n_in * 16 + 64))) return rc;
  const size_t tab_bytes = (size_t)(B + 1) * 4;
  const size_t pose_at = (4 * tab_bytes + 15) & ~(size_t)15;
  if ((rc = e->c_off.reserve(pose_at + (size_t)B * 56))) return rc;
  char *d_tab = e->c_off.as<char>();
  uint32_t *d_c_in_off = (uint32_t *)d_tab, *d_s_in_off = (uint32_t *)(d_tab + tab_bytes);
  int32_t *d_c_off = (int32_t *)(d_tab + 2 * tab_bytes), *d_s_off = (int32_t *)(d_tab + 3 * tab_bytes);
  double *d_poses = (double *)(d_tab + pose_at);
  MSFL_CUDA_OK(cudaMemcpyAsync(d_tab, tab.data(), 2 * tab_bytes, cudaMemcpyHostToDevice, st));  // pageable: staged before return
  MSFL_CUDA_OK(cudaMemcpyAsync(d_poses, poses_tq, (size_t)B * 56, cudaMemcpyHostToDevice, st));
  float4 *corner_in = e->c_in.as<float4>(), *surf_in = corner_in + nc_in;
  float4 *d_qc = e->c_q.as<float4>(), *d_qs = d_qc + nc_in;
  const uint32_t max_in = std::max(max_c, max_s);
  if (max_in > 0) {
    k_chain_gather<<<dim3((max_in + 255) / 256, (unsigned)B, 2), 256, 0, st>>>(fd.full_post, fd.soff, fd.o_less, fd.o_lf, d_c_in_off,
                                                                            d_s_in_off, corner_in, surf_in);
    e->launches += 1;
  }
  // 2. caller-side VoxelGrid of both feature classes (laser_mapping.cc:264-270), whole batch at once
  if ((rc = run_voxel_grid_batch(e, B, corner_in, d_c_in_off, nc_in, max_c, leaf_corner, d_qc, d_c_off, 0))) return rc;
  if ((rc = run_voxel_grid_batch(e, B, surf_in, d_s_in_off, ns_in, max_s, leaf_surf, d_qs, d_s_off, 1))) return rc;
  // the query counts size the matcher's launches (sync 2)
  std::vector<int32_t> q_off(2 * (size_t)(B + 1));
  MSFL_CUDA_OK(cudaMemcpyAsync(q_off.data(), d_c_off, 2 * tab_bytes, cudaMemcpyDeviceToHost, st));
  MSFL_CUDA_OK(cudaStreamSynchronize(st));
  const uint32_t nct = (uint32_t)q_off[B], nst = (uint32_t)q_off[2 * B + 1];
  // 3. scan-to-map of the batch (the surf queries must follow the corner queries of the matcher's layout: they do, d_qs
  //    starts at the corner INPUT count, which is >= nct -- the matcher takes the two arrays separately)
  msfl_stats *d_stats = nullptr;
  if (stats) {
    if ((rc = e->d_stats.reserve((size_t)B * sizeof(msfl_stats)))) return rc;
    d_stats = e->d_stats.as<msfl_stats>();
    MSFL_CUDA_OK(cudaMemsetAsync(d_stats, 0, (size_t)B * sizeof(msfl_stats), st));
  }
  if ((rc = scan2map_enqueue(e, B, d_qc, d_c_off, nct, d_qs, d_s_off, nst, d_poses, d_stats))) return rc;
  if ((rc = e->h_poses.reserve((size_t)B * 56))) return rc;
  MSFL_CUDA_OK(cudaMemcpyAsync(e->h_poses.p, d_poses, (size_t)B * 56, cudaMemcpyDeviceToHost, st));
  if (stats) {
    if ((rc = e->h_stats.reserve((size_t)B * sizeof(msfl_stats)))) return rc;
    MSFL_CUDA_OK(cudaMemcpyAsync(e->h_stats.p, d_stats, (size_t)B * sizeof(msfl_stats), cudaMemcpyDeviceToHost, st));
  }
  MSFL_CUDA_OK(cudaStreamSynchronize(st));  // sync 3: the poses
  memcpy(poses_tq, e->h_poses.p, (size_t)B * 56);
  if (stats) memcpy(stats, e->h_stats.p, (size_t)B * sizeof(msfl_stats));
  if (counts)
    for (int b = 0; b < B; ++b) {
      counts[b].n_full = hc[5 * b]; counts[b].n_sharp = hc[5 * b + 1]; counts[b].n_less_sharp = hc[5 * b + 2];
      counts[b].n_flat = hc[5 * b + 3]; counts[b].n_less_flat = hc[5 * b + 4];
      counts[b].n_corner_queries = q_off[b + 1] - q_off[b];
      counts[b].n_surf_queries = q_off[B + 1 + b + 1] - q_off[B + 1 + b];
    }
  return MSFL_OK;
}

extern "C" int msfl_register_and_match_batch(msfl_engine *e, int B, const msfl_cloud *raw, const double T_lidar2imu[7],
                                             float leaf_corner, float leaf_surf, double *poses_tq, msfl_chain_counts *counts,
                                             msfl_stats *stats) {
  return chain_impl(e, B, raw, T_lidar2imu, leaf_corner, leaf_surf, nullptr, nullptr, 0, poses_tq, counts, stats,
                    "msfl_register_and_match_batch");
}

extern "C" int msfl_replay_batch(msfl_engine *e, int B, const msfl_cloud *raw, const double T_lidar2imu[7], float leaf_corner,
                                 float leaf_surf, double *odom_tq, int32_t *odom_status, int compose, double *poses_tq,
                                 msfl_chain_counts *counts, msfl_stats *stats) {
  if (!odom_tq) { set_error("msfl_replay_batch: odom_tq is null"); return MSFL_ERR_ARG; }
  return chain_impl(e, B, raw, T_lidar2imu, leaf_corner, leaf_surf, odom_tq, odom_status, compose, poses_tq, counts, stats,
                    "msfl_replay_batch");
}
