// chain.cu -- raw clouds -> poses for a batch of independent scans (replay of a log / re-localisation against a frozen
// map): scan registration (msf_loam_node.cc:160-371), the caller-side VoxelGrid of the less-sharp / less-flat features
// (laser_mapping.cc:264-270) and MappingScanMatcher::MatchScan2Map (mapping_scan_matcher.cc:63-278) back to back on the
// GPU.  Every stage runs ONCE for the whole batch (grid.y = scan); the feature clouds and the down-sampled queries never
// leave HBM -- the host sees the raw clouds going in and B poses coming out, plus two small count read-backs that size
// the next stage's launches.
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

}  // namespace msfl

using namespace msfl;

extern "C" int msfl_register_and_match_batch(msfl_engine *e, int B, const msfl_cloud *raw, const double T_lidar2imu[7],
                                             float leaf_corner, float leaf_surf, double *poses_tq, msfl_chain_counts *counts,
                                             msfl_stats *stats) {
  if (!e || !raw || !poses_tq || B <= 0) { set_error("msfl_register_and_match_batch: bad argument"); return MSFL_ERR_ARG; }
  if (!e->has_submap) { set_error("msfl_register_and_match_batch: no submap set"); return MSFL_ERR_NOSUBMAP; }
  int rc;
  for (int b = 0; b < B; ++b) {
    if ((rc = check_cloud(&raw[b], true, "register_and_match_batch"))) return rc;
    if (raw[b].n == 0) { set_error("register_and_match_batch: scan %d is empty", b); return MSFL_ERR_EMPTY; }
  }
  MSFL_CUDA_OK(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  // 1. registration of the whole batch; the per-scan feature counts come back (sync 1) and size the gather
  FeatDevice fd;
  std::vector<uint32_t> h_off;
  std::vector<int32_t> hc;
  if ((rc = extract_batch_to_device(e, B, raw, T_lidar2imu, &fd, h_off, hc))) return rc;
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
