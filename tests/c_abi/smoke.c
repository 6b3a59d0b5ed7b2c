/* Plain-C host program over the C ABI (include/msfl.h): proves the boundary is usable without C++ or
 * Python.  Built and run by tests/test_c_abi_program.py on the GPU box:
 *     gcc -std=c11 -Iinclude tests/c_abi/smoke.c -Lmsf_loam_b200 -lmsfl -lm -o smoke
 * Scene: a 20 x 20 m floor (z = 0), two walls and a pole line; the scan is the submap moved by a known
 * rigid transform, so MatchScan2Map must recover that transform. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "msfl.h"

typedef struct { float x, y, z, pad, intensity, pad2[3]; } PointXYZI; /* pcl::PointXYZI layout, 32 bytes */

static msfl_cloud view(const PointXYZI *p, size_t n) {
  msfl_cloud c = {p, n, sizeof(PointXYZI), offsetof(PointXYZI, x), offsetof(PointXYZI, intensity), MSFL_NO_FIELD};
  return c;
}

int main(void) {
  size_t ns = 0, nc = 0, cap = 40000;
  PointXYZI *surf = calloc(cap, sizeof *surf), *corner = calloc(cap, sizeof *corner);
  for (int i = 0; i < 50; ++i)
    for (int j = 0; j < 50; ++j) { /* floor + two walls on a 0.4 m lattice */
      surf[ns].x = -10 + 0.4f * i; surf[ns].y = -10 + 0.4f * j; surf[ns].z = 0.01f * ((i * 7 + j * 3) % 5); ++ns;
      surf[ns].x = -10 + 0.4f * i; surf[ns].y = 10 + 0.01f * ((i + j) % 3); surf[ns].z = 0.12f * j; ++ns;
      surf[ns].x = 10 + 0.01f * ((i + 2 * j) % 3); surf[ns].y = -10 + 0.4f * i; surf[ns].z = 0.12f * j; ++ns;
    }
  for (int k = 0; k < 8; ++k)
    for (int j = 0; j < 30; ++j) { /* vertical poles = line features */
      corner[nc].x = -8 + 2.3f * k; corner[nc].y = -6 + 1.7f * (k % 4); corner[nc].z = 0.2f * j; ++nc;
    }
  /* scan = map seen from a sensor at T_true: p_scan = R^T (p_map - t) */
  const double yaw = 0.02, t[3] = {0.15, -0.08, 0.03};
  const double cy = cos(yaw), sy = sin(yaw);
  PointXYZI *ss = calloc(ns, sizeof *ss), *sc = calloc(nc, sizeof *sc);
  for (size_t i = 0; i < ns; ++i) {
    const double dx = surf[i].x - t[0], dy = surf[i].y - t[1], dz = surf[i].z - t[2];
    ss[i].x = (float)(cy * dx + sy * dy); ss[i].y = (float)(-sy * dx + cy * dy); ss[i].z = (float)dz;
  }
  for (size_t i = 0; i < nc; ++i) {
    const double dx = corner[i].x - t[0], dy = corner[i].y - t[1], dz = corner[i].z - t[2];
    sc[i].x = (float)(cy * dx + sy * dy); sc[i].y = (float)(-sy * dx + cy * dy); sc[i].z = (float)dz;
  }
  msfl_params P;
  msfl_default_params(&P);
  msfl_engine *e = NULL;
  if (msfl_create(&P, 0, &e) != MSFL_OK) { fprintf(stderr, "create: %s\n", msfl_last_error()); return 2; }
  msfl_cloud mc = view(corner, nc), ms = view(surf, ns), qc = view(sc, nc), qs = view(ss, ns);
  if (msfl_set_submap(e, &mc, &ms) != MSFL_OK) { fprintf(stderr, "set_submap: %s\n", msfl_last_error()); return 3; }
  double pose[7] = {0, 0, 0, 0, 0, 0, 1}; /* identity guess */
  msfl_stats st;
  const int rc = msfl_scan2map(e, &qc, &qs, pose, &st);
  if (rc != MSFL_OK) { fprintf(stderr, "scan2map: %d %s\n", rc, msfl_last_error()); return 4; }
  const double est_yaw = 2.0 * atan2(pose[5], pose[6]);
  const double et = sqrt((pose[0] - t[0]) * (pose[0] - t[0]) + (pose[1] - t[1]) * (pose[1] - t[1]) + (pose[2] - t[2]) * (pose[2] - t[2]));
  printf("pose t = %.5f %.5f %.5f yaw = %.6f | err %.2e m %.2e rad | corr %d edge %d plane | launches %llu\n", pose[0], pose[1],
         pose[2], est_yaw, et, fabs(est_yaw - yaw), st.n_edge[1], st.n_plane[1], (unsigned long long)msfl_launch_count(e));
  /* the asynchronous batch form: two tickets in flight, each batch = the same scan three times from the identity
   * guess; every pose must equal the synchronous result bit for bit */
  int async_ok = 1;
  {
    msfl_cloud bc[3] = {qc, qc, qc}, bs[3] = {qs, qs, qs};
    double init[21], out0[21], out1[21], out2[21];
    for (int b = 0; b < 3; ++b)
      for (int k = 0; k < 7; ++k) init[7 * b + k] = k == 6 ? 1.0 : 0.0;
    int t0 = -1, t1 = -1, t2 = -1, t3 = -1;
    if (msfl_scan2map_batch_submit(e, 3, bc, bs, init, 0, &t0) != MSFL_OK || msfl_scan2map_batch_submit(e, 3, bc, bs, init, 0, &t1) != MSFL_OK ||
        msfl_scan2map_batch_submit(e, 3, bc, bs, init, 0, &t2) != MSFL_OK) {
      fprintf(stderr, "submit: %s\n", msfl_last_error());
      return 5;
    }
    if (msfl_scan2map_batch_submit(e, 3, bc, bs, init, 0, &t3) == MSFL_OK) async_ok = 0; /* MSFL_MAX_INFLIGHT = 3: a fourth ticket must be refused */
    if (msfl_scan2map_batch_wait(e, t0, out0, NULL) != MSFL_OK || msfl_scan2map_batch_wait(e, t1, out1, NULL) != MSFL_OK ||
        msfl_scan2map_batch_wait(e, t2, out2, NULL) != MSFL_OK) {
      fprintf(stderr, "wait: %s\n", msfl_last_error());
      return 6;
    }
    for (int b = 0; b < 3; ++b)
      if (memcmp(out0 + 7 * b, pose, sizeof pose) != 0 || memcmp(out1 + 7 * b, pose, sizeof pose) != 0 ||
          memcmp(out2 + 7 * b, pose, sizeof pose) != 0)
        async_ok = 0;
    printf("async batches %s\n", async_ok ? "match the synchronous pose bit for bit" : "MISMATCH");
  }
  msfl_destroy(e);
  return (async_ok && et < 5e-3 && fabs(est_yaw - yaw) < 5e-4 && st.n_plane[1] > 1000) ? 0 : 1;
}
