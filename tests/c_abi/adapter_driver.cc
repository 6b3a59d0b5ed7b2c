// adapter_driver.cc -- compiles msf_loam_b200/adapter/gpu_scan_matchers.h against tests/adapter_stubs/ and drives both
// adapters the way LaserOdometry / LaserMapping do (laser_odometry.cc:75, laser_mapping.cc:304-311): the clouds are
// read from a flat binary case file (tests/test_adapter.py writes it), filled into pcl::PointCloud stand-ins, and
// the poses that come back are printed for comparison with the Python binding's.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#include "gpu_scan_matchers.h"

static bool read_f32(FILE *f, std::vector<float> &v, size_t n) { v.resize(n); return n == 0 || fread(v.data(), 4, n, f) == n; }

template <typename P>
static void fill(pcl::PointCloud<P> &c, const std::vector<float> &xyzi, const std::vector<float> *ring) {
  c.resize(xyzi.size() / 4);
  for (size_t i = 0; i < c.size(); ++i) {
    c[i].x = xyzi[4 * i]; c[i].y = xyzi[4 * i + 1]; c[i].z = xyzi[4 * i + 2]; c[i].intensity = xyzi[4 * i + 3];
  }
  (void)ring;
}
static void fill_rings(pcl::PointCloud<PointTypeOriginal> &c, const std::vector<float> &ring) {
  for (size_t i = 0; i < c.size(); ++i) c[i].ring = (std::uint16_t)ring[i];
}

int main(int argc, char **argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: adapter_driver case.bin\n"); return 2; }
  FILE *f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  // header: 8 counts (map corner, map surf, scan corner, scan surf, last corner, last surf, curr sharp, curr flat)
  int32_t n[8];
  if (fread(n, 4, 8, f) != 8) return 2;
  std::vector<float> a[8], ring_lc, ring_ls;
  for (int i = 0; i < 8; ++i) if (!read_f32(f, a[i], (size_t)n[i] * 4)) return 2;
  if (!read_f32(f, ring_lc, n[4]) || !read_f32(f, ring_ls, n[5])) return 2;
  double init_map[7], init_odo[7];
  if (fread(init_map, 8, 7, f) != 7 || fread(init_odo, 8, 7, f) != 7) return 2;
  std::fclose(f);

  // ---- LaserMapping::MatchScan2Map (LiDAR-only branch)
  TimestampedPointCloud<PointType> map, scan;
  fill(*map.cloud_corner_less_sharp, a[0], nullptr); fill(*map.cloud_surf_less_flat, a[1], nullptr);
  fill(*scan.cloud_corner_less_sharp, a[2], nullptr); fill(*scan.cloud_surf_less_flat, a[3], nullptr);
  GpuMappingScanMatcher mapping;
  MappingScanMatcher *mbase = &mapping;  // the drivers hold the base type (laser_mapping.h:65)
  Rigid3d pose = msfl_adapter::FromArray(init_map);
  Vector3d velocity;
  const bool ok_map = mbase->MatchScan2Map(map, scan, false, nullptr, Vector3d(0, 0, -9.8), RobotState{}, &pose, &velocity);
  double out[7];
  msfl_adapter::ToArray(pose, out);
  std::printf("MAP %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", (int)ok_map, out[0], out[1], out[2], out[3], out[4], out[5], out[6]);

  if (argc > 2 && !std::strcmp(argv[2], "bench")) {  // latency of the drop-in call, no Python in the loop
    const int reps = 200;
    for (int w = 0; w < 10; ++w) { pose = msfl_adapter::FromArray(init_map); mbase->MatchScan2Map(map, scan, false, nullptr, Vector3d(0, 0, -9.8), RobotState{}, &pose, &velocity); }
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) {
      pose = msfl_adapter::FromArray(init_map);
      mbase->MatchScan2Map(map, scan, false, nullptr, Vector3d(0, 0, -9.8), RobotState{}, &pose, &velocity);
    }
    const double us_full = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
    // the same without the per-frame submap upload + index build (a caller that keeps the map resident)
    msfl_adapter::Engine eng;
    const msfl_cloud mc = msfl_adapter::View(*map.cloud_corner_less_sharp, false), ms = msfl_adapter::View(*map.cloud_surf_less_flat, false);
    const msfl_cloud sc = msfl_adapter::View(*scan.cloud_corner_less_sharp, false), ss = msfl_adapter::View(*scan.cloud_surf_less_flat, false);
    msfl_set_submap(eng.get(), &mc, &ms);
    double p7[7];
    for (int w = 0; w < 10; ++w) { std::memcpy(p7, init_map, sizeof p7); msfl_scan2map(eng.get(), &sc, &ss, p7, nullptr); }
    t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) { std::memcpy(p7, init_map, sizeof p7); msfl_scan2map(eng.get(), &sc, &ss, p7, nullptr); }
    const double us_solve = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
    t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) msfl_set_submap(eng.get(), &mc, &ms);
    const double us_map = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
    std::printf("BENCH MatchScan2Map_us %.1f scan2map_us %.1f set_submap_us %.1f\n", us_full, us_solve, us_map);
    return 0;
  }

  // ---- LaserOdometry::AddLaserScan -> MatchScan2Scan
  TimestampedPointCloud<PointTypeOriginal> last, curr;
  fill(*last.cloud_corner_less_sharp, a[4], nullptr); fill_rings(*last.cloud_corner_less_sharp, ring_lc);
  fill(*last.cloud_surf_less_flat, a[5], nullptr); fill_rings(*last.cloud_surf_less_flat, ring_ls);
  fill(*curr.cloud_corner_sharp, a[6], nullptr); fill(*curr.cloud_surf_flat, a[7], nullptr);
  GpuOdometryScanMatcher odometry;
  OdometryScanMatcher *obase = &odometry;
  Rigid3d rel = msfl_adapter::FromArray(init_odo);
  const bool ok_odo = obase->MatchScan2Scan(last, curr, &rel);
  msfl_adapter::ToArray(rel, out);
  std::printf("ODO %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", (int)ok_odo, out[0], out[1], out[2], out[3], out[4], out[5], out[6]);

  // ---- the IMU-initialised branch: the stub Ceres leaves the predict at the previous state; two-sample preintegration
  auto pre = std::make_shared<IntegrationBase>();
  pre->sum_dt_buf_ = {0.0, 0.2};
  pre->delta_p_buf_ = {Vector3d(0, 0, 0), Vector3d(0, 0, 0)};
  pre->delta_q_buf_ = {Quaterniond(1, 0, 0, 0), Quaterniond(1, 0, 0, 0)};
  RobotState prev;
  prev.p = Vector3d(init_map[0], init_map[1], init_map[2]);
  prev.q = Quaterniond(init_map[6], init_map[3], init_map[4], init_map[5]);
  prev.imu_preintegration = pre;
  Rigid3d pose2;
  Vector3d vel2(0, 0, 0);
  const bool ok_dsk = mbase->MatchScan2Map(map, scan, true, pre, Vector3d(0, 0, 0), prev, &pose2, &vel2);
  msfl_adapter::ToArray(pose2, out);
  std::printf("DSK %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", (int)ok_dsk, out[0], out[1], out[2], out[3], out[4], out[5], out[6]);

  // ---- LaserMapping's whole frame (MatchScan2Map + InsertScan2Map, laser_mapping.cc:258-340) on GPU-resident maps:
  // frame 0 inserts the map clouds (world frame = identity pose; nothing to match against yet), frame 1 matches the scan
  GpuMappingFrame frame;
  Rigid3d world = msfl_adapter::FromArray(init_odo);  // identity
  const bool m0 = frame.MatchAndInsert(*map.cloud_corner_less_sharp, *map.cloud_surf_less_flat, &world);
  Rigid3d pose3 = msfl_adapter::FromArray(init_map);
  const bool m1 = frame.MatchAndInsert(*scan.cloud_corner_less_sharp, *scan.cloud_surf_less_flat, &pose3);
  msfl_adapter::ToArray(pose3, out);
  std::printf("FRM %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", (int)(!m0 && m1), out[0], out[1], out[2], out[3], out[4], out[5], out[6]);
  return 0;
}
