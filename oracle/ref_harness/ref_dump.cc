// ref_dump.cc -- runs the UNMODIFIED reference matchers (kekeliu-whu/MSF_LOAM @ 96924b3) on the arrays of
// tests/golden/ and prints what they compute, so that the oracle (oracle/msfl_oracle.c) and the CUDA engine can be
// pinned to the reference's own PCL + Ceres + Eigen arithmetic wherever those libraries exist.
//
// TEST INFRASTRUCTURE, not product code, and NOT buildable in the development image (no PCL / Ceres / Eigen / glog):
// build it on a machine that has the reference's dependencies (the reference's CI image, .github/workflows/Dockerfile)
// with the CMakeLists.txt next to this file, which compiles the reference's sources where they lie -- nothing is
// copied.  Usage:
//     python oracle/ref_harness/export_case.py tests/golden/vlp16_golden.npz /tmp/case.bin
//     ./ref_dump /tmp/case.bin > tests/golden/ref_dump_vlp16.txt 2> ref_dump.log
// tests/test_ref_dump.py consumes the dump when it is present (and says "parity unpinned" when it is not).
//
// Output lines (stdout; Ceres' own progress table of the mapping solve, minimizer_progress_to_stdout = true at
// mapping_scan_matcher.cc:253, is passed through untouched between them):
//     REF_MAP <ok> tx ty tz qx qy qz qw            MappingScanMatcher::MatchScan2Map, LiDAR-only branch
//     REF_ODO <ok> tx ty tz qx qy qz qw            OdometryScanMatcher::MatchScan2Scan
#include <glog/logging.h>

#include <cstdint>
#include <cstdio>
#include <memory>
#include <vector>

#include "slam/local/scan_matching/mapping_scan_matcher.h"
#include "slam/local/scan_matching/odometry_scan_matcher.h"

// defined in laser_mapping.cc (:36-38), which drags in ROS; integration_base.cc only needs the values
double ACC_N = 0.017, ACC_W = 0.007;
double GYR_N = 0.0033, GYR_W = 0.0012;
Eigen::Vector3d G;

namespace {

bool ReadF32(FILE *f, std::vector<float> *v, size_t n) {
  v->resize(n);
  return n == 0 || fread(v->data(), 4, n, f) == n;
}

template <typename P>
void Fill(const std::vector<float> &xyzi, pcl::PointCloud<P> *cloud) {
  cloud->resize(xyzi.size() / 4);
  for (size_t i = 0; i < cloud->size(); ++i) {
    auto &p = (*cloud)[i];
    p.x = xyzi[4 * i];
    p.y = xyzi[4 * i + 1];
    p.z = xyzi[4 * i + 2];
    p.intensity = xyzi[4 * i + 3];
  }
}

void Print(const char *tag, bool ok, Rigid3d pose) {
  const auto v = pose.ToVector7();  // t xyz, q xyzw (rigid_transform.h:59-64)
  std::printf("%s %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", tag, (int)ok, v[0], v[1], v[2], v[3], v[4], v[5], v[6]);
  std::fflush(stdout);
}

Rigid3d FromArray(const double *a) {
  return Rigid3d(Eigen::Vector3d(a[0], a[1], a[2]), Eigen::Quaterniond(a[6], a[3], a[4], a[5]));
}

}  // namespace

int main(int argc, char **argv) {
  google::InitGoogleLogging(argv[0]);
  if (argc < 2) {
    std::fprintf(stderr, "usage: ref_dump case.bin\n");
    return 2;
  }
  FILE *f = std::fopen(argv[1], "rb");
  CHECK(f != nullptr) << argv[1];
  // same layout as tests/c_abi/adapter_driver.cc: 8 int32 counts (map corner, map surf, scan corner, scan surf, last
  // corner, last surf, curr sharp, curr flat), the eight n x 4 float arrays, two ring arrays (float), two poses
  int32_t n[8];
  CHECK_EQ(fread(n, 4, 8, f), 8u);
  std::vector<float> a[8], ring_lc, ring_ls;
  for (int i = 0; i < 8; ++i) CHECK(ReadF32(f, &a[i], (size_t)n[i] * 4));
  CHECK(ReadF32(f, &ring_lc, n[4]));
  CHECK(ReadF32(f, &ring_ls, n[5]));
  double init_map[7], init_odo[7];
  CHECK_EQ(fread(init_map, 8, 7, f), 7u);
  CHECK_EQ(fread(init_odo, 8, 7, f), 7u);
  std::fclose(f);

  {  // ---- scan-to-map, is_initialized = false
    TimestampedPointCloud<PointType> map, scan;
    Fill(a[0], map.cloud_corner_less_sharp.get());
    Fill(a[1], map.cloud_surf_less_flat.get());
    Fill(a[2], scan.cloud_corner_less_sharp.get());
    Fill(a[3], scan.cloud_surf_less_flat.get());
    // MatchScan2Map calls GetDeltaQP for every point even in the LiDAR-only branch (:115, :185; its result is unused
    // there) and GetDeltaQP CHECKs the time range (scan_undistortion.cc:26): give it a preintegration that covers it
    auto pre = std::make_shared<IntegrationBase>(Eigen::Vector3d(0, 0, 9.8), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(),
                                                 Eigen::Vector3d::Zero());
    for (int k = 0; k < 40; ++k) pre->push_back(0.01, Eigen::Vector3d(0, 0, 9.8), Eigen::Vector3d::Zero());
    MappingScanMatcher matcher;
    Rigid3d pose = FromArray(init_map);
    Vector3d velocity = Vector3d::Zero();
    RobotState prev{};
    prev.p = Vector3d::Zero();
    prev.v = Vector3d::Zero();
    prev.q = Quaterniond::Identity();
    const bool ok = matcher.MatchScan2Map(map, scan, /*is_initialized=*/false, pre, Vector3d(0, 0, -9.8), prev, &pose, &velocity);
    Print("REF_MAP", ok, pose);
  }
  {  // ---- scan-to-scan
    TimestampedPointCloud<PointTypeOriginal> last, curr;
    Fill(a[4], last.cloud_corner_less_sharp.get());
    Fill(a[5], last.cloud_surf_less_flat.get());
    for (size_t i = 0; i < last.cloud_corner_less_sharp->size(); ++i) (*last.cloud_corner_less_sharp)[i].ring = (uint16_t)ring_lc[i];
    for (size_t i = 0; i < last.cloud_surf_less_flat->size(); ++i) (*last.cloud_surf_less_flat)[i].ring = (uint16_t)ring_ls[i];
    Fill(a[6], curr.cloud_corner_sharp.get());
    Fill(a[7], curr.cloud_surf_flat.get());
    OdometryScanMatcher matcher;
    Rigid3d pose = FromArray(init_odo);
    const bool ok = matcher.MatchScan2Scan(last, curr, &pose);
    Print("REF_ODO", ok, pose);
  }
  return 0;
}
