// Stand-in for the GENERATED proto/config.pb.h (TEST INFRASTRUCTURE; protoc output is not in the reference tree)
#ifndef MSFL_PROTO_CONFIG_STANDIN_H
#define MSFL_PROTO_CONFIG_STANDIN_H
#include <string>
namespace proto {
struct Rigid3d {};
struct MsfLoamConfig {
  std::string DebugString() const { return ""; }
  Rigid3d lidar2imu_extrinsic_parameters() const { return Rigid3d(); }
};
}  // namespace proto
#endif
