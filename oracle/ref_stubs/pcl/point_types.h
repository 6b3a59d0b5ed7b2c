// Stand-in for <pcl/point_types.h> (TEST INFRASTRUCTURE): the point layouts and the Eigen maps the reference's matchers
// touch (pcl::PointXYZI; the PCL_ADD_* / POINT_CLOUD_REGISTER_POINT_STRUCT macros common.h builds PointXYZIRT from).
#ifndef MSFL_PCL_POINT_TYPES_STANDIN_H
#define MSFL_PCL_POINT_TYPES_STANDIN_H
#include <cstdint>

#include "../Eigen/Core"

#define PCL_ADD_POINT4D                                                                              \
  union EIGEN_ALIGN16 {                                                                              \
    float data[4];                                                                                   \
    struct {                                                                                         \
      float x, y, z;                                                                                 \
    };                                                                                               \
  };                                                                                                 \
  inline Eigen::Map<Eigen::Vector3f> getVector3fMap() { return Eigen::Map<Eigen::Vector3f>(data); } \
  inline Eigen::Map<const Eigen::Vector3f> getVector3fMap() const { return Eigen::Map<const Eigen::Vector3f>(data); } \
  inline Eigen::Map<const Eigen::Array3f> getArray3fMap() const { return Eigen::Map<const Eigen::Array3f>(data); }
#define PCL_ADD_INTENSITY float intensity
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fields)

namespace pcl {
struct EIGEN_ALIGN16 PointXYZI {
  PCL_ADD_POINT4D
  float intensity;
  PointXYZI() : data{0.f, 0.f, 0.f, 1.f}, intensity(0.f) {}
};
}  // namespace pcl
#endif
