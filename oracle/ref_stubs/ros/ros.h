// Stand-in for the slice of roscpp / boost::bind that msf_loam_node.cc's main() and message handlers name
// (TEST INFRASTRUCTURE).  Nothing here does anything: the harness calls RealHandleLaserCloudMessage directly; main() only
// has to compile because the file is compiled whole and unmodified.
#ifndef MSFL_ROS_STANDIN_H
#define MSFL_ROS_STANDIN_H
#include <cstdint>
#include <fstream>
#include <functional>
#include <memory>
#include <string>
#include <vector>

namespace ros {
struct Time {
  std::uint32_t sec = 0, nsec = 0;
  std::uint64_t toNSec() const { return (std::uint64_t)sec * 1000000000ull + nsec; }
};
inline void init(int &, char **, const std::string &) {}
struct Subscriber {};
struct Publisher {};
struct NodeHandle {
  template <typename T>
  bool param(const std::string &, T &value, const T &fallback) const {
    value = fallback;
    return false;
  }
  template <typename M, typename F>
  Subscriber subscribe(const std::string &, int, F) {
    return Subscriber();
  }
};
struct AsyncSpinner {
  explicit AsyncSpinner(int) {}
  void start() {}
};
inline void waitForShutdown() {}
}  // namespace ros

namespace std_msgs {
struct Header {
  ros::Time stamp;
  std::string frame_id;
};
}  // namespace std_msgs

namespace boost {
using std::bind;
using std::ref;
}  // namespace boost
using namespace std::placeholders;  // _1, as <boost/bind.hpp> puts it in the global namespace
#endif
