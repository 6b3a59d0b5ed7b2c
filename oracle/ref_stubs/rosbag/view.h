// stand-in for <rosbag/view.h> (TEST INFRASTRUCTURE): an empty bag
#ifndef MSFL_ROSBAG_STANDIN_H
#define MSFL_ROSBAG_STANDIN_H
#include "../ros/ros.h"
namespace rosbag {
struct MessageInstance {
  template <typename T>
  bool isType() const { return false; }
  template <typename T>
  std::shared_ptr<const T> instantiate() const { return std::shared_ptr<const T>(); }
};
struct Bag {
  void open(const std::string &) {}
  void close() {}
};
struct View {
  explicit View(const Bag &) {}
  std::vector<MessageInstance>::iterator begin() { return none_.begin(); }
  std::vector<MessageInstance>::iterator end() { return none_.end(); }
  std::vector<MessageInstance> none_;
};
}  // namespace rosbag
#endif
