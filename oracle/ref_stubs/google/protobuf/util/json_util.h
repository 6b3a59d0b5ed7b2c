// stand-in for <google/protobuf/util/json_util.h> (TEST INFRASTRUCTURE): main() of msf_loam_node.cc names it
#ifndef MSFL_PROTOBUF_JSON_STANDIN_H
#define MSFL_PROTOBUF_JSON_STANDIN_H
#include <string>
namespace google {
namespace protobuf {
namespace util {
struct Status {
  bool ok() const { return true; }
  std::string error_message() const { return ""; }
};
template <typename M>
Status JsonStringToMessage(const std::string &, M *) { return Status(); }
}  // namespace util
}  // namespace protobuf
}  // namespace google
#endif
