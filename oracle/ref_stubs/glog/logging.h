// Stand-in for <glog/logging.h> (TEST INFRASTRUCTURE): log statements are swallowed, a failed CHECK aborts like glog's.
#ifndef MSFL_GLOG_STANDIN_H
#define MSFL_GLOG_STANDIN_H
#include <cstdio>
#include <cstdlib>
#include <ostream>
namespace msfl_glog {
struct NullStream {
  template <typename T>
  NullStream &operator<<(const T &) { return *this; }
  NullStream &operator<<(std::ostream &(*)(std::ostream &)) { return *this; }
};
struct FatalStream {
  FatalStream(const char *file, int line, const char *what) { std::fprintf(stderr, "%s:%d CHECK failed: %s\n", file, line, what); }
  [[noreturn]] ~FatalStream() { std::abort(); }
  template <typename T>
  FatalStream &operator<<(const T &) { return *this; }
};
}  // namespace msfl_glog
#define LOG(severity) ::msfl_glog::NullStream()
#define CHECK(c) \
  if (c) {       \
  } else         \
    ::msfl_glog::FatalStream(__FILE__, __LINE__, #c)
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define LOG_IF(severity, condition) ::msfl_glog::NullStream()
// debug checks: compiled, never evaluated (NDEBUG behaviour of glog)
#define DCHECK(c) \
  while (false) ::msfl_glog::NullStream()
#define DCHECK_LT(a, b) DCHECK((a) < (b))
#define DCHECK_GE(a, b) DCHECK((a) >= (b))
// the gflags slice msf_loam_node.cc uses (glog pulls gflags in upstream)
#include <string>
#define DEFINE_bool(name, value, help) bool FLAGS_##name = value
#define DEFINE_string(name, value, help) std::string FLAGS_##name = value
static bool FLAGS_alsologtostderr __attribute__((unused)) = false;
namespace google {
inline void InitGoogleLogging(const char *) {}
inline void ParseCommandLineFlags(int *, char ***, bool) {}
}  // namespace google
#endif
