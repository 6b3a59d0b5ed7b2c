// Which atan2 does `-atan2(point.y, point.x)` (msf_loam_node.cc:131, :139; float members, unqualified call) bind to?
//   g++ -std=c++14 atan2_overload.cc && ./a.out                 -> <cmath> only: arguments promoted, double result
//   g++ -std=c++14 -DWITH_MATH_H atan2_overload.cc && ./a.out    -> libstdc++'s <math.h> wrapper (`using std::atan2;`)
//                                                                  is in the include graph: float atan2(float, float)
// The reference translation unit includes ros/ros.h (ros/duration.h includes <math.h>), tf (LinearMath/Scalar.h
// includes <math.h>) and PCL, so the second case applies: the azimuth is a float.  tests/test_oracle.py runs both
// builds and checks the sizes; the oracle and k_feat_angles follow the float overload.
#ifdef WITH_MATH_H
#include <math.h>
#else
#include <cmath>
#endif
#include <cstdio>

int main() {
  float y = 0.3f, x = -1.7f;
  auto r = atan2(y, x);
  std::printf("%zu\n", sizeof(r));
  return 0;
}
