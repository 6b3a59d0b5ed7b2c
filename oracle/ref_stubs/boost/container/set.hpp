// stand-in for <boost/container/set.hpp> (TEST INFRASTRUCTURE): included by hybrid_grid.cc, not used
#include <set>
