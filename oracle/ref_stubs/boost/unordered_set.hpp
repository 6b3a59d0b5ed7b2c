// Stand-in for <boost/unordered_set.hpp> (TEST INFRASTRUCTURE): hybrid_grid.cc keeps the cells a scan touches in a
// boost::unordered_set of shared pointers; iteration order is hash (= heap address) order in either library.
#ifndef MSFL_BOOST_UNORDERED_SET_STANDIN_H
#define MSFL_BOOST_UNORDERED_SET_STANDIN_H
#include <memory>
#include <unordered_set>
namespace boost {
template <typename T>
using unordered_set = std::unordered_set<T>;
}
#endif
