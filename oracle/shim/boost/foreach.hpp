// Minimal stand-in for <boost/foreach.hpp>: BOOST_FOREACH over a container or
// an iterator pair, expressed with a C++11 range-for.
#pragma once
#include <utility>
namespace boost_shim {
template <class It> struct range_of_pair {
  It b, e;
  It begin() const { return b; }
  It end() const { return e; }
};
template <class C> inline C &as_range(C &c) { return c; }
template <class C> inline const C &as_range(const C &c) { return c; }
template <class It> inline range_of_pair<It> as_range(const std::pair<It, It> &p) {
  return range_of_pair<It>{p.first, p.second};
}
template <class It> inline range_of_pair<It> as_range(std::pair<It, It> &p) {
  return range_of_pair<It>{p.first, p.second};
}
}  // namespace boost_shim
#define BOOST_FOREACH(decl, col) for (decl : ::boost_shim::as_range(col))
