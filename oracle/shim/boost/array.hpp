// Minimal stand-in for <boost/array.hpp> so the reference REF runtime builds
// without Boost (test infrastructure only; see oracle/README.md).
#pragma once
#include <array>
#include <cstddef>
namespace boost {
template <class T, std::size_t N>
struct array : public std::array<T, N> {
  void assign(const T &v) { this->fill(v); }
};
}  // namespace boost
