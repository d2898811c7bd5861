// Minimal stand-in for <boost/function.hpp> (maps onto std::function).
#pragma once
#include <functional>
namespace boost {
template <class S> using function = std::function<S>;
}
