// helpers/hash.hpp -- hash functor for index pairs used as keys of the quadratic
// coefficient map.  Same name and behaviour as the reference helper
// (/root/reference/include/helpers/hash.hpp:19-35): hash(first) XOR hash(second).
#ifndef ONESOLVER_B200_HELPERS_HASH_HPP_
#define ONESOLVER_B200_HELPERS_HASH_HPP_

#include <cstddef>
#include <functional>
#include <utility>

namespace helpers {

struct hash_pair {
  template <class A, class B>
  std::size_t operator()(const std::pair<A, B> &key) const {
    return std::hash<A>{}(key.first) ^ std::hash<B>{}(key.second);
  }
};

}  // namespace helpers

#endif
