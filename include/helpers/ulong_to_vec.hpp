// helpers/ulong_to_vec.hpp -- integer state -> vector of 0/1 chars, bit i = variable i
// (reference: /root/reference/include/helpers/ulong_to_vec.hpp:23-32, which throws above
// 32 bits; this version carries 64).
#ifndef ONESOLVER_B200_HELPERS_ULONG_TO_VEC_HPP_
#define ONESOLVER_B200_HELPERS_ULONG_TO_VEC_HPP_

#include <stdexcept>
#include <vector>

namespace helpers {

inline std::vector<char> ulong_to_vec(unsigned long long val, unsigned int n_bits) {
  if (n_bits > 64) throw std::invalid_argument("state can be up to 64 bit");
  std::vector<char> bits(n_bits);
  for (unsigned int i = 0; i < n_bits; ++i) bits[i] = static_cast<char>((val >> i) & 1ull);
  return bits;
}

}  // namespace helpers

#endif
