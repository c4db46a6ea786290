// model/solution.hpp -- result of a solver run: a 0/1 assignment and its energy.
// API and CSV format follow the reference (/root/reference/include/model/solution.hpp:19-68;
// format pinned by tests/io_test.cpp:108-118): a header row "0,1,...,N-1,energy" and a value
// row "b0,b1,...,bN-1,<energy>", the energy written with the stream's default formatting.
#ifndef ONESOLVER_B200_MODEL_SOLUTION_HPP_
#define ONESOLVER_B200_MODEL_SOLUTION_HPP_

#include <cstddef>
#include <ostream>
#include <vector>

namespace qubo {

class Solution {
public:
  std::vector<char> state;  // state[i] in {0,1} is the value of variable i
  double energy;

  template <typename InputIt>
  Solution(InputIt begin, InputIt end, double energy) : state(begin, end), energy(energy) {}

  void save(std::ostream &stream) const {
    for (std::size_t i = 0; i < state.size(); ++i) stream << i << ',';
    stream << "energy" << std::endl;
    for (char bit : state) stream << static_cast<int>(bit) << ',';
    stream << energy << std::endl;
  }
};

}  // namespace qubo

#endif
