// app/schedules.hpp -- the two beta schedules of the annealing CLI.
// Values must be bit-identical to the reference builders
// (/root/reference/app/one-solver-anneal.cpp:23-39), including their quirks: the linear
// schedule ends at beta_min + beta_max, the geometric one is an iterated product, and
// num_iter == 1 divides by zero in both.
#ifndef ONESOLVER_B200_APP_SCHEDULES_HPP_
#define ONESOLVER_B200_APP_SCHEDULES_HPP_

#include <cmath>
#include <vector>

inline void construct_linear_beta_schedule(std::vector<double> &schedule, double beta_min,
                                           double beta_max, unsigned int num_iter) {
  const double last = static_cast<double>(num_iter - 1);
  for (unsigned int i = 0; i < num_iter; ++i) schedule[i] = beta_min + beta_max * i / last;
}

inline void construct_geometric_beta_schedule(std::vector<double> &schedule, double beta_min,
                                              double beta_max, unsigned int num_iter) {
  const double alpha = std::pow(beta_max / beta_min, 1.0 / (num_iter - 1));
  schedule[0] = beta_min;
  for (unsigned int i = 1; i < num_iter; ++i) schedule[i] = schedule[i - 1] * alpha;
}

#endif
