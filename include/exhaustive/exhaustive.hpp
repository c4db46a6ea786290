// exhaustive/exhaustive.hpp -- brute-force ground state of a small QUBO (verification tool).
//
// Same entry point as the reference (/root/reference/include/exhaustive/exhaustive.hpp:29-31):
//   qubo::Solution exhaustive::solve(queue &q, qubo::QUBOModel<Node, Coef> &qubos)
// The reference splits the 2^N states into max_compute_units contiguous ranges, keeps the first
// strict minimum of each range (:104-137) and takes the first minimum over ranges (:158-166), so
// the winner is the LOWEST state integer attaining the minimum; state bit i is variable i
// (helpers/ulong_to_vec.hpp).  Per-state energies use the reference's summation order (upper
// triangle, i then j), so ties resolve identically.  A "gpu" queue runs the CUDA search
// (osa_exhaustive_dense_f64, Gray-code incremental energies, onesolver_b200/csrc/osa_exhaustive.cu);
// "cpu"/"host" queues enumerate on the host threads.  Both lift the reference's `1 << n_bits` int
// limit (exhaustive.hpp:67) to 40 bits.
#ifndef ONESOLVER_B200_EXHAUSTIVE_HPP_
#define ONESOLVER_B200_EXHAUSTIVE_HPP_

#include <algorithm>
#include <limits>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "helpers/devices.hpp"
#include "helpers/qubo_helpers.hpp"
#include "helpers/ulong_to_vec.hpp"
#include "model/qubo.hpp"
#include "model/solution.hpp"

namespace exhaustive {

template <class NodeType, class CoefType>
qubo::Solution solve(devices::queue &q, qubo::QUBOModel<NodeType, CoefType> &qubos) {
  const unsigned n_bits = static_cast<unsigned>(qubos.get_nodes());
  if (n_bits == 0) throw std::invalid_argument("exhaustive: the model has no variables");
  if (n_bits > 40) throw std::invalid_argument("exhaustive: at most 40 variables are supported");

  if (q.is_gpu()) {
    qubo::QUBOModel<int, double> as_double;
    as_double.set_nodes(static_cast<int>(n_bits));
    for (unsigned i = 0; i < n_bits; ++i) {
      as_double.add_variable(static_cast<int>(i), qubos.get_variable(static_cast<NodeType>(i)));
      for (unsigned j = i + 1; j < n_bits; ++j) {
        const double c = qubos.get_connection(
            std::make_pair(static_cast<NodeType>(i), static_cast<NodeType>(j)));
        if (c != 0.0) as_double.add_connection(std::make_pair(static_cast<int>(i), static_cast<int>(j)), c);
      }
    }
    const auto flat = helpers::flatten_qubo(as_double);
    std::vector<unsigned char> state(n_bits);
    double energy = 0.0;
    if (osa_exhaustive_dense_f64(flat.data(), static_cast<int>(n_bits), q.cuda_device(),
                                 state.data(), &energy) != OSA_OK) {
      throw std::runtime_error(std::string("osa_exhaustive_dense_f64: ") + osa_last_error());
    }
    return qubo::Solution(state.begin(), state.end(), energy);
  }

  // upper-triangular coefficient table, as in the reference (:44-61)
  std::vector<double> upper(static_cast<std::size_t>(n_bits) * n_bits, 0.0);
  for (unsigned i = 0; i < n_bits; ++i) {
    upper[i * n_bits + i] = qubos.get_variable(static_cast<NodeType>(i));
    for (unsigned j = i + 1; j < n_bits; ++j)
      upper[i * n_bits + j] = qubos.get_connection(
          std::make_pair(static_cast<NodeType>(i), static_cast<NodeType>(j)));
  }

  const unsigned long long n_states = 1ull << n_bits;
  const unsigned hw = std::max(1u, q.max_compute_units());
  const unsigned workers = static_cast<unsigned>(std::min<unsigned long long>(hw, n_states));
  std::vector<double> energies(workers, std::numeric_limits<double>::max());
  std::vector<unsigned long long> states(workers, 0);

  auto scan = [&](unsigned item) {
    const unsigned long long per = n_states / workers, rem = n_states % workers;
    const unsigned long long begin = item * per + std::min<unsigned long long>(item, rem);
    const unsigned long long end = begin + per + (item < rem ? 1 : 0);
    double e_best = std::numeric_limits<double>::max();
    unsigned long long s_best = 0;
    for (unsigned long long state = begin; state < end; ++state) {
      double e = 0.0;
      for (unsigned i = 0; i < n_bits; ++i) {
        if (!((state >> i) & 1ull)) continue;
        const double *row = upper.data() + static_cast<std::size_t>(i) * n_bits;
        for (unsigned j = i; j < n_bits; ++j)
          if ((state >> j) & 1ull) e += row[j];
      }
      if (e < e_best) {
        e_best = e;
        s_best = state;
      }
    }
    energies[item] = e_best;
    states[item] = s_best;
  };
  std::vector<std::thread> pool;
  for (unsigned w = 1; w < workers; ++w) pool.emplace_back(scan, w);
  scan(0);
  for (auto &th : pool) th.join();

  const auto winner = std::min_element(energies.begin(), energies.end()) - energies.begin();
  const auto bits = helpers::ulong_to_vec(states[winner], n_bits);
  return qubo::Solution(bits.begin(), bits.end(), energies[winner]);
}

}  // namespace exhaustive

#endif
