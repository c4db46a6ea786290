// simulated_annealing/annealing.hpp -- sa::anneal and sa::energy with the reference's
// signatures, running on the B200 engine.
//
// Reference interface being replaced (/root/reference/include/simulated_annealing/
// annealing.hpp): sa::energy(flat_qubo, state, N) (:31-40) and
//   template <typename T> qubo::Solution anneal(qubo::QUBOModel<int, T> instance, sycl::queue q,
//       std::vector<double> &h_beta_schedule, int num_iter, unsigned int num_tries,
//       int sweeps_per_beta = 1)                                                  (:55-58)
// Same argument order and meaning, same return type, blocking, errors as C++ exceptions
// (std::runtime_error carrying osa_last_error()).  The sycl::queue is replaced by
// devices::queue (helpers/devices.hpp).  A trailing sa::Options argument exposes what the
// engine adds (sequential-sweep mode, Boltzmann rule, fp32 sweep arithmetic, CSR layout,
// seed, shard offset); its defaults reproduce the reference's behaviour: random-site
// attempts, acceptance exp((E_cur - E_new) / beta) > u, seed 1234, fp64.
// A queue that holds several CUDA devices (the default of "gpu": every visible one) makes the
// call shard the trajectories over them (osa_multi_anneal: one host thread and stream per GPU, Q
// replicated, one NCCL all-gather of the best records); the result does not depend on the
// number of devices.
#ifndef ONESOLVER_B200_SA_ANNEALING_HPP_
#define ONESOLVER_B200_SA_ANNEALING_HPP_

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "helpers/devices.hpp"
#include "helpers/qubo_helpers.hpp"
#include "model/qubo.hpp"
#include "model/solution.hpp"
#include "onesolver_b200.h"
#include "simulated_annealing/host_engine.hpp"

namespace sa {

// E(x) = sum_{i<=j} Q[i*N+j] x_i x_j over the upper triangle, fixed i-then-j order
template <typename QuboArray, typename StateArray>
double energy(const QuboArray &flat_qubo, const StateArray &state, int N) {
  double total = 0.0;
  for (int i = 0; i < N; ++i) {
    if (!state[i]) continue;  // x_i = 0 rows contribute exact zeros
    const std::size_t row = static_cast<std::size_t>(i) * N;
    for (int j = i; j < N; ++j) total += flat_qubo[row + j] * state[i] * state[j];
  }
  return total;
}

enum class Layout { automatic, dense, csr };

struct Options {
  int mode = OSA_MODE_RANDOM_SITE;
  int accept_rule = OSA_ACCEPT_REFERENCE;
  int sweep_precision = OSA_SWEEP_F64;
  std::uint64_t seed = 1234;  // reference annealing.hpp:87
  std::uint64_t first_try = 0;
  Layout layout = Layout::automatic;
  osa_stats *stats = nullptr;  // filled on the GPU path when non-null
  // a GPU takes part in a multi-device call only if its shard has at least this many
  // trajectories (a shard smaller than one wave of CTAs gains nothing from another device)
  std::uint64_t min_tries_per_device = 2048;
};

namespace detail {

inline void check(int rc, const char *what) {
  if (rc != OSA_OK) {
    throw std::runtime_error(std::string(what) + ": " + osa_last_error());
  }
}

struct ProblemGuard {
  osa_problem *p = nullptr;
  osa_multi *m = nullptr;
  ~ProblemGuard() {
    osa_problem_destroy(p);
    osa_multi_destroy(m);
  }
};

inline std::vector<double> threshold_scale(const std::vector<double> &beta, int num_iter,
                                           int accept_rule) {
  std::vector<double> ts(num_iter);
  for (int i = 0; i < num_iter; ++i)
    ts[i] = accept_rule == OSA_ACCEPT_REFERENCE ? beta[i] : 1.0 / beta[i];
  return ts;
}

}  // namespace detail

template <typename T>
qubo::Solution anneal(qubo::QUBOModel<int, T> instance, devices::queue q,
                      std::vector<double> &h_beta_schedule, int num_iter, unsigned int num_tries,
                      int sweeps_per_beta = 1, const Options &opt = Options()) {
  const int N = static_cast<int>(instance.get_nodes());
  if (N <= 0) throw std::invalid_argument("anneal: the model has no variables");
  if (num_iter <= 0 || static_cast<std::size_t>(num_iter) > h_beta_schedule.size())
    throw std::invalid_argument("anneal: num_iter must be in [1, beta_schedule.size()]");
  if (num_tries == 0) throw std::invalid_argument("anneal: num_tries must be positive");

  // one validation for every device type (the C ABI applies the same rule)
  for (int i = 0; i < num_iter; ++i)
    if (!(h_beta_schedule[i] > 0.0) || !std::isfinite(h_beta_schedule[i]))
      throw std::invalid_argument("anneal: beta_schedule[" + std::to_string(i) +
                                  "] is not a positive finite number");

  if (q.is_gpu()) {
    detail::ProblemGuard guard;
    // devices that take part: as many of the queue's GPUs as there are shards worth having
    const std::uint64_t per = opt.min_tries_per_device ? opt.min_tries_per_device : 1;
    std::vector<int> devs(q.cuda_devices());
    const std::size_t want = static_cast<std::size_t>(std::max<std::uint64_t>(1, num_tries / per));
    if (devs.size() > want) devs.resize(want);
    const bool multi = devs.size() > 1;
    const double density =
        static_cast<double>(instance.quadratic_terms().size()) * 2.0 / (static_cast<double>(N) * N);
    const bool use_csr = opt.layout == Layout::csr ||
                         (opt.layout == Layout::automatic && N > 2048 && density < 0.05);
    if (use_csr) {
      qubo::QUBOModel<int, double> as_double;
      as_double.set_nodes(N);
      for (const auto &t : instance.linear_terms()) as_double.add_variable(t.first, t.second);
      for (const auto &t : instance.quadratic_terms()) as_double.add_connection(t.first, t.second);
      const auto csr = helpers::build_csr(as_double);
      if (multi)
        detail::check(osa_multi_create_csr_f64(csr.rowptr.data(), csr.col.data(), csr.val.data(),
                                               csr.diag.data(), N, devs.data(),
                                               static_cast<int>(devs.size()), opt.sweep_precision,
                                               &guard.m),
                      "osa_multi_create_csr_f64");
      else
        detail::check(osa_problem_create_csr_f64(csr.rowptr.data(), csr.col.data(), csr.val.data(),
                                                 csr.diag.data(), N, devs[0], opt.sweep_precision,
                                                 &guard.p),
                      "osa_problem_create_csr_f64");
    } else {
      const auto flat = helpers::flatten_qubo(instance);
      std::vector<double> flat64(flat.begin(), flat.end());
      if (multi)
        detail::check(osa_multi_create_dense_f64(flat64.data(), N, devs.data(),
                                                 static_cast<int>(devs.size()),
                                                 opt.sweep_precision, &guard.m),
                      "osa_multi_create_dense_f64");
      else
        detail::check(osa_problem_create_dense_f64(flat64.data(), N, devs[0], opt.sweep_precision,
                                                   &guard.p),
                      "osa_problem_create_dense_f64");
    }
    osa_anneal_params prm{};
    prm.seed = opt.seed;
    prm.first_try = opt.first_try;
    prm.num_tries = num_tries;
    prm.num_iter = num_iter;
    prm.sweeps_per_beta = sweeps_per_beta;
    prm.mode = opt.mode;
    prm.accept_rule = opt.accept_rule;
    std::vector<std::uint8_t> state(N);
    double best_energy = 0.0;
    std::uint64_t best_index = 0;
    if (multi)
      detail::check(osa_multi_anneal(guard.m, h_beta_schedule.data(), &prm, nullptr, nullptr,
                                     state.data(), &best_energy, &best_index, opt.stats, nullptr),
                    "osa_multi_anneal");
    else
      detail::check(osa_anneal(guard.p, h_beta_schedule.data(), &prm, nullptr, nullptr,
                               state.data(), &best_energy, &best_index, opt.stats),
                    "osa_anneal");
    return qubo::Solution(state.begin(), state.end(), best_energy);
  }

  // ---- "cpu" / "host" device types: the host engine, trajectories spread over threads
  if (sweeps_per_beta <= 0) throw std::invalid_argument("anneal: sweeps_per_beta must be positive");
  if (opt.sweep_precision != OSA_SWEEP_F64)
    throw std::invalid_argument("anneal: fp32 sweep arithmetic runs on the gpu device type only "
                                "(the host engine computes in fp64)");
  const auto flat = helpers::flatten_qubo(instance);
  std::vector<double> qsym(flat.begin(), flat.end()), qoff(qsym), diag(N);
  for (int i = 0; i < N; ++i) {
    diag[i] = qsym[static_cast<std::size_t>(i) * N + i];
    qoff[static_cast<std::size_t>(i) * N + i] = 0.0;
  }
  const auto ts = detail::threshold_scale(h_beta_schedule, num_iter, opt.accept_rule);
  std::vector<double> energies(num_tries);
  std::vector<std::vector<std::uint32_t>> states(num_tries);
  const unsigned workers = std::max(1u, std::min<unsigned>(q.max_compute_units(), num_tries));
  std::atomic<unsigned> next{0};
  auto work = [&]() {
    std::vector<char> bits(N);
    for (unsigned t = next++; t < num_tries; t = next++) {
      auto tr = host::run_dense(qoff, diag, N, ts, num_iter, sweeps_per_beta, opt.mode, opt.seed,
                                opt.first_try + t);
      for (int i = 0; i < N; ++i) bits[i] = static_cast<char>((tr.best_state[i >> 5] >> (i & 31)) & 1u);
      energies[t] = energy(qsym, bits, N);  // exact recompute, like the GPU epilogue
      states[t] = std::move(tr.best_state);
    }
  };
  std::vector<std::thread> pool;
  for (unsigned w = 1; w < workers; ++w) pool.emplace_back(work);
  work();
  for (auto &th : pool) th.join();
  const auto best_idx = std::min_element(energies.begin(), energies.end()) - energies.begin();
  std::vector<char> bits(N);
  for (int i = 0; i < N; ++i)
    bits[i] = static_cast<char>((states[best_idx][i >> 5] >> (i & 31)) & 1u);
  return qubo::Solution(bits.begin(), bits.end(), energies[best_idx]);
}

// Parallel tempering on the GPU engine (osa_pt_anneal).  The reference has no such sampler; its
// benchmark report names it as the next step (benchmarks/annealing/performance.md:54-59).
// `betas` is the temperature ladder (strictly increasing); num_groups independent ladders are run
// for num_rounds rounds of sweeps_per_round sequential sweeps, with a replica-exchange step after
// every round.  opt.accept_rule selects the Boltzmann weight (exp(-beta E) or, with the
// reference's rule, exp(-E / beta)); opt.first_try is the id of the first GROUP.  GPU only.
template <typename T>
qubo::Solution parallel_tempering(qubo::QUBOModel<int, T> instance, devices::queue q,
                                  const std::vector<double> &betas, int num_rounds,
                                  int sweeps_per_round, unsigned int num_groups,
                                  const Options &opt = Options()) {
  const int N = static_cast<int>(instance.get_nodes());
  if (N <= 0) throw std::invalid_argument("parallel_tempering: the model has no variables");
  if (!q.is_gpu())
    throw std::runtime_error("parallel_tempering: only --device-type gpu runs parallel tempering");
  detail::ProblemGuard guard;
  const auto flat = helpers::flatten_qubo(instance);
  std::vector<double> flat64(flat.begin(), flat.end());
  detail::check(osa_problem_create_dense_f64(flat64.data(), N, q.cuda_device(), opt.sweep_precision,
                                             &guard.p),
                "osa_problem_create_dense_f64");
  osa_pt_params prm{};
  prm.seed = opt.seed;
  prm.first_group = opt.first_try;
  prm.num_groups = num_groups;
  prm.num_replicas = static_cast<std::int32_t>(betas.size());
  prm.num_rounds = num_rounds;
  prm.sweeps_per_round = sweeps_per_round;
  prm.accept_rule = opt.accept_rule;
  std::vector<std::uint8_t> state(N);
  double best_energy = 0.0;
  std::uint64_t best_index = 0;
  detail::check(osa_pt_anneal(guard.p, betas.data(), &prm, nullptr, nullptr, state.data(),
                              &best_energy, &best_index, opt.stats),
                "osa_pt_anneal");
  return qubo::Solution(state.begin(), state.end(), best_energy);
}

// Population annealing on the same sweep kernel (osa_pa_anneal).  The reference's report names it
// first among the samplers it recommends (benchmarks/annealing/performance.md:54-59); the
// signature follows anneal(): `betas` is the annealing schedule of a population (one temperature
// per step), `population_size` replicas per population, `num_populations` independent populations;
// between two temperatures a population is resampled with the Boltzmann weights of the step.
// opt.first_try is the id of the first POPULATION.  GPU only.
template <typename T>
qubo::Solution population_annealing(qubo::QUBOModel<int, T> instance, devices::queue q,
                                    const std::vector<double> &betas, int sweeps_per_step,
                                    unsigned int population_size, unsigned int num_populations,
                                    const Options &opt = Options()) {
  const int N = static_cast<int>(instance.get_nodes());
  if (N <= 0) throw std::invalid_argument("population_annealing: the model has no variables");
  if (!q.is_gpu())
    throw std::runtime_error(
        "population_annealing: only --device-type gpu runs population annealing");
  detail::ProblemGuard guard;
  const auto flat = helpers::flatten_qubo(instance);
  std::vector<double> flat64(flat.begin(), flat.end());
  detail::check(osa_problem_create_dense_f64(flat64.data(), N, q.cuda_device(), opt.sweep_precision,
                                             &guard.p),
                "osa_problem_create_dense_f64");
  osa_pa_params prm{};
  prm.seed = opt.seed;
  prm.first_population = opt.first_try;
  prm.num_populations = num_populations;
  prm.population_size = static_cast<std::int32_t>(population_size);
  prm.num_steps = static_cast<std::int32_t>(betas.size());
  prm.sweeps_per_step = sweeps_per_step;
  prm.accept_rule = opt.accept_rule;
  std::vector<std::uint8_t> state(N);
  double best_energy = 0.0;
  std::uint64_t best_index = 0;
  detail::check(osa_pa_anneal(guard.p, betas.data(), &prm, nullptr, nullptr, state.data(),
                              &best_energy, &best_index, opt.stats),
                "osa_pa_anneal");
  return qubo::Solution(state.begin(), state.end(), best_energy);
}

}  // namespace sa

#endif
