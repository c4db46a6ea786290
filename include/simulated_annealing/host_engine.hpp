// simulated_annealing/host_engine.hpp -- the annealing engine on the host ("cpu"/"host"
// device types, the counterpart of SYCL's host and CPU devices in the reference CLI).
//
// This is the SAME algorithm as the CUDA kernels (onesolver_b200/csrc): bit-packed spins,
// local field h, dE = (1 - 2 x_i) h_i, acceptance dE < tscale * (-ln u) with u from
// Philox4x32-10 keyed by (seed, trajectory, sweep, site), strict-improvement best
// tracking (reference annealing.hpp:115-121).  Every rounding-sensitive operation is an
// explicit std::fma or an exact product, so host and GPU runs of the same call return
// identical states.  It is selected explicitly by the caller's device type; the GPU
// path never falls back to it.
#ifndef ONESOLVER_B200_SA_HOST_ENGINE_HPP_
#define ONESOLVER_B200_SA_HOST_ENGINE_HPP_

#include <atomic>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

namespace sa {
namespace host {

struct Draw { std::uint32_t w[4]; };

inline Draw philox4x32_10(std::uint32_t c0, std::uint32_t c1, std::uint32_t c2, std::uint32_t c3,
                          std::uint32_t k0, std::uint32_t k1) {
  for (int round = 0; round < 10; ++round) {
    const std::uint64_t p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
    const std::uint32_t n0 = static_cast<std::uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const std::uint32_t n2 = static_cast<std::uint32_t>(p0 >> 32) ^ c3 ^ k1;
    c1 = static_cast<std::uint32_t>(p1);
    c3 = static_cast<std::uint32_t>(p0);
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Draw{{c0, c1, c2, c3}};
}

// streams: 0 = initial spins, 1 = sequential sweep (site>>2, sweep), 2 = random site (0, step)
inline Draw engine_draw(std::uint64_t seed, std::uint64_t traj, std::uint32_t stream,
                        std::uint32_t c0, std::uint32_t c1) {
  return philox4x32_10(c0, c1, static_cast<std::uint32_t>(traj),
                       (static_cast<std::uint32_t>(traj >> 32) & 0x3fffffffu) | (stream << 30),
                       static_cast<std::uint32_t>(seed), static_cast<std::uint32_t>(seed >> 32));
}

// -ln(u), u = (2w+1)/2^33: 24-bit mantissa + single-precision polynomial, fma only
inline float neglogf_det(std::uint32_t w) {
  const std::uint64_t v = (static_cast<std::uint64_t>(w) << 1) | 1u;
  const int p = 63 - __builtin_clzll(v);
  const std::uint32_t m24 = static_cast<std::uint32_t>((v << (63 - p)) >> 40);
  int e = p - 33;
  float mf = static_cast<float>(m24) * 1.1920928955078125e-07f;
  if (m24 > 0x00B504F3u) {
    mf = mf * 0.5f;
    e += 1;
  }
  const float f = mf - 1.0f;
  const float z = f * f;
  float y = 7.0376836292E-2f;
  y = std::fmaf(y, f, -1.1514610310E-1f);
  y = std::fmaf(y, f, 1.1676998740E-1f);
  y = std::fmaf(y, f, -1.2420140846E-1f);
  y = std::fmaf(y, f, 1.4249322787E-1f);
  y = std::fmaf(y, f, -1.6668057665E-1f);
  y = std::fmaf(y, f, 2.0000714765E-1f);
  y = std::fmaf(y, f, -2.4999993993E-1f);
  y = std::fmaf(y, f, 3.3333331174E-1f);
  volatile float yf = y * f;  // volatile: keep the two products un-fused with what follows
  volatile float yz = yf * z;
  const float fe = static_cast<float>(e);
  y = std::fmaf(fe, -2.12194440e-4f, yz);
  y = std::fmaf(-0.5f, z, y);
  volatile float r0 = f + y;
  return -std::fmaf(fe, 0.693359375f, r0);
}

struct Trajectory {
  std::vector<std::uint32_t> best_state;  // packed
  double best_rel = 0.0;
  std::uint64_t accepts = 0;
};

// One trajectory on a dense problem: qoff = symmetric matrix with zeroed diagonal
// (row-major, leading dimension n), diag = linear terms, tscale per iteration.
inline Trajectory run_dense(const std::vector<double> &qoff, const std::vector<double> &diag,
                            int n, const std::vector<double> &tscale, int num_iter,
                            int sweeps_per_beta, int mode, std::uint64_t seed,
                            std::uint64_t traj) {
  const int nw = (n + 31) / 32;
  std::vector<std::uint32_t> x(nw), xb;
  for (int k = 0; k < nw; ++k) {
    std::uint32_t word = engine_draw(seed, traj, 0u, static_cast<std::uint32_t>(k) >> 2, 0u).w[k & 3];
    const int valid = n - k * 32;
    if (valid < 32) word &= (1u << valid) - 1u;
    x[k] = word;
  }
  xb = x;
  std::vector<double> h(diag);
  auto add_row = [&](int k, double sgn) {
    const double *row = qoff.data() + static_cast<std::size_t>(k) * n;
    for (int j = 0; j < n; ++j) h[j] = std::fma(sgn, row[j], h[j]);
  };
  for (int i = 0; i < n; ++i)
    if ((x[i >> 5] >> (i & 31)) & 1u) add_row(i, 1.0);

  Trajectory out;
  double erel = 0.0, best = 0.0;
  std::uint32_t step = 0;
  for (int iter = 0; iter < num_iter; ++iter) {
    const double ts = tscale[iter];
    for (int sw = 0; sw < sweeps_per_beta; ++sw, ++step) {
      const int sites = mode == OSA_MODE_SEQUENTIAL_SWEEP ? n : 1;
      for (int s = 0; s < sites; ++s) {
        int k;
        std::uint32_t wu;
        if (mode == OSA_MODE_SEQUENTIAL_SWEEP) {
          k = s;
          wu = engine_draw(seed, traj, 1u, static_cast<std::uint32_t>(s) >> 2, step).w[s & 3];
        } else {
          const Draw d = engine_draw(seed, traj, 2u, 0u, step);
          k = static_cast<int>((static_cast<std::uint64_t>(d.w[0]) * static_cast<std::uint64_t>(n)) >> 32);
          wu = d.w[1];
        }
        const double theta = ts * static_cast<double>(neglogf_det(wu));
        const bool xk = (x[k >> 5] >> (k & 31)) & 1u;
        const double dE = xk ? -h[k] : h[k];
        if (dE < theta) {
          add_row(k, xk ? -1.0 : 1.0);
          x[k >> 5] ^= (1u << (k & 31));
          erel += dE;
          ++out.accepts;
          if (erel < best) {
            best = erel;
            xb = x;
          }
        }
      }
    }
  }
  out.best_state = xb;
  out.best_rel = best;
  return out;
}

}  // namespace host
}  // namespace sa

#endif
