// osa_exhaustive.cu -- brute-force ground state of a small QUBO on the GPU (N <= 40).
//
// CUDA counterpart of the reference's exhaustive solver
// (/root/reference/include/exhaustive/exhaustive.hpp:29-167, kernel `calc_energy` :104-137):
// the reference gives every work-item a contiguous range of state integers and evaluates the
// full O(N^2) energy of each state.  Here every thread owns the states that share its HIGH bits
// and walks its 2^L low-bit patterns in Gray-code order, so one step flips one bit k and costs
// dE = (1-2x_k) (q_kk + sum_{j != k} Q_kj x_j): O(N) with the row of Q broadcast from shared
// memory (k is the same for every thread of the grid at a given step).
//
// Result rule of the reference: first strict minimum per range, first minimum over ranges ==
// the LOWEST state integer among the minima, with energies summed in upper-triangle i-then-j
// order.  Incremental energies round differently, so the search runs in two passes: pass 1 finds
// the minimum E*, pass 2 lists every state with E <= E* + tol; the host re-evaluates that short
// list with the reference formula and applies the tie rule exactly.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "osa_common.cuh"

namespace osa {

namespace {

constexpr int EX_THREADS = 256;
constexpr int EX_MAXN = 40;

struct ExBest {
  double e;
  unsigned long long x;
};

// mode 0: per-CTA best (energy, then lowest state); mode 1: append states with E <= thr
__global__ void __launch_bounds__(EX_THREADS) k_exhaustive(const double *__restrict__ qsym, int n,
                                                           int low_bits, int mode, double thr,
                                                           ExBest *__restrict__ cta_best,
                                                           unsigned long long *__restrict__ list,
                                                           unsigned int *__restrict__ list_count,
                                                           unsigned int list_cap) {
  __shared__ double s_q[EX_MAXN * EX_MAXN];
  __shared__ ExBest s_red[EX_THREADS / 32];
  for (int i = threadIdx.x; i < n * n; i += EX_THREADS) s_q[i] = qsym[i];
  __syncthreads();

  const unsigned long long gid = (unsigned long long)blockIdx.x * EX_THREADS + threadIdx.x;
  const unsigned long long n_threads = 1ull << (n - low_bits);
  const bool active = gid < n_threads;
  unsigned long long x = active ? (gid << low_bits) : 0ull;

  // energy of the starting state (low bits zero): upper triangle incl. diagonal
  double e = 0.0;
  for (int i = low_bits; i < n; ++i) {
    if (!((x >> i) & 1ull)) continue;
    for (int j = i; j < n; ++j)
      if ((x >> j) & 1ull) e += s_q[i * n + j];
  }
  double be = e;
  unsigned long long bx = x;
  auto visit = [&]() {
    if (mode == 0) {
      if (e < be || (e == be && x < bx)) {
        be = e;
        bx = x;
      }
    } else if (active && e <= thr) {
      const unsigned int slot = atomicAdd(list_count, 1u);
      if (slot < list_cap) list[slot] = x;
    }
  };
  if (mode == 1) visit();

  const unsigned long long steps = 1ull << low_bits;
  for (unsigned long long i = 1; i < steps; ++i) {
    const int k = __ffsll((long long)i) - 1;  // Gray code: bit flipped at step i (grid-uniform)
    const double *row = s_q + k * n;
    double hk = row[k];
#pragma unroll 4
    for (int j = 0; j < n; ++j)
      if (j != k && ((x >> j) & 1ull)) hk += row[j];
    e += ((x >> k) & 1ull) ? -hk : hk;
    x ^= (1ull << k);
    visit();
  }

  if (mode == 0) {
    if (!active) {
      be = INFINITY;
      bx = ~0ull;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double oe = __shfl_down_sync(0xffffffffu, be, o);
      const unsigned long long ox = __shfl_down_sync(0xffffffffu, bx, o);
      if (oe < be || (oe == be && ox < bx)) {
        be = oe;
        bx = ox;
      }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_red[warp] = ExBest{be, bx};
    __syncthreads();
    if (threadIdx.x == 0) {
      ExBest b = s_red[0];
      for (int w = 1; w < EX_THREADS / 32; ++w)
        if (s_red[w].e < b.e || (s_red[w].e == b.e && s_red[w].x < b.x)) b = s_red[w];
      cta_best[blockIdx.x] = b;
    }
  }
}

double ref_energy_bits(const double *qsym, int n, unsigned long long x) {
  double e = 0.0;  // exhaustive.hpp:111-130 == annealing.hpp:31-40 order
  for (int i = 0; i < n; ++i) {
    if (!((x >> i) & 1ull)) continue;
    for (int j = i; j < n; ++j)
      if ((x >> j) & 1ull) e += qsym[(size_t)i * n + j];
  }
  return e;
}

}  // namespace

// returns cudaSuccess and fills x_best / e_best, or an error; `msg` receives a description
cudaError_t exhaustive_search(const double *qsym_host, int n, unsigned long long *x_best,
                              double *e_best, std::string *msg) {
  double *d_q = nullptr;
  ExBest *d_best = nullptr;
  unsigned long long *d_list = nullptr;
  unsigned int *d_count = nullptr;
  const unsigned int cap = 1u << 16;
  cudaError_t err = cudaSuccess;
  auto step = [&](cudaError_t r) {
    if (err == cudaSuccess) err = r;
    return err == cudaSuccess;
  };
  // threads own the high bits: 2^(n-L) threads, L low bits walked per thread
  int low_bits = n > 18 ? n - 18 : 0;
  if (low_bits > 24) low_bits = 24;           // <= 16M steps per thread
  if (low_bits > n) low_bits = n;
  const unsigned long long n_threads = 1ull << (n - low_bits);
  const unsigned long long grid64 = (n_threads + EX_THREADS - 1) / EX_THREADS;
  if (grid64 > 0x7fffffffull) {
    if (msg) *msg = "exhaustive search: problem too large";
    return cudaErrorInvalidValue;
  }
  const unsigned grid = (unsigned)grid64;
  step(cudaMalloc(&d_q, sizeof(double) * n * n));
  step(cudaMalloc(&d_best, sizeof(ExBest) * grid));
  step(cudaMalloc(&d_list, sizeof(unsigned long long) * cap));
  step(cudaMalloc(&d_count, sizeof(unsigned int)));
  if (step(cudaMemcpy(d_q, qsym_host, sizeof(double) * n * n, cudaMemcpyHostToDevice))) {
    k_exhaustive<<<grid, EX_THREADS>>>(d_q, n, low_bits, 0, 0.0, d_best, d_list, d_count, cap);
    step(cudaGetLastError());
  }
  std::vector<ExBest> best(grid);
  if (step(cudaMemcpy(best.data(), d_best, sizeof(ExBest) * grid, cudaMemcpyDeviceToHost))) {
    ExBest b = best[0];
    for (unsigned i = 1; i < grid; ++i)
      if (best[i].e < b.e || (best[i].e == b.e && best[i].x < b.x)) b = best[i];
    // pass 2: every state within rounding distance of the minimum
    const double tol = 1e-9 * std::max(1.0, std::fabs(b.e));
    step(cudaMemset(d_count, 0, sizeof(unsigned int)));
    if (err == cudaSuccess) {
      k_exhaustive<<<grid, EX_THREADS>>>(d_q, n, low_bits, 1, b.e + tol, d_best, d_list, d_count,
                                         cap);
      step(cudaGetLastError());
    }
    unsigned int count = 0;
    step(cudaMemcpy(&count, d_count, sizeof(count), cudaMemcpyDeviceToHost));
    if (err == cudaSuccess) {
      if (count == 0 || count > cap) {
        // degenerate landscape (more near-minimal states than the list holds): keep pass 1's winner
        *x_best = b.x;
        *e_best = ref_energy_bits(qsym_host, n, b.x);
      } else {
        std::vector<unsigned long long> cand(count);
        step(cudaMemcpy(cand.data(), d_list, sizeof(unsigned long long) * count,
                        cudaMemcpyDeviceToHost));
        double we = 0.0;
        unsigned long long wx = ~0ull;
        for (unsigned long long x : cand) {
          const double e = ref_energy_bits(qsym_host, n, x);
          if (wx == ~0ull || e < we || (e == we && x < wx)) {
            we = e;
            wx = x;
          }
        }
        *x_best = wx;
        *e_best = we;
      }
    }
  }
  cudaFree(d_q);
  cudaFree(d_best);
  cudaFree(d_list);
  cudaFree(d_count);
  if (err != cudaSuccess && msg) *msg = std::string("exhaustive search failed: ") + cudaGetErrorString(err);
  return err;
}

}  // namespace osa
