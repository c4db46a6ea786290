// osa_pt.cu -- parallel tempering (replica exchange) around the dense sweep kernel.
//
// The reference's own benchmark report recommends parallel tempering as the next sampler
// (/root/reference/benchmarks/annealing/performance.md:54-59); there is no reference code for it.
// A PT run is G independent groups of M replicas; replica slot k of group g is trajectory
// g*M + k of the sweep kernel.  Every round the sweep kernel resumes all trajectories for S
// sequential sweeps at their current temperature, the exact energies of the states they end in
// are recomputed in fp64, and neighbouring temperatures try to exchange their configurations.
// Exchanges move the TEMPERATURES between slots (the states stay where they are):
//   temp_of_slot[t] = rung of the ladder trajectory t currently runs at,
//   slot_of_temp[g*M + j] = slot of group g that currently holds rung j.
// The helpers here: initial states, the swap step, and best-so-far bookkeeping across rounds.
#include "osa_common.cuh"

namespace osa {

namespace {

// initial spins of every trajectory, exactly the bits the sweep kernels draw (STREAM_INIT)
__global__ void k_pt_init_states(uint64_t seed, uint64_t first_try, uint64_t num_tries, int n,
                                 int nw, uint32_t *__restrict__ states) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= num_tries * (uint64_t)nw) return;
  const uint64_t tl = idx / (uint64_t)nw;
  const int k = (int)(idx % (uint64_t)nw);
  const U4 d = engine_draw(seed, first_try + tl, STREAM_INIT, (uint32_t)k >> 2, 0u);
  uint32_t word = pick(d, (uint32_t)k & 3u);
  const int valid = n - k * 32;
  if (valid < 32) word &= (1u << valid) - 1u;
  states[idx] = word;
}

// best state seen by each trajectory over all rounds.  A round reports its best energy relative
// to the state it started from (best_rel <= 0, tracked in the sweep precision); e_start is the
// exact energy of that start state.  Strict improvement, like annealing.hpp:115-121.
__global__ void k_pt_track_best(const double *__restrict__ e_start,
                                const double *__restrict__ best_rel,
                                const uint32_t *__restrict__ round_best, uint64_t num_tries, int nw,
                                double *__restrict__ best_e, uint32_t *__restrict__ best_keep) {
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= num_tries) return;
  const double cand = det::add(e_start[warp], best_rel[warp]);
  const bool better = cand < best_e[warp];
  if (!better) return;  // warp-uniform
  for (int k = lane; k < nw; k += 32)
    best_keep[warp * (uint64_t)nw + k] = round_best[warp * (uint64_t)nw + k];
  __syncwarp();
  if (lane == 0) best_e[warp] = cand;
}

// One exchange step: in round t the pairs of rungs (j, j+1) with j = t mod 2, t mod 2 + 2, ...
// are tried.  With inverse temperatures b_j (b = beta for the Boltzmann rule, 1/beta for the
// reference rule) the exchange is accepted with probability min(1, exp(x)),
//   x = (b_j - b_{j+1}) * (E_a - E_b),   a / b = slots holding rungs j / j+1,
// decided as  x >= 0  or  -x < -ln(u)  with u from STREAM_PT (key: group, pair j, round t).
template <typename T>
__global__ void k_pt_swap(uint64_t seed, uint64_t first_group, uint64_t num_groups, int replicas,
                          uint32_t round, const double *__restrict__ e_cur,
                          const double *__restrict__ dinv, const T *__restrict__ tscale_of_temp,
                          int32_t *__restrict__ temp_of_slot, int32_t *__restrict__ slot_of_temp,
                          T *__restrict__ tscale_traj, unsigned long long *swap_count) {
  const int pairs = replicas / 2;  // upper bound of pairs per group and parity
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pairs == 0 || idx >= num_groups * (uint64_t)pairs) return;
  const uint64_t g = idx / (uint64_t)pairs;
  const int j = 2 * (int)(idx % (uint64_t)pairs) + (int)(round & 1u);
  if (j + 1 >= replicas) return;
  const uint64_t base = g * (uint64_t)replicas;
  const int a = slot_of_temp[base + j], b = slot_of_temp[base + j + 1];
  const double x = det::mul(dinv[j], det::add(e_cur[base + a], -e_cur[base + b]));
  bool accept = x >= 0.0;
  if (!accept) {
    const U4 d = engine_draw(seed, first_group + g, STREAM_PT, (uint32_t)j, round);
    accept = -x < (double)neglogf_det(d.x);
  }
  if (!accept) return;
  slot_of_temp[base + j] = b;
  slot_of_temp[base + j + 1] = a;
  temp_of_slot[base + a] = j + 1;
  temp_of_slot[base + b] = j;
  tscale_traj[base + a] = tscale_of_temp[j + 1];
  tscale_traj[base + b] = tscale_of_temp[j];
  atomicAdd(swap_count, 1ull);
}

}  // namespace

cudaError_t launch_pt_init_states(uint64_t seed, uint64_t first_try, uint64_t num_tries, int n,
                                  int nw, uint32_t *states, cudaStream_t s) {
  const uint64_t total = num_tries * (uint64_t)nw;
  const uint64_t grid = (total + 255) / 256;
  if (grid == 0 || grid > 0x7fffffffull) return cudaErrorInvalidValue;
  k_pt_init_states<<<(unsigned)grid, 256, 0, s>>>(seed, first_try, num_tries, n, nw, states);
  return cudaGetLastError();
}

cudaError_t launch_pt_track_best(const double *e_start, const double *best_rel,
                                 const uint32_t *round_best, uint64_t num_tries, int nw,
                                 double *best_e, uint32_t *best_keep, cudaStream_t s) {
  const uint64_t grid = (num_tries * 32 + 255) / 256;
  if (grid == 0 || grid > 0x7fffffffull) return cudaErrorInvalidValue;
  k_pt_track_best<<<(unsigned)grid, 256, 0, s>>>(e_start, best_rel, round_best, num_tries, nw,
                                                 best_e, best_keep);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_pt_swap(uint64_t seed, uint64_t first_group, uint64_t num_groups, int replicas,
                           uint32_t round, const double *e_cur, const double *dinv,
                           const T *tscale_of_temp, int32_t *temp_of_slot, int32_t *slot_of_temp,
                           T *tscale_traj, unsigned long long *swap_count, cudaStream_t s) {
  const int pairs = replicas / 2;
  if (pairs == 0) return cudaSuccess;
  const uint64_t total = num_groups * (uint64_t)pairs;
  const uint64_t grid = (total + 127) / 128;
  if (grid == 0 || grid > 0x7fffffffull) return cudaErrorInvalidValue;
  k_pt_swap<T><<<(unsigned)grid, 128, 0, s>>>(seed, first_group, num_groups, replicas, round, e_cur,
                                             dinv, tscale_of_temp, temp_of_slot, slot_of_temp,
                                             tscale_traj, swap_count);
  return cudaGetLastError();
}

template cudaError_t launch_pt_swap<float>(uint64_t, uint64_t, uint64_t, int, uint32_t,
                                           const double *, const double *, const float *, int32_t *,
                                           int32_t *, float *, unsigned long long *, cudaStream_t);
template cudaError_t launch_pt_swap<double>(uint64_t, uint64_t, uint64_t, int, uint32_t,
                                            const double *, const double *, const double *,
                                            int32_t *, int32_t *, double *, unsigned long long *,
                                            cudaStream_t);

}  // namespace osa
