// osa_sparse.cu -- K2: CSR annealing kernel for sparse (Chimera / Pegasus-like) QUBOs.
//
// Multi-spin coding across replicas: one warp anneals 32 trajectories, lane t owning
// trajectory t.  Word X[j] in shared memory holds spin j of all 32 trajectories
// (bit t = trajectory t), so the CSR row of the visited site is a warp-uniform
// (broadcast) load in the sequential-sweep mode and the local field
//     h_i = q_ii + sum_{j in nbr(i)} Q_ij x_j
// is recomputed on demand from <= deg gathers instead of being stored (N*4 bytes per
// trajectory would not fit on chip at N=5627).  Replaces the reference's dense
// O(N^2)-per-attempt kernel (/root/reference/include/simulated_annealing/
// annealing.hpp:85-126) for sparse instances; the reference itself has no sparse path.
//
// Best-state tracking (annealing.hpp:115-121) is lazy: the transposed best state XB
// (global workspace) is only rewritten, for the lanes concerned, when a trajectory
// LEAVES its best state.
#include "osa_common.cuh"

namespace osa {

namespace {

template <typename T>
__global__ void k_sparse(const SparseParams<T> p) {
  extern __shared__ uint32_t smem_x[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const uint64_t gw = (uint64_t)blockIdx.x * wpb + warp;
  const uint64_t tl0 = gw * 32ull;
  if (tl0 >= p.num_tries) return;  // whole warp leaves; no block-level barrier below
  const uint64_t tl = tl0 + lane;
  const bool tv = tl < p.num_tries;
  const uint64_t traj = p.first_try + tl;
  const int n = p.n;

  uint32_t *X = smem_x + (size_t)warp * n;
  uint32_t *XB = p.xbest_ws + gw * (uint64_t)n;

  // initial spins: lane draws its own packed word, the warp transposes it with ballots
  for (int j0 = 0; j0 < n; j0 += 32) {
    const int k = j0 >> 5;
    const U4 d = engine_draw(p.seed, traj, STREAM_INIT, (uint32_t)k >> 2, 0u);
    const uint32_t w = tv ? pick(d, (uint32_t)k & 3u) : 0u;
    uint32_t mine = 0;
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
      const uint32_t colw = __ballot_sync(0xffffffffu, (w >> jj) & 1u);
      if (lane == jj) mine = colw;
    }
    if (j0 + lane < n) {
      X[j0 + lane] = mine;
      XB[j0 + lane] = mine;
    }
  }
  __syncwarp();

  double erel = 0.0, best = 0.0;
  bool at_best = true;
  unsigned long long cnt_acc = 0;

  // lanes in `leaving` drop out of their best state: copy their (pre-flip) bits to XB
  auto snapshot = [&](uint32_t leaving) {
    for (int j = lane; j < n; j += 32) XB[j] = (XB[j] & ~leaving) | (X[j] & leaving);
  };

  if (p.mode == OSA_MODE_SEQUENTIAL_SWEEP) {
    uint32_t step = 0;
    for (int iter = 0; iter < p.num_iter; ++iter) {
      const T ts = p.tscale[iter];
      for (int sw = 0; sw < p.sweeps_per_beta; ++sw, ++step) {
        U4 d = U4{0, 0, 0, 0};
        for (int i = 0; i < n; ++i) {
          if ((i & 3) == 0) d = engine_draw(p.seed, traj, STREAM_SEQ, (uint32_t)i >> 2, step);
          const T theta = threshold<T>(ts, pick(d, (uint32_t)i & 3u));
          T hk = __ldg(p.diag + i);
          const int pb = __ldg(p.rowptr + i), pe = __ldg(p.rowptr + i + 1);
          for (int q = pb; q < pe; ++q) {
            const int c = __ldg(p.col + q);
            const T v = __ldg(p.val + q);
            if ((X[c] >> lane) & 1u) hk = det::add(hk, v);
          }
          const uint32_t xiw = X[i];
          const T dE = ((xiw >> lane) & 1u) ? -hk : hk;
          const bool acc = tv && (dE < theta);
          bool leave = false;
          if (acc) {
            const double e = det::add(erel, (double)dE);
            erel = e;
            ++cnt_acc;
            if (e < best) {
              best = e;
              at_best = true;
            } else if (at_best) {
              leave = true;
              at_best = false;
            }
          }
          const uint32_t leaving = __ballot_sync(0xffffffffu, leave);
          if (leaving) snapshot(leaving);
          const uint32_t bal = __ballot_sync(0xffffffffu, acc);
          if (bal) {
            __syncwarp();
            if (lane == 0) X[i] = xiw ^ bal;
            __syncwarp();
          }
        }
      }
    }
  } else {
    const uint64_t total = (uint64_t)p.num_iter * (uint64_t)p.sweeps_per_beta;
    for (uint64_t st = 0; st < total; ++st) {
      const int iter = (int)(st / (uint64_t)p.sweeps_per_beta);
      const T ts = p.tscale[iter];
      const U4 d = engine_draw(p.seed, traj, STREAM_RND, 0u, (uint32_t)st);
      const int k = (int)__umulhi(d.x, (uint32_t)n);  // per-lane site (annealing.hpp:101)
      const T theta = threshold<T>(ts, d.y);
      T hk = __ldg(p.diag + k);
      const int pb = __ldg(p.rowptr + k), pe = __ldg(p.rowptr + k + 1);
      for (int q = pb; q < pe; ++q) {
        const int c = __ldg(p.col + q);
        const T v = __ldg(p.val + q);
        if ((X[c] >> lane) & 1u) hk = det::add(hk, v);
      }
      const T dE = ((X[k] >> lane) & 1u) ? -hk : hk;
      const bool acc = tv && (dE < theta);
      bool leave = false;
      if (acc) {
        const double e = det::add(erel, (double)dE);
        erel = e;
        ++cnt_acc;
        if (e < best) {
          best = e;
          at_best = true;
        } else if (at_best) {
          leave = true;
          at_best = false;
        }
      }
      const uint32_t leaving = __ballot_sync(0xffffffffu, leave);
      if (leaving) snapshot(leaving);
      __syncwarp();
      if (acc) atomicXor(&X[k], 1u << lane);
      __syncwarp();
    }
  }

  const uint32_t still = __ballot_sync(0xffffffffu, at_best);
  if (still) snapshot(still);
  __syncwarp();

  // un-transpose: lane assembles its own packed words
  if (tv) {
    for (int k = 0; k < p.nw; ++k) {
      uint32_t word = 0;
      const int jmax = min(32, n - k * 32);
      for (int jj = 0; jj < jmax; ++jj) word |= ((XB[k * 32 + jj] >> lane) & 1u) << jj;
      p.best_states[tl * (uint64_t)p.nw + k] = word;
    }
    p.best_rel[tl] = best;
  }
  // warp-reduce the accept counter
  for (int o = 16; o > 0; o >>= 1) cnt_acc += __shfl_down_sync(0xffffffffu, cnt_acc, o);
  if (lane == 0) {
    atomicAdd(&p.counters->accepts, cnt_acc);
    atomicAdd(&p.counters->row_fetches, cnt_acc);
  }
}

constexpr size_t kMaxSmem = 227 * 1024;

int pick_wpb(int n, uint64_t num_tries, int sm_count) {
  const size_t per_warp = (size_t)n * sizeof(uint32_t);
  int wpb_max = (int)(kMaxSmem / per_warp);
  if (wpb_max < 1) return 0;
  if (wpb_max > 8) wpb_max = 8;
  const uint64_t warps = (num_tries + 31) / 32;
  // CTAs that fit per SM at wpb_max, then balance the warps over whole waves
  uint64_t ctas_per_sm = kMaxSmem / (per_warp * wpb_max);
  if (ctas_per_sm > (uint64_t)(64 / wpb_max)) ctas_per_sm = 64 / wpb_max;  // 64 warps per SM
  const uint64_t slots = (uint64_t)sm_count * ctas_per_sm * wpb_max;
  const uint64_t waves = (warps + slots - 1) / slots;
  const uint64_t ctas = (uint64_t)sm_count * ctas_per_sm * waves;
  int wpb = (int)((warps + ctas - 1) / ctas);
  if (wpb < 1) wpb = 1;
  if (wpb > wpb_max) wpb = wpb_max;
  return wpb;
}

template <typename T>
cudaError_t launch_impl(const SparseParams<T> &p, cudaStream_t s, LaunchInfo *info) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int wpb = pick_wpb(p.n, p.num_tries, sms);
  if (wpb < 1) return cudaErrorInvalidValue;
  const size_t smem = (size_t)wpb * p.n * sizeof(uint32_t);
  cudaError_t err =
      cudaFuncSetAttribute(k_sparse<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  const uint64_t warps = (p.num_tries + 31) / 32;
  const uint64_t grid64 = (warps + wpb - 1) / wpb;
  if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
  k_sparse<T><<<(unsigned)grid64, wpb * 32, smem, s>>>(p);
  if (info) {
    info->grid = (int)grid64;
    info->block = wpb * 32;
    info->traj_per_batch = 32;
    info->smem = smem;
  }
  return cudaGetLastError();
}

}  // namespace

size_t sparse_ws_words(int n, uint64_t num_tries) {
  // one transposed best-state array per warp; warps are indexed by gw = tl0/32 regardless of wpb
  return (size_t)((num_tries + 31) / 32) * (size_t)n;
}

template <>
cudaError_t launch_sparse<float>(const SparseParams<float> &p, cudaStream_t s, LaunchInfo *info) {
  return launch_impl<float>(p, s, info);
}
template <>
cudaError_t launch_sparse<double>(const SparseParams<double> &p, cudaStream_t s,
                                  LaunchInfo *info) {
  return launch_impl<double>(p, s, info);
}

}  // namespace osa
