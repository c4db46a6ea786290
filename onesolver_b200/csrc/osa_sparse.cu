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
// Best-state tracking (annealing.hpp:115-121, strict improvement) without copying states:
// when a trajectory leaves its best state it starts a short per-lane LOG of the sites it
// flips; the best state is then (current state) XOR (logged flips).  A new best clears the
// log.  Only when an excursion outgrows the log (SP_LOG entries) is the best state written
// out to the transposed workspace XB (one warp-cooperative column copy + the logged toggles),
// after which logging stops until the next new best.  Short excursions -- the common case
// while the walk descends -- therefore cost one 2-byte store per flip instead of an O(N) copy.
#include <cstdlib>

#include "osa_common.cuh"

namespace osa {

namespace {

// CSR entries of one block of 32 consecutive sites, staged per warp in shared memory
constexpr int SP_LOG = 64;   // flips remembered per trajectory after leaving a best state
constexpr int SP_CAP = 512;  // entries per staging buffer (32 sites x degree 16); larger blocks
                             // fall back to direct global loads

template <typename T>
struct SpStage {
  int32_t col[2][SP_CAP];
  T val[2][SP_CAP];
};

__device__ __forceinline__ uint32_t sp_smem_addr(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

template <typename T>
__global__ void k_sparse(const SparseParams<T> p, int x_words_per_warp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const uint64_t gw = (uint64_t)blockIdx.x * wpb + warp;
  const uint64_t tl0 = gw * 32ull;
  if (tl0 >= p.num_tries) return;  // whole warp leaves; no block-level barrier below
  const uint64_t tl = tl0 + lane;
  const bool tv = tl < p.num_tries;
  const uint64_t traj = p.first_try + tl;
  const int n = p.n;

  // shared memory: [wpb] staging structs, then [wpb][x_words_per_warp] spin words
  SpStage<T> *stage = reinterpret_cast<SpStage<T> *>(smem_raw) + warp;
  uint32_t *X = reinterpret_cast<uint32_t *>(smem_raw + (size_t)wpb * sizeof(SpStage<T>)) +
                (size_t)warp * x_words_per_warp;
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
  unsigned long long trace = TRACE_OFFSET;  // flip trace, see osa_common.cuh
  bool at_best = true;   // the current state IS the best state
  bool mat = false;      // XB holds the best state explicitly (no log needed)
  int log_len = 0;       // flips since the best state was left (valid while !at_best && !mat)
  unsigned long long cnt_acc = 0;
  // per-warp flip log in the workspace behind the N words of XB: [SP_LOG][32] uint16
  uint16_t *LOG = reinterpret_cast<uint16_t *>(p.xbest_ws + p.log_base) + gw * (uint64_t)(SP_LOG * 32);

  // write the best states of the lanes in `who` to XB: their column of X (the state BEFORE the
  // current step's flips) with the logged flips undone
  auto materialize = [&](uint32_t who, bool use_log) {
    if (p.debug_flags & 1) return;
    for (int j = lane; j < n; j += 32) XB[j] = (XB[j] & ~who) | (X[j] & who);
    __syncwarp();
    if (use_log && ((who >> lane) & 1u)) {
      for (int e = 0; e < log_len; ++e) atomicXor(&XB[LOG[e * 32 + lane]], 1u << lane);
    }
    __syncwarp();
  };
  // bookkeeping of one accepted flip of `site` with energy change dE (call BEFORE X is updated);
  // returns true when this lane's log overflowed and its best state must be materialised
  auto track = [&](int site, T dE) -> bool {
    const double e = det::add(erel, (double)dE);
    erel = e;
    ++cnt_acc;
    if (e < best) {
      best = e;
      at_best = true;
      mat = false;
      log_len = 0;
      return false;
    }
    if (at_best) {  // leaving the best state: it equals the state before this flip
      at_best = false;
      mat = false;
      log_len = 0;
    }
    if (mat) return false;
    if (log_len < SP_LOG) {
      LOG[log_len * 32 + lane] = (uint16_t)site;
      ++log_len;
      return false;
    }
    return true;
  };

  if (p.mode == OSA_MODE_SEQUENTIAL_SWEEP) {
    // The CSR arrays (~10 bytes per neighbour) live in L2; reading them with dependent loads
    // would put 2-3 L2 round trips on every site.  Instead the entries of the NEXT block of 32
    // sites are copied into a per-warp shared-memory buffer with cp.async while the current
    // block is processed (double buffer), and rowptr/diag of the next block are prefetched into
    // registers, so the per-site critical path only touches shared memory.
    const int nblk = (n + 31) >> 5;
    auto block_range = [&](int b, int &e0, int &e1) {
      e0 = __ldg(p.rowptr + b * 32);
      e1 = __ldg(p.rowptr + min(b * 32 + 32, n));
    };
    auto prefetch_entries = [&](int b, int buf) {  // always commits one group
      int e0, e1;
      block_range(b, e0, e1);
      if (e1 - e0 <= SP_CAP) {
        for (int q = e0 + lane; q < e1; q += 32) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(
                           sp_smem_addr(&stage->col[buf][q - e0])),
                       "l"(p.col + q)
                       : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(
                           sp_smem_addr(&stage->val[buf][q - e0])),
                       "l"(p.val + q), "n"((int)sizeof(T))
                       : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto load_meta = [&](int b, int &rp_lo, int &rp_hi, T &dg) {
      const int i = b * 32 + lane;
      rp_lo = __ldg(p.rowptr + min(i, n));
      rp_hi = __ldg(p.rowptr + min(i + 1, n));
      dg = (i < n) ? __ldg(p.diag + i) : (T)0;
    };

    int rp_lo, rp_hi, nrp_lo, nrp_hi;
    T dg, ndg;
    prefetch_entries(0, 0);
    load_meta(0, rp_lo, rp_hi, dg);
    long long gblock = 0;
    uint32_t step = 0;
    for (int iter = 0; iter < p.num_iter; ++iter) {
      const T ts = p.tscale[iter];
      for (int sw = 0; sw < p.sweeps_per_beta; ++sw, ++step) {
        U4 d = U4{0, 0, 0, 0};
        for (int b = 0; b < nblk; ++b, ++gblock) {
          const int buf = (int)(gblock & 1);
          const int bn = (b + 1 == nblk) ? 0 : b + 1;  // next block (wraps into the next sweep)
          prefetch_entries(bn, buf ^ 1);
          load_meta(bn, nrp_lo, nrp_hi, ndg);
          asm volatile("cp.async.wait_group 1;" ::: "memory");  // this block's entries landed
          __syncwarp();
          const int e0 = __shfl_sync(0xffffffffu, rp_lo, 0);
          const int e1 = __shfl_sync(0xffffffffu, rp_hi, min(31, n - b * 32 - 1));
          const bool staged = (e1 - e0) <= SP_CAP;
          const int i_end = min(32, n - b * 32);
          // one Philox block serves four consecutive sites: their four thresholds (a chain of
          // ~40 dependent operations each) are computed together so the chains overlap, and sit
          // off the critical path of three of the four site steps
          uint32_t blk_acc = 0u;  // sites of this block that this lane's trajectory flipped
          const uint32_t indep = p.indep ? __ldg(p.indep + b) : 0u;
          // decision of site i (field hk in hand): accept test, bookkeeping, spin update
          auto decide = [&](int s, int i, T hk, T theta) {
            const uint32_t xiw = X[i];
            const T dE = ((xiw >> lane) & 1u) ? -hk : hk;
            const bool acc = tv && (dE < theta);
            const bool overflow = acc ? track(i, dE) : false;
            const uint32_t spill = __ballot_sync(0xffffffffu, overflow);
            if (spill) {
              materialize(spill, true);
              if (overflow) mat = true;
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, acc);
            if (acc) blk_acc |= 1u << s;
            if (bal) {
              __syncwarp();
              if (lane == 0) X[i] = xiw ^ bal;
              __syncwarp();
            }
          };
          for (int s4 = 0; s4 < i_end; s4 += 4) {
            d = engine_draw(p.seed, traj, STREAM_SEQ, (uint32_t)(b * 32 + s4) >> 2, step);
            const T th4[4] = {threshold<T>(ts, d.x), threshold<T>(ts, d.y), threshold<T>(ts, d.z),
                              threshold<T>(ts, d.w)};
            if (staged && e1 > e0 && s4 + 4 <= i_end && ((indep >> (s4 >> 2)) & 1u)) {
              // The four sites are pairwise non-adjacent (precomputed per block at problem
              // creation), so none of their flips changes the field of another: the four
              // neighbour gathers are independent chains and run side by side, the decisions
              // follow in site order.  Rows are padded to the longest of the four with +0.0
              // (h + 0 = h), the additions of a row stay in CSR order -- the same bits as the
              // one-site-at-a-time path below.
              const int32_t *sc = stage->col[buf] - e0;
              const T *sv = stage->val[buf] - e0;
              int pb[4], pe[4];
              T hk[4];
              int len = 0;
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                pb[k4] = __shfl_sync(0xffffffffu, rp_lo, s4 + k4);
                pe[k4] = __shfl_sync(0xffffffffu, rp_hi, s4 + k4);
                hk[k4] = __shfl_sync(0xffffffffu, dg, s4 + k4);
                len = max(len, pe[k4] - pb[k4]);
              }
              for (int t = 0; t < len; ++t) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                  const int q = pb[k4] + t;
                  const int qc = min(q, e1 - 1);
                  const T v = q < pe[k4] ? sv[qc] : (T)0;
                  if ((X[sc[qc]] >> lane) & 1u) hk[k4] = det::add(hk[k4], v);
                }
              }
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) decide(s4 + k4, b * 32 + s4 + k4, hk[k4], th4[k4]);
              continue;
            }
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const int s = s4 + k4;
              if (s >= i_end) break;
              const int i = b * 32 + s;
              const int pb = __shfl_sync(0xffffffffu, rp_lo, s);
              const int pe = __shfl_sync(0xffffffffu, rp_hi, s);
              T hk = __shfl_sync(0xffffffffu, dg, s);
              if (staged) {
                const int32_t *sc = stage->col[buf] - e0;
                const T *sv = stage->val[buf] - e0;
  #pragma unroll 4
                for (int q = pb; q < pe; ++q) {
                  if ((X[sc[q]] >> lane) & 1u) hk = det::add(hk, sv[q]);
                }
              } else {
                for (int q = pb; q < pe; ++q) {
                  const int c = __ldg(p.col + q);
                  const T v = __ldg(p.val + q);
                  if ((X[c] >> lane) & 1u) hk = det::add(hk, v);
                }
              }
              decide(s, i, hk, th4[k4]);
            }
          }
          if (blk_acc != 0u) trace = trace_step(trace, step, (uint32_t)b, blk_acc);
          rp_lo = nrp_lo;
          rp_hi = nrp_hi;
          dg = ndg;
          __syncwarp();  // everyone is done with stage buffer `buf` before it is refilled
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
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
      const bool overflow = acc ? track(k, dE) : false;
      const uint32_t spill = __ballot_sync(0xffffffffu, overflow);
      if (spill) {
        materialize(spill, true);
        if (overflow) mat = true;
      }
      __syncwarp();
      if (acc) {
        atomicXor(&X[k], 1u << lane);
        trace = trace_step(trace, (uint32_t)st, (uint32_t)k >> 5, 1u << (k & 31));
      }
      __syncwarp();
    }
  }

  // final best states: current state (at_best), current state with the log undone, or XB as is
  {
    const uint32_t cur = __ballot_sync(0xffffffffu, at_best);
    if (cur) materialize(cur, false);
    const uint32_t logged = __ballot_sync(0xffffffffu, !at_best && !mat);
    if (logged) materialize(logged, true);
  }
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
    if (p.trace_hash) p.trace_hash[tl] = trace;
  }
  // warp-reduce the accept counter
  for (int o = 16; o > 0; o >>= 1) cnt_acc += __shfl_down_sync(0xffffffffu, cnt_acc, o);
  if (lane == 0) {
    atomicAdd(&p.counters->accepts, cnt_acc);
    atomicAdd(&p.counters->row_fetches, cnt_acc);
  }
}

constexpr size_t kMaxSmem = 227 * 1024;

template <typename T>
size_t sp_per_warp_bytes(int n) {
  return sizeof(SpStage<T>) + (size_t)((n + 3) / 4 * 4) * sizeof(uint32_t);
}

int pick_wpb(size_t per_warp, uint64_t num_tries, int sm_count) {
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
  const size_t per_warp = sp_per_warp_bytes<T>(p.n);
  SparseParams<T> pp = p;
  pp.log_base = (size_t)((p.num_tries + 31) / 32) * (size_t)p.n;  // words; see sparse_ws_words
  pp.debug_flags = probe_env_int("OSA_SP_DEBUG");  // timing experiments (probe builds only)
  const int wpb = pick_wpb(per_warp, p.num_tries, sms);
  if (wpb < 1) return cudaErrorInvalidValue;
  const size_t smem = (size_t)wpb * per_warp;
  const int x_words = (p.n + 3) / 4 * 4;
  cudaError_t err =
      cudaFuncSetAttribute(k_sparse<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  const uint64_t warps = (p.num_tries + 31) / 32;
  const uint64_t grid64 = (warps + wpb - 1) / wpb;
  if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
  k_sparse<T><<<(unsigned)grid64, wpb * 32, smem, s>>>(pp, x_words);
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
  // per warp (gw = tl0/32, independent of the CTA shape): N words of transposed best state, and
  // behind all of those the flip logs, SP_LOG*32 uint16 = SP_LOG*16 words per warp
  const size_t warps = (size_t)((num_tries + 31) / 32);
  return warps * (size_t)n + warps * (size_t)(SP_LOG * 16);
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
