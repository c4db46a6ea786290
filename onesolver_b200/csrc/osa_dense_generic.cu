// osa_dense_generic.cu -- K1r: warp-per-trajectory dense annealing kernel (sm_100a).
//
// Covers the reference-faithful RANDOM-SITE mode (one attempt per (iter, sweep) at a
// site drawn from the trajectory's own stream, /root/reference/include/
// simulated_annealing/annealing.hpp:97-101) and is the catch-all for the sequential
// mode when N exceeds the register-resident kernel (osa_dense_seq.cu).
//
// One warp owns one trajectory: local field h[N] and the bit-packed state live in
// shared memory.  Attempts are evaluated 32 at a time (one per lane: site and
// threshold come from the counter-based stream, so they are known up front); the warp
// then walks from accepted flip to accepted flip (ballot/ffs), re-evaluating the
// remaining lanes after every accepted flip, which reproduces the sequential chain
// exactly.  The walk of a batch needs the fields at the <=32 candidate sites only: they are kept
// in registers (one per lane) and follow an accepted flip k through ONE gathered element
// Q[k, site_l] per lane, with the same fma the row add applies to that element.  The rows of all
// accepted flips of the batch are then added to h in one pass: a piece of h is read from shared
// memory once, takes the rows in flip order (so every element sees the same fma sequence as with
// one pass per row) and is written back once, and the 16-byte loads of consecutive rows follow
// each other without a drain in between.
#include <algorithm>
#include <cstdlib>

#include "osa_common.cuh"

#ifndef OSA_GEN_U
#define OSA_GEN_U 8  // 16-byte loads of a row in flight per lane (A/B: 16)
#endif

namespace osa {

namespace {

// PIPE: rows of at least one round (U * 32 pieces of 16 bytes) get the software-pipelined row add;
// the instantiation for shorter rows does not carry its register buffers (64 instead of 127
// registers: twice the warps per SM where shared memory does not bound the residency anyway)
// BATCH: the rows of all accepted flips of a batch are added in one pass over h (add_rows below);
// otherwise they are added flip by flip -- a warp then has up to two rounds of its row in flight,
// and the gathered element per lane and flip costs more than the saved shared-memory passes
// unless many warps share the SM (launch_impl chooses; measured: N = 4096 fp32 +33 %, N = 1024
// fp64 -28 %)
template <typename T, bool PIPE, bool BATCH>
__global__ void k_dense_generic(const DenseParams<T> p, int n_pad, int per_warp_bytes) {
  using VecT = typename Vec16<T>::type;
  constexpr int V = Vec16<T>::V;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const uint64_t tl = (uint64_t)blockIdx.x * wpb + warp;
  if (tl >= p.num_tries) return;  // whole warp leaves; no block-level barrier is used below
  const uint64_t traj = p.first_try + tl;
  const int n = p.n, nw = p.nw;

  unsigned char *mine = smem_raw + (size_t)warp * per_warp_bytes;
  T *h = reinterpret_cast<T *>(mine);
  uint32_t *x = reinterpret_cast<uint32_t *>(mine + (size_t)n_pad * sizeof(T));
  uint32_t *xb = x + nw;

  // initial spins (replaces random.bit(), annealing.hpp:90-92)
  for (int k = lane; k < nw; k += 32) {
    const U4 d = engine_draw(p.seed, traj, STREAM_INIT, (uint32_t)k >> 2, 0u);
    uint32_t word = pick(d, (uint32_t)k & 3u);
    const int valid = n - k * 32;
    if (valid < 32) word &= (1u << valid) - 1u;
    x[k] = word;
    xb[k] = word;
  }
  // fields_in: the initial field of this trajectory, built with shared row fetches by
  // k_dense_init_fields (osa_dense_init.cu); otherwise it starts at the diagonal and is built below
  {
    const T *src = p.fields_in ? p.fields_in + tl * p.ld : p.diag;
    for (int j = lane * V; j < n_pad; j += 32 * V)
      *reinterpret_cast<VecT *>(h + j) = *reinterpret_cast<const VecT *>(src + j);
  }
  __syncwarp();

  // h += sgn * Q[k,:].  The row comes from L2 in 16-byte pieces, U per lane and round; the pieces
  // of round r+1 are requested BEFORE the arithmetic of round r (two register buffers), so that a
  // warp always has U..2U loads in flight -- with 13 warps per SM the kernel lives on that.  The
  // loads are volatile asm: the compiler keeps them in program order (as __ldg they were sunk next
  // to their uses, three in flight at a time).  h is read and written with 16-byte shared-memory
  // accesses.
  auto add_row = [&](int k, T sgn) {
    const T *row = p.qoff + (size_t)k * p.ld;
    constexpr int U = OSA_GEN_U;
    constexpr int ROUND = U * 32 * V;  // elements per round and warp
    auto request = [&](int j0, uint4 (&q)[U]) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + (u * 32 + lane) * V;
        if (j < n_pad)
          asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(q[u].x), "=r"(q[u].y), "=r"(q[u].z), "=r"(q[u].w)
                       : "l"(row + j));
      }
    };
    auto apply = [&](int j0, const uint4 (&q)[U]) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + (u * 32 + lane) * V;
        if (j < n_pad) {
          T qv[V], hv[V];
          vec_unpack<T>(*reinterpret_cast<const VecT *>(&q[u]), qv);
          VecT *hp = reinterpret_cast<VecT *>(h + j);
          vec_unpack<T>(*hp, hv);
#pragma unroll
          for (int e = 0; e < V; ++e) hv[e] = det::fma(sgn, qv[e], hv[e]);
          *hp = vec_pack<T>(hv);
        }
      }
    };
    if constexpr (!PIPE) {  // short rows (N < 1024 fp32 / 512 fp64): a plain loop, nothing to pipeline
      for (int j = lane * V; j < n_pad; j += 32 * V) {
        T qv[V], hv[V];
        vec_unpack<T>(__ldg(reinterpret_cast<const VecT *>(row + j)), qv);
        VecT *hp = reinterpret_cast<VecT *>(h + j);
        vec_unpack<T>(*hp, hv);
#pragma unroll
        for (int e = 0; e < V; ++e) hv[e] = det::fma(sgn, qv[e], hv[e]);
        *hp = vec_pack<T>(hv);
      }
    } else {
      uint4 qa[U], qb[U];
      request(0, qa);
      for (int j0 = 0; j0 < n_pad; j0 += 2 * ROUND) {
        request(j0 + ROUND, qb);  // (no-op past the end of the row)
        apply(j0, qa);
        request(j0 + 2 * ROUND, qa);
        apply(j0 + ROUND, qb);
      }
    }
  };

  // h += sgn_0 * Q[k_0,:], then sgn_1 * Q[k_1,:], ... for the cnt flips held by lanes 0..cnt-1
  // (myk, mysgn), element by element in that order.  h is walked in pieces of U * 32 16-byte
  // vectors; a piece stays in registers while the rows go through it.  The row data comes from L2
  // in 16-byte loads, U per lane in flight: the register of a load is refilled with the same piece
  // of the NEXT row (or the next piece of the first row) as soon as its value is consumed, so the
  // stream never drains between rows.  The loads are volatile asm: the compiler keeps them in
  // program order (as __ldg they were sunk next to their uses, three in flight at a time).
  // (ptxas issues the U refills of an item together and waits for them at the next item.)
  // Measured and rejected, all bit-exact (profiles/r02/ab_random_site_batched_rows_variants.txt):
  // refills in groups of 2 / 4 loads (-9 %), U = 6 / 10 / 12 / 16 (-14 % ... -52 %), two alternating
  // buffers with an unpredicated path for whole pieces (55 instead of 181 warp instructions per item:
  // -3 %, rejected_random_site_two_buffers_r3s.patch), a 13th warp per SM (+0.1 %), and walking
  // batch b+1 while the rows of batch b stream (-16 %, rejected_random_site_walk_ahead_r3p.patch).
  // The kernel moves 15.3 TB/s through the L2 at 96 % hits (ncu, 66 % of the lts peak) whatever is
  // done to its instruction stream.
  auto add_rows = [&](int cnt, int myk, T mysgn) {
    constexpr int U = OSA_GEN_U;
    constexpr int ROUND = U * 32 * V;  // elements per piece and warp
    uint4 q[U];
    auto request = [&](const T *at, int j0, int u) {
      const int j = j0 + (u * 32 + lane) * V;
      if (j < n_pad)
        asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(q[u].x), "=r"(q[u].y), "=r"(q[u].z), "=r"(q[u].w)
                     : "l"(at + j));
    };
    T sgn = __shfl_sync(0xffffffffu, mysgn, 0);
    {
      const T *row = p.qoff + (size_t)__shfl_sync(0xffffffffu, myk, 0) * p.ld;
#pragma unroll
      for (int u = 0; u < U; ++u) request(row, 0, u);
    }
    for (int j0 = 0; j0 < n_pad; j0 += ROUND) {
      T hv[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + (u * 32 + lane) * V;
        if (j < n_pad) vec_unpack<T>(*reinterpret_cast<const VecT *>(h + j), hv[u]);
      }
      for (int i = 0; i < cnt; ++i) {
        const bool last = i + 1 == cnt;
        const int ni = last ? 0 : i + 1;
        const int nj0 = last ? j0 + ROUND : j0;
        const T *nrow = p.qoff + (size_t)__shfl_sync(0xffffffffu, myk, ni) * p.ld;
        const T nsgn = __shfl_sync(0xffffffffu, mysgn, ni);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          T qv[V];
          vec_unpack<T>(*reinterpret_cast<const VecT *>(&q[u]), qv);
          request(nrow, nj0, u);  // refill: same piece of the next row / next piece of row 0
          if (j0 + (u * 32 + lane) * V < n_pad) {
#pragma unroll
            for (int e = 0; e < V; ++e) hv[u][e] = det::fma(sgn, qv[e], hv[u][e]);
          }
        }
        sgn = nsgn;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + (u * 32 + lane) * V;
        if (j < n_pad) *reinterpret_cast<VecT *>(h + j) = vec_pack<T>(hv[u]);
      }
    }
  };

  // initial local field: diag + rows of the set spins, in site order
  unsigned long long cnt_init = 0, cnt_acc = 0;
  if constexpr (BATCH) {
    for (int w = 0; w < (p.fields_in ? 0 : nw); ++w) {
      uint32_t bits = x[w];
      int c = 0, myk = 0;
      for (; bits; bits &= bits - 1, ++c)
        if (lane == c) myk = w * 32 + __ffs(bits) - 1;
      if (c) add_rows(c, myk, (T)1);
      cnt_init += c;
    }
  } else {
    for (int i = 0; i < (p.fields_in ? 0 : n); ++i) {
      if ((x[i >> 5] >> (i & 31)) & 1u) {
        add_row(i, (T)1);
        ++cnt_init;
      }
    }
  }
  __syncwarp();

  double erel = 0.0, best = 0.0;
  bool at_best = true;
  unsigned long long trace = TRACE_OFFSET;  // flip trace, see osa_common.cuh

  // walk one batch of <=32 attempts (lane l: site_l, theta_l, active).  Sequential mode: the
  // batch is block `blk` of sweep `step`; random-site mode (blk < 0): lane l is attempt step + l.
  auto run_batch_row = [&](int site_l, T theta_l, bool active, uint32_t step, int blk) {
    uint32_t from = 0xffffffffu, accepted = 0u;
    for (;;) {
      __syncwarp();
      uint32_t xl = 0;
      T dEl = (T)0;
      bool al = false;
      if (active) {
        xl = (x[site_l >> 5] >> (site_l & 31)) & 1u;
        const T hl = h[site_l];
        dEl = xl ? -hl : hl;
        al = dEl < theta_l;
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, al) & from;
      if (bal == 0) break;
      const int s = __ffs(bal) - 1;
      const int k = __shfl_sync(0xffffffffu, site_l, s);
      const T dEs = __shfl_sync(0xffffffffu, dEl, s);
      const uint32_t xk = __shfl_sync(0xffffffffu, xl, s);
      const double e = det::add(erel, (double)dEs);
      erel = e;
      if (e < best) {
        best = e;
        at_best = true;
      } else if (at_best) {
        for (int kk = lane; kk < nw; kk += 32) xb[kk] = x[kk];  // state before this flip
        at_best = false;
      }
      __syncwarp();
      if (lane == 0) x[k >> 5] ^= (1u << (k & 31));
      add_row(k, xk ? (T)-1 : (T)1);
      ++cnt_acc;
      accepted |= 1u << s;
      if (blk < 0) trace = trace_step(trace, step + (uint32_t)s, (uint32_t)k >> 5, 1u << (k & 31));
      from = (s == 31) ? 0u : (0xffffffffu << (s + 1));
    }
    if (blk >= 0 && accepted != 0u) trace = trace_step(trace, step, (uint32_t)blk, accepted);
  };

  auto run_batch_rows = [&](int site_l, T theta_l, bool active, uint32_t step, int blk) {
    __syncwarp();
    uint32_t xl = 0;
    T hl = (T)0;
    if (active) {
      xl = (x[site_l >> 5] >> (site_l & 31)) & 1u;
      hl = h[site_l];
    }
    uint32_t from = 0xffffffffu, accepted = 0u;
    int cnt = 0, myk = 0;
    T mysgn = (T)0;
    for (;;) {
      const T dEl = xl ? -hl : hl;
      const uint32_t bal = __ballot_sync(0xffffffffu, active && dEl < theta_l) & from;
      if (bal == 0) break;
      const int s = __ffs(bal) - 1;
      const int k = __shfl_sync(0xffffffffu, site_l, s);
      const T dEs = __shfl_sync(0xffffffffu, dEl, s);
      const uint32_t xk = __shfl_sync(0xffffffffu, xl, s);
      const T sgn = xk ? (T)-1 : (T)1;
      from = (s == 31) ? 0u : (0xffffffffu << (s + 1));
      // the lanes still to come follow the flip at their own site: one element of row k each
      // (h[site_l] += sgn * Q[k, site_l], the fma of the row add), and the spin itself where a
      // later attempt of the batch draws site k again
      T ql = (T)0;
      const bool later = active && ((from >> lane) & 1u);
      if (later) ql = __ldg(p.qoff + (size_t)k * p.ld + site_l);
      const double e = det::add(erel, (double)dEs);
      erel = e;
      if (e < best) {
        best = e;
        at_best = true;
      } else if (at_best) {
        for (int kk = lane; kk < nw; kk += 32) xb[kk] = x[kk];  // state before this flip
        at_best = false;
      }
      __syncwarp();
      if (lane == 0) x[k >> 5] ^= (1u << (k & 31));
      __syncwarp();
      if (lane == cnt) {
        myk = k;
        mysgn = sgn;
      }
      ++cnt;
      ++cnt_acc;
      accepted |= 1u << s;
      if (blk < 0) trace = trace_step(trace, step + (uint32_t)s, (uint32_t)k >> 5, 1u << (k & 31));
      if (later) {
        hl = det::fma(sgn, ql, hl);
        if (site_l == k) xl ^= 1u;
      }
    }
    if (blk >= 0 && accepted != 0u) trace = trace_step(trace, step, (uint32_t)blk, accepted);
    if (cnt) add_rows(cnt, myk, mysgn);
  };

  auto run_batch = [&](int site_l, T theta_l, bool active, uint32_t step, int blk) {
    if constexpr (BATCH) run_batch_rows(site_l, theta_l, active, step, blk);
    else run_batch_row(site_l, theta_l, active, step, blk);
  };

  if (p.mode == OSA_MODE_SEQUENTIAL_SWEEP) {
    uint32_t step = 0;
    for (int iter = 0; iter < p.num_iter; ++iter) {
      const T ts = p.tscale[iter];
      for (int sw = 0; sw < p.sweeps_per_beta; ++sw, ++step) {
        for (int i0 = 0; i0 < n; i0 += 32) {
          const int site = i0 + lane;
          const U4 d = engine_draw(p.seed, traj, STREAM_SEQ, (uint32_t)site >> 2, step);
          const T theta = threshold<T>(ts, pick(d, (uint32_t)site & 3u));
          run_batch(site, theta, site < n, step, i0 >> 5);
        }
      }
    }
  } else {
    const uint64_t total = (uint64_t)p.num_iter * (uint64_t)p.sweeps_per_beta;
    for (uint64_t s0 = 0; s0 < total; s0 += 32) {
      const uint64_t st = s0 + lane;
      const bool active = st < total;
      int site = 0;
      T theta = (T)0;
      if (active) {
        const int iter = (int)(st / (uint64_t)p.sweeps_per_beta);
        const T ts = p.tscale[iter];
        const U4 d = engine_draw(p.seed, traj, STREAM_RND, 0u, (uint32_t)st);
        site = (int)__umulhi(d.x, (uint32_t)n);  // bit_index(), annealing.hpp:101
        theta = threshold<T>(ts, d.y);
      }
      run_batch(site, theta, active, (uint32_t)s0, -1);
    }
  }

  __syncwarp();
  if (at_best)
    for (int k = lane; k < nw; k += 32) xb[k] = x[k];
  __syncwarp();
  for (int k = lane; k < nw; k += 32) p.best_states[tl * (uint64_t)nw + k] = xb[k];
  if (lane == 0) {
    p.best_rel[tl] = best;
    if (p.trace_hash) p.trace_hash[tl] = trace;
    atomicAdd(&p.counters->accepts, cnt_acc);
    atomicAdd(&p.counters->row_fetches, cnt_acc);
    atomicAdd(&p.counters->init_row_fetches, cnt_init);
  }
}

constexpr size_t kMaxSmem = 227 * 1024;

template <typename T>
size_t per_warp_bytes(int n) {
  constexpr int V = Vec16<T>::V;
  const size_t n_pad = ((size_t)n + V - 1) / V * V;
  const size_t nw = ((size_t)n + 31) / 32;
  size_t b = n_pad * sizeof(T) + 2 * nw * sizeof(uint32_t);
  return (b + 15) / 16 * 16;
}

template <typename T>
cudaError_t launch_impl(const DenseParams<T> &p, cudaStream_t s, LaunchInfo *info) {
  constexpr int V = Vec16<T>::V;
  const size_t pw = per_warp_bytes<T>(p.n);
  if (pw > kMaxSmem) return cudaErrorInvalidValue;
  const int n_pad = (p.n + V - 1) / V * V;
  const bool pipe = n_pad >= OSA_GEN_U * 32 * V;  // at least one round of the pipelined row add
  // rows added batch by batch: where it measured faster -- fp32 rows of 16 to 20 KiB (N = 4096 ...
  // 5120: +33 % at 4096), where 11 or more warps per SM hide a walk that waits for one gathered
  // element per flip.  Longer rows leave fewer warps (N = 6000: -4 %, 8192: -20 %), and fp64 fields
  // lose whatever the row length (N = 2048: -25 %, 4096: -53 %;
  // profiles/r02/ab_random_site_batched_rows_variants.txt).  OSA_GEN_BATCH = 0 / 1 forces the choice
  // for rows of at least four rounds (result-preserving; tests run both forms).
  bool batch = sizeof(T) == 4 && n_pad >= 4 * OSA_GEN_U * 32 * V && n_pad <= 5 * OSA_GEN_U * 32 * V;
  if (const char *e = getenv("OSA_GEN_BATCH"))
    batch = atoi(e) != 0 && n_pad >= 4 * OSA_GEN_U * 32 * V;
  auto kern = batch ? k_dense_generic<T, true, true>
                    : pipe ? k_dense_generic<T, true, false> : k_dense_generic<T, false, false>;
  // CTA size: at most four warps (several small CTAs per SM beat one big one for tiny N), and of
  // the sizes 4 ... 1 the one that keeps most warps resident -- the warps of a CTA do not interact,
  // and with fields of 32 KiB per warp (N = 4096 fp64, 8192 fp32) a four-warp CTA left room for
  // one CTA = 4 warps per SM where 6 fit.  OSA_GEN_WPB forces a size (result-preserving).
  int wpb = (int)std::min<size_t>(4, kMaxSmem / pw);
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(pw * (size_t)wpb));
  if (err != cudaSuccess) return err;
  {
    int best_warps = 0, best_wpb = wpb;
    for (int w = wpb; w >= 1; --w) {
      int ctas = 0;
      err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, w * 32, pw * (size_t)w);
      if (err != cudaSuccess) return err;
      if (ctas * w > best_warps) best_warps = ctas * w, best_wpb = w;
    }
    wpb = best_wpb;
    if (const char *e = getenv("OSA_GEN_WPB")) {
      const int w = atoi(e);
      if (w >= 1 && w <= 4 && (size_t)w * pw <= kMaxSmem) wpb = w;
    }
  }
  const size_t smem = pw * (size_t)wpb;
  const uint64_t grid64 = (p.num_tries + wpb - 1) / wpb;
  if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
  kern<<<(unsigned)grid64, wpb * 32, smem, s>>>(p, n_pad, (int)pw);
  if (info) {
    info->grid = (int)grid64;
    info->block = wpb * 32;
    info->traj_per_batch = 1;
    info->smem = smem;
  }
  return cudaGetLastError();
}

}  // namespace

bool dense_generic_supported(int n, int elem_bytes) {
  if (n < 1) return false;
  return (elem_bytes == 4 ? per_warp_bytes<float>(n) : per_warp_bytes<double>(n)) <= kMaxSmem;
}

template <>
cudaError_t launch_dense_generic<float>(const DenseParams<float> &p, cudaStream_t s,
                                        LaunchInfo *info) {
  return launch_impl<float>(p, s, info);
}
template <>
cudaError_t launch_dense_generic<double>(const DenseParams<double> &p, cudaStream_t s,
                                         LaunchInfo *info) {
  return launch_impl<double>(p, s, info);
}

}  // namespace osa
