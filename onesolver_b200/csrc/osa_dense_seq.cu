// osa_dense_seq.cu -- K1s: dense sequential-sweep annealing kernel for sm_100a.
//
// Replaces the reference's `class annealing` SYCL kernel
// (/root/reference/include/simulated_annealing/annealing.hpp:85-126) for the
// sequential-sweep mode.  One CTA anneals R trajectories from start to finish:
//
//   * local fields h[r][:] live in REGISTERS, column-sliced over the 256 threads
//     (thread t owns 16-byte column groups t, t+256, ...), so a Q row is fetched
//     from L2 once per CTA and applied to every trajectory that flipped that site;
//     rows travel through a per-thread cp.async ring in shared memory (see apply_rows)
//     and fp32 fields are updated with the packed fma.rn.f32x2 of sm_100;
//   * sites are processed in blocks of 32 (one bit-packed spin word):
//       P1 "decide": one warp per trajectory walks the block sequentially on a
//          32-value copy of h (one lane per site) and the 32x32 diagonal tile of Q,
//          jumping from accepted flip to accepted flip with ballot/ffs;
//       P2 "apply" : all threads stream the accepted rows (in site order) and do
//          h[r] += sign * Q[row] for the trajectories that accepted that site.
//     Applying rows in site order performs, per element of h, exactly the additions
//     of the plain sequential algorithm in the same order, so the result is
//     bit-identical to the scalar host replay (oracle/osa_oracle.c).
//   * dE = (1-2x_i) h_i, accept iff dE < tscale * (-ln u)  (equivalent to
//     dE < 0 || exp(-dE/beta) > u, annealing.hpp:106-108), u from Philox4x32-10
//     keyed by (seed, trajectory, sweep, site).
//   * best state: strict-improvement tracking (annealing.hpp:115-121) done lazily --
//     the packed state is copied only when the walk LEAVES a best state.
#include <cstdlib>

#include "osa_common.cuh"

namespace osa {

namespace {


// Local-field storage of one trajectory in one thread: N values.  For fp32 the values are kept as
// 64-bit register pairs so that the row update can use the packed FMA of sm_100
// (fma.rn.f32x2: two IEEE fp32 FMAs per instruction, bit-identical to two scalar fma.rn).
template <typename T, int N>
struct Field;

template <int N>
struct Field<double, N> {
  double v[N];
  __device__ __forceinline__ double get(int i) const { return v[i]; }
  __device__ __forceinline__ void set(int i, double x) { v[i] = x; }
  // v[i] += m * q[i]
  __device__ __forceinline__ void axpy(double m, const double (&q)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = det::fma(m, q[i], v[i]);
  }
};

template <int N>
struct Field<float, N> {
  static_assert(N % 2 == 0, "fp32 fields are stored as pairs");
  unsigned long long p[N / 2];
  __device__ __forceinline__ float get(int i) const {
    return __uint_as_float((i & 1) ? (uint32_t)(p[i >> 1] >> 32) : (uint32_t)p[i >> 1]);
  }
  __device__ __forceinline__ void set(int i, float x) {
    const unsigned long long b = __float_as_uint(x);
    p[i >> 1] = (i & 1) ? ((p[i >> 1] & 0xffffffffull) | (b << 32))
                        : ((p[i >> 1] & 0xffffffff00000000ull) | b);
  }
  __device__ __forceinline__ void axpy(float m, const float (&q)[N]) {
    const unsigned long long mb = __float_as_uint(m);
    const unsigned long long mm = (mb << 32) | mb;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const unsigned long long qq =
          ((unsigned long long)__float_as_uint(q[2 * i + 1]) << 32) | __float_as_uint(q[2 * i]);
      asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(qq), "l"(mm));
    }
  }
};

// TH threads per CTA; thread t owns NCH 16-byte column groups t, t+TH, ...
template <typename T, int NCH, int R, int TH>
struct Cfg {
  using VecT = typename Vec16<T>::type;
  static constexpr int V = Vec16<T>::V;
  static constexpr int WARPS = TH / 32;
  static constexpr int CPT = NCH * V;      // columns per thread
  static constexpr int CHW = TH * V;       // columns covered by one 16-byte group across the CTA
  static constexpr int MAXN = TH * CPT;
  static constexpr int NWP = MAXN / 32;    // state words (padded)
  static constexpr int TPW = (R + WARPS - 1) / WARPS;  // trajectories decided per warp
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// P2: stream the rows of block i0 whose site was accepted by at least one trajectory
// and apply them.  am[r] bit s: trajectory r flipped site i0+s; sm[r] bit s: the spin
// was 1 before the flip (sign -1).
//
// Loads: a per-thread cp.async (LDGSTS) ring in shared memory, K rows deep.  Every thread
// copies exactly the 16-byte pieces of a row that it will consume itself into its own ring
// slots, so the ring needs no barrier of any kind: cp.async.wait_group gives in-order
// completion, and the copies stay in flight while the thread runs the FMAs of earlier rows.
// (Alternatives measured in profiles/r01/microbench_l2_streaming*.log: LDG into registers is
// capped by the register file -- ptxas tracks all LDGs of a warp on one scoreboard, so loads
// cannot overlap the warp's own arithmetic -- and a TMA ring pays a per-stage mbarrier
// handshake; the cp.async ring streams at the full L2 rate including the read-back.)
// Rows are applied strictly in site order.
//
// Arithmetic: per row the trajectories are handled in groups of G: a group is skipped with one
// uniform branch when none of its members flipped the site, otherwise every member runs
// h += m * q with m in {-1, 0, +1} (m = 0 leaves h unchanged).  Returns the union mask.
template <typename T, int NCH, int R, int K, int TH, int G, int DBG = 0>
__device__ __forceinline__ uint32_t apply_rows(const T *__restrict__ qoff, unsigned char *ring,
                                               size_t ld, int i0, const uint32_t (&am)[R],
                                               const uint32_t (&sm)[R],
                                               Field<T, NCH * Vec16<T>::V> (&h)[R], int tid) {
  using C = Cfg<T, NCH, R, TH>;
  using VecT = typename C::VecT;
  constexpr int V = C::V, CHW = C::CHW;
  constexpr int ROW_VECS = NCH * TH;  // 16-byte pieces per row
  static_assert(R % G == 0, "R must be a multiple of the group size");

  uint32_t any = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) any |= am[r];
  if (any == 0) return 0;

  uint32_t pos[R], neg[R], grp[R / G];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    pos[r] = am[r] & ~sm[r];
    neg[r] = am[r] & sm[r];
  }
#pragma unroll
  for (int g = 0; g < R / G; ++g) {
    grp[g] = 0;
#pragma unroll
    for (int k = 0; k < G; ++k) grp[g] |= am[g * G + k];
  }

  const T *base = qoff + (size_t)i0 * ld + (size_t)tid * V;
  VecT *mine = reinterpret_cast<VecT *>(ring) + tid;  // slot s, piece c: mine[s*ROW_VECS + c*TH]
  const uint32_t mine_s = smem_addr(mine);
  uint32_t rem_issue = any, rem_apply = any;
  int slot_w = 0, slot_r = 0;

  auto issue_next = [&]() {
    if (rem_issue) {
      const int s = __ffs(rem_issue) - 1;
      rem_issue &= rem_issue - 1;
      const T *rp = base + (size_t)s * ld;
      const uint32_t dst = mine_s + (uint32_t)slot_w * (ROW_VECS * 16);
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + c * (TH * 16)),
                     "l"(rp + c * CHW)
                     : "memory");
      slot_w = (slot_w + 1 == K) ? 0 : slot_w + 1;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // one group per step, empty or not
  };

#pragma unroll 1
  for (int k = 0; k < K - 1; ++k) issue_next();
#pragma unroll 1
  while (rem_apply) {
    issue_next();
    asm volatile("cp.async.wait_group %0;" ::"n"(K - 1) : "memory");
    const int s = __ffs(rem_apply) - 1;
    rem_apply &= rem_apply - 1;
    const uint32_t bit = 1u << s;
    T qv[NCH * V];
#pragma unroll
    for (int c = 0; c < NCH; ++c) vec_unpack<T>(mine[slot_r * ROW_VECS + c * TH], &qv[c * V]);
    slot_r = (slot_r + 1 == K) ? 0 : slot_r + 1;
    if (DBG == 1) {  // timing experiment: touch the data, skip the arithmetic
      T acc = (T)0;
#pragma unroll
      for (int e = 0; e < NCH * V; ++e) acc += qv[e];
      if (acc == (T)123456789) h[0].set(0, acc);
      continue;
    }
#pragma unroll
    for (int g = 0; g < R / G; ++g) {
      if (grp[g] & bit) {
#pragma unroll
        for (int k = 0; k < G; ++k) {
          const int r = g * G + k;
          const T m = (pos[r] & bit) ? (T)1 : ((neg[r] & bit) ? (T)-1 : (T)0);
          h[r].axpy(m, qv);
        }
      }
    }
  }
  return any;
}

template <typename T, int NCH, int R, int K, int TH, int G, int DBG = 0>
__global__ void __launch_bounds__(TH, 1) k_dense_seq(const DenseParams<T> p) {
  extern __shared__ __align__(128) unsigned char s_ring[];  // K rows, thread-private slots
  using C = Cfg<T, NCH, R, TH>;
  using VecT = typename C::VecT;
  constexpr int V = C::V, CPT = C::CPT, CHW = C::CHW, NWP = C::NWP, TPW = C::TPW;
  constexpr int DS_WARPS = C::WARPS;
  constexpr int TILE_VECS = 32 * 32 / V;  // 16-byte pieces of the diagonal tile
  constexpr int TILE_PER_THREAD = (TILE_VECS + TH - 1) / TH;

  __shared__ __align__(16) T s_panel[R][32];
  __shared__ __align__(16) T s_tile[32][32];
  __shared__ uint32_t s_acc[R];
  __shared__ uint32_t s_sign[R];
  __shared__ uint32_t s_x[R][NWP];
  __shared__ uint32_t s_xb[R][NWP];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.n;
  const int nblk = (n + 31) >> 5;
  const uint64_t batch0 = (uint64_t)blockIdx.x * R;
  const uint64_t left = p.num_tries - batch0;
  const int nvalid = left < (uint64_t)R ? (int)left : R;

  // ---- local fields start at the diagonal (linear terms); host guarantees ld == NCH*CHW ----
  Field<T, CPT> h[R];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    T dv[V];
    const VecT v = *reinterpret_cast<const VecT *>(p.diag + c * CHW + tid * V);
    vec_unpack<T>(v, dv);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int e = 0; e < V; ++e) h[r].set(c * V + e, dv[e]);
  }

  // ---- initial spins (replaces random.bit(), annealing.hpp:90-92) ----
#pragma unroll
  for (int rr = 0; rr < TPW; ++rr) {
    const int r = warp + rr * DS_WARPS;
    if (r < R) {
      const bool tv = r < nvalid;
      const uint64_t traj = p.first_try + batch0 + (uint64_t)r;
      for (int k = lane; k < NWP; k += 32) {
        uint32_t word = 0;
        if (tv && k < nblk) {
          const U4 d = engine_draw(p.seed, traj, STREAM_INIT, (uint32_t)k >> 2, 0u);
          word = pick(d, (uint32_t)k & 3u);
          const int valid = n - k * 32;
          if (valid < 32) word &= (1u << valid) - 1u;
        }
        s_x[r][k] = word;
        s_xb[r][k] = word;
      }
    }
  }
  __syncthreads();

  unsigned long long cnt_rows = 0, cnt_init_rows = 0, cnt_acc = 0;

  // the 32x32 diagonal tile of the NEXT block is fetched while the current block's rows
  // stream, so its L2 latency never sits on the decision path
  VecT tile_next[TILE_PER_THREAD];
  auto fetch_tile = [&](int i0) {
#pragma unroll
    for (int t = 0; t < TILE_PER_THREAD; ++t) {
      const int qi = tid + t * TH;
      const int row = qi / (32 / V), cv = qi % (32 / V);
      if (qi < TILE_VECS)
        tile_next[t] = __ldg(reinterpret_cast<const VecT *>(p.qoff + (size_t)(i0 + row) * p.ld + i0 + cv * V));
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int t = 0; t < TILE_PER_THREAD; ++t) {
      const int qi = tid + t * TH;
      const int row = qi / (32 / V), cv = qi % (32 / V);
      if (qi < TILE_VECS) *reinterpret_cast<VecT *>(&s_tile[row][cv * V]) = tile_next[t];
    }
  };

  long long t_decide = 0, t_apply = 0, t_stage = 0, t_init = 0, t_mark = clock64();
  auto lap = [&](long long &acc) {
    const long long now = clock64();
    acc += now - t_mark;
    t_mark = now;
  };

  // ---- initial local fields: add the rows of the set spins, in site order ----
  fetch_tile(0);
  for (int b = 0; b < nblk; ++b) {
    uint32_t am[R], sm[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      am[r] = s_x[r][b];
      sm[r] = 0u;
    }
    const uint32_t any = apply_rows<T, NCH, R, K, TH, G, DBG>(p.qoff, s_ring, p.ld, b * 32, am, sm, h, tid);
    cnt_init_rows += (unsigned)__popc(any);
  }

  lap(t_init);

  // ---- annealing ----
  double erel[TPW], best[TPW];
  bool at_best[TPW];
#pragma unroll
  for (int rr = 0; rr < TPW; ++rr) {
    erel[rr] = 0.0;
    best[rr] = 0.0;
    at_best[rr] = true;
  }

  uint32_t step = 0;
  for (int iter = 0; iter < p.num_iter; ++iter) {
    const T ts = p.tscale[iter];
    for (int sw = 0; sw < p.sweeps_per_beta; ++sw, ++step) {
      for (int b = 0; b < nblk; ++b) {
        const int i0 = b * 32;
        // -- stage the (prefetched) diagonal tile and the 32-column panel of h --
        store_tile();
        {
          const int cb = i0 / CHW;
          const int rel = tid * V - (i0 % CHW);
          if (rel >= 0 && rel < 32) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              if (c == cb) {
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                  for (int e = 0; e < V; ++e) s_panel[r][rel + e] = h[r].get(c * V + e);
              }
            }
          }
        }
        __syncthreads();
        lap(t_stage);

        // -- P1: sequential decisions, one warp per trajectory --
#pragma unroll
        for (int rr = 0; rr < TPW; ++rr) {
          const int r = warp + rr * DS_WARPS;
          if (r < R) {
            const bool tv = r < nvalid;
            const uint64_t traj = p.first_try + batch0 + (uint64_t)r;
            const int site = i0 + lane;
            T hl = s_panel[r][lane];
            uint32_t xw = s_x[r][b];
            const U4 d = engine_draw(p.seed, traj, STREAM_SEQ, (uint32_t)site >> 2, step);
            const T theta = threshold<T>(ts, pick(d, (uint32_t)site & 3u));
            const bool lane_ok = tv && site < n;
            uint32_t acc = 0, sg = 0, from = 0xffffffffu;
            for (;;) {
              const uint32_t xl = (xw >> lane) & 1u;
              const T dEl = xl ? -hl : hl;
              const uint32_t bal = __ballot_sync(0xffffffffu, lane_ok && (dEl < theta)) & from;
              if (bal == 0) break;
              const int s = __ffs(bal) - 1;
              const T dEs = __shfl_sync(0xffffffffu, dEl, s);
              const uint32_t xbit = (xw >> s) & 1u;
              const T sgn = xbit ? (T)-1 : (T)1;
              hl = det::fma(sgn, s_tile[s][lane], hl);
              const double e = det::add(erel[rr], (double)dEs);
              erel[rr] = e;
              if (e < best[rr]) {
                best[rr] = e;
                at_best[rr] = true;
              } else if (at_best[rr]) {
                // leaving the best state: snapshot the state as it was BEFORE this flip
                for (int k = lane; k < nblk; k += 32) s_xb[r][k] = s_x[r][k];
                __syncwarp();
                if (lane == 0) s_xb[r][b] = xw;
                __syncwarp();
                at_best[rr] = false;
              }
              xw ^= (1u << s);
              acc |= (1u << s);
              sg |= (xbit << s);
              from = (s == 31) ? 0u : (0xffffffffu << (s + 1));
            }
            if (lane == 0) {
              s_x[r][b] = xw;
              s_acc[r] = acc;
              s_sign[r] = sg;
              cnt_acc += (unsigned)__popc(acc);
            }
          }
        }
        __syncthreads();
        lap(t_decide);

        // -- P2: stream accepted rows, update all local fields --
        {
          const int bn = (b + 1 == nblk) ? 0 : b + 1;  // next block (wraps into the next sweep)
          fetch_tile(bn * 32);
        }
        uint32_t am[R], sm[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          am[r] = s_acc[r];
          sm[r] = s_sign[r];
        }
        const uint32_t any = apply_rows<T, NCH, R, K, TH, G, DBG>(p.qoff, s_ring, p.ld, i0, am, sm, h, tid);
        cnt_rows += (unsigned)__popc(any);
        lap(t_apply);
      }
    }
  }
  __syncthreads();

  // ---- results ----
#pragma unroll
  for (int rr = 0; rr < TPW; ++rr) {
    const int r = warp + rr * DS_WARPS;
    if (r < R && r < nvalid) {
      if (at_best[rr]) {
        for (int k = lane; k < nblk; k += 32) s_xb[r][k] = s_x[r][k];
      }
      __syncwarp();
      const uint64_t tl = batch0 + (uint64_t)r;
      for (int k = lane; k < p.nw; k += 32) p.best_states[tl * (uint64_t)p.nw + k] = s_xb[r][k];
      if (lane == 0) p.best_rel[tl] = best[rr];
    }
  }
  if (lane == 0 && cnt_acc) atomicAdd(&p.counters->accepts, cnt_acc);
  if (tid == 0) {
    atomicAdd(&p.counters->row_fetches, cnt_rows);
    atomicAdd(&p.counters->init_row_fetches, cnt_init_rows);
    atomicAdd(&p.counters->cyc_decide, (unsigned long long)t_decide);
    atomicAdd(&p.counters->cyc_apply, (unsigned long long)t_apply);
    atomicAdd(&p.counters->cyc_stage, (unsigned long long)t_stage);
    atomicAdd(&p.counters->cyc_init, (unsigned long long)t_init);
  }
}

template <typename T, int NCH, int R, int K, int TH, int G, int DBG = 0>
cudaError_t launch_cfg(const DenseParams<T> &p, cudaStream_t s, LaunchInfo *info) {
  const uint64_t grid64 = (p.num_tries + R - 1) / R;
  if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
  const size_t smem = (size_t)K * NCH * TH * 16;  // the cp.async ring
  auto kern = k_dense_seq<T, NCH, R, K, TH, G, DBG>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  kern<<<(unsigned)grid64, TH, smem, s>>>(p);
  if (info) {
    info->grid = (int)grid64;
    info->block = TH;
    info->traj_per_batch = R;
    info->smem = smem;
  }
  return cudaGetLastError();
}

}  // namespace

// column capacity: 256 threads * NCH 16-byte groups, NCH <= 8
bool dense_seq_supported(int n, int elem_bytes) {
  if (n < 1) return false;
  return elem_bytes == 4 ? n <= 8192 : n <= 4096;
}

// ld contract: ld is a multiple of 1024 floats / 512 doubles (one 16-byte group per thread of a
// 256-thread CTA).  Per shape: R trajectories per CTA (h = R*NCH*16 bytes of registers per thread)
// and a ring of K rows (K*NCH*4 KiB of shared memory, <= 192 KiB).
template <>
cudaError_t launch_dense_seq<float>(const DenseParams<float> &p, cudaStream_t s, LaunchInfo *info) {
  if (p.ld % 1024 != 0) return cudaErrorInvalidValue;
  switch (p.ld / 1024) {
    case 1: return launch_cfg<float, 1, 16, 16, 256, 4>(p, s, info);
    case 2: return launch_cfg<float, 2, 16, 16, 256, 4>(p, s, info);
    case 3: return launch_cfg<float, 3, 12, 12, 256, 4>(p, s, info);
    case 4: {
      // tuning knob used by tools/probe.py (R*1000 + K*10 + G); the default is the measured best
      const char *e = getenv("OSA_DS_CFG");
      switch (e ? atoi(e) : 0) {
        case 8081: return launch_cfg<float, 4, 8, 8, 256, 1>(p, s, info);
        case 8122: return launch_cfg<float, 4, 8, 12, 256, 2>(p, s, info);
        case 8128: return launch_cfg<float, 4, 8, 12, 256, 8>(p, s, info);
        case 12122: return launch_cfg<float, 4, 12, 12, 256, 2>(p, s, info);
        case 81211: return launch_cfg<float, 4, 8, 12, 256, 1, 1>(p, s, info);  // loads only (timing)
        default: return launch_cfg<float, 4, 8, 12, 256, 1>(p, s, info);
      }
    }
    case 5: return launch_cfg<float, 5, 8, 9, 256, 2>(p, s, info);
    case 6: return launch_cfg<float, 6, 8, 8, 256, 2>(p, s, info);
    case 7: return launch_cfg<float, 7, 6, 6, 256, 2>(p, s, info);
    case 8: return launch_cfg<float, 8, 6, 6, 256, 2>(p, s, info);
    default: return cudaErrorInvalidValue;
  }
}

template <>
cudaError_t launch_dense_seq<double>(const DenseParams<double> &p, cudaStream_t s,
                                     LaunchInfo *info) {
  if (p.ld % 512 != 0) return cudaErrorInvalidValue;
  switch (p.ld / 512) {
    case 1: return launch_cfg<double, 1, 16, 16, 256, 4>(p, s, info);
    case 2: return launch_cfg<double, 2, 16, 16, 256, 4>(p, s, info);
    case 3: return launch_cfg<double, 3, 12, 12, 256, 4>(p, s, info);
    case 4: return launch_cfg<double, 4, 8, 12, 256, 2>(p, s, info);
    case 5: return launch_cfg<double, 5, 6, 9, 256, 2>(p, s, info);
    case 6: return launch_cfg<double, 6, 6, 8, 256, 2>(p, s, info);
    case 7: return launch_cfg<double, 7, 4, 6, 256, 2>(p, s, info);
    case 8: return launch_cfg<double, 8, 4, 6, 256, 2>(p, s, info);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace osa
