// osa_dense_seq.cu -- K1s: dense sequential-sweep annealing kernel for sm_100a.
//
// Replaces the reference's `class annealing` SYCL kernel
// (/root/reference/include/simulated_annealing/annealing.hpp:85-126) for the
// sequential-sweep mode.  One CTA anneals R trajectories from start to finish:
//
//   * local fields h[r][:] live in REGISTERS, column-sliced over the 256 threads
//     (thread t owns 16-byte column groups t, t+256, ...), so a Q row is fetched
//     from L2 once per CTA and applied to every trajectory that flipped that site;
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
#include "osa_common.cuh"

namespace osa {

namespace {

constexpr int DS_THREADS = 256;
constexpr int DS_WARPS = DS_THREADS / 32;

template <typename VecT>
__device__ __forceinline__ VecT ldg_stream(const VecT *p);
template <>
__device__ __forceinline__ float4 ldg_stream<float4>(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
template <>
__device__ __forceinline__ double2 ldg_stream<double2>(const double2 *p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
               : "=d"(v.x), "=d"(v.y)
               : "l"(p));
  return v;
}

template <typename T, int CPT, int R>
struct Cfg {
  using VecT = typename Vec16<T>::type;
  static constexpr int V = Vec16<T>::V;
  static constexpr int NCH = CPT / V;          // 16-byte column groups per thread
  static constexpr int CHW = DS_THREADS * V;   // columns covered by one group across the CTA
  static constexpr int MAXN = DS_THREADS * CPT;
  static constexpr int NWP = MAXN / 32;        // state words (padded)
  static constexpr int TPW = (R + DS_WARPS - 1) / DS_WARPS;  // trajectories decided per warp
};

// P2: stream the rows of block i0 whose site was accepted by at least one trajectory
// and apply them.  am[r] bit s: trajectory r flipped site i0+s; sm[r] bit s: the spin
// was 1 before the flip (sign -1).  Returns the union mask (rows fetched).
template <typename T, int CPT, int R>
__device__ __forceinline__ uint32_t apply_rows(const T *__restrict__ qoff, size_t ld, int i0,
                                               const uint32_t (&am)[R], const uint32_t (&sm)[R],
                                               T (&h)[R][CPT], int tid, int nch_act) {
  using C = Cfg<T, CPT, R>;
  using VecT = typename C::VecT;
  constexpr int V = C::V, NCH = C::NCH, CHW = C::CHW;

  uint32_t any = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) any |= am[r];
  if (any == 0) return 0;

  const T *base = qoff + (size_t)i0 * ld + (size_t)tid * V;
  uint32_t rem = any;
  auto next_row = [&]() -> int {
    if (rem == 0) return -1;
    const int s = __ffs(rem) - 1;
    rem &= rem - 1;
    return s;
  };
  auto load_row = [&](VecT (&q)[NCH], int s) {
    const T *rp = base + (size_t)s * ld;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      if (c < nch_act) q[c] = ldg_stream(reinterpret_cast<const VecT *>(rp + c * CHW));
  };
  auto apply_row = [&](const VecT (&q)[NCH], int s) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if ((am[r] >> s) & 1u) {
        const T sg = ((sm[r] >> s) & 1u) ? (T)-1 : (T)1;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (c < nch_act) {
            T qv[V];
            vec_unpack<T>(q[c], qv);
#pragma unroll
            for (int e = 0; e < V; ++e) h[r][c * V + e] = det::fma(sg, qv[e], h[r][c * V + e]);
          }
        }
      }
    }
  };

  // three rows in flight per thread, applied strictly in site order
  VecT q0[NCH], q1[NCH], q2[NCH];
  int s0 = next_row(), s1 = next_row(), s2 = next_row();
  if (s0 >= 0) load_row(q0, s0);
  if (s1 >= 0) load_row(q1, s1);
  if (s2 >= 0) load_row(q2, s2);
  for (;;) {
    if (s0 < 0) break;
    apply_row(q0, s0);
    s0 = next_row();
    if (s0 >= 0) load_row(q0, s0);
    if (s1 < 0) break;
    apply_row(q1, s1);
    s1 = next_row();
    if (s1 >= 0) load_row(q1, s1);
    if (s2 < 0) break;
    apply_row(q2, s2);
    s2 = next_row();
    if (s2 >= 0) load_row(q2, s2);
  }
  return any;
}

template <typename T, int CPT, int R>
__global__ void __launch_bounds__(DS_THREADS, 1) k_dense_seq(const DenseParams<T> p) {
  using C = Cfg<T, CPT, R>;
  using VecT = typename C::VecT;
  constexpr int V = C::V, NCH = C::NCH, CHW = C::CHW, NWP = C::NWP, TPW = C::TPW;

  __shared__ __align__(16) T s_panel[R][32];
  __shared__ __align__(16) T s_tile[32][32];
  __shared__ uint32_t s_acc[R];
  __shared__ uint32_t s_sign[R];
  __shared__ uint32_t s_x[R][NWP];
  __shared__ uint32_t s_xb[R][NWP];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.n;
  const int nblk = (n + 31) >> 5;
  const int nch_act = (int)(p.ld / CHW);  // host guarantees ld % CHW == 0 and ld <= MAXN
  const uint64_t batch0 = (uint64_t)blockIdx.x * R;
  const uint64_t left = p.num_tries - batch0;
  const int nvalid = left < (uint64_t)R ? (int)left : R;

  // ---- local fields start at the diagonal (linear terms) ----
  T h[R][CPT];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    T dv[V];
#pragma unroll
    for (int e = 0; e < V; ++e) dv[e] = (T)0;
    if (c < nch_act) {
      const VecT v = *reinterpret_cast<const VecT *>(p.diag + c * CHW + tid * V);
      vec_unpack<T>(v, dv);
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int e = 0; e < V; ++e) h[r][c * V + e] = dv[e];
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

  // ---- initial local fields: add the rows of the set spins, in site order ----
  for (int b = 0; b < nblk; ++b) {
    uint32_t am[R], sm[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      am[r] = s_x[r][b];
      sm[r] = 0u;
    }
    const uint32_t any = apply_rows<T, CPT, R>(p.qoff, p.ld, b * 32, am, sm, h, tid, nch_act);
    cnt_init_rows += (unsigned)__popc(any);
  }

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
        // -- stage the 32x32 diagonal tile and the 32-column panel of h --
        for (int q = tid; q < 32 * 32 / V; q += DS_THREADS) {
          const int row = q / (32 / V), cv = q % (32 / V);
          const VecT v = *reinterpret_cast<const VecT *>(p.qoff + (size_t)(i0 + row) * p.ld + i0 +
                                                         cv * V);
          *reinterpret_cast<VecT *>(&s_tile[row][cv * V]) = v;
        }
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
                  for (int e = 0; e < V; ++e) s_panel[r][rel + e] = h[r][c * V + e];
              }
            }
          }
        }
        __syncthreads();

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

        // -- P2: stream accepted rows, update all local fields --
        uint32_t am[R], sm[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          am[r] = s_acc[r];
          sm[r] = s_sign[r];
        }
        const uint32_t any = apply_rows<T, CPT, R>(p.qoff, p.ld, i0, am, sm, h, tid, nch_act);
        cnt_rows += (unsigned)__popc(any);
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
  }
}

template <typename T, int CPT, int R>
cudaError_t launch_cfg(const DenseParams<T> &p, cudaStream_t s, LaunchInfo *info) {
  const uint64_t grid64 = (p.num_tries + R - 1) / R;
  if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
  k_dense_seq<T, CPT, R><<<(unsigned)grid64, DS_THREADS, 0, s>>>(p);
  if (info) {
    info->grid = (int)grid64;
    info->block = DS_THREADS;
    info->traj_per_batch = R;
    info->smem = 0;
  }
  return cudaGetLastError();
}

}  // namespace

// column capacity per config: 256 threads * CPT
bool dense_seq_supported(int n, int elem_bytes) {
  if (n < 1) return false;
  return elem_bytes == 4 ? n <= 8192 : n <= 4096;
}

// ld contract: multiple of CHW (1024 floats / 512 doubles) and <= MAXN of the chosen config
template <>
cudaError_t launch_dense_seq<float>(const DenseParams<float> &p, cudaStream_t s, LaunchInfo *info) {
  if (p.ld % 1024 != 0) return cudaErrorInvalidValue;
  if (p.ld <= 1024) return launch_cfg<float, 4, 16>(p, s, info);
  if (p.ld <= 2048) return launch_cfg<float, 8, 16>(p, s, info);
  if (p.ld <= 4096) return launch_cfg<float, 16, 8>(p, s, info);
  if (p.ld <= 8192) return launch_cfg<float, 32, 4>(p, s, info);
  return cudaErrorInvalidValue;
}

template <>
cudaError_t launch_dense_seq<double>(const DenseParams<double> &p, cudaStream_t s,
                                     LaunchInfo *info) {
  if (p.ld % 512 != 0) return cudaErrorInvalidValue;
  if (p.ld <= 512) return launch_cfg<double, 2, 16>(p, s, info);
  if (p.ld <= 1024) return launch_cfg<double, 4, 16>(p, s, info);
  if (p.ld <= 2048) return launch_cfg<double, 8, 8>(p, s, info);
  if (p.ld <= 4096) return launch_cfg<double, 16, 4>(p, s, info);
  return cudaErrorInvalidValue;
}

}  // namespace osa
