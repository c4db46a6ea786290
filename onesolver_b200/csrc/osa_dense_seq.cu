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

#include "osa_dense_seq.cuh"

namespace osa {

using namespace dseq;

namespace {


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
  unsigned long long trace[TPW];  // flip trace, see osa_common.cuh
#pragma unroll
  for (int rr = 0; rr < TPW; ++rr) {
    erel[rr] = 0.0;
    best[rr] = 0.0;
    at_best[rr] = true;
    trace[rr] = TRACE_OFFSET;
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
            if (acc != 0u) trace[rr] = trace_step(trace[rr], step, (uint32_t)b, acc);
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
      if (lane == 0 && p.trace_hash) p.trace_hash[tl] = trace[rr];
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
// OSA_DS_WS=0 selects the single-role kernel of this file; the default is the warp-specialised
// variant (osa_dense_seq_ws.cu), which overlaps the decisions of block g+1 with the rows of block g.
static bool use_ws() {
  const char *e = getenv("OSA_DS_WS");
  return !(e && e[0] == '0');
}
// ... and of those, the free-running variant (osa_dense_seq_ws2.cu) where it is the faster one:
// its flip-to-flip decide walk wins in the cold sweeps of a schedule and loses when most sites
// flip, and it has no CTA barrier per block.  Measured per shape on a bench-like schedule
// (profiles/r02/flow_shapes_probe.txt, flow_check4_parity_and_probes.txt): the shapes with
// R <= 12 trajectories per CTA gain 6 % (fp32 N = 5120, fp64 N = 1536 / 2048) to 29 % (fp64
// N = 4096, R = 4), fp32 N = 4096 gains 2.6 % on the bench; the R = 16 shapes (fp32 N <= 2048,
// fp64 N <= 1024) lose 5-16 % and fp32 N = 3072 loses 3 %, so they keep the lock-step kernel.
// OSA_WS_FLOW=0/1 forces one of the two for A/B runs; both give the same results bit for bit.
static bool use_flow(size_t ld, int elem_bytes) {
  const char *e = getenv("OSA_WS_FLOW");
  if (e) return e[0] != '0';
  return elem_bytes == 4 ? ld >= 4096 : ld >= 1536;
}
template <typename T>
static cudaError_t launch_ws_any(const DenseParams<T> &p, cudaStream_t s, LaunchInfo *info) {
  return use_flow(p.ld, (int)sizeof(T)) ? launch_dense_seq_flow<T>(p, s, info)
                                        : launch_dense_seq_ws<T>(p, s, info);
}

template <>
cudaError_t launch_dense_seq<float>(const DenseParams<float> &p, cudaStream_t s, LaunchInfo *info) {
  if (p.ld % 1024 != 0) return cudaErrorInvalidValue;
  if (use_ws() && !getenv("OSA_DS_CFG")) return launch_ws_any<float>(p, s, info);
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
#ifdef OSA_PROBE  // loads only (timing; results are meaningless): probe builds only
        case 81211: return launch_cfg<float, 4, 8, 12, 256, 1, 1>(p, s, info);
#endif
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
  if (use_ws()) return launch_ws_any<double>(p, s, info);
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
