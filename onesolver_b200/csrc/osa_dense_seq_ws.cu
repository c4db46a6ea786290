// osa_dense_seq_ws.cu -- K1s/ws: the dense sequential-sweep kernel with warp-specialised
// decide/apply overlap (same arithmetic and results as osa_dense_seq.cu, bit for bit).
//
// In osa_dense_seq.cu every block of 32 sites runs stage -> P1 decide -> P2 apply back to back,
// and the decide phase (a latency chain of ballot/shfl/fma steps on one warp per trajectory)
// leaves the row streaming idle for ~20% of the time.  Here the CTA has two roles:
//
//   apply warps  (8 warps, 256 threads, ~224 registers each): own the local fields h[r][:] in
//                registers and stream/apply the accepted rows of block g  (P2 of osa_dense_seq.cu);
//   decide warps (4 warps, 128 threads, 56 registers each): run the decisions of block g+1
//                WHILE block g is being applied.
//
// Deciding block g+1 needs h on its 32 columns *after* the rows of block g have been applied,
// which the apply warps have not finished yet.  The decide warps therefore rebuild those 32
// values themselves: they start from a snapshot of the columns taken after block g-1 was applied
// and add the accepted rows of block g restricted to the 32 columns (a 32x32 tile of Q), in site
// order -- the same fma sequence the apply warps perform on those columns, hence the same bits.
// One CTA-wide barrier per block hands over {accept masks of block g+1} and {snapshot of the
// columns of block g+2}; both are double buffered.  Registers are moved between the roles with
// setmaxnreg (the kernel is launched with 384 threads, 168 registers each).
#include <cstdlib>

#include "osa_dense_seq.cuh"

namespace osa {

using namespace dseq;

namespace {

constexpr int WS_APPLY_THREADS = 256;

// Register budgets of the two roles for DW decide warps.  The kernel is launched with
// LAUNCH = 65536 / threads registers per thread (multiple of 8); the decide warps give up
// LAUNCH - DECIDE each, and setmaxnreg.inc can only draw from what the CTA's own warps released,
// so the apply warps get exactly LAUNCH + (LAUNCH - DECIDE) * decide_threads / apply_threads:
//   DW = 4: 384 threads, 168 -> apply 224 / decide 56   (shapes with >= 192 field registers)
//   DW = 8: 512 threads, 128 -> apply 200 / decide 56   (small N: the decisions are the bottleneck)
template <int DW>
struct WsRegs {
  static constexpr int THREADS = WS_APPLY_THREADS + DW * 32;
  static constexpr int LAUNCH = (65536 / THREADS) / 8 * 8;
  static constexpr int DECIDE = 56;
  static constexpr int APPLY_RAW = LAUNCH + (LAUNCH - DECIDE) * (DW * 32) / WS_APPLY_THREADS;
  static constexpr int APPLY = (APPLY_RAW > 232 ? 232 : APPLY_RAW) / 8 * 8;
};
static_assert(WsRegs<4>::APPLY == 224 && WsRegs<8>::APPLY == 200, "register split");

// SM clock, read only after `dep` is available.  A clock read placed right after bar.sync can
// execute before the barrier has released the warp; making it depend on a shared-memory load
// issued after the barrier gives the release time (used by the role timers below).
__device__ __forceinline__ long long clock_after(uint32_t dep) {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(dep) : "memory");
  return t;
}

template <int THREADS>
__device__ __forceinline__ void bar_named(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(THREADS) : "memory");
}

// PT: resumable launch for parallel tempering (start spins from memory, per-trajectory threshold
// scale, final spins written back, running sweep counter); a separate instantiation so that the
// plain annealing kernel carries none of it.
template <typename T, int NCH, int R, int K, int G, bool PT, int DW>
__global__ void __launch_bounds__(WsRegs<DW>::THREADS, 1) k_dense_seq_ws(const DenseParams<T> p) {
  constexpr int TH = WS_APPLY_THREADS;
  constexpr int WS_DECIDE_WARPS = DW;
  constexpr int WS_THREADS = WsRegs<DW>::THREADS;
  auto bar_cta = [] { bar_named<WS_THREADS>(1); };
  auto bar_decide = [] { bar_named<DW * 32>(2); };
  using C = Cfg<T, NCH, R, TH>;
  using VecT = typename C::VecT;
  constexpr int V = C::V, CPT = C::CPT, CHW = C::CHW, NWP = C::NWP;
  constexpr int TPD = (R + WS_DECIDE_WARPS - 1) / WS_DECIDE_WARPS;  // trajectories per decide warp
  constexpr int TILE_VECS = 32 * 32 / V;
  // trajectories decided together by one decide warp (register budget of the decide role: 56)
  constexpr int IL = TPD < (sizeof(T) == 4 ? 3 : 2) ? TPD : (sizeof(T) == 4 ? 3 : 2);
  constexpr bool ALLV = (R % WS_DECIDE_WARPS == 0) && (TPD % IL == 0);  // every slot exists

  extern __shared__ __align__(128) unsigned char s_ring[];  // apply warps: K rows, private slots
  __shared__ __align__(16) T s_snap[2][R][32];   // columns of block g (parity g&1) after block g-2
  __shared__ __align__(16) T s_tile_d[32][32];   // diagonal tile of the block being decided
  __shared__ __align__(16) T s_tile_x[32][32];   // rows of the previous block x columns of this one
  __shared__ uint32_t s_acc[2][R];               // accept / sign masks of block g (parity g&1)
  __shared__ uint32_t s_sign[2][R];
  // per-trajectory walk state of the decide warps (indexed by a run-time trajectory number, so it
  // lives here rather than in a register array that the compiler would demote to local memory)
  __shared__ double s_erel[R], s_best[R];
  __shared__ T s_ts[R];  // per-trajectory threshold scale (only with p.tscale_traj)
  __shared__ uint32_t s_atbest[R];
  __shared__ uint32_t s_x[R][NWP];
  __shared__ uint32_t s_xb[R][NWP];

  const int tid = threadIdx.x;
  const int n = p.n;
  const int nblk = (n + 31) >> 5;
  const uint64_t batch0 = (uint64_t)blockIdx.x * R;
  const uint64_t left = p.num_tries - batch0;
  const int nvalid = left < (uint64_t)R ? (int)left : R;
  const long long total_blocks = (long long)p.num_iter * p.sweeps_per_beta * nblk;

  if (tid < WS_APPLY_THREADS) {
    // =============================== APPLY ROLE ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WsRegs<DW>::APPLY));
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
    // snapshot of the 32 columns of block b into s_snap[par]
    auto snapshot = [&](int b, int par) {
      const int i0 = b * 32;
      const int cb = i0 / CHW;
      const int rel = tid * V - (i0 % CHW);
      if (rel >= 0 && rel < 32) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (c == cb) {
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
              for (int e = 0; e < V; ++e) s_snap[par][r][rel + e] = h[r].get(c * V + e);
          }
        }
      }
    };
    unsigned long long cnt_rows = 0, cnt_init_rows = 0;
    // role timers (osa_stats): cyc_init = initial fields, cyc_apply = streaming/applying rows
    // from the release of barrier B(g) on, cyc_stage = the rest of the main loop (snapshots and
    // waiting for the decide warps)
    long long t_apply = 0, t_init = clock64();

    bar_cta();  // #0: initial spins are in s_x
    for (int b = 0; b < nblk; ++b) {
      uint32_t am[R], sm[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        am[r] = s_x[r][b];
        sm[r] = 0u;
      }
      const uint32_t any = apply_rows<T, NCH, R, K, TH, G>(p.qoff, s_ring, p.ld, b * 32, am, sm, h, tid);
      cnt_init_rows += (unsigned)__popc(any);
    }
    snapshot(0, 0);
    snapshot(nblk > 1 ? 1 : 0, 1);
    t_init = clock64() - t_init;
    bar_cta();  // B(-1): snapshots of blocks 0 and 1 are ready
    bar_cta();  // B(0) : masks of block 0 are ready
    const long long t_main = clock64();
    int b = 0;
    for (long long g = 0; g < total_blocks; ++g) {
      const int par = (int)(g & 1);
      uint32_t am[R], sm[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        am[r] = s_acc[par][r];
        sm[r] = s_sign[par][r];
      }
      const long long t0 = clock_after(am[0]);
      if (p.debug_flags & 1) am[0] = 0u;
      const uint32_t any = (p.debug_flags & 1) ? 0u :
          apply_rows<T, NCH, R, K, TH, G>(p.qoff, s_ring, p.ld, b * 32, am, sm, h, tid);
      cnt_rows += (unsigned)__popc(any);
      t_apply += clock64() - t0;
      if (g + 1 < total_blocks) {
        // columns of block g+2 as they are now (after block g): consumed by the decision of g+2
        int b2 = b + 2;
        if (b2 >= nblk) b2 -= nblk;
        if (b2 >= nblk) b2 -= nblk;  // nblk == 1
        snapshot(b2, par);
        bar_cta();  // B(g+1)
      }
      b = (b + 1 == nblk) ? 0 : b + 1;
    }
    if (tid == 0) {
      atomicAdd(&p.counters->row_fetches, cnt_rows);
      atomicAdd(&p.counters->init_row_fetches, cnt_init_rows);
      atomicAdd(&p.counters->cyc_apply, (unsigned long long)t_apply);
      atomicAdd(&p.counters->cyc_stage, (unsigned long long)(clock64() - t_main - t_apply));
      if (!(p.debug_flags & 8)) atomicAdd(&p.counters->cyc_init, (unsigned long long)t_init);
    }
  } else {
    // =============================== DECIDE ROLE ===============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WsRegs<DW>::DECIDE));
    const int dt = tid - WS_APPLY_THREADS;  // 0..127
    const int lane = dt & 31, dwarp = dt >> 5;

    // initial spins (replaces random.bit(), annealing.hpp:90-92)
#pragma unroll 1
    for (int rr = 0; rr < TPD; ++rr) {
      const int r = dwarp + rr * WS_DECIDE_WARPS;
      if (r < R) {
        const bool tv = r < nvalid;
        const uint64_t traj = p.first_try + batch0 + (uint64_t)r;
        for (int k = lane; k < NWP; k += 32) {
          uint32_t word = 0;
          if (tv && k < nblk) {
            if (PT && p.init_states) {  // resume: the spins a previous launch left in final_states
              word = p.init_states[(batch0 + (uint64_t)r) * (uint64_t)p.nw + k];
            } else {
              const U4 d = engine_draw(p.seed, traj, STREAM_INIT, (uint32_t)k >> 2, 0u);
              word = pick(d, (uint32_t)k & 3u);
            }
            const int valid = n - k * 32;
            if (valid < 32) word &= (1u << valid) - 1u;
          }
          s_x[r][k] = word;
          s_xb[r][k] = word;
        }
        if (PT && lane == 0)
          s_ts[r] = (p.tscale_traj && tv) ? p.tscale_traj[batch0 + (uint64_t)r] : (T)0;
      }
    }
    for (int r = dt; r < R; r += WS_DECIDE_WARPS * 32) {
      s_erel[r] = 0.0;
      s_best[r] = 0.0;
      s_atbest[r] = 1u;
    }
    unsigned long long cnt_acc = 0, n_walk = 0;
    long long t_decide = 0;

    bar_cta();  // #0
    bar_cta();  // B(-1): snapshots 0 and 1 ready

    // decision of block (iter, sw, b) = global block g.  The trajectories of this warp are
    // decided IL at a time, interleaved: every step below is written branch-free over the IL
    // slots (warp-uniform selects instead of branches), so the latency chains of the slots
    // (ballot -> ffs -> tile row from shared memory -> fma) overlap.  The walk itself only
    // carries what the next decision needs; the energy bookkeeping of annealing.hpp:115-121
    // (running energy, strict-< best, state at the best) is done after the walk from the
    // per-lane dE values, in the same site order, so the sums see the same fp64 additions.
    auto decide = [&](long long g, int b, uint32_t step, T ts) {
      const int par = (int)(g & 1);
      // cyc_decide: from the release of barrier B(g-1) to the end of the decision of block g
      const long long t0 = clock_after(*(volatile uint32_t *)&s_atbest[0]);
      if (p.debug_flags & 2) {  // timing experiment: masks of density 3/16, no decisions
        if (dt < R) {
          const U4 d = engine_draw(p.seed, batch0 + (uint64_t)dt, STREAM_SEQ, (uint32_t)g, step);
          s_acc[par][dt] = d.x & d.y & (d.z | d.w);
          s_sign[par][dt] = d.w;
        }
        t_decide += clock64() - t0;
        return;
      }
      const int i0 = b * 32;
      const int bp = (b == 0) ? nblk - 1 : b - 1;  // previous block (cyclic)
      // stage the two tiles: diagonal tile of block b, and rows of block bp x columns of block b
      for (int qi = dt; qi < TILE_VECS; qi += WS_DECIDE_WARPS * 32) {
        const int row = qi / (32 / V), cv = qi % (32 / V);
        *reinterpret_cast<VecT *>(&s_tile_d[row][cv * V]) =
            __ldg(reinterpret_cast<const VecT *>(p.qoff + (size_t)(i0 + row) * p.ld + i0 + cv * V));
        if (g > 0)
          *reinterpret_cast<VecT *>(&s_tile_x[row][cv * V]) = __ldg(reinterpret_cast<const VecT *>(
              p.qoff + (size_t)(bp * 32 + row) * p.ld + i0 + cv * V));
      }
      bar_decide();
      const int site = i0 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < TPD; c0 += IL) {
        int rj[IL];       // trajectory slot of the CTA
        bool vj[IL];      // slot exists in this CTA (R not a multiple of the interleave)
        bool okj[IL];     // this lane may flip: trajectory exists in the batch and site < n
        T hl[IL], theta[IL], myd[IL];
        uint32_t xw[IL], acc[IL], from[IL];
#pragma unroll
        for (int j = 0; j < IL; ++j) {
          const int r = dwarp + (c0 + j) * WS_DECIDE_WARPS;
          vj[j] = ALLV || ((c0 + j < TPD) && (r < R));
          rj[j] = vj[j] ? r : dwarp;
          okj[j] = vj[j] && (rj[j] < nvalid) && (site < n);
          hl[j] = s_snap[par][rj[j]][lane];
          xw[j] = s_x[rj[j]][b];
          acc[j] = 0u;
          from[j] = 0xffffffffu;
          myd[j] = (T)0;
        }
        if (g > 0) {
          // bring the snapshots up to date: rows of block g-1 that the trajectory flipped
          uint32_t pa[IL], ps[IL], left = 0u;
#pragma unroll
          for (int j = 0; j < IL; ++j) {
            pa[j] = vj[j] ? s_acc[par ^ 1][rj[j]] : 0u;
            ps[j] = s_sign[par ^ 1][rj[j]];
            left |= pa[j];
          }
          while (left) {
            left = 0u;
#pragma unroll
            for (int j = 0; j < IL; ++j) {
              const bool on = pa[j] != 0u;
              const int s = on ? __ffs(pa[j]) - 1 : 0;
              pa[j] &= pa[j] - 1u;  // 0 stays 0
              const T sgn = ((ps[j] >> s) & 1u) ? (T)-1 : (T)1;
              const T up = det::fma(sgn, s_tile_x[s][lane], hl[j]);
              hl[j] = on ? up : hl[j];
              left |= pa[j];
            }
          }
        }
#pragma unroll
        for (int j = 0; j < IL; ++j) {
          const uint64_t traj = p.first_try + batch0 + (uint64_t)rj[j];
          const U4 d = engine_draw(p.seed, traj, STREAM_SEQ, (uint32_t)site >> 2, step);
          theta[j] = threshold<T>(PT ? s_ts[rj[j]] : ts, pick(d, (uint32_t)site & 3u));
        }
        // the walk: from accepted flip to accepted flip
        for (;;) {
          ++n_walk;
          uint32_t bal[IL], anyb = 0u;
          T dE[IL];
#pragma unroll
          for (int j = 0; j < IL; ++j) {
            const uint32_t xl = (xw[j] >> lane) & 1u;
            dE[j] = xl ? -hl[j] : hl[j];
            bal[j] = __ballot_sync(0xffffffffu, okj[j] && (dE[j] < theta[j])) & from[j];
            anyb |= bal[j];
          }
          if (anyb == 0u) break;
#pragma unroll
          for (int j = 0; j < IL; ++j) {
            const bool on = bal[j] != 0u;
            const int s = on ? __ffs(bal[j]) - 1 : 0;
            const uint32_t bit = on ? (1u << s) : 0u;
            const uint32_t xbit = (xw[j] >> s) & 1u;
            const T sgn = xbit ? (T)-1 : (T)1;
            const T up = det::fma(sgn, s_tile_d[s][lane], hl[j]);
            hl[j] = on ? up : hl[j];
            myd[j] = (on && lane == s) ? dE[j] : myd[j];
            xw[j] ^= bit;
            acc[j] |= bit;
            from[j] = on ? (0xfffffffeu << s) : from[j];
          }
        }
        // energies along the walk (all lanes run the same additions): erel += dE in site order,
        // kb = site of the last flip that set a new best
        double erel[IL], best[IL];
        int kb[IL];
        uint32_t rem[IL], left = 0u;
#pragma unroll
        for (int j = 0; j < IL; ++j) {
          erel[j] = s_erel[rj[j]];
          best[j] = s_best[rj[j]];
          kb[j] = -1;
          rem[j] = acc[j];
          left |= rem[j];
        }
        while (left) {
          left = 0u;
#pragma unroll
          for (int j = 0; j < IL; ++j) {
            const bool on = rem[j] != 0u;
            const int s = on ? __ffs(rem[j]) - 1 : 0;
            rem[j] &= rem[j] - 1u;
            const T dEs = __shfl_sync(0xffffffffu, myd[j], s);
            const double e = det::add(erel[j], (double)dEs);
            erel[j] = on ? e : erel[j];
            const bool nb = on && (e < best[j]);
            best[j] = nb ? e : best[j];
            kb[j] = nb ? s : kb[j];
            left |= rem[j];
          }
        }
        // state at the best (annealing.hpp:115-121, kept lazily: s_xb is only written when the
        // walk has left the best state by the end of the block)
#pragma unroll
        for (int j = 0; j < IL; ++j) {
          if (vj[j]) {
            const int r = rj[j];
            // a site flips at most once per block: the word before the block and the spins
            // that were 1 before their flip (sign -1) follow from the final word
            const uint32_t xw0 = xw[j] ^ acc[j];
            const uint32_t sg = acc[j] & xw0;
            bool at_best = s_atbest[r] != 0u;
            uint32_t wb = xw0;
            bool copy = false;
            if (kb[j] >= 0) {
              const uint32_t le = (2u << kb[j]) - 1u;  // flips up to and including the best one
              at_best = (acc[j] & ~le) == 0u;
              copy = !at_best;
              wb = xw0 ^ (acc[j] & le);
            } else if (at_best && acc[j] != 0u) {
              copy = true;  // the first flip of the block left the best state
              at_best = false;
            }
            if (copy) {
              for (int k = lane; k < nblk; k += 32) s_xb[r][k] = s_x[r][k];
              __syncwarp();
              if (lane == 0) s_xb[r][b] = wb;
            }
            __syncwarp();
            if (lane == 0) {
              s_erel[r] = erel[j];
              s_best[r] = best[j];
              s_atbest[r] = at_best ? 1u : 0u;
              s_x[r][b] = xw[j];
              s_acc[par][r] = acc[j];
              s_sign[par][r] = sg;
              cnt_acc += (unsigned)__popc(acc[j]);
            }
          }
        }
      }
      bar_decide();  // the tiles may be overwritten by the next call
      t_decide += clock64() - t0;
    };

    long long g = 0;
    uint32_t step = PT ? p.step_base : 0u;
    for (int iter = 0; iter < p.num_iter; ++iter) {
      const T ts = PT ? (T)0 : p.tscale[iter];
      for (int sw = 0; sw < p.sweeps_per_beta; ++sw, ++step) {
        for (int b = 0; b < nblk; ++b, ++g) {
          decide(g, b, step, ts);
          bar_cta();  // B(g): masks of block g ready (g = 0: the apply warps have been waiting)
        }
      }
    }
    // ---- results ----
#pragma unroll 1
    for (int rr = 0; rr < TPD; ++rr) {
      const int r = dwarp + rr * WS_DECIDE_WARPS;
      if (r < R && r < nvalid) {
        if (s_atbest[r]) {
          for (int k = lane; k < nblk; k += 32) s_xb[r][k] = s_x[r][k];
        }
        __syncwarp();
        const uint64_t tl = batch0 + (uint64_t)r;
        for (int k = lane; k < p.nw; k += 32) p.best_states[tl * (uint64_t)p.nw + k] = s_xb[r][k];
        if (PT && p.final_states)
          for (int k = lane; k < p.nw; k += 32) p.final_states[tl * (uint64_t)p.nw + k] = s_x[r][k];
        if (lane == 0) p.best_rel[tl] = s_best[r];
      }
    }
    if (lane == 0) {
      if (cnt_acc) atomicAdd(&p.counters->accepts, cnt_acc);
      if (dwarp == 0) atomicAdd(&p.counters->cyc_decide, (unsigned long long)t_decide);
      if (dwarp == 0 && (p.debug_flags & 8)) atomicAdd(&p.counters->pad, n_walk);
    }
  }
}

template <typename T, int NCH, int R, int K, int G, int DW = 4>
cudaError_t launch_ws(const DenseParams<T> &p, cudaStream_t s, LaunchInfo *info) {
  constexpr int WS_THREADS = WsRegs<DW>::THREADS;
  const uint64_t grid64 = (p.num_tries + R - 1) / R;
  if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
  const size_t smem = (size_t)K * NCH * WS_APPLY_THREADS * 16;
  // the resumable instantiation only when a per-trajectory scale is given (osa_pt_anneal)
  auto kern = p.tscale_traj ? k_dense_seq_ws<T, NCH, R, K, G, true, DW>
                            : k_dense_seq_ws<T, NCH, R, K, G, false, DW>;
  DenseParams<T> pd = p;
  const char *dbg = getenv("OSA_WS_DEBUG");  // timing experiments, see DenseParams::debug_flags
  pd.debug_flags = dbg ? atoi(dbg) : 0;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  kern<<<(unsigned)grid64, WS_THREADS, smem, s>>>(pd);
  if (info) {
    info->grid = (int)grid64;
    info->block = WS_THREADS;
    info->traj_per_batch = R;
    info->smem = smem;
  }
  return cudaGetLastError();
}

// decide warps for the shapes with <= 3 column groups per thread (N <= 3072 fp32 / 1536 fp64),
// where the decisions, not the row streaming, bound the kernel; OSA_WS_DW=4 for A/B runs
int decide_warps_small_n() {
  static const int dw = [] {
    const char *e = getenv("OSA_WS_DW");
    return (e && atoi(e) == 4) ? 4 : 8;
  }();
  return dw;
}

}  // namespace

// same shape table as launch_dense_seq (osa_dense_seq.cu)
template <typename T>
cudaError_t launch_dense_seq_ws(const DenseParams<T> &p, cudaStream_t s, LaunchInfo *info);

template <>
cudaError_t launch_dense_seq_ws<float>(const DenseParams<float> &p, cudaStream_t s,
                                       LaunchInfo *info) {
  if (p.ld % 1024 != 0) return cudaErrorInvalidValue;
  const bool dw8 = decide_warps_small_n() == 8;
  switch (p.ld / 1024) {
    case 1: return dw8 ? launch_ws<float, 1, 16, 16, 4, 8>(p, s, info) : launch_ws<float, 1, 16, 16, 4>(p, s, info);
    case 2: return dw8 ? launch_ws<float, 2, 16, 16, 4, 8>(p, s, info) : launch_ws<float, 2, 16, 16, 4>(p, s, info);
    case 3: return dw8 ? launch_ws<float, 3, 12, 12, 4, 8>(p, s, info) : launch_ws<float, 3, 12, 12, 4>(p, s, info);
    case 4: {
      const char *e = getenv("OSA_WS_R");  // tuning knob (tools/probe.py): trajectories per CTA
      const int r = e ? atoi(e) : 12;
      if (r == 8) return launch_ws<float, 4, 8, 12, 1>(p, s, info);
      if (r == 10) return launch_ws<float, 4, 10, 12, 1>(p, s, info);
      return launch_ws<float, 4, 12, 12, 2>(p, s, info);  // measured best (profiles/r01)
    }
    case 5: return launch_ws<float, 5, 8, 9, 2>(p, s, info);
    case 6: return launch_ws<float, 6, 8, 8, 2>(p, s, info);
    case 7: return launch_ws<float, 7, 4, 6, 2>(p, s, info);
    case 8: return launch_ws<float, 8, 4, 6, 2>(p, s, info);
    default: return cudaErrorInvalidValue;
  }
}

template <>
cudaError_t launch_dense_seq_ws<double>(const DenseParams<double> &p, cudaStream_t s,
                                        LaunchInfo *info) {
  if (p.ld % 512 != 0) return cudaErrorInvalidValue;
  const bool dw8 = decide_warps_small_n() == 8;
  switch (p.ld / 512) {
    case 1: return dw8 ? launch_ws<double, 1, 16, 16, 4, 8>(p, s, info) : launch_ws<double, 1, 16, 16, 4>(p, s, info);
    case 2: return dw8 ? launch_ws<double, 2, 16, 16, 4, 8>(p, s, info) : launch_ws<double, 2, 16, 16, 4>(p, s, info);
    case 3: return dw8 ? launch_ws<double, 3, 12, 12, 4, 8>(p, s, info) : launch_ws<double, 3, 12, 12, 4>(p, s, info);
    case 4: return launch_ws<double, 4, 8, 12, 2>(p, s, info);
    case 5: return launch_ws<double, 5, 6, 9, 2>(p, s, info);
    case 6: return launch_ws<double, 6, 6, 8, 2>(p, s, info);
    case 7: return launch_ws<double, 7, 4, 6, 2>(p, s, info);
    case 8: return launch_ws<double, 8, 4, 6, 2>(p, s, info);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace osa
