// osa_dense_seq_ws2.cu -- K1s/flow: the dense sequential-sweep kernel with free-running decide and
// apply roles (same arithmetic and results as osa_dense_seq.cu / osa_dense_seq_ws.cu, bit for bit).
//
// osa_dense_seq_ws.cu runs the two roles in lock step: one CTA barrier per block of 32 sites and a
// decide role whose cost (a 32-site walk) does not depend on how many flips a block has.  Here:
//
//   * no CTA barrier in the main loop.  The roles hand blocks over through sequence numbers in
//     shared memory (st.release / ld.acquire): accept masks of block u in one of NS slots
//     (decide -> apply), the snapshot of the 32 columns of block u+2 (apply -> decide, written by
//     the one warp that owns those columns), and per slot a count of the apply warps that have
//     read its masks (apply -> decide, slot re-use).  The eight apply warps are not synchronised
//     with each other at all: every thread owns its columns and its row buffers.
//     The initial fields are the first nblk blocks of the same stream (masks = the initial spins).
//   * the decide role walks a block from accepted flip to accepted flip (lane = site: compare,
//     ballot, first set bit, add that row of the diagonal tile to the later sites), the
//     trajectories of a warp side by side, so a cold block costs a few steps instead of 32;
//     thresholds stay in registers (one Philox block per four sites, handed out by shuffles).
//   * the rows of a block are streamed by apply_rows (osa_dense_seq.cuh), the cp.async ring of the
//     other dense kernels.  Two other feeds were built and measured in round 2 and rejected
//     (profiles/r02): a ring that streams across block boundaries (more instructions per row than
//     it saves in refills), and rows staged in registers with plain 16-byte loads (one pass
//     through the load/store unit instead of two, but only four rows in flight per warp:
//     latency-bound at ~555 clk per row).
#include <cstdlib>

#include "osa_dense_seq_ws.cuh"

namespace osa {

using namespace dseq;
using namespace dsws;

namespace {

#ifndef OSA_FLOW_SLEEP
#define OSA_FLOW_SLEEP 100  // ns between two polls of a sequence number
#endif
constexpr int NS = 4;  // mask slots between the roles
constexpr int FLOW_DW = 4;

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_addr(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(uint32_t *p, uint32_t v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_addr(p)), "r"(v) : "memory");
}
// A value that all lanes of the warp read from one shared-memory word with one instruction: the
// broadcast from lane 0 tells the compiler that it is warp-uniform, so that the loops controlled by
// it (and the accept masks behind them) stay on the uniform datapath.
__device__ __forceinline__ uint32_t uni(uint32_t v) { return __reduce_or_sync(0xffffffffu, v); }
// wait until the sequence number at p has reached `need` (wrap-safe)
__device__ __forceinline__ void wait_seq(const uint32_t *p, uint32_t need) {
  while ((int)(uni(ld_acquire(p)) - need) < 0) __nanosleep(OSA_FLOW_SLEEP);
}

template <typename T, int R, int NWP, int TPW>
struct FlowShared {
  alignas(16) T snap[2][32][R];       // columns of block u (parity u&1) after block u-2, [column][traj]
  T dE[FLOW_DW][TPW][32];             // dE of the accepted flips of the block a decide warp just walked
  T ts[R];                            // per-trajectory threshold scale (only with p.tscale_traj)
  uint32_t x[NWP][R];                 // current spins, [word][traj]
  uint32_t anyw[NS][FLOW_DW];         // union of the accept masks of a decide warp's trajectories
};

template <typename T, int NCH, int R, int K>
struct FlowRing {
  using C = Cfg<T, NCH, R, WS_APPLY_THREADS>;
  static constexpr int TPW = (R + FLOW_DW - 1) / FLOW_DW;
  static constexpr int ROW_BYTES = NCH * WS_APPLY_THREADS * 16;
  static constexpr int FIT = (WS_SMEM_LIMIT - (int)sizeof(FlowShared<T, R, C::NWP, TPW>) -
                              (2 * NS * R + 2 * NS + 16) * 4 - (int)sizeof(WsTiles<T>)) / ROW_BYTES;
  static constexpr int KE = K < FIT ? K : FIT;  // rows in flight
  static_assert(KE >= 3, "row ring too small");
};


template <typename T, int NCH, int R, int K, int G, bool PT>
__global__ void __launch_bounds__(WsRegs<FLOW_DW>::THREADS, 1) k_dense_seq_flow(const DenseParams<T> p) {
  constexpr int DW = FLOW_DW;
  constexpr int TH = WS_APPLY_THREADS;
  using C = Cfg<T, NCH, R, TH>;
  using VecT = typename C::VecT;
  constexpr int V = C::V, CPT = C::CPT, CHW = C::CHW, NWP = C::NWP;
  constexpr int KE = FlowRing<T, NCH, R, K>::KE;
  constexpr int ROW_BYTES = FlowRing<T, NCH, R, K>::ROW_BYTES;
  constexpr int TPW = FlowRing<T, NCH, R, K>::TPW;
  constexpr int TILE_VECS = 32 * 32 / V;
  constexpr int DT = DW * 32;
  static_assert(TPW <= 4, "a decide warp walks at most four trajectories (Philox lanes)");

  extern __shared__ __align__(128) unsigned char s_ring[];
  __shared__ FlowShared<T, R, NWP, TPW> sh;
  // accept / sign masks of block u in slot u % NS; plain arrays, read by the apply warps with
  // scalar loads at warp-uniform addresses (they stay in uniform registers through the row loop)
  __shared__ uint32_t s_acc[NS][R], s_sign[NS][R], s_any[NS];
  __shared__ uint32_t s_seq_masks;    // blocks whose masks are published
  __shared__ uint32_t s_seq_snap[2];  // 1 + block whose snapshot is in sh.snap[parity]
  __shared__ uint32_t s_consumed[NS];  // per slot: reads of its masks by apply warps (8 per block)
  // role timers of thread 0 (osa_stats), kept out of the registers of the row loop: start of the
  // stream, end of the initial fields, time spent waiting for masks
  __shared__ long long s_t_start, s_t_init, s_t_wait;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int n = p.n;
  const int nblk = (n + 31) >> 5;
  const uint64_t batch0 = (uint64_t)blockIdx.x * R;
  const uint64_t left = p.num_tries - batch0;
  const int nvalid = left < (uint64_t)R ? (int)left : R;
  const long long total_blocks = (long long)p.num_iter * p.sweeps_per_beta * nblk;
  // unified block numbering: u < nblk are the blocks of the initial fields (masks = initial spins,
  // every multiplier +1), u = nblk + j is block j of the schedule
  const long long total_u = total_blocks + nblk;
#ifdef OSA_PROBE  // timing experiments (probe builds only, results are meaningless): 1 = the apply
  const int dbg = p.debug_flags;  // warps stream no rows, 2 = pseudo-random masks instead of the walk
#else
  constexpr int dbg = 0;
#endif

  if (tid == 0) {
    s_seq_masks = 0u;
    s_seq_snap[0] = s_seq_snap[1] = 0u;
    for (int k = 0; k < NS; ++k) s_consumed[k] = 0u;
  }
  __syncthreads();

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
    // snapshot of the 32 columns of block bcol, for block uu of the stream
    auto snapshot = [&](int bcol, long long uu) {
      const int par = (int)(uu & 1);
      const int i0 = bcol * 32;
      const int cb = i0 / CHW;
      const int rel = tid * V - (i0 % CHW);
      const bool own = rel >= 0 && rel < 32;
      if (own) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (c == cb) {
#pragma unroll
            for (int e = 0; e < V; ++e)
#pragma unroll
              for (int r = 0; r < R; ++r) sh.snap[par][rel + e][r] = h[r].get(c * V + e);
          }
        }
      }
      if (__any_sync(0xffffffffu, own)) {  // the owners of 32 consecutive columns sit in one warp
        __syncwarp();
        if (lane == 0) st_release(&s_seq_snap[par], (uint32_t)uu + 1u);
      }
    };

    // ---- the row stream ----
    uint32_t am[R], sm[R];
    unsigned long long cnt_rows = 0, cnt_init_rows = 0;
    if (tid == 0) {
      s_t_start = clock64();
      s_t_init = s_t_wait = 0;
    }
    if (nblk == 1) snapshot(0, 1);  // "after block -1": the first block of the schedule starts from diag

    // Block u is opened by OPEN_BLOCK: wait for its masks, bring them into (uniform) registers and
    // count this warp as a reader of the slot.  The masks are in registers once `any` is computed
    // and the decide warps re-use the slot a full decide step after the last warp has counted
    // itself, so no fence is needed (a fence here would also wait for the row loads in flight).
    // CLOSE_BLOCK: count the rows, hand the columns of block u+2 to the decide warps.
#define OPEN_BLOCK()                                                                    \
  _Pragma("unroll") for (int r = 0; r < R; ++r) {                                       \
    am[r] = s_acc[slot][r];                                                               \
    sm[r] = s_sign[slot][r];                                                             \
  }                                                                                     \
  uint32_t any = 0u;                                                                    \
  _Pragma("unroll") for (int r = 0; r < R; ++r) any |= am[r];                           \
  if (dbg & 1) any = 0u; /* timing experiment: the decide warps alone */                \
  if (lane == 0) atomicAdd(&s_consumed[slot], 1u);
#define CLOSE_BLOCK()                                                                   \
  if (u < nblk) cnt_init_rows += (unsigned)__popc(any);                                 \
  else cnt_rows += (unsigned)__popc(any);                                               \
  if (u + 2 >= nblk && u + 2 < total_u) {                                               \
    int b2 = b + 2; /* columns of block u+2 as they are now */                          \
    if (b2 >= nblk) b2 -= nblk;                                                         \
    if (b2 >= nblk) b2 -= nblk; /* nblk == 1 */                                         \
    snapshot(b2, u + 2);                                                                \
  }                                                                                     \
  if (u + 1 == nblk && tid == 0) s_t_init = clock64() - s_t_start;                      \
  if (u + 1 < total_u) { /* the masks of the next block */                              \
    if (tid == 0) s_t_wait -= clock64();                                                \
    wait_seq(&s_seq_masks, (uint32_t)u + 2u);                                           \
    if (tid == 0) s_t_wait += clock64();                                                \
  }

    {
      // The cp.async ring of apply_rows (osa_dense_seq.cuh), restarted at every block.
      int b = 0, slot = 0;
      wait_seq(&s_seq_masks, 1u);
      for (long long u = 0; u < total_u; ++u) {
        OPEN_BLOCK()
        if (any != 0u) apply_rows<T, NCH, R, KE, TH, G>(p.qoff, s_ring, p.ld, b * 32, am, sm, h, tid);
        CLOSE_BLOCK()
        b = (b + 1 == nblk) ? 0 : b + 1;
        slot = (slot + 1) % NS;
      }
    }
#undef OPEN_BLOCK
#undef CLOSE_BLOCK
    if (tid == 0) {
      atomicAdd(&p.counters->row_fetches, cnt_rows);
      atomicAdd(&p.counters->init_row_fetches, cnt_init_rows);
      atomicAdd(&p.counters->cyc_apply, (unsigned long long)(clock64() - s_t_start - s_t_init - s_t_wait));
      atomicAdd(&p.counters->cyc_stage, (unsigned long long)s_t_wait);
      atomicAdd(&p.counters->cyc_init, (unsigned long long)s_t_init);
    }
  } else {
    // =============================== DECIDE ROLE ===============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WsRegs<DW>::DECIDE));
    auto bar_decide = [] { bar_named<DT>(2); };
    const int dt = tid - WS_APPLY_THREADS;
    const int dwarp = dt >> 5;
    WsTiles<T> &tl = *reinterpret_cast<WsTiles<T> *>(s_ring + (size_t)KE * ROW_BYTES);

    // initial spins (replaces random.bit(), annealing.hpp:90-92)
    for (int q = dt; q < R * NWP; q += DT) {
      const int r = q / NWP, k = q % NWP;
      uint32_t word = 0;
      if (r < nvalid && k < nblk) {
        if (PT && p.init_states) {  // resume: the spins a previous launch left in final_states
          word = p.init_states[(batch0 + (uint64_t)r) * (uint64_t)p.nw + k];
        } else {
          const U4 d = engine_draw(p.seed, p.first_try + batch0 + (uint64_t)r, STREAM_INIT,
                                   (uint32_t)k >> 2, 0u);
          word = pick(d, (uint32_t)k & 3u);
        }
        const int valid = n - k * 32;
        if (valid < 32) word &= (1u << valid) - 1u;
      }
      sh.x[k][r] = word;
    }
    for (int r = dt; r < R; r += DT)
      sh.ts[r] = (PT && p.tscale_traj && r < nvalid) ? p.tscale_traj[batch0 + (uint64_t)r] : (T)0;
    bar_decide();

    // position of a block in the schedule
    struct BlockIt {
      int iter, sw, b;
      uint32_t step;
    };
    auto advance = [&](BlockIt &c) {
      if (++c.b == nblk) {
        c.b = 0;
        ++c.step;
        if (++c.sw == p.sweeps_per_beta) {
          c.sw = 0;
          ++c.iter;
        }
      }
    };
    // the two tiles of Q for block c of the sweep, straight from L2 to shared memory (cp.async):
    // tile_d = the diagonal tile, tile_x = rows of the previous block x columns of this block
    auto prepare_tiles = [&](int cb, int buf) {
      const int i0 = cb * 32;
      const int bp = (cb == 0) ? nblk - 1 : cb - 1;  // previous block (cyclic)
      for (int qi = dt; qi < TILE_VECS; qi += DT) {
        const int row = qi / (32 / V), cv = qi % (32 / V);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(&tl.tile_d[buf][row][cv * V])),
                     "l"(p.qoff + (size_t)(i0 + row) * p.ld + i0 + cv * V)
                     : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(&tl.tile_x[buf][row][cv * V])),
                     "l"(p.qoff + (size_t)(bp * 32 + row) * p.ld + i0 + cv * V)
                     : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // the trajectories of this warp: slot t is trajectory r = dwarp + DW * t.  Lane = site in the
    // walk; lane t (< TPW) also keeps the scalars of trajectory slot t.
    int rr[TPW];
    bool tv[TPW];
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      const int r = dwarp + DW * t;
      rr[t] = r < R ? r : 0;
      tv[t] = r < nvalid;
    }
    const int my_r = dwarp + DW * lane;             // lane t: its trajectory
    const bool my_has = lane < TPW && my_r < R;
    const bool my_tv = my_has && my_r < nvalid;
    double erel = 0.0, best = 0.0;                   // energy relative to the start: now / best
    bool at_best = true;
    uint32_t naccept = 0u;
    unsigned long long trace = TRACE_OFFSET;

    // thresholds of block c for the sites of this warp's trajectories, into th[]: lane (t, g)
    // draws the Philox block of sites 4g..4g+3 of trajectory slot t (STREAM_SEQ: c0 = site >> 2),
    // the lane of site s takes word s & 3 of lane (t, s >> 2)
    T th[TPW];
    auto thresholds = [&](const BlockIt &c) {
      const int tt = lane >> 3, gg = lane & 7;
      const int r = dwarp + DW * tt;
      U4 d = U4{0u, 0u, 0u, 0u};
      if (tt < TPW && r < R)
        d = engine_draw(p.seed, p.first_try + batch0 + (uint64_t)r, STREAM_SEQ,
                        (uint32_t)(c.b * 8) + gg, c.step);
#pragma unroll
      for (int t = 0; t < TPW; ++t) {
        const int src = t * 8 + (lane >> 2);
        U4 e;
        e.x = __shfl_sync(0xffffffffu, d.x, src);
        e.y = __shfl_sync(0xffffffffu, d.y, src);
        e.z = __shfl_sync(0xffffffffu, d.z, src);
        e.w = __shfl_sync(0xffffffffu, d.w, src);
        const T ts = PT ? sh.ts[rr[t]] : p.tscale[c.iter];
        th[t] = threshold<T>(ts, pick(e, (uint32_t)lane & 3u));
      }
    };

    uint32_t pa[TPW], ps[TPW];  // accept / sign masks of the previous block
    long long t_busy = 0;

    // the blocks of the initial fields: masks = the initial spins, published by decide warp 0
    BlockIt cur{0, 0, 0, PT ? p.step_base : 0u};
    prepare_tiles(0, (int)(nblk & 1));
    thresholds(cur);
    if (dwarp == 0) {
      for (long long u = 0; u < nblk; ++u) {
        const int slot = (int)(u % NS);
        if (u >= NS) wait_seq(&s_consumed[slot], (uint32_t)(8 * (u / NS)));
        const uint32_t v = lane < R ? sh.x[u][lane] : 0u;
        if (lane < R) {
          s_acc[slot][lane] = v;
          s_sign[slot][lane] = 0u;
        }
        const uint32_t any = __reduce_or_sync(0xffffffffu, v);
        __syncwarp();
        if (lane == 0) {
          s_any[slot] = any;
          st_release(&s_seq_masks, (uint32_t)u + 1u);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      pa[t] = sh.x[nblk - 1][rr[t]];
      ps[t] = 0u;
    }

    for (long long u = nblk; u < total_u; ++u) {
      const int par = (int)(u & 1), slot = (int)(u % NS);
      const int b = cur.b;
      const int i0 = b * 32;
      asm volatile("cp.async.wait_all;" ::: "memory");
      bar_decide();  // tiles of block u are complete; every decide warp is done with block u-1
      BlockIt nxt = cur;
      advance(nxt);
      if (u + 1 < total_u) prepare_tiles(nxt.b, par ^ 1);
      wait_seq(&s_seq_snap[par], (uint32_t)u + 1u);
      if (u >= NS) wait_seq(&s_consumed[slot], (uint32_t)(8 * (u / NS)));
      const long long t0 = clock64();

      const bool site_ok = i0 + lane < n;
      T h[TPW];
      uint32_t xw[TPW];
#pragma unroll
      for (int t = 0; t < TPW; ++t) {
        h[t] = sh.snap[par][lane][rr[t]];
        xw[t] = sh.x[b][rr[t]];
      }
      // bring the snapshot up to date: the rows of block u-1 that the trajectory flipped, in site
      // order -- the same fma sequence the apply warps run on these columns, hence the same bits.
      // Four flips at a time so that the loads run ahead of the fma chain; past the last flip the
      // multiplier is zero.
      {
        const T(*tx)[WsTiles<T>::TP] = tl.tile_x[par];
#pragma unroll
        for (int t = 0; t < TPW; ++t) {
          uint32_t m = pa[t];
          const uint32_t sg = ps[t];
          while (m != 0u) {
            T qv[4], ml[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int s = __ffs(m | 0x80000000u) - 1;  // 31 when none is left (multiplier 0)
              ml[q] = Bits<T>::unit((sg >> s) & 1u, m != 0u ? 0xffffffffu : 0u);
              qv[q] = tx[s][lane];
              m &= m - 1u;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) h[t] = det::fma(ml[q], qv[q], h[t]);
          }
        }
      }
      // the walk, from accepted flip to accepted flip
      uint32_t acc[TPW], live[TPW];
      if (dbg & 2) {  // masks of density 3/16 (1/16 with flag 64, 1/32: 128, 1/64: 192), no decisions
#pragma unroll
        for (int t = 0; t < TPW; ++t) {
          const U4 d = engine_draw(p.seed, batch0 + (uint64_t)rr[t], STREAM_SEQ, (uint32_t)u, 0u);
          uint32_t m = d.x & d.y & (d.z | d.w);
          if (dbg & 192) m = d.x & d.y & d.z & d.w;
          if (dbg & 128) m &= __funnelshift_l(d.x, d.x, 11);
          if ((dbg & 192) == 192) m &= __funnelshift_l(d.y, d.y, 13);
          acc[t] = tv[t] ? m : 0u;
        }
      } else {
      T dE[TPW];
      bool ok[TPW];
#pragma unroll
      for (int t = 0; t < TPW; ++t) {
        acc[t] = 0u;
        live[t] = 0xffffffffu;
        dE[t] = Bits<T>::neg_if(h[t], (xw[t] >> lane) & 1u);
        ok[t] = tv[t] && site_ok && dE[t] < th[t];
      }
      {
        const T(*td)[WsTiles<T>::TP] = tl.tile_d[par];
        for (;;) {
          uint32_t bal[TPW], some = 0u;
#pragma unroll
          for (int t = 0; t < TPW; ++t) {
            bal[t] = __ballot_sync(0xffffffffu, ok[t]) & live[t];
            some |= bal[t];
          }
          if (some == 0u) break;
#pragma unroll
          for (int t = 0; t < TPW; ++t) {
            const uint32_t bt = bal[t];
            const int f = __ffs(bt | 0x80000000u) - 1;  // 31 when the trajectory is done
            const uint32_t fb = bt & (0u - bt);         // the bit of f, 0 when done
            const T q = td[f][lane];
            if (fb >> lane & 1u) sh.dE[dwarp][t][lane] = dE[t];  // the accepted flip's dE
            const T m = Bits<T>::unit((xw[t] >> f) & 1u, (fb != 0u && lane > f) ? 0xffffffffu : 0u);
            h[t] = det::fma(m, q, h[t]);
            dE[t] = Bits<T>::neg_if(h[t], (xw[t] >> lane) & 1u);
            ok[t] = tv[t] && site_ok && dE[t] < th[t];
            acc[t] |= fb;
            live[t] = fb != 0u ? ~((fb << 1) - 1u) : live[t];  // the sites after f
          }
        }
      }
      }
      __syncwarp();  // sh.dE of the block is complete
      // energies along the walk (annealing.hpp:115-121), by lane t for trajectory slot t:
      // erel += dE in site order, kb = site of the last flip that set a new best (strict <)
      uint32_t my_acc = 0u, my_xw = 0u;
#pragma unroll
      for (int t = 0; t < TPW; ++t) {
        if (lane == t) {
          my_acc = acc[t];
          my_xw = xw[t];
        }
      }
      bool copy = false;
      uint32_t wb = my_xw;
      if (my_has) {
        int kb = -1;
        for (uint32_t m = my_acc; m != 0u; m &= m - 1u) {
          const int s = __ffs(m) - 1;
          const double e = det::add(erel, (double)sh.dE[dwarp][lane][s]);
          erel = e;
          if (e < best) {
            best = e;
            kb = s;
          }
        }
        // state at the best, kept lazily: written out (to the trajectory's row of best_states)
        // only when the walk has left the best state by the end of the block
        if (kb >= 0) {
          const uint32_t le = (2u << kb) - 1u;  // flips up to and including the best one
          at_best = (my_acc & ~le) == 0u;
          copy = !at_best;
          wb = my_xw ^ (my_acc & le);
        } else if (at_best && my_acc != 0u) {
          copy = true;  // the first flip of the block left the best state
          at_best = false;
        }
        if (my_acc != 0u) trace = trace_step(trace, cur.step, (uint32_t)b, my_acc);
        naccept += (uint32_t)__popc(my_acc);
        if (naccept >= 0x80000000u) {
          atomicAdd(&p.counters->accepts, (unsigned long long)naccept);
          naccept = 0u;
        }
        s_acc[slot][my_r] = my_acc;
        s_sign[slot][my_r] = my_acc & my_xw;  // spins that were 1 before their flip: sign -1
      }
      // the copies are made by the whole warp, one trajectory after the other (lane = word);
      // sh.x[b] still holds the spins before this block
      uint32_t cm = __ballot_sync(0xffffffffu, copy && my_tv);
      while (cm) {
        const int cl = __ffs(cm) - 1;
        cm &= cm - 1u;
        const int cr = dwarp + DW * cl;
        const uint32_t wbr = __shfl_sync(0xffffffffu, wb, cl);
        uint32_t *const xb = p.best_states + (batch0 + (uint64_t)cr) * (uint64_t)p.nw;
        for (int k = lane; k < nblk; k += 32) xb[k] = (k == b) ? wbr : sh.x[k][cr];
      }
      __syncwarp();
      if (my_has) sh.x[b][my_r] = my_xw ^ my_acc;
      uint32_t wany = 0u;
#pragma unroll
      for (int t = 0; t < TPW; ++t) {
        wany |= acc[t];
        pa[t] = acc[t];
        ps[t] = acc[t] & xw[t];
      }
      if (lane == 0) sh.anyw[slot][dwarp] = wany;
      t_busy += clock64() - t0;
      bar_decide();  // the masks of all trajectories are written
      if (dt == 0) {
        uint32_t any = 0u;
#pragma unroll
        for (int w = 0; w < DW; ++w) any |= sh.anyw[slot][w];
        s_any[slot] = any;
        st_release(&s_seq_masks, (uint32_t)u + 1u);
      }
      cur = nxt;
      if (u + 1 < total_u) {
        const long long t1 = clock64();
        thresholds(cur);
        t_busy += clock64() - t1;
      }
    }
    // ---- results ----
    __syncwarp();
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      const int r = dwarp + DW * t;
      const bool ab = __shfl_sync(0xffffffffu, at_best ? 1 : 0, t) != 0;
      if (r < nvalid) {
        const uint64_t row = batch0 + (uint64_t)r;
        if (ab)
          for (int k = lane; k < p.nw; k += 32) p.best_states[row * (uint64_t)p.nw + k] = sh.x[k][r];
        if (PT && p.final_states)
          for (int k = lane; k < p.nw; k += 32) p.final_states[row * (uint64_t)p.nw + k] = sh.x[k][r];
      }
    }
    if (my_tv) {
      const uint64_t row = batch0 + (uint64_t)my_r;
      p.best_rel[row] = best;
      if (p.trace_hash) p.trace_hash[row] = trace;
      if (naccept) atomicAdd(&p.counters->accepts, (unsigned long long)naccept);
    }
    if (dt == 0) atomicAdd(&p.counters->cyc_decide, (unsigned long long)t_busy);
  }
}

template <typename T, int NCH, int R, int K, int G>
cudaError_t launch_flow(const DenseParams<T> &p, cudaStream_t s, LaunchInfo *info) {
  constexpr int THREADS = WsRegs<FLOW_DW>::THREADS;
  const uint64_t grid64 = (p.num_tries + R - 1) / R;
  if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
  const size_t smem = (size_t)FlowRing<T, NCH, R, K>::KE * NCH * WS_APPLY_THREADS * 16 + sizeof(WsTiles<T>);
  // the resumable instantiation only when a per-trajectory scale is given (osa_pt_anneal)
  auto kern = p.tscale_traj ? k_dense_seq_flow<T, NCH, R, K, G, true> : k_dense_seq_flow<T, NCH, R, K, G, false>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  DenseParams<T> pd = p;
  pd.debug_flags = probe_env_int("OSA_WS_DEBUG");  // timing experiments (probe builds only)
  kern<<<(unsigned)grid64, THREADS, smem, s>>>(pd);
  if (info) {
    info->grid = (int)grid64;
    info->block = THREADS;
    info->traj_per_batch = R;
    info->smem = smem;
  }
  return cudaGetLastError();
}

}  // namespace

// same shape table as launch_dense_seq_ws (osa_dense_seq_ws.cu)
template <typename T>
cudaError_t launch_dense_seq_flow(const DenseParams<T> &p, cudaStream_t s, LaunchInfo *info);

template <>
cudaError_t launch_dense_seq_flow<float>(const DenseParams<float> &p, cudaStream_t s, LaunchInfo *info) {
  if (p.ld % 1024 != 0) return cudaErrorInvalidValue;
  switch (p.ld / 1024) {
#ifndef OSA_WS_ONLY_F32_4  // (SASS inspection builds compile the N = 4096 fp32 shape alone)
    case 1: return launch_flow<float, 1, 16, 16, 4>(p, s, info);
    case 2: return launch_flow<float, 2, 16, 16, 4>(p, s, info);
    case 3: return launch_flow<float, 3, 12, 12, 4>(p, s, info);
#endif
    case 4: {
      const char *e = getenv("OSA_FLOW_R");  // tuning knob (result-preserving): trajectories per CTA
      const int r = e ? atoi(e) : 12;
#ifndef OSA_WS_ONLY_F32_4
      if (r == 8) return launch_flow<float, 4, 8, 12, 2>(p, s, info);
      if (r == 10) return launch_flow<float, 4, 10, 12, 2>(p, s, info);
      if (r == 6) return launch_flow<float, 4, 6, 12, 2>(p, s, info);
#endif
      return launch_flow<float, 4, 12, 12, 2>(p, s, info);
    }
#ifndef OSA_WS_ONLY_F32_4
    case 5: return launch_flow<float, 5, 8, 9, 2>(p, s, info);
    case 6: return launch_flow<float, 6, 8, 8, 2>(p, s, info);
    case 7: return launch_flow<float, 7, 4, 6, 2>(p, s, info);
    case 8: return launch_flow<float, 8, 4, 6, 2>(p, s, info);
#endif
    default: return cudaErrorInvalidValue;
  }
}

template <>
cudaError_t launch_dense_seq_flow<double>(const DenseParams<double> &p, cudaStream_t s, LaunchInfo *info) {
  if (p.ld % 512 != 0) return cudaErrorInvalidValue;
  switch (p.ld / 512) {
#ifndef OSA_WS_ONLY_F32_4
    case 1: return launch_flow<double, 1, 16, 16, 4>(p, s, info);
    case 2: return launch_flow<double, 2, 16, 16, 4>(p, s, info);
    case 3: return launch_flow<double, 3, 12, 12, 4>(p, s, info);
    case 4: return launch_flow<double, 4, 8, 12, 2>(p, s, info);
    case 5: return launch_flow<double, 5, 6, 9, 2>(p, s, info);
    case 6: return launch_flow<double, 6, 6, 8, 2>(p, s, info);
    case 7: return launch_flow<double, 7, 4, 6, 2>(p, s, info);
    case 8: return launch_flow<double, 8, 4, 6, 2>(p, s, info);
#endif
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace osa
