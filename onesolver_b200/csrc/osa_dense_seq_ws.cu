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

#include "osa_dense_seq_ws.cuh"

namespace osa {

using namespace dseq;
using namespace dsws;

namespace {

// PT: resumable launch for parallel tempering (start spins from memory, per-trajectory threshold
// scale, final spins written back, running sweep counter); a separate instantiation so that the
// plain annealing kernel carries none of it.
//
// Decide role.  Every decide warp walks R/DW trajectories of the CTA side by side: NSEG lanes per
// trajectory, each holding the local fields of HS = 32/NSEG sites of the block in registers
// (R = 12, DW = 4: three trajectories x eight lanes x four sites).  The sites are walked in order
// -- compare, and on an accepted flip add the row of the diagonal tile to the fields of the later
// sites, under a per-lane predicate; the lanes of the later segments learn the decision through a
// ballot.  The cost does not depend on how many flips are accepted, and everything around the
// walk (bringing the snapshot up to date, the energy bookkeeping) is branch-free over the 32
// sites so that its shared-memory loads run ahead of the arithmetic.  After its walk a warp helps
// to prepare the next block: Philox draws and thresholds of all (site, trajectory) pairs and the
// two tiles of Q (cp.async), into the other half of double buffers; the CTA barrier that ends
// the block publishes them.
template <typename T, int NCH, int R, int K, int G, bool PT, int DW>
__global__ void __launch_bounds__(WsRegs<DW>::THREADS, 1) k_dense_seq_ws(const DenseParams<T> p) {
  constexpr int TH = WS_APPLY_THREADS;
  constexpr int WS_THREADS = WsRegs<DW>::THREADS;
  auto bar_cta = [] { bar_named<WS_THREADS>(1); };
  using C = Cfg<T, NCH, R, TH>;
  using VecT = typename C::VecT;
  constexpr int V = C::V, CPT = C::CPT, CHW = C::CHW, NWP = C::NWP;
  constexpr int KE = WsRing<T, NCH, R, K>::KE;
  constexpr int TILE_VECS = 32 * 32 / V;
  constexpr int DT = DW * 32;         // threads of the decide role
  // every decide warp walks TPW trajectories, NSEG lanes per trajectory, each lane with the
  // fields of HS sites in registers
  constexpr int TPW = (R + DW - 1) / DW;
  constexpr int NSEG = TPW <= 1 ? 32 : TPW <= 2 ? 16 : TPW <= 4 ? 8 : TPW <= 8 ? 4 : 2;
  constexpr int HS = 32 / NSEG;
  static_assert(TPW * NSEG <= 32, "lanes of a decide warp");

  // dynamic shared memory: the row ring of the apply warps (KE rows, private slots), then WsTiles
  extern __shared__ __align__(128) unsigned char s_ring[];
  __shared__ WsShared<T, R, NWP> sh;
  // accept / sign masks of block j (parity j&1).  Kept as plain arrays: the apply warps read them
  // with scalar loads at warp-uniform addresses, which keeps the 2R masks in uniform registers
  // through apply_rows (as members of the 16-byte aligned struct they were read with vector loads
  // into ordinary registers, and the kernel spilled)
  __shared__ uint32_t s_acc[2][R], s_sign[2][R];

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
    const bool resume = PT && p.fields_in != nullptr;  // fields of a previous launch (CTA-uniform)
    if (resume) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const T *src = p.fields_in + (batch0 + (uint64_t)(r < nvalid ? r : nvalid - 1)) * p.ld;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          T dv[V];
          vec_unpack<T>(*reinterpret_cast<const VecT *>(src + c * CHW + tid * V), dv);
#pragma unroll
          for (int e = 0; e < V; ++e) h[r].set(c * V + e, dv[e]);
        }
      }
    } else {
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
    }
    // snapshot of the 32 columns of block b into sh.snap[par]
    auto snapshot = [&](int b, int par) {
      const int i0 = b * 32;
      const int cb = i0 / CHW;
      const int rel = tid * V - (i0 % CHW);
      if (rel >= 0 && rel < 32) {
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
    };
    unsigned long long cnt_rows = 0, cnt_init_rows = 0;
    // role timers (osa_stats): cyc_init = initial fields, cyc_apply = streaming/applying rows
    // from the release of barrier B(g) on, cyc_stage = the rest of the main loop (snapshots and
    // waiting for the decide warps)
    long long t_apply = 0, t_init = clock64();

    bar_cta();  // #0: initial spins are in sh.x
    for (int b = 0; b < (resume ? 0 : nblk); ++b) {
      uint32_t am[R], sm[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        am[r] = sh.x[b][r];
        sm[r] = 0u;
      }
      const uint32_t any = apply_rows<T, NCH, R, KE, TH, G>(p.qoff, s_ring, p.ld, b * 32, am, sm, h, tid);
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
          apply_rows<T, NCH, R, KE, TH, G>(p.qoff, s_ring, p.ld, b * 32, am, sm, h, tid);
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
    if (PT && p.fields_out) {  // the fields after the last sweep, for the next launch
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r < nvalid) {
          T *dst = p.fields_out + (batch0 + (uint64_t)r) * p.ld;
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            T dv[V];
#pragma unroll
            for (int e = 0; e < V; ++e) dv[e] = h[r].get(c * V + e);
            *reinterpret_cast<VecT *>(dst + c * CHW + tid * V) = vec_pack<T>(dv);
          }
        }
      }
    }
    if (tid == 0) {
      atomicAdd(&p.counters->row_fetches, cnt_rows);
      atomicAdd(&p.counters->init_row_fetches, cnt_init_rows);
      atomicAdd(&p.counters->cyc_apply, (unsigned long long)t_apply);
      if (!(p.debug_flags & 8)) {
        atomicAdd(&p.counters->cyc_stage, (unsigned long long)(clock64() - t_main - t_apply));
        atomicAdd(&p.counters->cyc_init, (unsigned long long)t_init);
      }
    }
  } else {
    // =============================== DECIDE ROLE ===============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WsRegs<DW>::DECIDE));
    const int dt = tid - WS_APPLY_THREADS;
    const int lane = dt & 31, dwarp = dt >> 5;
    WsTiles<T> &tl = *reinterpret_cast<WsTiles<T> *>(s_ring + (size_t)KE * WsRing<T, NCH, R, K>::ROW_BYTES);

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
    for (int r = dt; r < R; r += DT) {
      sh.erel[r] = 0.0;
      sh.best[r] = 0.0;
      sh.atbest[r] = 1u;
      sh.naccept[r] = 0u;
      sh.ts[r] = (PT && p.tscale_traj && r < nvalid) ? p.tscale_traj[batch0 + (uint64_t)r] : (T)0;
    }

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

    // thresholds and tiles of block c into the buffers of parity `buf` (all decide threads).  The
    // tiles go straight from L2 to shared memory (cp.async, issued before the walk of the current
    // block); the thresholds are computed after the walk.
    auto prepare_tiles = [&](const BlockIt &c, int buf) {
      const int i0 = c.b * 32;
      const int bp = (c.b == 0) ? nblk - 1 : c.b - 1;  // previous block (cyclic)
      for (int qi = dt; qi < TILE_VECS; qi += DT) {
        const int row = qi / (32 / V), cv = qi % (32 / V);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(&tl.tile_d[buf][row][cv * V])),
                     "l"(p.qoff + (size_t)(i0 + row) * p.ld + i0 + cv * V)
                     : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(&tl.tile_x[buf][row][cv * V])),
                     "l"(p.qoff + (size_t)(bp * 32 + row) * p.ld + i0 + cv * V)
                     : "memory");
      }
    };
    auto prepare_thresholds = [&](const BlockIt &c, int buf) {
      const int i0 = c.b * 32;
      // one Philox block serves four consecutive sites of a trajectory (STREAM_SEQ: c0 = site>>2)
      const T ts = PT ? (T)0 : p.tscale[c.iter];
      for (int q = dt; q < R * 8; q += DT) {
        const int r = q >> 3, grp = q & 7;
        const uint64_t traj = p.first_try + batch0 + (uint64_t)r;
        const U4 d = engine_draw(p.seed, traj, STREAM_SEQ, (uint32_t)(i0 >> 2) + grp, c.step);
        const T tsr = PT ? sh.ts[r] : ts;
        sh.theta[buf][grp * 4 + 0][r] = threshold<T>(tsr, d.x);
        sh.theta[buf][grp * 4 + 1][r] = threshold<T>(tsr, d.y);
        sh.theta[buf][grp * 4 + 2][r] = threshold<T>(tsr, d.z);
        sh.theta[buf][grp * 4 + 3][r] = threshold<T>(tsr, d.w);
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    };

    // walk state: lane = (slot t of the warp's trajectories, column segment seg); the lanes of a
    // trajectory are NSEG consecutive lanes, each with the fields of HS sites in registers
    const int t = lane / NSEG, seg = lane % NSEG;
    const int r = dwarp + DW * t;
    const bool has = (t < TPW) && (r < R);
    // lanes without a trajectory read along with the warp's own first trajectory (dwarp < DW <= R):
    // their values are never used, but reading another warp's slot would be a (benign) race with
    // that warp's writes -- compute-sanitizer racecheck flags it (profiles/r02/racecheck_*)
    const int rr = has ? r : dwarp;
    const bool tv = has && r < nvalid;
    const bool lead = has && seg == 0;
    const int c0 = seg * HS;
    uint32_t pa = 0u, ps = 0u;  // accept / sign masks of the previous block (every lane of r)
    unsigned long long trace = TRACE_OFFSET;  // flip trace of trajectory r (lead lane), osa_common.cuh
    long long t_decide = 0, t_catch = 0, t_sites = 0;  // the last two only with debug flag 8

    // decision of block j of the schedule (block b of the sweep) for the trajectories of this warp
    auto walk = [&](long long j, int b, uint32_t step) {
      const int par = (int)(j & 1);
      const int i0 = b * 32;
      const uint32_t xw0 = sh.x[b][rr];
      // cyc_decide: from the release of barrier B(j-1) to the arrival at B(j)
      const long long t0 = clock_after(xw0);
      if (p.debug_flags & 2) {  // timing experiment: masks of density 3/16, no decisions
        if (lead) {
          const U4 d = engine_draw(p.seed, batch0 + (uint64_t)r, STREAM_SEQ, (uint32_t)j, 0u);
          // density of the pseudo-random masks: 3/16, or 1/16 (flag 64), 1/32 (128), 1/64 (192)
          uint32_t m = d.x & d.y & (d.z | d.w);
          if (p.debug_flags & 192) m = d.x & d.y & d.z & d.w;
          if (p.debug_flags & 128) m &= __funnelshift_l(d.x, d.x, 11);
          if ((p.debug_flags & 192) == 192) m &= __funnelshift_l(d.y, d.y, 13);
          s_acc[par][r] = m;
          s_sign[par][r] = d.w & m;
        }
        return t0;
      }
      long long tp = (p.debug_flags & 8) ? clock64() : 0;
      T hs[HS];
#pragma unroll
      for (int k = 0; k < HS; ++k) hs[k] = sh.snap[par][c0 + k][rr];
      // bring the snapshot up to date: the rows of block j-1 that the trajectory flipped, in site
      // order -- the same fma sequence the apply warps run on these columns, hence the same bits.
      // Branch-free over the 32 sites, so that the loads run ahead of the arithmetic: a lane
      // whose trajectory did not flip the site multiplies the row by zero.  (Visiting only the
      // flipped sites, a warp-uniform ffs loop, was measured slower: profiles/r01/probe_v30*.)
      {
        const T(*tx)[WsTiles<T>::TP] = tl.tile_x[par];
#pragma unroll
        for (int s = 0; s < 32; ++s) {
          const T m = Bits<T>::unit((ps >> s) & 1u, 0u - ((pa >> s) & 1u));  // +-1 if flipped, else 0
          T qv[HS];
          load_seg<T, HS>(&tx[s][c0], qv);
#pragma unroll
          for (int k = 0; k < HS; ++k) hs[k] = det::fma(m, qv[k], hs[k]);
        }
      }
      if (p.debug_flags & 8) {
        const long long now = clock_after(__float_as_uint((float)hs[0]));
        t_catch += now - tp;
        tp = now;
      }
      // the walk: phase ph decides the sites of segment ph, site by site.  An accepted flip adds
      // the row of the diagonal tile to the fields of the later sites: at once in the deciding
      // lane (its own segment), at the end of the phase in the lanes of the later segments, which
      // receive the multipliers of the phase by shuffles that were issued along the way.
      uint32_t acc = 0u;
      const T(*td)[WsTiles<T>::TP] = tl.tile_d[par];
#pragma unroll
      for (int ph = 0; ph < NSEG; ++ph) {
        const uint32_t mine = (tv && seg == ph) ? 0xffffffffu : 0u;
        const uint32_t later = (seg > ph) ? 0xffffffffu : 0u;
        T qv[HS][HS], mo[HS];
#pragma unroll
        for (int s = 0; s < HS; ++s) {
          const int site = ph * HS + s;
          const T th = sh.theta[par][site][rr];
          const uint32_t xl = (xw0 >> site) & 1u;  // a site flips at most once per block
          const T dE = Bits<T>::neg_if(hs[s], xl);
          const uint32_t ok = Bits<T>::neg_mask(det::add(dE, -th)) & mine &
                              ((i0 + site < n) ? 0xffffffffu : 0u);
          const T m = Bits<T>::unit(xl, ok);
          if (has && seg == ph) sh.dE[site][rr] = Bits<T>::masked(dE, ok);  // 0 when rejected
          acc |= ok & (1u << site);
          if (ph + 1 < NSEG || s + 1 < HS) load_seg<T, HS>(&td[site][c0], qv[s]);
#pragma unroll
          for (int k = s + 1; k < HS; ++k) hs[k] = det::fma(m, qv[s][k], hs[k]);
          if (ph + 1 < NSEG) mo[s] = Bits<T>::bcast(m, t * NSEG + ph, later);
        }
        if (ph + 1 < NSEG) {
#pragma unroll
          for (int s = 0; s < HS; ++s)
#pragma unroll
            for (int k = 0; k < HS; ++k) hs[k] = det::fma(mo[s], qv[s][k], hs[k]);
        }
      }
      // every lane of a trajectory gets the whole accept mask
#pragma unroll
      for (int d = 1; d < NSEG; d *= 2) acc |= __shfl_xor_sync(0xffffffffu, acc, d);
      if (p.debug_flags & 8) t_sites += clock_after(acc) - tp;
      // energies along the walk (annealing.hpp:115-121), in the lead lane of the trajectory:
      // erel += dE in site order, kb = site of the last flip that set a new best (strict <)
      double erel = sh.erel[rr], best = sh.best[rr];
      int kb = -1;
      __syncwarp();  // sh.dE of the block is complete
#pragma unroll
      for (int s = 0; s < 32; ++s) {
        const double e = det::add(erel, (double)sh.dE[s][rr]);  // dE is +0 for a rejected site
        erel = e;
        const bool nb = e < best;  // never true without a flip: best <= erel
        best = nb ? e : best;
        kb = nb ? s : kb;
      }
      // state at the best, kept lazily: written out (to the trajectory's row of best_states) only
      // when the walk has left the best state by the end of the block
      bool at_best = sh.atbest[rr] != 0u;
      bool copy = false;
      uint32_t wb = xw0;
      if (kb >= 0) {
        const uint32_t le = (2u << kb) - 1u;  // flips up to and including the best one
        at_best = (acc & ~le) == 0u;
        copy = !at_best;
        wb = xw0 ^ (acc & le);
      } else if (at_best && acc != 0u) {
        copy = true;  // the first flip of the block left the best state
        at_best = false;
      }
      // the copies are made by the whole warp, one trajectory after the other (lane = word)
      uint32_t cm = __ballot_sync(0xffffffffu, copy && tv && lead);
      while (cm) {
        const int cl = __ffs(cm) - 1;
        cm &= cm - 1u;
        const int cr = dwarp + DW * (cl / NSEG);
        const uint32_t wbr = __shfl_sync(0xffffffffu, wb, cl);
        uint32_t *const xb = p.best_states + (batch0 + (uint64_t)cr) * (uint64_t)p.nw;
        for (int k = lane; k < nblk; k += 32)
          xb[k] = (k == b) ? wbr : sh.x[k][cr];  // sh.x[b] still holds the spins before this block
      }
      __syncwarp();
      if (lead) {
        sh.erel[r] = erel;
        sh.best[r] = best;
        sh.atbest[r] = at_best ? 1u : 0u;
        sh.x[b][r] = xw0 ^ acc;
        s_acc[par][r] = acc;
        s_sign[par][r] = acc & xw0;  // spins that were 1 before their flip: sign -1
        if (acc != 0u) trace = trace_step(trace, step, (uint32_t)b, acc);
        const uint32_t na = sh.naccept[r] + (uint32_t)__popc(acc);
        sh.naccept[r] = na;
        if (na >= 0x80000000u) {
          atomicAdd(&p.counters->accepts, (unsigned long long)na);
          sh.naccept[r] = 0u;
        }
      }
      pa = acc;
      ps = acc & xw0;
      return t0;
    };

    // Barriers: #0 initial spins and scales are in shared memory; B(-1) snapshots of blocks 0
    // and 1, thresholds and tiles of block 0 are ready; B(j) masks of block j are ready (j = 0:
    // the apply warps have been waiting), thresholds and tiles of block j+1 are ready.
    bar_cta();  // #0
    BlockIt nxt{0, 0, 0, PT ? p.step_base : 0u};
    prepare_tiles(nxt, 0);
    prepare_thresholds(nxt, 0);
    bar_cta();  // B(-1)
    {
      int b = 0;
      for (long long j = 0; j < total_blocks; ++j) {
        const uint32_t step = nxt.step;  // sweep number of block j
        advance(nxt);
        const bool more = j + 1 < total_blocks;
        if (more) prepare_tiles(nxt, (int)((j + 1) & 1));
        const long long t0 = walk(j, b, step);
        b = (b + 1 == nblk) ? 0 : b + 1;
        if (more) prepare_thresholds(nxt, (int)((j + 1) & 1));
        t_decide += clock64() - t0;
        bar_cta();  // B(j)
      }
    }
    // ---- results ----
    for (int q = 0; q < nvalid; ++q) {
      const uint64_t row = batch0 + (uint64_t)q;
      if (sh.atbest[q])
        for (int k = dt; k < p.nw; k += DT) p.best_states[row * (uint64_t)p.nw + k] = sh.x[k][q];
      if (PT && p.final_states)
        for (int k = dt; k < p.nw; k += DT) p.final_states[row * (uint64_t)p.nw + k] = sh.x[k][q];
      if (dt == 0) p.best_rel[row] = sh.best[q];
    }
    if (dt < R && sh.naccept[dt])
      atomicAdd(&p.counters->accepts, (unsigned long long)sh.naccept[dt]);
    if (p.trace_hash && lead && tv) p.trace_hash[batch0 + (uint64_t)r] = trace;
    if (dt == 0) {
      atomicAdd(&p.counters->cyc_decide, (unsigned long long)t_decide);
      if (p.debug_flags & 8) {  // walk phases of decide warp 0 instead of the apply-role timers
        atomicAdd(&p.counters->cyc_init, (unsigned long long)t_catch);
        atomicAdd(&p.counters->cyc_stage, (unsigned long long)t_sites);
      }
    }
  }
}

template <typename T, int NCH, int R, int K, int G, int DW = 4>
cudaError_t launch_ws(const DenseParams<T> &p, cudaStream_t s, LaunchInfo *info) {
  constexpr int WS_THREADS = WsRegs<DW>::THREADS;
  const uint64_t grid64 = (p.num_tries + R - 1) / R;
  if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
  const size_t smem = (size_t)WsRing<T, NCH, R, K>::KE * NCH * WS_APPLY_THREADS * 16 +
                      sizeof(WsTiles<T>);
  // the resumable instantiation only when a per-trajectory scale is given (osa_pt_anneal)
  auto kern = p.tscale_traj ? k_dense_seq_ws<T, NCH, R, K, G, true, DW>
                            : k_dense_seq_ws<T, NCH, R, K, G, false, DW>;
  DenseParams<T> pd = p;
  pd.debug_flags = probe_env_int("OSA_WS_DEBUG");  // timing experiments (probe builds only)
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

// decide warps for the shapes with <= 3 column groups per thread (N <= 3072 fp32 / 1536 fp64):
// 4 like everywhere else.  With the walk-per-lane decide role the SM is issue-bound at these
// sizes and 8 decide warps only add instructions (N = 1024 fp64: 2.40 -> 2.30 ms for the probe,
// profiles/r01/probe_v31*.log); OSA_WS_DW=8 keeps the old choice for A/B runs.
int decide_warps_small_n() {
  const char *e = getenv("OSA_WS_DW");
  return (e && atoi(e) == 8) ? 8 : 4;
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
#ifndef OSA_WS_ONLY_F32_4  // (SASS inspection builds compile the N = 4096 fp32 shape alone)
    case 1: return dw8 ? launch_ws<float, 1, 16, 16, 4, 8>(p, s, info) : launch_ws<float, 1, 16, 16, 4>(p, s, info);
    case 2: return dw8 ? launch_ws<float, 2, 16, 16, 4, 8>(p, s, info) : launch_ws<float, 2, 16, 16, 4>(p, s, info);
    case 3: return dw8 ? launch_ws<float, 3, 12, 12, 4, 8>(p, s, info) : launch_ws<float, 3, 12, 12, 4>(p, s, info);
#endif
    case 4: {
      const char *e = getenv("OSA_WS_R");  // tuning knob (tools/probe.py): trajectories per CTA
      const int r = e ? atoi(e) : 12;
#ifndef OSA_WS_ONLY_F32_4
      if (r == 8) return launch_ws<float, 4, 8, 12, 1>(p, s, info);
      if (r == 10) return launch_ws<float, 4, 10, 12, 1>(p, s, info);
#endif
#ifndef OSA_WS_G4
#define OSA_WS_G4 2
#endif
      return launch_ws<float, 4, 12, 12, OSA_WS_G4>(p, s, info);  // G = 2: measured best (profiles/r01)
    }
#ifndef OSA_WS_ONLY_F32_4
    case 5: return launch_ws<float, 5, 8, 9, 2>(p, s, info);
    case 6: return launch_ws<float, 6, 8, 8, 2>(p, s, info);
    case 7: return launch_ws<float, 7, 4, 6, 2>(p, s, info);
    case 8: return launch_ws<float, 8, 4, 6, 2>(p, s, info);
#endif
    default: return cudaErrorInvalidValue;
  }
}

template <>
cudaError_t launch_dense_seq_ws<double>(const DenseParams<double> &p, cudaStream_t s,
                                        LaunchInfo *info) {
  if (p.ld % 512 != 0) return cudaErrorInvalidValue;
  const bool dw8 = decide_warps_small_n() == 8;
  switch (p.ld / 512) {
#ifndef OSA_WS_ONLY_F32_4
    case 1: return dw8 ? launch_ws<double, 1, 16, 16, 4, 8>(p, s, info) : launch_ws<double, 1, 16, 16, 4>(p, s, info);
    case 2: return dw8 ? launch_ws<double, 2, 16, 16, 4, 8>(p, s, info) : launch_ws<double, 2, 16, 16, 4>(p, s, info);
    case 3: return dw8 ? launch_ws<double, 3, 12, 12, 4, 8>(p, s, info) : launch_ws<double, 3, 12, 12, 4>(p, s, info);
    case 4: return launch_ws<double, 4, 8, 12, 2>(p, s, info);
    case 5: return launch_ws<double, 5, 6, 9, 2>(p, s, info);
    case 6: return launch_ws<double, 6, 6, 8, 2>(p, s, info);
    case 7: return launch_ws<double, 7, 4, 6, 2>(p, s, info);
    case 8: return launch_ws<double, 8, 4, 6, 2>(p, s, info);
#endif
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace osa
