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

// Sequential sweeps read the couplings from a GROUPED layout built at problem creation
// (osa_api.cu): the sites are taken in groups of G = 4 or 8 consecutive ones (8 when at least
// three quarters of the 8-groups of the instance are independent, see below); a group stores
// len x G entries, entry (t, k) = the t-th neighbour of site G g + k as
// {byte offset of its spin word in X, coupling}, rows shorter than the longest of the four padded
// with {offset of an always-zero word, 0}.  A pad never adds anything (its spin bit is 0), so the
// additions of a row stay the CSR ones in CSR order.  One 16-byte shared-memory load brings two
// (fp32) entries, the spin word address needs no index arithmetic, and for groups whose sites are
// pairwise non-adjacent (flagged at creation) the G fields are independent chains that run side by
// side and the G decisions are taken together: about 5 instructions per neighbour instead of 12
// for the plain CSR walk, and G-fold instruction-level parallelism in a kernel whose residency
// (7-8 warps per SM at N = 5627: the spin words of 32 trajectories take 22.5 KB) cannot hide latency.
constexpr int SP_LOG = 64;        // flips remembered per trajectory after leaving a best state
constexpr int SP_HALF_CAP = 256;  // entries per staging buffer: half a block = 16 sites x degree 16

template <typename T> struct SpEnt;
template <> struct __align__(8) SpEnt<float> {
  uint32_t xoff;
  float val;
};
template <> struct __align__(16) SpEnt<double> {
  uint32_t xoff, pad;
  double val;
};

__device__ __forceinline__ uint32_t sp_smem_addr(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ONE_WARP: the CTA is a single warp, so every shared-memory base address is CTA-uniform (the spin
// word of a neighbour is then one load at [offset register + uniform base]); used whenever shared
// memory, not the 32-CTA limit, bounds the residency.  STAGED: the entries of the next half block
// are copied into shared memory with cp.async while the current one is processed.
template <typename T, int G, bool ONE_WARP, bool STAGED>
__global__ void k_sparse(const SparseParams<T> p, int x_words_per_warp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Ent = SpEnt<T>;
  const int lane = threadIdx.x & 31;
  const int warp = ONE_WARP ? 0 : (int)(threadIdx.x >> 5), wpb = ONE_WARP ? 1 : (int)(blockDim.x >> 5);
  const uint64_t gw = (uint64_t)blockIdx.x * wpb + warp;
  const uint64_t tl0 = gw * 32ull;
  if (tl0 >= p.num_tries) return;  // whole warp leaves; no block-level barrier below
  const uint64_t tl = tl0 + lane;
  const bool tv = tl < p.num_tries;
  const uint64_t traj = p.first_try + tl;
  const int n = p.n;

  // shared memory per warp: [2][SP_HALF_CAP] staged entries, then x_words_per_warp spin words
  // (>= n + 1: the words from n on stay zero, the pads of the grouped layout point at word n)
  constexpr size_t STAGE_BYTES = 2 * SP_HALF_CAP * sizeof(Ent);
  const size_t per_warp = STAGE_BYTES + (size_t)x_words_per_warp * sizeof(uint32_t);
  unsigned char *mine = smem_raw + (size_t)warp * per_warp;
  Ent *stage = reinterpret_cast<Ent *>(mine);
  uint32_t *X = reinterpret_cast<uint32_t *>(mine + STAGE_BYTES);
  uint32_t *XB = p.xbest_ws + gw * (uint64_t)n;
  for (int j = n + lane; j < x_words_per_warp; j += 32) X[j] = 0u;

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
    // read-modify-write of the transposed workspace (other lanes' bits stay): eight independent
    // words in flight per lane -- one by one this loop is a chain of N/32 global round trips
    int j = lane;
    for (; j + 7 * 32 < n; j += 8 * 32) {
      uint32_t w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = XB[j + u * 32];
#pragma unroll
      for (int u = 0; u < 8; ++u) XB[j + u * 32] = (w[u] & ~who) | (X[j + u * 32] & who);
    }
    for (; j < n; j += 32) XB[j] = (XB[j] & ~who) | (X[j] & who);
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

  // the same bookkeeping without branches, for a decision that cannot overflow the log (the
  // batched path checks that before a group): every lane runs the same ~20 instructions, the lanes
  // that did not flip keep their state through selects
  auto track_flat = [&](bool acc, int site, T dE) {
    const double e = det::add(erel, (double)dE);
    const bool nb = acc && (e < best);            // new best (strict)
    const bool leave = acc && !nb && at_best;     // the flip leaves the best state
    const bool fresh = nb || leave;               // the log restarts
    const bool do_log = acc && !nb && (leave || !mat);
    erel = acc ? e : erel;
    best = nb ? e : best;
    log_len = fresh ? 0 : log_len;
    mat = fresh ? false : mat;
    at_best = nb ? true : (acc ? false : at_best);
    if (do_log) LOG[log_len * 32 + lane] = (uint16_t)site;
    log_len += do_log ? 1 : 0;
  };

  if (p.mode == OSA_MODE_SEQUENTIAL_SWEEP) {
    // Blocks of 32 sites (one Philox block per 4 sites, one trace update per block) = NG groups of
    // G sites, staged in halves of 16 sites.  Lane l <= NG holds the entry base of group l of the
    // block (lane NG: its end), lane l < NG the group's {len, independent}, lane l the diagonal of
    // site l; the same for the next block is loaded one block ahead.
    constexpr int NG = 32 / G, NGH = NG / 2;
    const int nblk = (n + 31) >> 5;
    const uint32_t lanebit = 1u << lane;
    const Ent *gent = reinterpret_cast<const Ent *>(p.gent);
    auto load_meta = [&](int b, uint32_t &gb, uint32_t &gi, T &dg) {
      gb = __ldg(p.gbase + b * NG + min(lane, NG));
      gi = lane < NG ? __ldg(p.ginfo + b * NG + lane) : 0u;
      const int i = b * 32 + lane;
      dg = (i < n) ? __ldg(p.diag + i) : (T)0;
    };
    auto prefetch = [&](uint32_t s0, uint32_t s1, int buf) {  // always commits one group
      if (STAGED) {
        for (uint32_t q = s0 + lane; q < s1; q += 32)
          asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(
                           sp_smem_addr(stage + buf * SP_HALF_CAP + (q - s0))),
                       "l"(gent + q), "n"((int)sizeof(Ent))
                       : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // Loads of the gather as volatile asm: the compiler keeps them in program order, i.e. BATCHED
    // (all entries of a row, then all spin words of the row).  Written as plain C++ the loads were
    // sunk next to their uses and a row became G/2 serial {LDS.128, LDS, FADD} round trips
    // (profiles/r02/ncu_k_sparse_grouped8_serial_r2e_summary.txt: 58 % of the samples on two
    // shared-memory latencies per pair of entries).
    const uint32_t xbase = sp_smem_addr(X);
    auto xword = [&](uint32_t off) {  // spin word of a neighbour: X + byte offset
      uint32_t w;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(xbase + off));
      return w;
    };
    auto ld16 = [&](const void *q) {
      uint4 v;
      if (STAGED)
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(sp_smem_addr(q)));
      else
        asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(q));
      return v;
    };
    // the G entries (t, 0..G-1) of a group / one entry, from the staging buffer or global memory
    auto load_row = [&](const Ent *e, Ent (&o)[G]) {
      if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int k = 0; k < G; k += 2) {
          const uint4 a = ld16(e + k);
          o[k].xoff = a.x; o[k].val = __uint_as_float(a.y);
          o[k + 1].xoff = a.z; o[k + 1].val = __uint_as_float(a.w);
        }
      } else {
#pragma unroll
        for (int k = 0; k < G; ++k) {
          const uint4 a = ld16(e + k);
          o[k].xoff = a.x;
          o[k].val = __hiloint2double((int)a.w, (int)a.z);
        }
      }
    };
    auto load1 = [&](const Ent *e, Ent &o) {
      if constexpr (sizeof(T) == 4) {
        uint2 a;
        if (STAGED)
          asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a.x), "=r"(a.y) : "r"(sp_smem_addr(e)));
        else
          asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(a.x), "=r"(a.y) : "l"(e));
        o.xoff = a.x; o.val = __uint_as_float(a.y);
      } else {
        const uint4 a = ld16(e);
        o.xoff = a.x;
        o.val = __hiloint2double((int)a.w, (int)a.z);
      }
    };

    uint32_t gb, gi, ngb, ngi;
    T dg, ndg;
    load_meta(0, gb, gi, dg);
    prefetch(__shfl_sync(0xffffffffu, gb, 0), __shfl_sync(0xffffffffu, gb, NGH), 0);
    unsigned hcount = 0;
    uint32_t step = 0;
    for (int iter = 0; iter < p.num_iter; ++iter) {
      const T ts = p.tscale[iter];
      for (int sw = 0; sw < p.sweeps_per_beta; ++sw, ++step) {
        for (int b = 0; b < nblk; ++b) {
          const int bn = (b + 1 == nblk) ? 0 : b + 1;  // next block (wraps into the next sweep)
          load_meta(bn, ngb, ngi, ndg);
          const int i_end = min(32, n - b * 32);
          uint32_t blk_acc = 0u;  // sites of this block that this lane's trajectory flipped
          // decision of site i (field hk in hand): accept test, bookkeeping, spin update
          auto decide = [&](int s, int i, T hk, T theta) {
            const uint32_t xiw = X[i];
            const T dE = (xiw & lanebit) ? -hk : hk;
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
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            const int buf = (int)(hcount & 1u);
            ++hcount;
            // entries of the half after this one: second half of this block / first of the next
            const uint32_t n0 = hh == 0 ? __shfl_sync(0xffffffffu, gb, NGH) : __shfl_sync(0xffffffffu, ngb, 0);
            const uint32_t n1 = hh == 0 ? __shfl_sync(0xffffffffu, gb, NG) : __shfl_sync(0xffffffffu, ngb, NGH);
            prefetch(n0, n1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");  // this half's entries landed
            __syncwarp();
            const uint32_t h0 = __shfl_sync(0xffffffffu, gb, hh * NGH);
#pragma unroll 1
            for (int g = hh * NGH; g < hh * NGH + NGH; ++g) {
              const int s0 = g * G;
              if (s0 >= i_end) break;
              // one Philox block serves four consecutive sites; the thresholds of the whole group
              // (chains of ~40 dependent operations each) are computed together so that they overlap
              T th[G];
#pragma unroll
              for (int k = 0; k < G; k += 4) {
                const U4 d = engine_draw(p.seed, traj, STREAM_SEQ, (uint32_t)(b * 32 + s0 + k) >> 2, step);
                th[k] = threshold<T>(ts, d.x);
                th[k + 1] = threshold<T>(ts, d.y);
                th[k + 2] = threshold<T>(ts, d.z);
                th[k + 3] = threshold<T>(ts, d.w);
              }
              const uint32_t base = __shfl_sync(0xffffffffu, gb, g);
              const uint32_t info = __shfl_sync(0xffffffffu, gi, g);
              const int len = (int)(info & 0xffffu);
              const Ent *e = STAGED ? stage + buf * SP_HALF_CAP + (base - h0) : gent + base;
              if (info & 0x10000u) {
                // pairwise non-adjacent sites: none of their flips changes the field of another,
                // so the G gathers run side by side and the G decisions are taken together
                T hk[G];
#pragma unroll
                for (int k = 0; k < G; ++k) hk[k] = __shfl_sync(0xffffffffu, dg, s0 + k);
                // software pipeline: the entries of row t + 1 are requested before the spin words
                // of row t (the row after the last one is read and dropped: it lies inside the
                // staging buffer / the padded entry array)
                Ent en[G];
                load_row(e, en);
                for (int t = 0; t < len; ++t) {
                  Ent nx[G];
                  load_row(e + (t + 1) * G, nx);
                  uint32_t w[G];
#pragma unroll
                  for (int k = 0; k < G; ++k) w[k] = xword(en[k].xoff);
#pragma unroll
                  for (int k = 0; k < G; ++k)
                    if (w[k] & lanebit) hk[k] = det::add(hk[k], en[k].val);
#pragma unroll
                  for (int k = 0; k < G; ++k) en[k] = nx[k];
                }
                // a log that could overflow inside this group is written out now (the state before
                // the group's flips, log undone) -- when it happens does not change any result
                const uint32_t spill = __ballot_sync(0xffffffffu, !at_best && !mat && log_len > SP_LOG - G);
                if (spill) {
                  materialize(spill, true);
                  if ((spill >> lane) & 1u) mat = true;
                }
                uint32_t xiw[G], bal[G];
#pragma unroll
                for (int k = 0; k < G; ++k) {
                  xiw[k] = X[b * 32 + s0 + k];
                  const T dE = (xiw[k] & lanebit) ? -hk[k] : hk[k];
                  const bool acc = tv && (dE < th[k]);
                  track_flat(acc, b * 32 + s0 + k, dE);
                  blk_acc |= acc ? (1u << (s0 + k)) : 0u;
                  bal[k] = __ballot_sync(0xffffffffu, acc);
                }
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                  for (int k = 0; k < G; ++k) X[b * 32 + s0 + k] = xiw[k] ^ bal[k];
                }
                __syncwarp();
              } else {
#pragma unroll
                for (int k = 0; k < G; ++k) {
                  const int s = s0 + k;
                  if (s >= i_end) break;
                  T hk = __shfl_sync(0xffffffffu, dg, s);
#pragma unroll 4
                  for (int t = 0; t < len; ++t) {
                    Ent en;
                    load1(e + t * G + k, en);
                    if (xword(en.xoff) & lanebit) hk = det::add(hk, en.val);
                  }
                  decide(s, b * 32 + s, hk, th[k]);
                }
              }
            }
            __syncwarp();  // everyone is done with stage buffer `buf` before it is refilled
          }
          if (blk_acc != 0u) trace = trace_step(trace, step, (uint32_t)b, blk_acc);
          cnt_acc += (unsigned)__popc(blk_acc);
          gb = ngb;
          gi = ngi;
          dg = ndg;
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
        ++cnt_acc;
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

int sp_x_words(int n) { return (n + 1 + 3) / 4 * 4; }  // >= n + 1: word n is the always-zero pad word

template <typename T>
size_t sp_per_warp_bytes(int n) {
  return 2 * SP_HALF_CAP * sizeof(SpEnt<T>) + (size_t)sp_x_words(n) * sizeof(uint32_t);
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

template <typename T, int G, bool ONE_WARP, bool STAGED>
cudaError_t launch_one(const SparseParams<T> &pp, int wpb, size_t smem, unsigned grid, int x_words,
                       cudaStream_t s) {
  cudaError_t err = cudaFuncSetAttribute(k_sparse<T, G, ONE_WARP, STAGED>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  k_sparse<T, G, ONE_WARP, STAGED><<<grid, wpb * 32, smem, s>>>(pp, x_words);
  return cudaGetLastError();
}

template <typename T, int G>
cudaError_t launch_g(const SparseParams<T> &pp, bool one_warp, bool staged, int wpb, size_t smem,
                     unsigned grid, int x_words, cudaStream_t s) {
  if (one_warp)
    return staged ? launch_one<T, G, true, true>(pp, wpb, smem, grid, x_words, s)
                  : launch_one<T, G, true, false>(pp, wpb, smem, grid, x_words, s);
  return staged ? launch_one<T, G, false, true>(pp, wpb, smem, grid, x_words, s)
                : launch_one<T, G, false, false>(pp, wpb, smem, grid, x_words, s);
}

template <typename T>
cudaError_t launch_impl(const SparseParams<T> &p, cudaStream_t s, LaunchInfo *info) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t per_warp = sp_per_warp_bytes<T>(p.n);
  if (per_warp > kMaxSmem) return cudaErrorInvalidValue;
  SparseParams<T> pp = p;
  pp.log_base = (size_t)((p.num_tries + 31) / 32) * (size_t)p.n;  // words; see sparse_ws_words
  pp.debug_flags = probe_env_int("OSA_SP_DEBUG");  // timing experiments (probe builds only)
  // one-warp CTAs while shared memory (not the limit of 32 resident CTAs) bounds the residency
  const bool one_warp = per_warp * 32 >= kMaxSmem;
  const int wpb = one_warp ? 1 : pick_wpb(per_warp, p.num_tries, sms);
  if (wpb < 1) return cudaErrorInvalidValue;
  const size_t smem = (size_t)wpb * per_warp;
  const int x_words = sp_x_words(p.n);
  const uint64_t warps = (p.num_tries + 31) / 32;
  const uint64_t grid64 = (warps + wpb - 1) / wpb;
  if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)grid64;
  const bool staged = p.stage_ok != 0;
  if (p.group != 4 && p.group != 8) return cudaErrorInvalidValue;
  const cudaError_t err = p.group == 8
                              ? launch_g<T, 8>(pp, one_warp, staged, wpb, smem, grid, x_words, s)
                              : launch_g<T, 4>(pp, one_warp, staged, wpb, smem, grid, x_words, s);
  if (info) {
    info->grid = (int)grid64;
    info->block = wpb * 32;
    info->traj_per_batch = 32;
    info->smem = smem;
  }
  return err;
}

}  // namespace

bool sparse_supported(int n, int elem_bytes) {
  if (n < 1) return false;
  return (elem_bytes == 4 ? sp_per_warp_bytes<float>(n) : sp_per_warp_bytes<double>(n)) <= kMaxSmem;
}

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
