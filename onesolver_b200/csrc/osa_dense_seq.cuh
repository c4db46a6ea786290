// osa_dense_seq.cuh -- pieces shared by the dense sequential-sweep kernels
// (osa_dense_seq.cu: single-role CTA; osa_dense_seq_ws.cu: warp-specialised decide/apply overlap).
#pragma once

#include <type_traits>

#include "osa_common.cuh"

#ifndef OSA_APPLY_VARIANT
#define OSA_APPLY_VARIANT 0
#endif

namespace osa {
namespace dseq {

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
  // if (flag) v[i] += m * q[i], as predicated instructions (no branch)
  __device__ __forceinline__ void axpy_if(uint32_t flag, double m, const double (&q)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i)
      asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p fma.rn.f64 %0, %1, %2, %0;\n\t}"
          : "+d"(v[i]) : "d"(q[i]), "d"(m), "r"(flag));
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
  // if (flag) v[i] += m * q[i], as predicated instructions (no branch)
  __device__ __forceinline__ void axpy_if(uint32_t flag, float m, const float (&q)[N]) {
    const unsigned long long mb = __float_as_uint(m);
    const unsigned long long mm = (mb << 32) | mb;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const unsigned long long qq =
          ((unsigned long long)__float_as_uint(q[2 * i + 1]) << 32) | __float_as_uint(q[2 * i]);
      asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p fma.rn.f32x2 %0, %1, %2, %0;\n\t}"
          : "+l"(p[i]) : "l"(qq), "l"(mm), "r"(flag));
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

// NCH cp.async copies of 16 bytes: shared address + c*SSTEP  <-  global address + c*GSTEP, with
// the offsets as immediates of the instruction.
template <int NCH, int SSTEP, int GSTEP, int C = 0>
struct CopyPieces {
  static __device__ __forceinline__ void issue(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0+%2], [%1+%3], 16;" ::"r"(dst), "l"(src),
                 "n"(C * SSTEP), "n"(C * GSTEP)
                 : "memory");
    CopyPieces<NCH, SSTEP, GSTEP, C + 1>::issue(dst, src);
  }
};
template <int NCH, int SSTEP, int GSTEP>
struct CopyPieces<NCH, SSTEP, GSTEP, NCH> {
  static __device__ __forceinline__ void issue(uint32_t, const void *) {}
};

// NCH reads of 16 bytes from the shared-memory address + c*SSTEP (offsets as immediates)
__device__ __forceinline__ void lds16(uint32_t addr, int off, float *o) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3])
               : "r"(addr + off));
}
__device__ __forceinline__ void lds16(uint32_t addr, int off, double *o) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "r"(addr + off));
}
template <typename T, int NCH, int SSTEP>
struct LoadPieces {
  static __device__ __forceinline__ void load(uint32_t addr, T *out) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) lds16(addr, c * SSTEP, out + c * Vec16<T>::V);
  }
};

template <typename T, int N>
__device__ __forceinline__ void keep_alive(const T (&q)[N]);
template <>
__device__ __forceinline__ void keep_alive<float, 16>(const float (&q)[16]) {
  asm volatile("" ::"f"(q[0]), "f"(q[1]), "f"(q[2]), "f"(q[3]), "f"(q[4]), "f"(q[5]), "f"(q[6]), "f"(q[7]),
               "f"(q[8]), "f"(q[9]), "f"(q[10]), "f"(q[11]), "f"(q[12]), "f"(q[13]), "f"(q[14]), "f"(q[15]));
}
template <typename T, int N>
__device__ __forceinline__ void keep_alive(const T (&)[N]) {}

// One streamed row applied to the fields of the R trajectories of a CTA.  bit: the site of the row
// within its block (one bit set); am[r] / sm[r]: accept and sign masks of the block (see apply_rows).
// The trajectories are handled in groups of G: a group is skipped with one uniform branch when none
// of its members flipped the site.  OSA_APPLY_VARIANT 0: every member of an active group runs
// h += m * q with m in {-1, 0, +1}; 1: one more uniform branch per member, no FMAs with m = 0.
template <typename T, int NCH, int R, int G>
__device__ __forceinline__ void apply_row(uint32_t bit, const uint32_t (&am)[R], const uint32_t (&sm)[R],
                                          const T (&qv)[NCH * Vec16<T>::V],
                                          Field<T, NCH * Vec16<T>::V> (&h)[R]) {
  static_assert(R % G == 0, "R must be a multiple of the group size");
#pragma unroll
  for (int gg = 0; gg < R / G; ++gg) {
#ifdef OSA_APPLY_REVERSE
    const int g = R / G - 1 - gg;
#else
    const int g = gg;
#endif
    uint32_t gm = 0u;
#pragma unroll
    for (int k = 0; k < G; ++k) gm |= am[g * G + k];
    if (gm & bit) {
#pragma unroll
      for (int k = 0; k < G; ++k) {
        const int r = g * G + k;
#if OSA_APPLY_VARIANT == 1
        if (am[r] & bit) h[r].axpy((sm[r] & bit) ? (T)-1 : (T)1, qv);
#else
        const T m = (am[r] & ~sm[r] & bit) ? (T)1 : ((sm[r] & bit) ? (T)-1 : (T)0);  // sm is a subset of am
        h[r].axpy(m, qv);
#endif
      }
    }
  }
  // keep the row alive past the last FMA: otherwise the FMAs of the last trajectory are allocated
  // onto the registers of the row and moved back to those of the field afterwards
  keep_alive<T, NCH * Vec16<T>::V>(qv);
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
  constexpr int ROW_BYTES = NCH * TH * 16;  // one ring slot
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

  // addresses: one 64-bit base per block, a 32-bit row offset per row, immediates per piece;
  // the ring is walked with wrapping 32-bit shared-memory addresses (no per-row multiplications).
  // The two bases go through an empty asm so that the compiler keeps them in registers: left to
  // itself it recomputes both from the kernel parameters and the thread id for every row.
  const unsigned char *base =
      reinterpret_cast<const unsigned char *>(qoff + (size_t)i0 * ld + (size_t)tid * V);
  uint32_t ring_lo = smem_addr(ring + (size_t)tid * 16);
  asm volatile("" : "+l"(base), "+r"(ring_lo));
  const uint32_t row_bytes = (uint32_t)(ld * sizeof(T));
  const uint32_t ring_hi = ring_lo + (uint32_t)K * ROW_BYTES;
  uint32_t r_addr = ring_lo, w_addr = ring_lo;
  uint32_t rem_issue = any, rem_apply = any;

  // ring fill: the first K-1 requests of the block (fewer when the block has fewer rows)
#pragma unroll 1
  for (int k = 0; k < K - 1; ++k) {
    if (rem_issue) {
      const int s = __ffs(rem_issue) - 1;
      rem_issue &= rem_issue - 1;
      const unsigned char *rp = base + (uint32_t)s * row_bytes;
      CopyPieces<NCH, TH * 16, CHW * (int)sizeof(T)>::issue(w_addr, rp);
      w_addr += ROW_BYTES;
      if (w_addr == ring_hi) w_addr = ring_lo;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // one group per step, empty or not
  }

  // In the loop the address of a request is prepared one step ahead (rp_next), so that the
  // ffs -> multiply -> add chain is off the path of the copies.  (Doing the same in the fill
  // was measured slower: profiles/r01/probe_v51*.log.)
  const unsigned char *rp_next = base + (uint32_t)(__ffs(rem_issue) - 1) * row_bytes;
  auto request = [&]() {  // requires rem_issue != 0
    CopyPieces<NCH, TH * 16, CHW * (int)sizeof(T)>::issue(w_addr, rp_next);
    w_addr += ROW_BYTES;
    if (w_addr == ring_hi) w_addr = ring_lo;
    rem_issue &= rem_issue - 1;
    rp_next = base + (uint32_t)(__ffs(rem_issue) - 1) * row_bytes;  // unused when none is left
  };

  // one step: wait for the oldest row, read it, (request one more row,) apply it
  auto step = [&](auto more) {
    // row i has landed when at most K-2 younger groups are pending; read it into registers
    // first, so that the latency of the shared-memory loads hides behind the address
    // arithmetic of the next copies
    asm volatile("cp.async.wait_group %0;" ::"n"(K - 2) : "memory");
    T qv[NCH * V];
    LoadPieces<T, NCH, TH * 16>::load(r_addr, qv);
    r_addr += ROW_BYTES;
    if (r_addr == ring_hi) r_addr = ring_lo;
    if constexpr (decltype(more)::value) request();
    asm volatile("cp.async.commit_group;" ::: "memory");  // one group per step, empty or not
    const uint32_t bit = rem_apply & (0u - rem_apply);  // lowest site not yet applied
    rem_apply ^= bit;
    if (DBG == 1) {  // timing experiment: touch the data, skip the arithmetic
      T acc = (T)0;
#pragma unroll
      for (int e = 0; e < NCH * V; ++e) acc += qv[e];
      if (acc == (T)123456789) h[0].set(0, acc);
      return;
    }
#pragma unroll
    for (int g = 0; g < R / G; ++g) {
      if (grp[g] & bit) {
#pragma unroll
        for (int k = 0; k < G; ++k) {
          const int r = g * G + k;
#if OSA_APPLY_VARIANT == 1    // one branch per member: no FMAs with a zero multiplier
          if (am[r] & bit) h[r].axpy((neg[r] & bit) ? (T)-1 : (T)1, qv);
#elif OSA_APPLY_VARIANT == 2  // predicated members: no FMAs with a zero multiplier, no branch
          h[r].axpy_if(am[r] & bit, (neg[r] & bit) ? (T)-1 : (T)1, qv);
#else
          const T m = (pos[r] & bit) ? (T)1 : ((neg[r] & bit) ? (T)-1 : (T)0);
          h[r].axpy(m, qv);
#endif
        }
      }
    }
  };
  // while rows are left to request every step requests one; the last K-1 rows only drain
#pragma unroll 1
  while (rem_issue) step(std::true_type{});
#pragma unroll 1
  while (rem_apply) step(std::false_type{});
  return any;
}


}  // namespace dseq
}  // namespace osa
