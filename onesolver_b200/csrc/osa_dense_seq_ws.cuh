// osa_dense_seq_ws.cuh -- helpers shared by the warp-specialised dense sweep kernels
// (osa_dense_seq_ws.cu: lock-step decide/apply overlap; osa_dense_seq_ws2.cu: free-running roles).
#pragma once

#include "osa_dense_seq.cuh"

namespace osa {
namespace dsws {

using namespace dseq;

constexpr int WS_APPLY_THREADS = 256;

// Register budgets of the two roles for DW decide warps.  The kernel is launched with
// LAUNCH = 65536 / threads registers per thread (multiple of 8); the decide warps give up
// LAUNCH - DECIDE each, and setmaxnreg.inc can only draw from what the CTA's own warps released,
// so the apply warps get exactly LAUNCH + (LAUNCH - DECIDE) * decide_threads / apply_threads:
//   DW = 4: 384 threads, 168 -> apply 224 / decide 56   (shapes with >= 192 field registers)
//   DW = 8: 512 threads, 128 -> apply 200 / decide 56   (small N: the decisions are the bottleneck)
// (Round 2, profiles/r02/ab_config3_two_ctas.txt: two CTAs per SM -- 384 threads at 80 registers,
// apply 88 / decide 56, R = 8 or 6 at N = 1024 fp64 -- are 8 % / 18 % slower than one CTA with
// R = 16: every CTA streams its own rows.  DW has to stay a multiple of 4: setmaxnreg is executed
// by whole warpgroups, and a decide role of two warps -- half a warpgroup -- hangs the kernel.)
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

// Sign / mask arithmetic on the bit patterns of the fields.  The walk of the decide warps keeps
// its dependency chain free of predicates (a predicate costs ~13 cycles from its compare to the
// instruction it guards): "dE < theta" is the sign bit of the rounded difference dE - theta
// (exact in sign: without flush-to-zero a difference of two different numbers never rounds to
// zero, and theta > 0), turned into an all-ones / all-zeros mask; an accepted flip multiplies the
// tile row by +-1.0, a rejected one by +0.0 (h + 0*q == h).
template <typename T>
struct Bits;
template <>
struct Bits<float> {
  static __device__ __forceinline__ float neg_if(float x, uint32_t bit) {  // bit ? -x : x
    return __uint_as_float(__float_as_uint(x) ^ (bit << 31));
  }
  static __device__ __forceinline__ uint32_t neg_mask(float d) {  // sign bit set ? ~0 : 0
    return (uint32_t)((int)__float_as_uint(d) >> 31);
  }
  static __device__ __forceinline__ float unit(uint32_t bit, uint32_t mask) {  // mask ? (bit ? -1 : 1) : 0
    return __uint_as_float((0x3f800000u | (bit << 31)) & mask);
  }
  static __device__ __forceinline__ float masked(float x, uint32_t mask) {
    return __uint_as_float(__float_as_uint(x) & mask);
  }
  static __device__ __forceinline__ float bcast(float m, int src, uint32_t mask) {
    return __uint_as_float(__shfl_sync(0xffffffffu, __float_as_uint(m), src) & mask);
  }
};
template <>
struct Bits<double> {
  static __device__ __forceinline__ double neg_if(double x, uint32_t bit) {
    return __hiloint2double(__double2hiint(x) ^ (int)(bit << 31), __double2loint(x));
  }
  static __device__ __forceinline__ uint32_t neg_mask(double d) {
    return (uint32_t)(__double2hiint(d) >> 31);
  }
  static __device__ __forceinline__ double unit(uint32_t bit, uint32_t mask) {
    return __hiloint2double((int)((0x3ff00000u | (bit << 31)) & mask), 0);
  }
  static __device__ __forceinline__ double masked(double x, uint32_t mask) {
    return __hiloint2double((int)((uint32_t)__double2hiint(x) & mask),
                            (int)((uint32_t)__double2loint(x) & mask));
  }
  static __device__ __forceinline__ double bcast(double m, int src, uint32_t mask) {  // m is +-1 or 0
    return __hiloint2double((int)(__shfl_sync(0xffffffffu, (uint32_t)__double2hiint(m), src) & mask), 0);
  }
};

// HS consecutive elements from a (HS * sizeof(T))-byte aligned shared-memory address
template <typename T, int HS>
__device__ __forceinline__ void load_seg(const T *src, T (&out)[HS]) {
  constexpr int BYTES = HS * (int)sizeof(T);
  if constexpr (BYTES % 16 == 0) {
    constexpr int V = 16 / (int)sizeof(T);
#pragma unroll
    for (int i = 0; i < HS; i += V)
      vec_unpack<T>(*reinterpret_cast<const typename Vec16<T>::type *>(src + i), &out[i]);
  } else if constexpr (BYTES == 8 && sizeof(T) == 4) {
    const float2 v = *reinterpret_cast<const float2 *>(src);
    out[0] = v.x;
    out[1] = v.y;
  } else {
#pragma unroll
    for (int i = 0; i < HS; ++i) out[i] = src[i];
  }
}

// Shared memory of the CTA next to the row ring, as structs, so that the launcher can size the
// ring (K rows) from what is left of the 227 KiB.  WsTiles follows the ring in dynamic shared
// memory (statically allocated shared memory is limited to 48 KiB).
template <typename T>
struct WsTiles {
  static constexpr int TP = 32 + 16 / (int)sizeof(T);  // tile row pitch: 16-byte aligned rows whose
                                                       // 16-byte chunks rotate through the banks
  alignas(16) T tile_d[2][32][TP];  // diagonal tile of block j (parity j&1)
  alignas(16) T tile_x[2][32][TP];  // rows of block j-1 x columns of block j
};

template <typename T, int R, int NWP>
struct WsShared {
  alignas(16) T snap[2][32][R];     // columns of block j (parity j&1) after block j-2, [column][traj]
  alignas(16) T theta[2][32][R];    // acceptance thresholds of block j, [site][traj]
  T dE[32][R];                      // dE of the accepted flips of the block being decided
  double erel[R], best[R];          // running / best energy relative to the start
  T ts[R];                          // per-trajectory threshold scale (only with p.tscale_traj)
  uint32_t atbest[R];
  uint32_t naccept[R];              // accepted flips (32-bit; flushed to the 64-bit counter)
  uint32_t x[NWP][R];               // current spins, [word][traj]
};

constexpr int WS_SMEM_LIMIT = 232448 - 1024 - 256;  // 227 KiB minus the per-CTA reservation

template <typename T, int NCH, int R, int K>
struct WsRing {
  using C = Cfg<T, NCH, R, WS_APPLY_THREADS>;
  static constexpr int ROW_BYTES = NCH * WS_APPLY_THREADS * 16;
  static constexpr int FIT =
      (WS_SMEM_LIMIT - (int)sizeof(WsShared<T, R, C::NWP>) - 16 * R - (int)sizeof(WsTiles<T>)) / ROW_BYTES;
  static constexpr int KE = K < FIT ? K : FIT;  // rows in flight
  static_assert(KE >= 3, "row ring too small");
};

}  // namespace dsws
}  // namespace osa
