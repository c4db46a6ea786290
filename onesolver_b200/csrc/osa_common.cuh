// osa_common.cuh -- shared device helpers: Philox4x32-10, deterministic -ln(u),
// rounding-explicit arithmetic, internal problem/launch structs.
//
// Every operation whose rounding matters for the bit-exact host replay goes
// through det::fma / det::mul (explicit __f*_rn intrinsics, never contracted).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/onesolver_b200.h"

namespace osa {

// ---------------------------------------------------------------------------
// RNG streams.  key = (seed lo, seed hi); ctr = (c0, c1, traj lo, traj hi | stream<<30)
//   STREAM_INIT: c0 = j>>7, c1 = 0        -> word (j>>5)&3, bit j&31 = initial spin j
//                (replaces random.bit(), reference annealing.hpp:90-92)
//   STREAM_SEQ : c0 = site>>2, c1 = sweep -> word site&3 = uniform for (sweep, site)
//   STREAM_RND : c0 = 0, c1 = step        -> word0 = site draw, word1 = uniform
//   STREAM_PT  : c0 = pair, c1 = round (traj = group id) -> word0 = uniform of a replica swap
//                (replaces bit_index()/uniform(), reference annealing.hpp:101,108)
// ---------------------------------------------------------------------------
enum : uint32_t { STREAM_INIT = 0, STREAM_SEQ = 1, STREAM_RND = 2, STREAM_PT = 3 };

struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}

__device__ __forceinline__ U4 engine_draw(uint64_t seed, uint64_t traj, uint32_t stream,
                                          uint32_t c0, uint32_t c1) {
  return philox4x32_10(c0, c1, (uint32_t)traj,
                       ((uint32_t)(traj >> 32) & 0x3fffffffu) | (stream << 30), (uint32_t)seed,
                       (uint32_t)(seed >> 32));
}

__device__ __forceinline__ uint32_t pick(const U4 &v, uint32_t i) {
  return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// -ln(u), u = (2w+1)/2^33 in (0,1).  24-bit truncated mantissa + the classic
// single-precision minimax polynomial for ln(1+f); explicit rn intrinsics only.
__device__ __forceinline__ float neglogf_det(uint32_t w) {
  const uint64_t v = ((uint64_t)w << 1) | 1ull;
  const int p = 63 - __clzll((long long)v);
  const uint32_t m24 = (uint32_t)((v << (63 - p)) >> 40);
  int e = p - 33;
  float mf = __fmul_rn(__uint2float_rn(m24), 1.1920928955078125e-07f);  // * 2^-23 (exact)
  if (m24 > 0x00B504F3u) {
    mf = __fmul_rn(mf, 0.5f);
    e += 1;
  }
  const float f = __fsub_rn(mf, 1.0f);
  const float z = __fmul_rn(f, f);
  float y = 7.0376836292E-2f;
  y = __fmaf_rn(y, f, -1.1514610310E-1f);
  y = __fmaf_rn(y, f, 1.1676998740E-1f);
  y = __fmaf_rn(y, f, -1.2420140846E-1f);
  y = __fmaf_rn(y, f, 1.4249322787E-1f);
  y = __fmaf_rn(y, f, -1.6668057665E-1f);
  y = __fmaf_rn(y, f, 2.0000714765E-1f);
  y = __fmaf_rn(y, f, -2.4999993993E-1f);
  y = __fmaf_rn(y, f, 3.3333331174E-1f);
  y = __fmul_rn(y, f);
  y = __fmul_rn(y, z);
  const float fe = __int2float_rn(e);
  y = __fmaf_rn(fe, -2.12194440e-4f, y);
  y = __fmaf_rn(-0.5f, z, y);
  float r = __fadd_rn(f, y);
  r = __fmaf_rn(fe, 0.693359375f, r);
  return -r;
}

namespace det {
__device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
}  // namespace det

// threshold theta = tscale * (T)(-ln u)
template <typename T>
__device__ __forceinline__ T threshold(T ts, uint32_t w) {
  return det::mul(ts, (T)neglogf_det(w));
}

// 16-byte vector type per element type
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  using type = float4;
  static constexpr int V = 4;
};
template <> struct Vec16<double> {
  using type = double2;
  static constexpr int V = 2;
};

// Flip trace of a trajectory (osa_anneal_traced): 64-bit FNV-1a over the accepted flips, one update
// per (step, block of 32 sites) with at least one accepted flip, in the order they happen:
//   h = (h ^ step) * P;  h = (h ^ ((block << 32) | mask)) * P
// step = the counter of the random stream (sequential mode: the sweep number; random-site mode: the
// attempt number, block = site / 32 and mask = the one bit of the site), mask = the accepted sites
// of the block.  Within a block of a sweep the sites are visited in ascending order, so the hash
// pins the complete spin sequence.  The host replay (oracle/osa_oracle.c) computes the same value.
constexpr unsigned long long TRACE_OFFSET = 0xcbf29ce484222325ull, TRACE_PRIME = 0x100000001b3ull;
__host__ __device__ __forceinline__ unsigned long long trace_step(unsigned long long h, uint32_t step,
                                                                  uint32_t block, uint32_t mask) {
  h = (h ^ (unsigned long long)step) * TRACE_PRIME;
  return (h ^ (((unsigned long long)block << 32) | mask)) * TRACE_PRIME;
}

template <typename T>
__device__ __forceinline__ void vec_unpack(const typename Vec16<T>::type &v, T *out);
template <>
__device__ __forceinline__ void vec_unpack<float>(const float4 &v, float *o) {
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <>
__device__ __forceinline__ void vec_unpack<double>(const double2 &v, double *o) {
  o[0] = v.x; o[1] = v.y;
}
template <typename T>
__device__ __forceinline__ typename Vec16<T>::type vec_pack(const T *v);
template <>
__device__ __forceinline__ float4 vec_pack<float>(const float *v) {
  return make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ double2 vec_pack<double>(const double *v) {
  return make_double2(v[0], v[1]);
}

// ---------------------------------------------------------------------------
// launch parameter blocks (host fills, kernels read)
// ---------------------------------------------------------------------------
struct Counters {  // device-side accumulators, one struct per call
  unsigned long long accepts;
  unsigned long long row_fetches;
  unsigned long long init_row_fetches;
  unsigned long long cyc_decide;  // diagnostics: SM cycles (thread 0 of each CTA, summed) in P1
  unsigned long long cyc_apply;   //              ... in P2 (row streaming)
  unsigned long long cyc_stage;   //              ... staging tile/panel + barriers
  unsigned long long cyc_init;    //              ... building the initial fields
  unsigned long long pad;
};

template <typename T>
struct DenseParams {
  const T *qoff;   // [n_rows_pad][ld], zero diagonal, symmetric, zero padded
  const T *diag;   // [ld], zero padded
  const T *tscale; // [num_iter]
  size_t ld;
  int n;
  int num_iter;
  int sweeps_per_beta;
  int mode;
  uint64_t seed;
  uint64_t first_try;
  uint64_t num_tries;
  double *best_rel;         // [num_tries]
  uint32_t *best_states;    // [num_tries][nw]
  int nw;                   // words per state = ceil(n/32)
  Counters *counters;
  // resumable rounds (parallel tempering, osa_pt_anneal); all optional, zero = plain annealing.
  // Honoured by k_dense_seq_ws only.
  const uint32_t *init_states;  // [num_tries][nw]: start from these spins instead of STREAM_INIT
  const T *tscale_traj;         // [num_tries]: per-trajectory threshold scale, replaces tscale[iter]
  uint32_t *final_states;       // [num_tries][nw]: spins after the last sweep (may alias init_states)
  uint32_t step_base;           // first sweep number of this launch in the STREAM_SEQ counter
  // local fields carried from launch to launch, [num_tries][ld] in the sweep precision: with
  // fields_in the launch starts from them instead of rebuilding the fields from the spins (the
  // rebuild streams every row once: more arithmetic than two sweeps); fields_out receives the
  // fields after the last sweep (may alias fields_in)
  const T *fields_in;
  T *fields_out;
  // optional [num_tries]: FNV-1a hash of the trajectory's accepted flips (osa_anneal_traced)
  unsigned long long *trace_hash;
  // timing experiments only (OSA_WS_DEBUG, tools/probe.py; results are meaningless when set):
  // 1 = the apply warps skip the row streaming (decide warps alone), 2 = the decide warps emit
  // pseudo-random accept masks of density 0.19 instead of deciding (apply warps alone),
  // 8 = cyc_init / cyc_stage report two phases of the walker warp (bringing the snapshot up to
  // date, walking the sites) instead of the apply-role timers
  int debug_flags;
};

template <typename T>
struct SparseParams {
  const int32_t *rowptr;
  const int32_t *col;
  const T *val;
  const T *diag;
  const T *tscale;
  int n;
  int num_iter;
  int sweeps_per_beta;
  int mode;
  uint64_t seed;
  uint64_t first_try;
  uint64_t num_tries;
  double *best_rel;
  uint32_t *best_states;  // [num_tries][nw]
  uint32_t *xbest_ws;     // [n_warps_total][n] transposed best-state workspace
  size_t log_base;        // word offset of the flip logs inside xbest_ws (set by the launcher)
  int debug_flags;        // timing experiments only (tools/probe.py): 1 = skip best-state snapshots
  int nw;
  Counters *counters;
  unsigned long long *trace_hash;  // optional [num_tries], see trace_step()
  // grouped layout of the sequential sweeps (osa_sparse.cu): groups of `group` consecutive sites
  int group;              // 4 or 8
  const uint32_t *gbase;  // [(32/group) * ceil(n/32) + 1] first entry of a group
  const uint32_t *ginfo;  // [(32/group) * ceil(n/32)] len | independent << 16
  const void *gent;       // SpEnt<T>[gbase[last]]: entry (t, k) of group g at gbase[g] + group t + k
  int stage_ok;           // every half block (4 groups) fits the staging buffer
};

// Timing-experiment switches that make the results of a call meaningless (skipped row streaming,
// faked accept masks, loads-only instantiations) exist only in probe builds: -DOSA_PROBE, set by
// tools/build_variant.sh for the libraries under build/ab/.  The in-tree library ignores them, so a
// stray environment variable cannot corrupt results; only result-preserving tuning knobs
// (OSA_DS_WS, OSA_WS_DW, OSA_WS_R, OSA_ENERGY_MMA) are read from the environment there.
inline int probe_env_int(const char *name) {
#ifdef OSA_PROBE
  const char *e = getenv(name);
  return e ? atoi(e) : 0;
#else
  (void)name;
  return 0;
#endif
}

// kernel ids reported in osa_stats.kernel_id
enum KernelId : int {
  KID_AUTO = 0,
  KID_DENSE_SEQ = 1,      // CTA-per-batch sequential sweep, h in registers
  KID_DENSE_GENERIC = 2,  // warp-per-trajectory, h in shared memory (both modes)
  KID_SPARSE = 3          // lane-per-trajectory CSR kernel (both modes)
};

// launchers implemented in the per-kernel .cu files; all return cudaError_t
struct LaunchInfo { int grid; int block; int traj_per_batch; size_t smem; };

template <typename T>
cudaError_t launch_dense_seq(const DenseParams<T> &p, cudaStream_t s, LaunchInfo *info);
template <typename T>
cudaError_t launch_dense_seq_ws(const DenseParams<T> &p, cudaStream_t s, LaunchInfo *info);
template <typename T>
cudaError_t launch_dense_seq_flow(const DenseParams<T> &p, cudaStream_t s, LaunchInfo *info);
template <typename T>
cudaError_t launch_dense_generic(const DenseParams<T> &p, cudaStream_t s, LaunchInfo *info);
// initial fields of all trajectories into p.fields_out with shared row fetches (osa_dense_init.cu);
// n within dense_seq_supported.  *traj_per_batch: trajectories that share a fetch.
template <typename T>
cudaError_t launch_dense_init_fields(const DenseParams<T> &p, cudaStream_t s, int *traj_per_batch);
template <typename T>
cudaError_t launch_sparse(const SparseParams<T> &p, cudaStream_t s, LaunchInfo *info);
cudaError_t launch_pt_init_states(uint64_t seed, uint64_t first_try, uint64_t num_tries, int n,
                                  int nw, uint32_t *states, cudaStream_t s);
cudaError_t launch_pt_track_best(const double *e_start, const double *best_rel,
                                 const uint32_t *round_best, uint64_t num_tries, int nw,
                                 double *best_e, uint32_t *best_keep, cudaStream_t s);
template <typename T>
cudaError_t launch_pt_swap(uint64_t seed, uint64_t first_group, uint64_t num_groups, int replicas,
                           uint32_t round, const double *e_cur, const double *dinv,
                           const T *tscale_of_temp, int32_t *temp_of_slot, int32_t *slot_of_temp,
                           T *tscale_traj, unsigned long long *swap_count, cudaStream_t s);
// population annealing (osa_pa.cu): integer weights + prefix sums, then the resampling copy
cudaError_t launch_pa_resample(uint64_t seed, uint64_t first_pop, uint64_t num_pops, int M,
                               uint32_t step, double neg_db, const uint32_t *cur,
                               const double *e_cur, int nw, unsigned long long *cum, uint32_t *nxt,
                               double *e_nxt, int32_t *src_out, unsigned long long *replaced,
                               const char *fields, char *fields_nxt, size_t field_bytes,
                               cudaStream_t s);
template <typename T>
cudaError_t launch_pa_fill(T *dst, T value, uint64_t count, cudaStream_t s);
bool dense_seq_supported(int n, int elem_bytes);
bool dense_generic_supported(int n, int elem_bytes);
size_t sparse_ws_words(int n, uint64_t num_tries);
bool sparse_supported(int n, int elem_bytes);
constexpr int SPARSE_HALF_CAP = 256;  // = SP_HALF_CAP of osa_sparse.cu

// exact fp64 energies (reference formula) of packed states
cudaError_t launch_energy_dense(const double *q64, size_t ld64, int n, const uint32_t *states,
                                int nw, uint64_t count, double *out, cudaStream_t s);
cudaError_t launch_energy_csr(const int32_t *rowptr, const int32_t *col, const double *val64,
                              const double *diag64, int n, const uint32_t *states, int nw,
                              uint64_t count, double *out, cudaStream_t s);
// argmin with lowest-index tie break; result: out_idx[0] = local index, out_e[0] = energy
cudaError_t launch_argmin(const double *e, uint64_t count, unsigned long long *out_idx,
                          double *out_e, cudaStream_t s);
cudaError_t launch_read_bw(const uint4 *buf, size_t n_vec, int iters, unsigned int *sink,
                           int grid, cudaStream_t s);

}  // namespace osa
