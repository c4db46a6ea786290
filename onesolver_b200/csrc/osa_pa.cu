// osa_pa.cu -- population annealing around the dense sweep kernel.
//
// The reference's benchmark report names population annealing first among the samplers it
// recommends (/root/reference/benchmarks/annealing/performance.md:54-59); there is no reference
// code for it.  A run is P independent populations of M replicas; replica slot k of population g
// is trajectory g*M + k of the resumable sweep kernel (the one parallel tempering uses).  Step t:
// every replica does S sequential sweeps at betas[t]; the exact fp64 energies of the states they
// end in are recomputed; then the population is RESAMPLED for the next temperature with weights
//     w_i = exp(-(b_{t+1} - b_t) * E_i),   b = inverse temperature of the acceptance rule,
// keeping its size: systematic resampling, one uniform per (population, step).
//
// Everything is arranged so that the result does not depend on the order of any floating-point
// reduction (and the host restatement, oracle/osa_oracle.c orc_pa_resample, is a plain loop):
//   x_i = -(db) * E_i,  x_max = max_i x_i            (a maximum is order independent)
//   q_i = floor(2^40 * det_exp(x_i - x_max))         (integer weight, 0 .. 2^40)
//   C_i = q_0 + .. + q_i                             (integer prefix sums: any scan order)
//   r   = floor(W * u32 / 2^32),  W = C_{M-1}        (offset from Philox, < W)
//   slot j continues from replica src_j = min{ i : C_i > floor((j * W + r) / M) }
// det_exp is a fixed sequence of correctly rounded operations (range reduction + degree-13
// polynomial), identical on host and device.
#include "osa_common.cuh"

namespace osa {

namespace {

constexpr uint32_t PA_C0 = 0x50410000u;  // c0 of STREAM_PT draws that belong to resampling steps

// exp(x) for x <= 0 to ~1 ulp, deterministic: k = rint(x * log2 e), r = x - k ln2 (two-part ln2),
// Horner with fma, exact scaling by 2^k.  Below -60 the 40-bit weight is 0 anyway.
__device__ __forceinline__ double det_exp(double x) {
  if (!(x >= -60.0)) return 0.0;
  const double kf = rint(det::mul(x, 1.4426950408889634074));
  double r = det::fma(-kf, 6.93147180369123816490e-01, x);
  r = det::fma(-kf, 1.90821492927058770002e-10, r);
  double p = 1.6059043836821613e-10;         // 1/13!
  p = det::fma(p, r, 2.08767569878681e-09);  // 1/12!
  p = det::fma(p, r, 2.505210838544172e-08); // 1/11!
  p = det::fma(p, r, 2.755731922398589e-07); // 1/10!
  p = det::fma(p, r, 2.7557319223985893e-06);  // 1/9!
  p = det::fma(p, r, 2.48015873015873e-05);    // 1/8!
  p = det::fma(p, r, 1.984126984126984e-04);   // 1/7!
  p = det::fma(p, r, 1.388888888888889e-03);   // 1/6!
  p = det::fma(p, r, 8.333333333333333e-03);   // 1/5!
  p = det::fma(p, r, 4.1666666666666664e-02);  // 1/4!
  p = det::fma(p, r, 1.6666666666666666e-01);  // 1/3!
  p = det::fma(p, r, 0.5);
  p = det::fma(p, r, 1.0);
  p = det::fma(p, r, 1.0);
  const long long k = (long long)kf;  // -87 .. 0
  return det::mul(p, __longlong_as_double((1023ll + k) << 52));
}

// One CTA per population: integer weights and their inclusive prefix sums.
__global__ void __launch_bounds__(1024) k_pa_weights(const double *__restrict__ e_cur, int M,
                                                     double neg_db,
                                                     unsigned long long *__restrict__ cum) {
  __shared__ double s_max[32];
  __shared__ unsigned long long s_part[32];
  __shared__ unsigned long long s_carry;
  const uint64_t base = (uint64_t)blockIdx.x * (uint64_t)M;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // x_max
  double mx = -INFINITY;
  for (int i = tid; i < M; i += blockDim.x) mx = fmax(mx, det::mul(neg_db, e_cur[base + i]));
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_max[warp] = mx;
  if (tid == 0) s_carry = 0ull;
  __syncthreads();
  mx = s_max[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, s_max[w]);
  // chunks of blockDim.x replicas: block-wide inclusive scan of the integer weights
  for (int i0 = 0; i0 < M; i0 += blockDim.x) {
    const int i = i0 + tid;
    unsigned long long q = 0ull;
    if (i < M) {
      const double x = det::add(det::mul(neg_db, e_cur[base + i]), -mx);
      q = __double2ull_rz(det::mul(det_exp(x), 1099511627776.0));  // 2^40
    }
    unsigned long long v = q;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) s_part[warp] = v;
    __syncthreads();
    unsigned long long before = s_carry;
    for (int w = 0; w < warp; ++w) before += s_part[w];
    if (i < M) cum[base + i] = before + v;
    __syncthreads();
    if (tid == blockDim.x - 1) s_carry = before + v;
    __syncthreads();
  }
}

// One warp per replica slot j: source replica by binary search, then the copy of its state.
__global__ void k_pa_resample(uint64_t seed, uint64_t first_pop, uint64_t num_pops, int M,
                              uint32_t step, const unsigned long long *__restrict__ cum,
                              const uint32_t *__restrict__ cur, const double *__restrict__ e_cur,
                              int nw, uint32_t *__restrict__ nxt, double *__restrict__ e_nxt,
                              int32_t *__restrict__ src_out, unsigned long long *replaced,
                              const uint4 *__restrict__ fields, uint4 *__restrict__ fields_nxt,
                              size_t field_vecs) {
  const uint64_t slot = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (slot >= num_pops * (uint64_t)M) return;
  const uint64_t g = slot / (uint64_t)M;
  const uint32_t j = (uint32_t)(slot % (uint64_t)M);
  const uint64_t base = g * (uint64_t)M;
  const unsigned long long W = cum[base + M - 1];  // >= 2^40: the replica at x_max weighs 2^40
  const U4 d = engine_draw(seed, first_pop + g, STREAM_PT, PA_C0, step);
  // r = floor(W * u32 / 2^32) < W;  pos = floor((j * W + r) / M) < W   (128-bit intermediate)
  const unsigned long long r = (__umul64hi(W, (unsigned long long)d.x) << 32) |
                               ((W * (unsigned long long)d.x) >> 32);
  unsigned long long lo = (unsigned long long)j * W, hi = __umul64hi((unsigned long long)j, W);
  lo += r;
  hi += (lo < r) ? 1ull : 0ull;
  // (hi:lo) / M with hi < M  (j < M and W < 2^61): two-step long division in base 2^32
  const unsigned long long m = (unsigned long long)M;
  unsigned long long rem = hi;  // < M <= 2^31
  const unsigned long long n1 = (rem << 32) | (lo >> 32);
  const unsigned long long q1 = n1 / m;
  rem = n1 % m;
  const unsigned long long n0 = (rem << 32) | (lo & 0xffffffffull);
  const unsigned long long pos = (q1 << 32) | (n0 / m);
  // smallest i with C_i > pos
  int a = 0, b = M - 1;
  while (a < b) {
    const int mid = (a + b) >> 1;
    if (cum[base + mid] > pos) b = mid; else a = mid + 1;
  }
  const uint64_t from = base + (uint64_t)a;
  for (int k = lane; k < nw; k += 32) nxt[slot * (uint64_t)nw + k] = cur[from * (uint64_t)nw + k];
  // the local fields travel with the replica (16-byte pieces, eight in flight per lane)
  if (fields) {
    const uint4 *src = fields + from * field_vecs;
    uint4 *dst = fields_nxt + slot * field_vecs;
    size_t k = lane;
    for (; k + 7 * 32 < field_vecs; k += 8 * 32) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[k + u * 32];
#pragma unroll
      for (int u = 0; u < 8; ++u) dst[k + u * 32] = v[u];
    }
    for (; k < field_vecs; k += 32) dst[k] = src[k];
  }
  if (lane == 0) {
    e_nxt[slot] = e_cur[from];
    if (src_out) src_out[slot] = a;
    if ((uint32_t)a != j) atomicAdd(replaced, 1ull);
  }
}

template <typename T>
__global__ void k_pa_fill(T *__restrict__ dst, T value, uint64_t count) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = value;
}

}  // namespace

// threshold scale of a step for every replica (the resumable sweep kernel takes it per trajectory)
template <typename T>
cudaError_t launch_pa_fill(T *dst, T value, uint64_t count, cudaStream_t s) {
  const uint64_t grid = (count + 255) / 256;
  if (grid == 0 || grid > 0x7fffffffull) return cudaErrorInvalidValue;
  k_pa_fill<T><<<(unsigned)grid, 256, 0, s>>>(dst, value, count);
  return cudaGetLastError();
}
template cudaError_t launch_pa_fill<float>(float *, float, uint64_t, cudaStream_t);
template cudaError_t launch_pa_fill<double>(double *, double, uint64_t, cudaStream_t);

cudaError_t launch_pa_resample(uint64_t seed, uint64_t first_pop, uint64_t num_pops, int M,
                               uint32_t step, double neg_db, const uint32_t *cur,
                               const double *e_cur, int nw, unsigned long long *cum, uint32_t *nxt,
                               double *e_nxt, int32_t *src_out, unsigned long long *replaced,
                               const char *fields, char *fields_nxt, size_t field_bytes,
                               cudaStream_t s) {
  if (num_pops == 0 || num_pops > 0x7fffffffull) return cudaErrorInvalidValue;
  int threads = 32;
  while (threads < M && threads < 1024) threads <<= 1;
  k_pa_weights<<<(unsigned)num_pops, threads, 0, s>>>(e_cur, M, neg_db, cum);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const uint64_t grid = (num_pops * (uint64_t)M * 32 + 255) / 256;
  if (grid > 0x7fffffffull) return cudaErrorInvalidValue;
  k_pa_resample<<<(unsigned)grid, 256, 0, s>>>(seed, first_pop, num_pops, M, step, cum, cur, e_cur,
                                              nw, nxt, e_nxt, src_out, replaced,
                                              reinterpret_cast<const uint4 *>(fields),
                                              reinterpret_cast<uint4 *>(fields_nxt), field_bytes / 16);
  return cudaGetLastError();
}

}  // namespace osa
