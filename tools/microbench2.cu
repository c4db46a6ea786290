// tools/microbench2.cu -- design probe (not product code): a row ring fed by a producer warp with TMA
// bulk copies (cp.async.bulk + mbarrier) against the self-fed cp.async ring of the sweep kernel, both
// with the consumer pattern of the apply warps: read the 16-byte pieces back, W packed FMAs per row.
// Consumers wait with mbarrier.test_wait (non-blocking, spin) -- try_wait suspends the warp for a
// scheduler quantum and was what made the TMA ring of round 1 slow.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/microbench2 tools/microbench2.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, unsigned parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ffma2(unsigned long long &acc, unsigned long long q, unsigned long long m) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(q), "l"(m));
}

// W = packed FMAs per row and thread, in units of 8 (one "flipper")
template <int K, int W>
__global__ void __launch_bounds__(288) k_tma_fed(const unsigned char *q, int rows, int iters, float *sink) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ uint64_t full[K], empty[K];
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < K; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int total = rows * iters;
  if (tid >= 256) {
    if (tid == 256) {
      int s = 0; unsigned ph = 0;
      for (int k = 0; k < total; ++k) {
        if (k >= K) while (!mbar_test(&empty[s], ph ^ 1)) {}
        mbar_expect_tx(&full[s], 16384);
        tma_load_1d(ring + (size_t)s * 16384, q + (size_t)(k % rows) * 16384, 16384, &full[s]);
        if (++s == K) { s = 0; ph ^= 1; }
      }
    }
  } else {
    unsigned long long acc[W > 0 ? W : 1][8];
    for (int w = 0; w < W; ++w) for (int i = 0; i < 8; ++i) acc[w][i] = 0ull;
    const unsigned long long one = 0x3f8000003f800000ull;
    int s = 0; unsigned ph = 0;
    for (int k = 0; k < total; ++k) {
      while (!mbar_test(&full[s], ph)) {}
      const uint4 *row = reinterpret_cast<const uint4 *>(ring + (size_t)s * 16384);
      uint4 x[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) x[v] = row[v * 256 + tid];
      unsigned long long qq[8];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        qq[2 * v] = ((unsigned long long)x[v].y << 32) | x[v].x;
        qq[2 * v + 1] = ((unsigned long long)x[v].w << 32) | x[v].z;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
#pragma unroll
      for (int w = 0; w < W; ++w)
#pragma unroll
        for (int i = 0; i < 8; ++i) ffma2(acc[w][i], qq[i], one);
      if (W == 0) acc[0][0] ^= qq[0] ^ qq[3] ^ qq[5] ^ qq[7];
      if (++s == K) { s = 0; ph ^= 1; }
    }
    unsigned long long t = 0;
    for (int w = 0; w < (W > 0 ? W : 1); ++w) for (int i = 0; i < 8; ++i) t ^= acc[w][i];
    if (t == 0x12345678u) *sink = (float)t;
  }
}

template <int K, int W>
__global__ void __launch_bounds__(256) k_self_fed(const unsigned char *q, int rows, int iters, float *sink) {
  extern __shared__ __align__(128) unsigned char ring[];
  const int tid = threadIdx.x;
  unsigned long long acc[W > 0 ? W : 1][8];
  for (int w = 0; w < W; ++w) for (int i = 0; i < 8; ++i) acc[w][i] = 0ull;
  const unsigned long long one = 0x3f8000003f800000ull;
  const int total = rows * iters;
  auto issue = [&](int k, int slot) {
#pragma unroll
    for (int v = 0; v < 4; ++v)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(ring + (size_t)slot * 16384 + (v * 256 + tid) * 16)),
                   "l"(q + (size_t)(k % rows) * 16384 + (v * 256 + tid) * 16) : "memory");
  };
  for (int k = 0; k < K - 1; ++k) { issue(k, k); asm volatile("cp.async.commit_group;" ::: "memory"); }
  int s = 0, sw = K - 1;
  for (int k = 0; k < total; ++k) {
    asm volatile("cp.async.wait_group %0;" ::"n"(K - 2) : "memory");
    const uint4 *row = reinterpret_cast<const uint4 *>(ring + (size_t)s * 16384);
    uint4 x[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) x[v] = row[v * 256 + tid];
    if (k + K - 1 < total) issue(k + K - 1, sw);
    asm volatile("cp.async.commit_group;" ::: "memory");
    unsigned long long qq[8];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      qq[2 * v] = ((unsigned long long)x[v].y << 32) | x[v].x;
      qq[2 * v + 1] = ((unsigned long long)x[v].w << 32) | x[v].z;
    }
#pragma unroll
    for (int w = 0; w < W; ++w)
#pragma unroll
      for (int i = 0; i < 8; ++i) ffma2(acc[w][i], qq[i], one);
    if (W == 0) acc[0][0] ^= qq[0] ^ qq[3] ^ qq[5] ^ qq[7];
    if (++s == K) s = 0;
    if (++sw == K) sw = 0;
  }
  unsigned long long t = 0;
  for (int w = 0; w < (W > 0 ? W : 1); ++w) for (int i = 0; i < 8; ++i) t ^= acc[w][i];
  if (t == 0x12345678u) *sink = (float)t;
}


// Register-staged rows: two batches of D rows per thread, filled with plain LDG (no shared memory);
// while one batch is consumed (W x 8 packed FMAs per row into NACC x 8 accumulator pairs, like the
// fields of NACC trajectories) the other is in flight.
__device__ __forceinline__ uint4 ldg_nc(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
template <int D, int W, int NACC>
__global__ void __launch_bounds__(256) k_ldg_db(const unsigned char *q, int rows, int iters, float *sink) {
  const int tid = threadIdx.x;
  unsigned long long acc[NACC][8];
  for (int w = 0; w < NACC; ++w) for (int i = 0; i < 8; ++i) acc[w][i] = 0ull;
  const unsigned long long one = 0x3f8000003f800000ull;
  const int total = rows * iters;
  const uint4 *base = reinterpret_cast<const uint4 *>(q) + tid;
  auto load = [&](uint4 (&buf)[D][4], int k0) {
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int v = 0; v < 4; ++v) buf[d][v] = ldg_nc(base + (size_t)((k0 + d) % rows) * 1024 + v * 256);
  };
  auto consume = [&](uint4 (&buf)[D][4], int k0) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      unsigned long long qq[8];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        qq[2 * v] = ((unsigned long long)buf[d][v].y << 32) | buf[d][v].x;
        qq[2 * v + 1] = ((unsigned long long)buf[d][v].w << 32) | buf[d][v].z;
      }
#pragma unroll
      for (int w = 0; w < W; ++w)
#pragma unroll
        for (int i = 0; i < 8; ++i) ffma2(acc[(w + d) % NACC][i], qq[i], one);
    }
  };
  uint4 A[D][4], B[D][4];
  load(A, 0);
  load(B, D);
  for (int k = 0; k < total; k += 2 * D) {
    consume(A, k);
    load(A, k + 2 * D);
    consume(B, k + D);
    load(B, k + 3 * D);
  }
  unsigned long long t = 0;
  for (int w = 0; w < NACC; ++w) for (int i = 0; i < 8; ++i) t ^= acc[w][i];
  if (t == 0x12345678u) *sink = (float)t;
}

int main(int argc, char **argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  int idx = 0;
  setvbuf(stdout, NULL, _IONBF, 0);
  const int rows = 4096, iters = 2, grid = 148;
  unsigned char *q; float *sink;
  cudaMalloc(&q, (size_t)rows * 16384);
  cudaMemset(q, 0, (size_t)rows * 16384);
  cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](auto kern, int threads, int k, const char *name) {
    if (only >= 0 && idx++ != only) return;
    printf("running %s\n", name);
    const size_t smem = (size_t)k * 16384;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      kern<<<grid, threads, smem>>>(q, rows, iters, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    printf("{\"probe\":\"%s\",\"ms\":%.3f,\"clk_per_row_at_1.93GHz\":%.0f,\"err\":\"%s\"}\n", name, best,
           best * 1e-3 * 1.93e9 / (rows * iters), cudaGetErrorString(err));
  };
  run(k_ldg_db<2, 2, 8>, 256, 0, "ldg_d2_w2");
  run(k_ldg_db<2, 3, 8>, 256, 0, "ldg_d2_w3");
  run(k_ldg_db<2, 4, 8>, 256, 0, "ldg_d2_w4");
  run(k_ldg_db<2, 6, 8>, 256, 0, "ldg_d2_w6");
  run(k_ldg_db<1, 3, 8>, 256, 0, "ldg_d1_w3");
  run(k_ldg_db<3, 3, 8>, 256, 0, "ldg_d3_w3");
  run(k_ldg_db<2, 1, 8>, 256, 0, "ldg_d2_w1");
  run(k_self_fed<12, 1>, 256, 12, "self_k12_w1");
  run(k_self_fed<12, 2>, 256, 12, "self_k12_w2");
  run(k_self_fed<12, 4>, 256, 12, "self_k12_w4");
  run(k_self_fed<12, 6>, 256, 12, "self_k12_w6");

  run(k_tma_fed<12, 2>, 288, 12, "tma_k12_w2");
  run(k_tma_fed<12, 4>, 288, 12, "tma_k12_w4");
  run(k_tma_fed<12, 6>, 288, 12, "tma_k12_w6");
  run(k_tma_fed<6, 4>, 288, 6, "tma_k6_w4");
  run(k_tma_fed<13, 4>, 288, 13, "tma_k13_w4");
  return 0;
}
