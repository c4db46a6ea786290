// tools/microbench_tmem.cu -- design probe (not product code): can tensor memory (TMEM, 256 KB per
// SM, tcgen05.ld / tcgen05.st) serve as a DYNAMICALLY INDEXABLE home for the local fields h[r][:] of
// the dense sweep kernel?  Registers cannot be indexed by the id of the flipping trajectory, which is
// what forces the group tests / zero-multiplier FMAs of the apply loop; a TMEM column address can.
// Layout probed: trajectory r owns TMEM columns [32r, 32r+32) of all 128 lanes (16 KiB = one fp32
// field vector of N = 4096); apply warp w (0..7) works on lanes 32(w&3).., columns 16(w>>2)..+15 of
// that slab, i.e. 16 floats per thread and flipper -- the same 16 columns per thread as today.
// Per "row": F flippers (distinct trajectories), each  h[r] += m * row  (tcgen05.ld.x16, 8 packed
// FMAs, tcgen05.st.x16).  Reports clocks per row and per flipper, and checks the final TMEM content.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/bin/microbench_tmem tools/microbench_tmem.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define LD16_REGS(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                     "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
#define ST16_REGS(v) "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), \
                     "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : LD16_REGS(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      :: "r"(taddr), ST16_REGS(v) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void fma16(uint32_t (&h)[16], const uint32_t (&q)[16], float m) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    unsigned long long acc, qq, mm;
    const uint32_t mb = __float_as_uint(m);
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "r"(h[i]), "r"(h[i + 1]));
    asm("mov.b64 %0, {%1, %2};" : "=l"(qq) : "r"(q[i]), "r"(q[i + 1]));
    asm("mov.b64 %0, {%1, %1};" : "=l"(mm) : "r"(mb));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(qq), "l"(mm));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(h[i]), "=r"(h[i + 1]) : "l"(acc));
  }
}

// flipper f of row k: trajectory (k * 5 + f * 3) % R  (distinct within a row for F <= R/3 .. and R coprime
// with 3; the host check recomputes the same sequence)
__device__ __host__ inline int flipper(int k, int f, int R) { return (k * 5 + f * 3) % R; }

// MODE 0: ld, wait::ld, fma, st                      (no wait::st inside the loop)
// MODE 1: MODE 0 + wait::st after every row
// MODE 2: software pipeline inside a row: ld(f+1) issued before fma(f); wait::st after every row
template <int MODE>
__global__ void __launch_bounds__(384, 1) k_tmem(int rows, int F, int R, float *out, long long *clk) {
  __shared__ uint32_t s_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(&s_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_base;
  long long t0 = 0, t1 = 0;
  if (warp < 8) {
    const uint32_t my = base + ((uint32_t)(32 * (warp & 3)) << 16) + 16u * (uint32_t)(warp >> 2);
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0u;
    for (int r = 0; r < 16; ++r) tmem_st16(my + 32u * r, z);
    tmem_wait_st();
    // the "row": 16 small integers per thread, different per row (exact in fp32, any order)
    uint32_t q[16];
    asm volatile("bar.sync 1, 256;" ::: "memory");
    t0 = clock64();
    for (int k = 0; k < rows; ++k) {
#pragma unroll
      for (int i = 0; i < 16; ++i) q[i] = __float_as_uint((float)(((tid * 16 + i) + k) & 7));
      if (MODE == 2) {
        uint32_t cur[16], nxt[16];
        tmem_ld16(my + 32u * flipper(k, 0, R), nxt);
        for (int f = 0; f < F; ++f) {
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
          const int r = flipper(k, f, R);
          if (f + 1 < F) tmem_ld16(my + 32u * flipper(k, f + 1, R), nxt);
          fma16(cur, q, (r & 1) ? -1.0f : 1.0f);
          tmem_st16(my + 32u * r, cur);
        }
        tmem_wait_st();
      } else {
        for (int f = 0; f < F; ++f) {
          const int r = flipper(k, f, R);
          uint32_t h[16];
          tmem_ld16(my + 32u * r, h);
          tmem_wait_ld();
          fma16(h, q, (r & 1) ? -1.0f : 1.0f);
          tmem_st16(my + 32u * r, h);
        }
        if (MODE == 1) tmem_wait_st();
      }
    }
    tmem_wait_st();
    t1 = clock64();
    if (blockIdx.x == 0) {
      for (int r = 0; r < R; ++r) {
        uint32_t h[16];
        tmem_ld16(my + 32u * r, h);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) out[(size_t)r * 4096 + tid * 16 + i] = __uint_as_float(h[i]);
      }
    }
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(base) : "memory");
}

template <int MODE>
static void run(int rows, int F, int R, float *d_out, long long *d_clk, int grid) {
  cudaMemset(d_out, 0, (size_t)16 * 4096 * sizeof(float));
  k_tmem<MODE><<<grid, 384>>>(rows, F, R, d_out, d_clk);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mode %d F %d: %s\n", MODE, F, cudaGetErrorString(e)); exit(1); }
  static float out[16 * 4096];
  static long long clk[1024];
  cudaMemcpy(out, d_out, sizeof(out), cudaMemcpyDeviceToHost);
  cudaMemcpy(clk, d_clk, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  // expected content
  static float ex[16 * 4096];
  for (int i = 0; i < 16 * 4096; ++i) ex[i] = 0.f;
  for (int k = 0; k < rows; ++k)
    for (int f = 0; f < F; ++f) {
      const int r = flipper(k, f, R);
      const float m = (r & 1) ? -1.f : 1.f;
      for (int c = 0; c < 4096; ++c) ex[r * 4096 + c] += m * (float)((c + k) & 7);
    }
  long bad = 0;
  for (int i = 0; i < R * 4096; ++i) bad += (out[i] != ex[i]);
  double avg = 0;
  for (int b = 0; b < grid; ++b) avg += (double)clk[b];
  avg /= grid;
  printf("mode %d  F=%d R=%d rows=%d: %.1f clk/row, %.1f clk/flipper (per SM, 8 warps)  mismatches=%ld\n", MODE, F, R,
         rows, avg / rows, avg / rows / F, bad);
}

int main(int argc, char **argv) {
  const int rows = argc > 1 ? atoi(argv[1]) : 4000;
  int grid = 148;
  float *d_out;
  long long *d_clk;
  cudaMalloc(&d_out, (size_t)16 * 4096 * sizeof(float));
  cudaMalloc(&d_clk, 1024 * sizeof(long long));
  for (int R : {16, 13}) {  // 13: coprime with 3 and 5 -> flippers of a row are distinct up to F = 4
    for (int F = 1; F <= 4; ++F) {
      run<0>(rows, F, R, d_out, d_clk, grid);
      run<1>(rows, F, R, d_out, d_clk, grid);
      run<2>(rows, F, R, d_out, d_clk, grid);
    }
  }
  return 0;
}
