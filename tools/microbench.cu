// tools/microbench.cu -- design probes for the dense sweep kernel (not product code).
// Measures how fast 148 CTAs can stream the SAME sequence of Q rows out of L2:
//   ldg      : 256 threads, 16-byte LDG into registers, DEPTH rows in flight
//   ldg-skew : same, CTA b starts SKEW*b rows later (decorrelates the hot lines)
//   tma      : one thread issues cp.async.bulk (TMA) row copies into a smem ring, others consume
//   tma-mc   : clusters of C CTAs, each CTA fetches 1/C of the row and multicasts it
// Usage: microbench [row_bytes=16384] [rows=4096] [iters=2]
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint4 ldg_na(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

template <int VPT, int DEPTH>
__global__ void __launch_bounds__(256) k_ldg(const uint4 *q, int rows, int row_vec, int iters, int skew,
                                             unsigned *sink) {
  unsigned acc = 0;
  const int start = (int)(((long long)blockIdx.x * skew) % rows);
  for (int it = 0; it < iters; ++it) {
    uint4 buf[DEPTH][VPT];
    // prologue
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      const int r = (start + d) % rows;
#pragma unroll
      for (int v = 0; v < VPT; ++v) buf[d][v] = ldg_na(q + (size_t)r * row_vec + v * 256 + threadIdx.x);
    }
    for (int r0 = 0; r0 < rows; r0 += DEPTH) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
#pragma unroll
        for (int v = 0; v < VPT; ++v) acc ^= buf[d][v].x + buf[d][v].y + buf[d][v].z + buf[d][v].w;
        const int rn = (start + r0 + d + DEPTH) % rows;
        if (r0 + d + DEPTH < rows) {
#pragma unroll
          for (int v = 0; v < VPT; ++v) buf[d][v] = ldg_na(q + (size_t)rn * row_vec + v * 256 + threadIdx.x);
        }
      }
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

// LDG with an L1 prefetch of a later row: prefetch.global.L1 has no destination register, so
// it is not tracked by the (single) LDG scoreboard; the demand loads then hit in L1.
__device__ __forceinline__ uint4 ldg_ca(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l1(const void *p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// batches of DEPTH rows; before consuming batch k the lines of batch k+AHEAD are prefetched
template <int VPT, int DEPTH, int AHEAD, int WORK>
__global__ void __launch_bounds__(256) k_ldg_pf(const uint4 *q, int rows, int row_vec, int iters, unsigned *sink) {
  unsigned acc = 0;
  const bool pf_lane = (threadIdx.x & 7) == 0;  // one prefetch per 128-byte line
  for (int it = 0; it < iters; ++it) {
    for (int r0 = 0; r0 < rows; r0 += DEPTH) {
      if (AHEAD > 0 && pf_lane) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
          const int rp = (r0 + AHEAD * DEPTH + d) % rows;
#pragma unroll
          for (int v = 0; v < VPT; ++v) prefetch_l1(q + (size_t)rp * row_vec + v * 256 + threadIdx.x);
        }
      }
      uint4 buf[DEPTH][VPT];
#pragma unroll
      for (int d = 0; d < DEPTH; ++d)
#pragma unroll
        for (int v = 0; v < VPT; ++v)
          buf[d][v] = AHEAD > 0 ? ldg_ca(q + (size_t)(r0 + d) * row_vec + v * 256 + threadIdx.x)
                                : ldg_na(q + (size_t)(r0 + d) * row_vec + v * 256 + threadIdx.x);
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
          unsigned x = buf[d][v].x + buf[d][v].y + buf[d][v].z + buf[d][v].w;
#pragma unroll
          for (int w = 0; w < WORK; ++w) x = x * 1664525u + 1013904223u;  // dependent ALU work
          acc ^= x;
        }
      }
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

// Dual-path double buffering: buffer A is filled with LDG (its own scoreboard), buffer B with
// texture fetches TLD (a different scoreboard), so the warp can compute on one while the other
// is in flight.  WORK = dependent ALU ops per 16 bytes, to mimic the FMA phase.
template <int D, int WORK, bool DUAL>
__global__ void __launch_bounds__(256) k_dual(const uint4 *q, cudaTextureObject_t tex, int rows, int row_vec,
                                              int iters, unsigned *sink) {
  unsigned acc = 0;
  const int t = threadIdx.x;
  auto consume = [&](uint4 (&buf)[D][4]) {
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        unsigned x = buf[d][v].x + buf[d][v].y + buf[d][v].z + buf[d][v].w;
#pragma unroll
        for (int w = 0; w < WORK; ++w) x = x * 1664525u + 1013904223u;
        acc ^= x;
      }
  };
  auto load_a = [&](uint4 (&buf)[D][4], int r0) {
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int v = 0; v < 4; ++v) buf[d][v] = ldg_na(q + (size_t)((r0 + d) % rows) * row_vec + v * 256 + t);
  };
  auto load_b = [&](uint4 (&buf)[D][4], int r0) {
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        if (DUAL) buf[d][v] = tex1Dfetch<uint4>(tex, ((r0 + d) % rows) * row_vec + v * 256 + t);
        else buf[d][v] = ldg_na(q + (size_t)((r0 + d) % rows) * row_vec + v * 256 + t);
      }
  };
  for (int it = 0; it < iters; ++it) {
    uint4 A[D][4], B[D][4];
    load_a(A, 0);
    load_b(B, D);
    for (int r0 = 0; r0 < rows; r0 += 2 * D) {
      consume(A);
      load_a(A, r0 + 2 * D);
      consume(B);
      load_b(B, r0 + 3 * D);
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

// cp.async (LDGSTS) ring: every thread copies its own 4 x 16 bytes of each row into its own
// shared-memory slots and later reads exactly those back, so no barrier of any kind is needed;
// cp.async.wait_group gives in-order completion K-1 rows deep.
template <int K>
__global__ void __launch_bounds__(256) k_cpasync(const uint4 *q, int rows, int row_vec, int iters, unsigned *sink) {
  extern __shared__ __align__(128) unsigned char ring[];
  const int t = threadIdx.x;
  unsigned acc = 0;
  uint4 *mine = reinterpret_cast<uint4 *>(ring);
  auto issue = [&](int r, int slot) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const uint32_t dst = smem_u32(mine + (size_t)slot * 1024 + v * 256 + t);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(q + (size_t)r * row_vec + v * 256 + t) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int it = 0; it < iters; ++it) {
    for (int k = 0; k < K - 1; ++k) issue(k, k);
    for (int j = 0; j < rows; ++j) {
      if (j + K - 1 < rows) issue(j + K - 1, (j + K - 1) % K);
      else asm volatile("cp.async.commit_group;" ::: "memory");  // keep the group count uniform
      asm volatile("cp.async.wait_group %0;" ::"n"(K - 1) : "memory");
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const uint4 x = mine[(size_t)(j % K) * 1024 + v * 256 + t];
        acc ^= x.x + x.y + x.z + x.w;
      }
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

// The same ring, but every CTA walks the rows in its own order (row = (a*j + b) mod rows with a
// per-CTA odd multiplier; rows must be a power of two): the CTAs of the sweep kernel do not ask for
// the same row at the same time, which is what the "same sequence" probes above measure.
template <int K>
__global__ void __launch_bounds__(256) k_cpasync_perm(const uint4 *q, int rows, int row_vec, int iters, unsigned *sink) {
  extern __shared__ __align__(128) unsigned char ring[];
  const int t = threadIdx.x;
  unsigned acc = 0;
  uint4 *mine = reinterpret_cast<uint4 *>(ring);
  const unsigned a = (blockIdx.x * 2654435761u) | 1u, b = blockIdx.x * 40503u + 17u, mask = rows - 1;
  auto issue = [&](int j, int slot) {
    const unsigned r = (a * (unsigned)j + b) & mask;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const uint32_t dst = smem_u32(mine + (size_t)slot * 1024 + v * 256 + t);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(q + (size_t)r * row_vec + v * 256 + t) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int it = 0; it < iters; ++it) {
    for (int k = 0; k < K - 1; ++k) issue(k, k);
    for (int j = 0; j < rows; ++j) {
      if (j + K - 1 < rows) issue(j + K - 1, (j + K - 1) % K);
      else asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(K - 1) : "memory");
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const uint4 x = mine[(size_t)(j % K) * 1024 + v * 256 + t];
        acc ^= x.x + x.y + x.z + x.w;
      }
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

// ---- TMA ring ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- per-warp TMA ring ----
// Every warp runs its own ring over its own 4 x 512 bytes of each row: lane 0 issues four bulk
// copies per row onto the warp's own mbarrier of the slot, all lanes wait on it and read their
// 16-byte pieces back.  The slot is re-used by the same warp, so "slot free" is a __syncwarp and
// there is no cross-warp handshake (the handshake is what the CTA-wide TMA ring below pays for).
// WORK: dependent-free FMAs per piece to mimic the arithmetic of the sweep kernel.
template <int K, int WORK>
__global__ void __launch_bounds__(256) k_tma_warp(const unsigned char *q, int rows, int iters, unsigned *sink) {
  extern __shared__ __align__(128) unsigned char ring[];  // K slots of 16 KiB, rows in natural layout
  __shared__ uint64_t full[8][K];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (lane == 0) {
    for (int s = 0; s < K; ++s) mbar_init(&full[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int total = rows * iters;
  auto issue = [&](int k) {  // lane 0 only
    const int s = k % K;
    const unsigned char *src = q + (size_t)(k % rows) * 16384 + warp * 512;
    unsigned char *dst = ring + (size_t)s * 16384 + warp * 512;
    mbar_expect_tx(&full[warp][s], 2048);
#pragma unroll
    for (int c = 0; c < 4; ++c) tma_load_1d(dst + c * 4096, src + c * 4096, 512, &full[warp][s]);
  };
  if (lane == 0)
    for (int k = 0; k < K - 1 && k < total; ++k) issue(k);
  float f[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = (float)i;
  unsigned acc = 0;
  for (int k = 0; k < total; ++k) {
    if (lane == 0 && k + K - 1 < total) issue(k + K - 1);  // slot of row k-1: read in the last step
    const int s = k % K, ph = (k / K) & 1;
    mbar_wait(&full[warp][s], ph);
    const uint4 *row = reinterpret_cast<const uint4 *>(ring + (size_t)s * 16384 + warp * 512 + lane * 16);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 x = row[c * 256];
      acc ^= x.x + x.y + x.z + x.w;
      if (WORK > 0) {
        const float m = __uint_as_float(x.x);
#pragma unroll
        for (int w = 0; w < WORK / 4; ++w) f[(c * (WORK / 4) + w) & 15] = fmaf(m, f[(c * (WORK / 4) + w) & 15], 1.0f);
      }
    }
    __syncwarp();  // every lane has read the slot before lane 0 refills it in the next step
  }
  float fs = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) fs += f[i];
  if (acc == 0x12345678u || fs == 123.456f) *sink = acc;
}

template <int STAGES>
__global__ void __launch_bounds__(288) k_tma(const unsigned char *q, int rows, int row_bytes, int iters,
                                             int read_smem, unsigned *sink) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int total = rows * iters;
  unsigned acc = 0;
  if (tid >= 256) {  // producer warp
    if (tid == 256) {
      for (int k = 0; k < total; ++k) {
        const int s = k % STAGES, ph = (k / STAGES) & 1;
        if (k >= STAGES) mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], row_bytes);
        tma_load_1d(ring + (size_t)s * row_bytes, q + (size_t)(k % rows) * row_bytes, row_bytes, &full[s]);
      }
    }
  } else {
    const int lane = tid & 31;
    for (int k = 0; k < total; ++k) {
      const int s = k % STAGES, ph = (k / STAGES) & 1;
      mbar_wait(&full[s], ph);
      if (read_smem) {
        const uint4 *row = reinterpret_cast<const uint4 *>(ring + (size_t)s * row_bytes);
        for (int v = tid; v < row_bytes / 16; v += 256) { uint4 x = row[v]; acc ^= x.x + x.y + x.z + x.w; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

// TMA ring, tuned consumer: stages of B rows on one mbarrier, one polling lane per warp,
// optional dependent ALU work per element to mimic the FMA phase
template <int STAGES, int B, int WORK>
__global__ void __launch_bounds__(288) k_tma2(const unsigned char *q, int rows, int row_bytes, int iters,
                                              unsigned *sink) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int total = rows * iters / B;  // stage-sized steps
  const unsigned stage_bytes = (unsigned)row_bytes * B;
  unsigned acc = 0;
  if (tid >= 256) {
    if (tid == 256) {
      for (int k = 0; k < total; ++k) {
        const int s = k % STAGES, ph = (k / STAGES) & 1;
        if (k >= STAGES) mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], stage_bytes);
#pragma unroll
        for (int b = 0; b < B; ++b)
          tma_load_1d(ring + (size_t)s * stage_bytes + (size_t)b * row_bytes,
                      q + (size_t)((k * B + b) % rows) * row_bytes, row_bytes, &full[s]);
      }
    }
  } else {
    const int lane = tid & 31;
    for (int k = 0; k < total; ++k) {
      const int s = k % STAGES, ph = (k / STAGES) & 1;
      if (lane == 0) mbar_wait(&full[s], ph);
      __syncwarp();
      const uint4 *stage = reinterpret_cast<const uint4 *>(ring + (size_t)s * stage_bytes);
#pragma unroll
      for (int b = 0; b < B; ++b) {
        uint4 x[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) x[v] = stage[b * (row_bytes / 16) + v * 256 + tid];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          unsigned y = x[v].x + x[v].y + x[v].z + x[v].w;
#pragma unroll
          for (int w = 0; w < WORK; ++w) y = y * 1664525u + 1013904223u;
          acc ^= y;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

// multicast: cluster of C CTAs; CTA rank c loads slice c of each row and multicasts to all
template <int STAGES, int C>
__global__ void __launch_bounds__(288) k_tma_mc(const unsigned char *q, int rows, int row_bytes, int iters,
                                                int read_smem, unsigned *sink) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8 * C); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster.sync();
  const int total = rows * iters;
  const unsigned slice = row_bytes / C;
  unsigned acc = 0;
  if (tid >= 256) {
    if (tid == 256) {
      for (int k = 0; k < total; ++k) {
        const int s = k % STAGES, ph = (k / STAGES) & 1;
        if (k >= STAGES) mbar_wait(&empty[s], ph ^ 1);  // all C CTAs' consumers released slot s
        mbar_expect_tx(&full[s], row_bytes);            // my smem receives the C slices
        const unsigned short mask = (unsigned short)((1u << C) - 1);
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
            ::"r"(smem_u32(ring + (size_t)s * row_bytes + rank * slice)),
              "l"(q + (size_t)(k % rows) * row_bytes + rank * slice), "r"(slice), "r"(smem_u32(&full[s])), "h"(mask)
            : "memory");
      }
    }
  } else {
    const int lane = tid & 31;
    for (int k = 0; k < total; ++k) {
      const int s = k % STAGES, ph = (k / STAGES) & 1;
      mbar_wait(&full[s], ph);
      if (read_smem) {
        const uint4 *row = reinterpret_cast<const uint4 *>(ring + (size_t)s * row_bytes);
        for (int v = tid; v < row_bytes / 16; v += 256) { uint4 x = row[v]; acc ^= x.x + x.y + x.z + x.w; }
      }
      __syncwarp();
      if (lane == 0) {
        // release slot s in EVERY CTA of the cluster (each producer writes into all of them)
        for (unsigned c = 0; c < C; ++c) {
          uint32_t remote;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&empty[s])), "r"(c));
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
        }
      }
    }
  }
  cluster.sync();
  if (acc == 0x12345678u) *sink = acc;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f();  // warm
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms;
}

int main(int argc, char **argv) {
  const int row_bytes = argc > 1 ? atoi(argv[1]) : 16384;
  const int rows = argc > 2 ? atoi(argv[2]) : 4096;
  const int iters = argc > 3 ? atoi(argv[3]) : 2;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  unsigned char *q; unsigned *sink;
  CK(cudaMalloc(&q, (size_t)rows * row_bytes)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(q, 1, (size_t)rows * row_bytes));
  const double bytes_per_cta = (double)rows * row_bytes * iters;
  auto report = [&](const char *name, int grid, float ms) {
    printf("{\"probe\":\"%s\",\"row_bytes\":%d,\"rows\":%d,\"grid\":%d,\"ms\":%.3f,\"agg_gbs\":%.1f,\"clk_per_row_at_1.93GHz\":%.0f}\n",
           name, row_bytes, rows, grid, ms, bytes_per_cta * grid / (ms * 1e-3) / 1e9,
           ms * 1e-3 * 1.93e9 / ((double)rows * iters));
    fflush(stdout);
  };
  const int row_vec = row_bytes / 16;
  if (getenv("MB_ONLY") && row_bytes == 16384) {  // MB_ONLY=tma_warp: only the per-warp TMA ring and its cp.async twin
    auto runw = [&](auto kern, int k, const char *name) {
      const size_t smem = (size_t)k * 16384;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      report(name, sms, time_ms([&] { kern<<<sms, 256, smem>>>(q, rows, iters, sink); }));
    };
    auto runc = [&](auto kern, int k, const char *name) {
      const size_t smem = (size_t)k * 16384;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      report(name, sms, time_ms([&] { kern<<<sms, 256, smem>>>((uint4 *)q, rows, row_vec, iters, sink); }));
    };
    runc(k_cpasync<12>, 12, "cpasync_k12");
    runw(k_tma_warp<6, 0>, 6, "tma_warp_k6");
    runw(k_tma_warp<12, 0>, 12, "tma_warp_k12");
    runw(k_tma_warp<12, 32>, 12, "tma_warp_k12_work32");
    runw(k_tma_warp<12, 64>, 12, "tma_warp_k12_work64");
    return 0;
  }
  if (row_bytes == 16384) {
    report("ldg_depth2", sms, time_ms([&] { k_ldg<4, 2><<<sms, 256>>>((uint4 *)q, rows, row_vec, iters, 0, sink); }));
    report("ldg_depth3", sms, time_ms([&] { k_ldg<4, 3><<<sms, 256>>>((uint4 *)q, rows, row_vec, iters, 0, sink); }));
    report("ldg_depth4", sms, time_ms([&] { k_ldg<4, 4><<<sms, 256>>>((uint4 *)q, rows, row_vec, iters, 0, sink); }));
    report("ldg_depth8", sms, time_ms([&] { k_ldg<4, 8><<<sms, 256>>>((uint4 *)q, rows, row_vec, iters, 0, sink); }));
    report("ldg_depth4_skew7", sms, time_ms([&] { k_ldg<4, 4><<<sms, 256>>>((uint4 *)q, rows, row_vec, iters, 7, sink); }));
    report("ldg_depth4_skew27", sms, time_ms([&] { k_ldg<4, 4><<<sms, 256>>>((uint4 *)q, rows, row_vec, iters, 27, sink); }));
    report("ldg_depth8_skew27", sms, time_ms([&] { k_ldg<4, 8><<<sms, 256>>>((uint4 *)q, rows, row_vec, iters, 27, sink); }));
    report("batch4_nopf_work0", sms, time_ms([&] { k_ldg_pf<4, 4, 0, 0><<<sms, 256>>>((uint4 *)q, rows, row_vec, iters, sink); }));
    report("batch4_nopf_work24", sms, time_ms([&] { k_ldg_pf<4, 4, 0, 24><<<sms, 256>>>((uint4 *)q, rows, row_vec, iters, sink); }));
    report("batch4_pf1_work0", sms, time_ms([&] { k_ldg_pf<4, 4, 1, 0><<<sms, 256>>>((uint4 *)q, rows, row_vec, iters, sink); }));
    report("batch4_pf1_work24", sms, time_ms([&] { k_ldg_pf<4, 4, 1, 24><<<sms, 256>>>((uint4 *)q, rows, row_vec, iters, sink); }));
    report("batch4_pf2_work24", sms, time_ms([&] { k_ldg_pf<4, 4, 2, 24><<<sms, 256>>>((uint4 *)q, rows, row_vec, iters, sink); }));
    report("batch2_pf2_work24", sms, time_ms([&] { k_ldg_pf<4, 2, 2, 24><<<sms, 256>>>((uint4 *)q, rows, row_vec, iters, sink); }));
    report("batch2_pf4_work24", sms, time_ms([&] { k_ldg_pf<4, 2, 4, 24><<<sms, 256>>>((uint4 *)q, rows, row_vec, iters, sink); }));
    report("batch1_pf8_work24", sms, time_ms([&] { k_ldg_pf<4, 1, 8, 24><<<sms, 256>>>((uint4 *)q, rows, row_vec, iters, sink); }));
    {
      cudaResourceDesc rd = {};
      rd.resType = cudaResourceTypeLinear;
      rd.res.linear.devPtr = q;
      rd.res.linear.desc = cudaCreateChannelDesc<uint4>();
      rd.res.linear.sizeInBytes = (size_t)rows * row_bytes;
      cudaTextureDesc td = {};
      td.readMode = cudaReadModeElementType;
      cudaTextureObject_t tex = 0;
      CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
      report("single_d2_work0", sms, time_ms([&] { k_dual<2, 0, false><<<sms, 256>>>((uint4 *)q, tex, rows, row_vec, iters, sink); }));
      report("dual_d2_work0", sms, time_ms([&] { k_dual<2, 0, true><<<sms, 256>>>((uint4 *)q, tex, rows, row_vec, iters, sink); }));
      report("single_d2_work24", sms, time_ms([&] { k_dual<2, 24, false><<<sms, 256>>>((uint4 *)q, tex, rows, row_vec, iters, sink); }));
      report("dual_d2_work24", sms, time_ms([&] { k_dual<2, 24, true><<<sms, 256>>>((uint4 *)q, tex, rows, row_vec, iters, sink); }));
      report("single_d2_work64", sms, time_ms([&] { k_dual<2, 64, false><<<sms, 256>>>((uint4 *)q, tex, rows, row_vec, iters, sink); }));
      report("dual_d2_work64", sms, time_ms([&] { k_dual<2, 64, true><<<sms, 256>>>((uint4 *)q, tex, rows, row_vec, iters, sink); }));
      report("dual_d3_work24", sms, time_ms([&] { k_dual<3, 24, true><<<sms, 256>>>((uint4 *)q, tex, rows, row_vec, iters, sink); }));
      report("dual_d3_work64", sms, time_ms([&] { k_dual<3, 64, true><<<sms, 256>>>((uint4 *)q, tex, rows, row_vec, iters, sink); }));
      report("dual_d4_work24", sms, time_ms([&] { k_dual<4, 24, true><<<sms, 256>>>((uint4 *)q, tex, rows, row_vec, iters, sink); }));
      report("dual_d1_work24", sms, time_ms([&] { k_dual<1, 24, true><<<sms, 256>>>((uint4 *)q, tex, rows, row_vec, iters, sink); }));
      CK(cudaDestroyTextureObject(tex));
    }
    {
      auto runc = [&](auto kern, int k, const char *name) {
        const size_t smem = (size_t)k * 16384;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        report(name, sms, time_ms([&] { kern<<<sms, 256, smem>>>((uint4 *)q, rows, row_vec, iters, sink); }));
      };
      runc(k_cpasync<4>, 4, "cpasync_k4");
      runc(k_cpasync<6>, 6, "cpasync_k6");
      runc(k_cpasync<8>, 8, "cpasync_k8");
      runc(k_cpasync<12>, 12, "cpasync_k12");
      if ((rows & (rows - 1)) == 0) {
        runc(k_cpasync_perm<6>, 6, "cpasync_k6_own_order");
        runc(k_cpasync_perm<12>, 12, "cpasync_k12_own_order");
      }
    }
    report("ldg_depth4_2cta_per_sm", 2 * sms, time_ms([&] { k_ldg<4, 4><<<2 * sms, 256>>>((uint4 *)q, rows, row_vec, iters, 0, sink); }));
  }
  {
    const size_t smem4 = 4 * (size_t)row_bytes, smem8 = 8 * (size_t)row_bytes;
    CK(cudaFuncSetAttribute(k_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
    CK(cudaFuncSetAttribute(k_tma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
    report("tma_stages4_noread", sms, time_ms([&] { k_tma<4><<<sms, 288, smem4>>>(q, rows, row_bytes, iters, 0, sink); }));
    report("tma_stages8_noread", sms, time_ms([&] { k_tma<8><<<sms, 288, smem8>>>(q, rows, row_bytes, iters, 0, sink); }));
    report("tma_stages8_read", sms, time_ms([&] { k_tma<8><<<sms, 288, smem8>>>(q, rows, row_bytes, iters, 1, sink); }));
    report("tma_stages4_read", sms, time_ms([&] { k_tma<4><<<sms, 288, smem4>>>(q, rows, row_bytes, iters, 1, sink); }));
    auto run2 = [&](auto kern, int stages, int b, const char *name) {
      const size_t smem = (size_t)stages * b * row_bytes;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      report(name, sms, time_ms([&] { kern<<<sms, 288, smem>>>(q, rows, row_bytes, iters, sink); }));
    };
    run2(k_tma2<8, 1, 0>, 8, 1, "tma2_st8_b1_work0");
    run2(k_tma2<4, 2, 0>, 4, 2, "tma2_st4_b2_work0");
    run2(k_tma2<3, 4, 0>, 3, 4, "tma2_st3_b4_work0");
    run2(k_tma2<6, 2, 0>, 6, 2, "tma2_st6_b2_work0");
    run2(k_tma2<6, 2, 8>, 6, 2, "tma2_st6_b2_work8");
    run2(k_tma2<6, 2, 24>, 6, 2, "tma2_st6_b2_work24");
    run2(k_tma2<12, 1, 24>, 12, 1, "tma2_st12_b1_work24");
  }
  {
    auto run_mc = [&](auto kern, int C, int stages, int read, const char *name) {
      const size_t smem = (size_t)stages * row_bytes;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(sms / C * C);
      cfg.blockDim = dim3(288);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      float ms = time_ms([&] { CK(cudaLaunchKernelEx(&cfg, kern, (const unsigned char *)q, rows, row_bytes, iters, read, sink)); });
      report(name, sms / C * C, ms);
    };
    if (argc > 4) run_mc(k_tma_mc<8, 2>, 2, 8, 0, "tma_mc2_stages8_noread");
    if (argc > 4) run_mc(k_tma_mc<8, 2>, 2, 8, 1, "tma_mc2_stages8_read");
    if (argc > 4) run_mc(k_tma_mc<8, 4>, 4, 8, 0, "tma_mc4_stages8_noread");
    if (argc > 4) run_mc(k_tma_mc<8, 4>, 4, 8, 1, "tma_mc4_stages8_read");
  }
  return 0;
}
