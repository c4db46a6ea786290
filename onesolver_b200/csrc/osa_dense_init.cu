// osa_dense_init.cu -- initial local fields of many trajectories with SHARED row fetches.
//
// The reference's own loop (random-site attempts, annealing.hpp:97-101; k_dense_generic here) keeps
// one trajectory per warp, so nothing is shared while it runs.  But every trajectory first needs
// its initial field h = q_ii + sum_{j: x_j = 1} Q_ij (the engine's replacement of the reference's
// initial energy evaluation, annealing.hpp:94-95): N/2 rows of Q per trajectory -- at the
// reference's typical run lengths (100 .. 1000 attempts) more row traffic than the whole walk that
// follows.  This kernel builds the fields of R trajectories per CTA with the streaming machinery of
// the sequential-sweep kernels (apply_rows, osa_dense_seq.cuh): a row is fetched once for the R
// trajectories, and every element is the same chain of additions in site order as in the
// one-trajectory build (fma(1, q, h) == h + q), so the bits are those of the host replay.
// The fields go to global memory, [num_tries][ld] in the sweep precision; k_dense_generic loads
// its row instead of building it (DenseParams::fields_in).
#include "osa_dense_seq.cuh"

namespace osa {

using namespace dseq;

namespace {

template <typename T, int NCH, int R, int K, int G>
__global__ void __launch_bounds__(256, 1) k_dense_init_fields(const DenseParams<T> p) {
  constexpr int TH = 256;
  extern __shared__ __align__(128) unsigned char s_ring[];  // K rows, thread-private slots
  using C = Cfg<T, NCH, R, TH>;
  using VecT = typename C::VecT;
  constexpr int V = C::V, CPT = C::CPT, CHW = C::CHW, NWP = C::NWP;
  __shared__ uint32_t s_x[NWP][R];

  const int tid = threadIdx.x;
  const int n = p.n;
  const int nblk = (n + 31) >> 5;
  const uint64_t batch0 = (uint64_t)blockIdx.x * R;
  const uint64_t left = p.num_tries - batch0;
  const int nvalid = left < (uint64_t)R ? (int)left : R;

  // initial spins, exactly the bits the annealing kernels draw (STREAM_INIT)
  for (int q = tid; q < R * NWP; q += TH) {
    const int r = q / NWP, k = q % NWP;
    uint32_t word = 0;
    if (r < nvalid && k < nblk) {
      const U4 d = engine_draw(p.seed, p.first_try + batch0 + (uint64_t)r, STREAM_INIT, (uint32_t)k >> 2, 0u);
      word = pick(d, (uint32_t)k & 3u);
      const int valid = n - k * 32;
      if (valid < 32) word &= (1u << valid) - 1u;
    }
    s_x[k][r] = word;
  }

  Field<T, CPT> h[R];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    T dv[V];
    vec_unpack<T>(*reinterpret_cast<const VecT *>(p.diag + c * CHW + tid * V), dv);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int e = 0; e < V; ++e) h[r].set(c * V + e, dv[e]);
  }
  __syncthreads();

  unsigned long long cnt_init_rows = 0;
  for (int b = 0; b < nblk; ++b) {
    uint32_t am[R], sm[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      am[r] = s_x[b][r];
      sm[r] = 0u;
    }
    const uint32_t any = apply_rows<T, NCH, R, K, TH, G>(p.qoff, s_ring, p.ld, b * 32, am, sm, h, tid);
    cnt_init_rows += (unsigned)__popc(any);
  }

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
  if (tid == 0) atomicAdd(&p.counters->init_row_fetches, cnt_init_rows);
}

template <typename T, int NCH, int R, int K, int G>
cudaError_t launch_init_cfg(const DenseParams<T> &p, cudaStream_t s, int *traj_per_batch) {
  const uint64_t grid64 = (p.num_tries + R - 1) / R;
  if (grid64 == 0 || grid64 > 0x7fffffffull) return cudaErrorInvalidValue;
  const size_t smem = (size_t)K * NCH * 256 * 16;  // the cp.async ring
  auto kern = k_dense_init_fields<T, NCH, R, K, G>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  kern<<<(unsigned)grid64, 256, smem, s>>>(p);
  if (traj_per_batch) *traj_per_batch = R;
  return cudaGetLastError();
}

}  // namespace

// shapes as in launch_dense_seq (osa_dense_seq.cu): R trajectories per CTA, a ring of K rows
template <>
cudaError_t launch_dense_init_fields<float>(const DenseParams<float> &p, cudaStream_t s, int *tpb) {
  if (p.ld % 1024 != 0 || !p.fields_out) return cudaErrorInvalidValue;
  switch (p.ld / 1024) {
    case 1: return launch_init_cfg<float, 1, 16, 16, 4>(p, s, tpb);
    case 2: return launch_init_cfg<float, 2, 16, 16, 4>(p, s, tpb);
    case 3: return launch_init_cfg<float, 3, 12, 12, 4>(p, s, tpb);
    case 4: return launch_init_cfg<float, 4, 12, 12, 2>(p, s, tpb);
    case 5: return launch_init_cfg<float, 5, 8, 9, 2>(p, s, tpb);
    case 6: return launch_init_cfg<float, 6, 8, 8, 2>(p, s, tpb);
    case 7: return launch_init_cfg<float, 7, 6, 6, 2>(p, s, tpb);
    case 8: return launch_init_cfg<float, 8, 6, 6, 2>(p, s, tpb);
    default: return cudaErrorInvalidValue;
  }
}

template <>
cudaError_t launch_dense_init_fields<double>(const DenseParams<double> &p, cudaStream_t s, int *tpb) {
  if (p.ld % 512 != 0 || !p.fields_out) return cudaErrorInvalidValue;
  switch (p.ld / 512) {
    case 1: return launch_init_cfg<double, 1, 16, 16, 4>(p, s, tpb);
    case 2: return launch_init_cfg<double, 2, 16, 16, 4>(p, s, tpb);
    case 3: return launch_init_cfg<double, 3, 12, 12, 4>(p, s, tpb);
    case 4: return launch_init_cfg<double, 4, 8, 12, 2>(p, s, tpb);
    case 5: return launch_init_cfg<double, 5, 6, 9, 2>(p, s, tpb);
    case 6: return launch_init_cfg<double, 6, 6, 8, 2>(p, s, tpb);
    case 7: return launch_init_cfg<double, 7, 4, 6, 2>(p, s, tpb);
    case 8: return launch_init_cfg<double, 8, 4, 6, 2>(p, s, tpb);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace osa
