// osa_api.cu -- the C ABI declared in include/onesolver_b200.h.
//
// Host orchestration of the annealing hot path: device-resident problem layouts
// (replacing the sycl::buffer set-up of /root/reference/include/simulated_annealing/
// annealing.hpp:59-72), kernel selection, the exact-energy epilogue and the argmin
// (annealing.hpp:134-139).  No CPU compute path exists here: every failure to reach a
// CUDA device is reported as an error.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <limits>
#include <mutex>
#include <vector>

#include "osa_internal.h"

using namespace osa;

namespace osa {
cudaError_t exhaustive_search(const double *qsym_host, int n, unsigned long long *x_best,
                              double *e_best, std::string *msg);
}

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local std::string g_last_error;

int osa_fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
#define fail osa_fail

#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      int _code = (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver)             \
                      ? OSA_ERR_NO_DEVICE                                                    \
                      : (_e == cudaErrorMemoryAllocation ? OSA_ERR_NOMEM : OSA_ERR_CUDA);    \
      return fail(_code, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                  __LINE__);                                                                 \
    }                                                                                        \
  } while (0)

// ---------------------------------------------------------------------------
// problem handle
// ---------------------------------------------------------------------------
namespace {

size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

// Device memory comes from a PRIVATE stream-ordered pool per device with an unlimited release
// threshold, so the create -> anneal -> destroy cycle of one sa::anneal call reuses pooled memory
// instead of paying cudaMalloc/cudaFree (tens of milliseconds for the 64-128 MiB arrays) every
// time -- without touching the default pool of the host application.
std::mutex g_pool_mutex;
cudaMemPool_t g_pools[64] = {nullptr};

cudaError_t device_pool(int device, cudaMemPool_t *out) {
  if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  if (!g_pools[device]) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    cudaMemPool_t pool = nullptr;
    cudaError_t e = cudaMemPoolCreate(&pool, &props);
    if (e != cudaSuccess) return e;
    unsigned long long threshold = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    g_pools[device] = pool;
  }
  *out = g_pools[device];
  return cudaSuccess;
}

template <typename P>
cudaError_t dev_alloc(P **ptr, size_t bytes, cudaStream_t stream) {
  int device = 0;
  cudaError_t e = cudaGetDevice(&device);
  cudaMemPool_t pool = nullptr;
  if (e == cudaSuccess) e = device_pool(device, &pool);
  if (e != cudaSuccess) return e;
  return cudaMallocFromPoolAsync(reinterpret_cast<void **>(ptr), bytes, pool, stream);
}
template <typename P>
void dev_free(P *ptr, cudaStream_t stream) {
  if (ptr) cudaFreeAsync(const_cast<void *>(static_cast<const void *>(ptr)), stream);
}

}  // namespace

cudaError_t osa_pool_alloc(void **ptr, size_t bytes, cudaStream_t stream) {
  return dev_alloc(reinterpret_cast<char **>(ptr), bytes, stream);
}
void osa_pool_free(void *ptr, cudaStream_t stream) { dev_free(static_cast<char *>(ptr), stream); }

namespace {

// build the sweep-precision layouts from a dense upload
template <typename TIn, typename T>
__global__ void k_prep_dense(const TIn *__restrict__ in, int n, T *__restrict__ qoff, size_t ld,
                             T *__restrict__ diag, double *__restrict__ q64, size_t ld64) {
  const size_t total = (size_t)n * n;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / n), j = (int)(idx % n);
    const TIn v = in[idx];
    q64[(size_t)i * ld64 + j] = (double)v;
    if (i == j) {
      diag[i] = (T)v;
    } else {
      qoff[(size_t)i * ld + j] = (T)v;
    }
  }
}

template <typename TIn>
__global__ void k_check_symmetric(const TIn *__restrict__ in, int n, int *__restrict__ bad) {
  const size_t total = (size_t)n * n;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / n), j = (int)(idx % n);
    if (j > i) {
      const TIn a = in[idx], b = in[(size_t)j * n + i];
      if (!(a == b)) atomicExch(bad, 1);
    }
  }
}

template <typename T>
__global__ void k_convert(const double *__restrict__ in, T *__restrict__ out, size_t count) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < count;
       idx += (size_t)gridDim.x * blockDim.x)
    out[idx] = (T)in[idx];
}

int init_exec(osa_problem *p) {
  CUDA_TRY(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  for (auto &e : p->ev) CUDA_TRY(cudaEventCreate(&e));
  CUDA_TRY(dev_alloc(&p->d_counters, sizeof(Counters), p->stream));
  CUDA_TRY(dev_alloc(&p->d_arg_idx, sizeof(unsigned long long), p->stream));
  CUDA_TRY(dev_alloc(&p->d_arg_e, sizeof(double), p->stream));
  return OSA_OK;
}

int select_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(OSA_ERR_NO_DEVICE, "no CUDA device available: %s",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= count)
    return fail(OSA_ERR_INVALID, "device %d out of range [0, %d)", device, count);
  CUDA_TRY(cudaSetDevice(device));
  return OSA_OK;
}

template <typename TIn>
int create_dense(const TIn *qsym, int n, int device, int prec, osa_problem **out) {
  if (!qsym || !out) return fail(OSA_ERR_INVALID, "null argument");
  if (n < 1) return fail(OSA_ERR_INVALID, "n must be >= 1 (got %d)", n);
  if (prec != OSA_SWEEP_F64 && prec != OSA_SWEEP_F32)
    return fail(OSA_ERR_INVALID, "unknown sweep precision %d", prec);
  DeviceGuard guard;
  int rc = select_device(device);
  if (rc) return rc;

  osa_problem *p = new (std::nothrow) osa_problem();
  if (!p) return fail(OSA_ERR_NOMEM, "out of host memory");
  p->device = device;
  p->n = n;
  p->nw = (n + 31) / 32;
  p->prec = prec;
  auto bail = [&](int code) {
    osa_problem_destroy(p);
    return code;
  };
  rc = init_exec(p);
  if (rc) return bail(rc);

  const size_t esz = prec == OSA_SWEEP_F32 ? 4 : 8;
  p->ld = round_up((size_t)n, prec == OSA_SWEEP_F32 ? 1024 : 512);
  p->rows_pad = round_up((size_t)n, 32);
  p->ld64 = round_up((size_t)n, 32);  // zero padded: the MMA energy kernel reads whole 32x32 tiles

  TIn *d_in = nullptr;
  int *d_bad = nullptr;
#define TRY_B(expr)                                                                          \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      dev_free(d_in, p->stream);                                                             \
      dev_free(d_bad, p->stream);                                                            \
      fail(_e == cudaErrorMemoryAllocation ? OSA_ERR_NOMEM : OSA_ERR_CUDA, "%s failed: %s",  \
           #expr, cudaGetErrorString(_e));                                                   \
      return bail(_e == cudaErrorMemoryAllocation ? OSA_ERR_NOMEM : OSA_ERR_CUDA);           \
    }                                                                                        \
  } while (0)

  const size_t total = (size_t)n * n;
  // The upload buffer is released at the end of this function: it is allocated AFTER the arrays
  // that stay, so that the hole it leaves lies at the end of what the pool has handed out.
  // Allocated first, the hole sat in front of the resident arrays, the 268 MB of shared initial
  // fields of a random-site anneal did not fit it, and in a create / anneal / destroy loop the
  // pool re-mapped its memory every cycle (3-50 ms per allocation, one of a full second:
  // tools/pool_probe.cu, profiles/r02/pool_probe_allocation_order.txt).
  TRY_B(dev_alloc(&p->d_qoff, p->rows_pad * p->ld * esz, p->stream));
  TRY_B(dev_alloc(&p->d_diag, p->ld * esz, p->stream));
  TRY_B(dev_alloc(&p->d_q64, p->rows_pad * p->ld64 * sizeof(double), p->stream));
  TRY_B(dev_alloc(&d_in, total * sizeof(TIn), p->stream));
  TRY_B(dev_alloc(&d_bad, sizeof(int), p->stream));
  TRY_B(cudaMemcpyAsync(d_in, qsym, total * sizeof(TIn), cudaMemcpyHostToDevice, p->stream));
  TRY_B(cudaMemsetAsync(d_bad, 0, sizeof(int), p->stream));
  TRY_B(cudaMemsetAsync(p->d_qoff, 0, p->rows_pad * p->ld * esz, p->stream));
  TRY_B(cudaMemsetAsync(p->d_diag, 0, p->ld * esz, p->stream));
  TRY_B(cudaMemsetAsync(p->d_q64, 0, p->rows_pad * p->ld64 * sizeof(double), p->stream));
  const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  k_check_symmetric<TIn><<<grid, 256, 0, p->stream>>>(d_in, n, d_bad);
  if (prec == OSA_SWEEP_F32)
    k_prep_dense<TIn, float><<<grid, 256, 0, p->stream>>>(d_in, n, (float *)p->d_qoff, p->ld,
                                                          (float *)p->d_diag, p->d_q64, p->ld64);
  else
    k_prep_dense<TIn, double><<<grid, 256, 0, p->stream>>>(d_in, n, (double *)p->d_qoff, p->ld,
                                                           (double *)p->d_diag, p->d_q64, p->ld64);
  TRY_B(cudaGetLastError());
  int bad = 0;
  TRY_B(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
  TRY_B(cudaStreamSynchronize(p->stream));
  dev_free(d_in, p->stream);
  dev_free(d_bad, p->stream);
  d_in = nullptr;
  d_bad = nullptr;
#undef TRY_B
  if (bad) {
    fail(OSA_ERR_INVALID, "Q is not symmetric (expected helpers::flatten_qubo layout)");
    return bail(OSA_ERR_INVALID);
  }
  *out = p;
  return OSA_OK;
}

int ensure_workspace(osa_problem *p, uint64_t num_tries, int num_iter) {
  if (num_tries > p->cap_tries) {
    dev_free(p->d_best_rel, p->stream);
    dev_free(p->d_energy, p->stream);
    p->d_best_rel = nullptr;
    p->d_energy = nullptr;
    p->cap_tries = 0;
    CUDA_TRY(dev_alloc(&p->d_best_rel, num_tries * sizeof(double), p->stream));
    CUDA_TRY(dev_alloc(&p->d_energy, num_tries * sizeof(double), p->stream));
    p->cap_tries = num_tries;
  }
  const size_t words = (size_t)num_tries * p->nw;
  if (words > p->cap_states_words) {
    dev_free(p->d_states, p->stream);
    p->d_states = nullptr;
    p->cap_states_words = 0;
    CUDA_TRY(dev_alloc(&p->d_states, words * sizeof(uint32_t), p->stream));
    p->cap_states_words = words;
  }
  if (p->sparse) {
    const size_t ws = sparse_ws_words(p->n, num_tries);
    if (ws > p->cap_ws_words) {
      dev_free(p->d_xbest_ws, p->stream);
      p->d_xbest_ws = nullptr;
      p->cap_ws_words = 0;
      CUDA_TRY(dev_alloc(&p->d_xbest_ws, ws * sizeof(uint32_t), p->stream));
      p->cap_ws_words = ws;
    }
  }
  const size_t tb = (size_t)num_iter * 8;
  if (tb > p->cap_tscale_bytes) {
    dev_free(p->d_tscale, p->stream);
    p->d_tscale = nullptr;
    p->cap_tscale_bytes = 0;
    CUDA_TRY(dev_alloc(&p->d_tscale, tb, p->stream));
    p->cap_tscale_bytes = tb;
  }
  return OSA_OK;
}

int exact_energies(osa_problem *p, const uint32_t *d_states, uint64_t count, double *d_out) {
  if (p->sparse) {
    CUDA_TRY(launch_energy_csr(p->d_rowptr, p->d_col, p->d_val64, p->d_diag64, p->n, d_states,
                               p->nw, count, d_out, p->stream));
  } else {
    CUDA_TRY(launch_energy_dense(p->d_q64, p->ld64, p->n, d_states, p->nw, count, d_out,
                                 p->stream));
  }
  return OSA_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int osa_abi_version(void) { return OSA_ABI_VERSION; }

const char *osa_last_error(void) { return g_last_error.c_str(); }

int osa_device_count(int *count) {
  if (!count) return fail(OSA_ERR_INVALID, "null argument");
  *count = 0;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(OSA_ERR_NO_DEVICE, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
  }
  return OSA_OK;
}

int osa_device_name(int device, char *buf, size_t buflen) {
  if (!buf || buflen == 0) return fail(OSA_ERR_INVALID, "null argument");
  DeviceGuard guard;
  int rc = select_device(device);
  if (rc) return rc;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  snprintf(buf, buflen, "%s", prop.name);
  return OSA_OK;
}

const char *osa_kernel_name(int kernel_id) {
  switch (kernel_id) {
    case KID_DENSE_SEQ: return "dense_seq";
    case KID_DENSE_GENERIC: return "dense_generic";
    case KID_SPARSE: return "sparse_csr";
    default: return "auto";
  }
}

int osa_problem_create_dense_f64(const double *qsym, int n, int device, int sweep_precision,
                                 osa_problem **out) {
  return create_dense<double>(qsym, n, device, sweep_precision, out);
}

int osa_problem_create_dense_f32(const float *qsym, int n, int device, osa_problem **out) {
  return create_dense<float>(qsym, n, device, OSA_SWEEP_F32, out);
}

int osa_problem_create_csr_f64(const int32_t *rowptr, const int32_t *col, const double *val,
                               const double *diag, int n, int device, int sweep_precision,
                               osa_problem **out) {
  if (!rowptr || !diag || !out) return fail(OSA_ERR_INVALID, "null argument");
  if (n < 1) return fail(OSA_ERR_INVALID, "n must be >= 1 (got %d)", n);
  if (sweep_precision != OSA_SWEEP_F64 && sweep_precision != OSA_SWEEP_F32)
    return fail(OSA_ERR_INVALID, "unknown sweep precision %d", sweep_precision);
  if (rowptr[0] != 0) return fail(OSA_ERR_INVALID, "rowptr[0] must be 0");
  const int64_t nnz = rowptr[n];
  if (nnz < 0 || (nnz > 0 && (!col || !val))) return fail(OSA_ERR_INVALID, "bad CSR arrays");
  if (!sparse_supported(n, sweep_precision == OSA_SWEEP_F32 ? 4 : 8))
    return fail(OSA_ERR_UNSUPPORTED,
                "sparse kernel supports n <= about %d (spin words of 32 trajectories in shared memory)",
                (227 * 1024 - 8192) / 4);
  for (int i = 0; i < n; ++i) {
    if (rowptr[i + 1] < rowptr[i]) return fail(OSA_ERR_INVALID, "rowptr not monotone at %d", i);
    for (int32_t q = rowptr[i]; q < rowptr[i + 1]; ++q) {
      if (col[q] < 0 || col[q] >= n || col[q] == i)
        return fail(OSA_ERR_INVALID, "row %d: bad column %d", i, col[q]);
      if (q > rowptr[i] && col[q] <= col[q - 1])
        return fail(OSA_ERR_INVALID, "row %d: columns must be strictly ascending", i);
    }
  }
  DeviceGuard guard;
  // both directions of every coupling, with the same value: the sweep reads row i for the local
  // field of i while the exact energies sum the entries with col > i, so a one-sided or
  // asymmetric upload would make the two disagree without any error (the dense path rejects an
  // asymmetric Q in the same way, k_check_symmetric)
  for (int i = 0; i < n; ++i) {
    for (int32_t q = rowptr[i]; q < rowptr[i + 1]; ++q) {
      const int j = col[q];
      const int32_t *lo = col + rowptr[j], *hi = col + rowptr[j + 1];
      const int32_t *it = std::lower_bound(lo, hi, (int32_t)i);
      if (it == hi || *it != i)
        return fail(OSA_ERR_INVALID, "CSR is not symmetric: entry (%d, %d) has no (%d, %d)", i, j, j, i);
      if (!(val[it - col] == val[q]))
        return fail(OSA_ERR_INVALID, "CSR is not symmetric: value(%d, %d) != value(%d, %d)", i, j, j, i);
    }
  }
  int rc = select_device(device);
  if (rc) return rc;
  osa_problem *p = new (std::nothrow) osa_problem();
  if (!p) return fail(OSA_ERR_NOMEM, "out of host memory");
  p->device = device;
  p->n = n;
  p->nw = (n + 31) / 32;
  p->sparse = true;
  p->prec = sweep_precision;
  p->nnz = nnz;
  rc = init_exec(p);
  if (rc) {
    osa_problem_destroy(p);
    return rc;
  }
  const size_t esz = sweep_precision == OSA_SWEEP_F32 ? 4 : 8;
  const size_t nnz_a = (size_t)(nnz > 0 ? nnz : 1);
  cudaError_t e = cudaSuccess;
  auto step = [&](cudaError_t r) {
    if (e == cudaSuccess) e = r;
  };
  step(dev_alloc(&p->d_rowptr, (size_t)(n + 1) * sizeof(int32_t), p->stream));
  step(dev_alloc(&p->d_col, nnz_a * sizeof(int32_t), p->stream));
  step(dev_alloc(&p->d_val64, nnz_a * sizeof(double), p->stream));
  step(dev_alloc(&p->d_diag64, (size_t)n * sizeof(double), p->stream));
  step(dev_alloc(&p->d_val, (nnz_a + (size_t)n) * esz, p->stream));
  // grouped layout of the sequential sweeps (osa_sparse.cu): groups of G consecutive sites,
  // entry (t, k) = t-th neighbour of site G g + k as {byte offset of its spin word, coupling in the
  // sweep precision}, short rows padded with {offset of the zero word n, 0}; a group is flagged
  // independent when no coupling joins two of its sites.  G = 8 when at least three quarters of
  // the 8-groups are independent (random graphs), else 4 (Chimera / Pegasus numbering: the four
  // qubits of a shore are independent, the eight of a cell are not).
  const int nblk32 = (n + 31) / 32;
  auto independent = [&](int i0, int g) {
    if (i0 + g > n) return false;
    for (int a = i0; a < i0 + g; ++a)
      for (int32_t q = rowptr[a]; q < rowptr[a + 1]; ++q)
        if (col[q] >= i0 && col[q] < i0 + g) return false;
    return true;
  };
  int free8 = 0;
  for (int i0 = 0; i0 + 8 <= n; i0 += 8) free8 += independent(i0, 8) ? 1 : 0;
  const int G = (n >= 8 && free8 * 4 >= (n / 8) * 3) ? 8 : 4;
  p->group = G;
  const int ngroups = nblk32 * (32 / G);
  std::vector<uint32_t> gbase((size_t)ngroups + 1), ginfo((size_t)ngroups), ent_off;
  std::vector<double> ent_val;
  ent_off.reserve((size_t)nnz + (size_t)n);
  ent_val.reserve((size_t)nnz + (size_t)n);
  for (int g = 0; g < ngroups; ++g) {
    const int i0 = g * G;
    gbase[(size_t)g] = (uint32_t)ent_off.size();
    int len = 0;
    for (int a = i0; a < i0 + G && a < n; ++a) len = std::max(len, (int)(rowptr[a + 1] - rowptr[a]));
    const bool indep = independent(i0, G);
    if (len > 0xffff) {
      osa_problem_destroy(p);
      return fail(OSA_ERR_UNSUPPORTED, "sparse kernel supports at most 65535 neighbours per site");
    }
    for (int t = 0; t < len; ++t)
      for (int k = 0; k < G; ++k) {
        const int a = i0 + k;
        const bool real = a < n && t < rowptr[a + 1] - rowptr[a];
        ent_off.push_back(real ? (uint32_t)col[rowptr[a] + t] * 4u : (uint32_t)n * 4u);
        ent_val.push_back(real ? val[rowptr[a] + t] : 0.0);
      }
    ginfo[(size_t)g] = (uint32_t)len | (indep ? 0x10000u : 0u);
  }
  gbase[(size_t)ngroups] = (uint32_t)ent_off.size();
  p->stage_ok = 1;
  const int gph = 16 / G;  // groups per staged half block
  for (int h = 0; h < nblk32 * 2; ++h)
    if (gbase[(size_t)(h + 1) * gph] - gbase[(size_t)h * gph] > (uint32_t)SPARSE_HALF_CAP)
      p->stage_ok = 0;
  // 8 entries of slack: the gather requests one row past the end of a group (and drops it)
  const size_t nent = ent_off.size(), nent_a = nent + 8;
  const size_t ent_bytes = sweep_precision == OSA_SWEEP_F32 ? 8 : 16;
  std::vector<unsigned char> ent(nent_a * ent_bytes, 0);
  for (size_t q = 0; q < nent; ++q) {
    unsigned char *dst = ent.data() + q * ent_bytes;
    memcpy(dst, &ent_off[q], 4);
    if (sweep_precision == OSA_SWEEP_F32) {
      const float v = (float)ent_val[q];  // round to nearest, like k_convert
      memcpy(dst + 4, &v, 4);
    } else {
      memcpy(dst + 8, &ent_val[q], 8);
    }
  }
  step(dev_alloc(&p->d_gbase, ((size_t)ngroups + 1) * sizeof(uint32_t), p->stream));
  step(dev_alloc(&p->d_ginfo, (size_t)ngroups * sizeof(uint32_t), p->stream));
  step(dev_alloc(&p->d_gent, ent.size(), p->stream));
  if (e == cudaSuccess) {
    step(cudaMemcpyAsync(p->d_gbase, gbase.data(), gbase.size() * sizeof(uint32_t),
                         cudaMemcpyHostToDevice, p->stream));
    step(cudaMemcpyAsync(p->d_ginfo, ginfo.data(), ginfo.size() * sizeof(uint32_t),
                         cudaMemcpyHostToDevice, p->stream));
    step(cudaMemcpyAsync(p->d_gent, ent.data(), ent.size(), cudaMemcpyHostToDevice, p->stream));
    step(cudaMemcpyAsync(p->d_rowptr, rowptr, (size_t)(n + 1) * sizeof(int32_t),
                         cudaMemcpyHostToDevice, p->stream));
    if (nnz > 0) {
      step(cudaMemcpyAsync(p->d_col, col, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice,
                           p->stream));
      step(cudaMemcpyAsync(p->d_val64, val, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice,
                           p->stream));
    }
    step(cudaMemcpyAsync(p->d_diag64, diag, (size_t)n * sizeof(double), cudaMemcpyHostToDevice,
                         p->stream));
    // sweep-precision copies: val followed by diag in one allocation
    if (sweep_precision == OSA_SWEEP_F32) {
      if (nnz > 0) k_convert<float><<<64, 256, 0, p->stream>>>(p->d_val64, (float *)p->d_val, nnz);
      k_convert<float><<<64, 256, 0, p->stream>>>(p->d_diag64, (float *)p->d_val + nnz_a, n);
    } else {
      if (nnz > 0)
        k_convert<double><<<64, 256, 0, p->stream>>>(p->d_val64, (double *)p->d_val, nnz);
      k_convert<double><<<64, 256, 0, p->stream>>>(p->d_diag64, (double *)p->d_val + nnz_a, n);
    }
    step(cudaGetLastError());
    step(cudaStreamSynchronize(p->stream));
  }
  if (e != cudaSuccess) {
    osa_problem_destroy(p);
    return fail(e == cudaErrorMemoryAllocation ? OSA_ERR_NOMEM : OSA_ERR_CUDA,
                "CSR upload failed: %s", cudaGetErrorString(e));
  }
  p->d_diag = (char *)p->d_val + nnz_a * esz;  // alias inside d_val; not freed separately
  *out = p;
  return OSA_OK;
}

int osa_problem_destroy(osa_problem *p) {
  if (!p) return OSA_OK;
  DeviceGuard guard;
  cudaSetDevice(p->device);
  cudaStream_t st = p->stream;
  if (p->sparse) {
    dev_free(p->d_rowptr, st);
    dev_free(p->d_col, st);
    dev_free(p->d_val, st);
    dev_free(p->d_val64, st);
    dev_free(p->d_diag64, st);
    dev_free(p->d_gbase, st);
    dev_free(p->d_ginfo, st);
    dev_free(p->d_gent, st);
  } else {
    dev_free(p->d_qoff, st);
    dev_free(p->d_diag, st);
    dev_free(p->d_q64, st);
  }
  dev_free(p->d_best_rel, st);
  dev_free(p->d_energy, st);
  dev_free(p->d_states, st);
  dev_free(p->d_xbest_ws, st);
  dev_free(p->d_trace, st);
  dev_free(p->d_fields, st);
  dev_free(p->d_tscale, st);
  dev_free(p->d_counters, st);
  dev_free(p->d_arg_idx, st);
  dev_free(p->d_arg_e, st);
  if (st) cudaStreamSynchronize(st);
  for (auto &e : p->ev)
    if (e) cudaEventDestroy(e);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
  return OSA_OK;
}

int osa_problem_size(const osa_problem *p, int *n, int *is_sparse, int *sweep_precision) {
  if (!p) return fail(OSA_ERR_INVALID, "null problem");
  if (n) *n = p->n;
  if (is_sparse) *is_sparse = p->sparse ? 1 : 0;
  if (sweep_precision) *sweep_precision = p->prec;
  return OSA_OK;
}

int osa_anneal(osa_problem *p, const double *beta_schedule, const osa_anneal_params *prm,
               double *best_energies, uint32_t *best_states_packed, uint8_t *best_state,
               double *best_energy, uint64_t *best_index, osa_stats *stats) {
  return osa_anneal_traced(p, beta_schedule, prm, best_energies, best_states_packed, best_state,
                           best_energy, best_index, nullptr, stats);
}

int osa_anneal_traced(osa_problem *p, const double *beta_schedule, const osa_anneal_params *prm,
                      double *best_energies, uint32_t *best_states_packed, uint8_t *best_state,
                      double *best_energy, uint64_t *best_index, uint64_t *trace_hash,
                      osa_stats *stats) {
  if (!p || !beta_schedule || !prm) return fail(OSA_ERR_INVALID, "null argument");
  if (prm->num_iter < 1) return fail(OSA_ERR_INVALID, "num_iter must be >= 1");
  if (prm->sweeps_per_beta < 1) return fail(OSA_ERR_INVALID, "sweeps_per_beta must be >= 1");
  if (prm->num_tries < 1) return fail(OSA_ERR_INVALID, "num_tries must be >= 1");
  if (prm->mode != OSA_MODE_RANDOM_SITE && prm->mode != OSA_MODE_SEQUENTIAL_SWEEP)
    return fail(OSA_ERR_INVALID, "unknown mode %d", prm->mode);
  if (prm->flags != 0) return fail(OSA_ERR_INVALID, "unknown flags 0x%x", prm->flags);
  if (prm->accept_rule != OSA_ACCEPT_REFERENCE && prm->accept_rule != OSA_ACCEPT_BOLTZMANN)
    return fail(OSA_ERR_INVALID, "unknown accept rule %d", prm->accept_rule);
  if ((uint64_t)prm->num_iter * (uint64_t)prm->sweeps_per_beta >= (1ull << 32))
    return fail(OSA_ERR_INVALID, "num_iter * sweeps_per_beta must be < 2^32");
  if (prm->first_try + prm->num_tries < prm->first_try ||
      (prm->first_try + prm->num_tries) >> 62)
    return fail(OSA_ERR_INVALID, "trajectory ids must stay below 2^62");
  for (int i = 0; i < prm->num_iter; ++i)
    if (!(beta_schedule[i] > 0.0) || !std::isfinite(beta_schedule[i]))
      return fail(OSA_ERR_INVALID, "beta_schedule[%d] = %g is not a positive finite number", i,
                  beta_schedule[i]);

  DeviceGuard guard;
  int rc = select_device(p->device);
  if (rc) return rc;
  rc = ensure_workspace(p, prm->num_tries, prm->num_iter);
  if (rc) return rc;
  if (trace_hash && prm->num_tries > p->cap_trace) {
    dev_free(p->d_trace, p->stream);
    p->d_trace = nullptr;
    p->cap_trace = 0;
    CUDA_TRY(dev_alloc(&p->d_trace, prm->num_tries * sizeof(unsigned long long), p->stream));
    p->cap_trace = prm->num_tries;
  }

  // threshold scale per iteration: accept iff dE < tscale * (-ln u)
  const bool f32 = p->prec == OSA_SWEEP_F32;
  std::vector<double> ts64(prm->num_iter);
  std::vector<float> ts32(f32 ? prm->num_iter : 0);
  for (int i = 0; i < prm->num_iter; ++i) {
    ts64[i] = prm->accept_rule == OSA_ACCEPT_REFERENCE ? beta_schedule[i] : 1.0 / beta_schedule[i];
    if (f32) ts32[i] = (float)ts64[i];
  }
  if (f32)
    CUDA_TRY(cudaMemcpyAsync(p->d_tscale, ts32.data(), ts32.size() * sizeof(float),
                             cudaMemcpyHostToDevice, p->stream));
  else
    CUDA_TRY(cudaMemcpyAsync(p->d_tscale, ts64.data(), ts64.size() * sizeof(double),
                             cudaMemcpyHostToDevice, p->stream));
  CUDA_TRY(cudaMemsetAsync(p->d_counters, 0, sizeof(Counters), p->stream));

  int kid = prm->kernel_variant;
  if (p->sparse) {
    if (kid != KID_AUTO && kid != KID_SPARSE)
      return fail(OSA_ERR_UNSUPPORTED, "kernel %d cannot run a CSR problem", kid);
    kid = KID_SPARSE;
  } else {
    const int esz = f32 ? 4 : 8;
    if (kid == KID_AUTO)
      kid = (prm->mode == OSA_MODE_SEQUENTIAL_SWEEP && dense_seq_supported(p->n, esz))
                ? KID_DENSE_SEQ
                : KID_DENSE_GENERIC;
    if (kid == KID_DENSE_SEQ &&
        (prm->mode != OSA_MODE_SEQUENTIAL_SWEEP || !dense_seq_supported(p->n, esz)))
      return fail(OSA_ERR_UNSUPPORTED, "dense_seq kernel needs sequential mode and n <= %d",
                  f32 ? 8192 : 4096);
    if (kid == KID_DENSE_GENERIC && !dense_generic_supported(p->n, esz))
      return fail(OSA_ERR_UNSUPPORTED, "n = %d exceeds the shared-memory resident kernel", p->n);
    if (kid != KID_DENSE_SEQ && kid != KID_DENSE_GENERIC)
      return fail(OSA_ERR_UNSUPPORTED, "kernel %d cannot run a dense problem", kid);
  }

  LaunchInfo info = {0, 0, 0, 0};
  int launches = 0;
  int shared_init = 0;  // trajectories per shared row fetch of the initial-field kernel (0: not used)
  CUDA_TRY(cudaEventRecord(p->ev[0], p->stream));
  if (p->sparse) {
    auto run = [&](auto tag) -> cudaError_t {
      using T = decltype(tag);
      SparseParams<T> sp;
      sp.rowptr = p->d_rowptr;
      sp.col = p->d_col;
      sp.val = (const T *)p->d_val;
      sp.diag = (const T *)p->d_diag;
      sp.tscale = (const T *)p->d_tscale;
      sp.n = p->n;
      sp.num_iter = prm->num_iter;
      sp.sweeps_per_beta = prm->sweeps_per_beta;
      sp.mode = prm->mode;
      sp.seed = prm->seed;
      sp.first_try = prm->first_try;
      sp.num_tries = prm->num_tries;
      sp.best_rel = p->d_best_rel;
      sp.best_states = p->d_states;
      sp.xbest_ws = p->d_xbest_ws;
      sp.debug_flags = 0;
      sp.log_base = 0;
      sp.nw = p->nw;
      sp.counters = p->d_counters;
      sp.trace_hash = trace_hash ? p->d_trace : nullptr;
      sp.gbase = p->d_gbase;
      sp.ginfo = p->d_ginfo;
      sp.gent = p->d_gent;
      sp.stage_ok = p->stage_ok;
      sp.group = p->group;
      return launch_sparse<T>(sp, p->stream, &info);
    };
    CUDA_TRY(f32 ? run(float()) : run(double()));
  } else {
    auto run = [&](auto tag) -> cudaError_t {
      using T = decltype(tag);
      DenseParams<T> dp{};
      dp.qoff = (const T *)p->d_qoff;
      dp.diag = (const T *)p->d_diag;
      dp.tscale = (const T *)p->d_tscale;
      dp.ld = p->ld;
      dp.n = p->n;
      dp.num_iter = prm->num_iter;
      dp.sweeps_per_beta = prm->sweeps_per_beta;
      dp.mode = prm->mode;
      dp.seed = prm->seed;
      dp.first_try = prm->first_try;
      dp.num_tries = prm->num_tries;
      dp.best_rel = p->d_best_rel;
      dp.best_states = p->d_states;
      dp.nw = p->nw;
      dp.counters = p->d_counters;
      dp.trace_hash = trace_hash ? p->d_trace : nullptr;
      if (kid == KID_DENSE_SEQ) return launch_dense_seq<T>(dp, p->stream, &info);
      // warp-per-trajectory kernel: the initial fields of all trajectories are built first, with
      // row fetches shared by a CTA's trajectories (osa_dense_init.cu), when the shape is one the
      // streaming machinery covers and the field buffer stays within 8 GiB
      const size_t field_bytes = (size_t)prm->num_tries * p->ld * sizeof(T);
      // (not below N = 256: the rows of the field buffer are padded to ld >= 1024 floats, and a
      // trajectory's own build reads N/2 short rows only)
      if (p->n >= 256 && dense_seq_supported(p->n, (int)sizeof(T)) && field_bytes <= (8ull << 30)) {
        if (field_bytes > p->cap_fields_bytes) {
          dev_free(p->d_fields, p->stream);
          p->d_fields = nullptr;
          p->cap_fields_bytes = 0;
          cudaError_t ea = dev_alloc(&p->d_fields, field_bytes, p->stream);
          if (ea != cudaSuccess) return ea;
          p->cap_fields_bytes = field_bytes;
        }
        dp.fields_out = (T *)p->d_fields;
        cudaError_t ei = launch_dense_init_fields<T>(dp, p->stream, &shared_init);
        if (ei != cudaSuccess) return ei;
        ++launches;
        dp.fields_in = (const T *)p->d_fields;
        dp.fields_out = nullptr;
      }
      return launch_dense_generic<T>(dp, p->stream, &info);
    };
    CUDA_TRY(f32 ? run(float()) : run(double()));
  }
  ++launches;
  CUDA_TRY(cudaEventRecord(p->ev[1], p->stream));

  // exact energies of the per-trajectory best states (annealing.hpp:125 semantics)
  rc = exact_energies(p, p->d_states, prm->num_tries, p->d_energy);
  if (rc) return rc;
  ++launches;
  CUDA_TRY(cudaEventRecord(p->ev[2], p->stream));
  CUDA_TRY(launch_argmin(p->d_energy, prm->num_tries, p->d_arg_idx, p->d_arg_e, p->stream));
  ++launches;
  CUDA_TRY(cudaEventRecord(p->ev[3], p->stream));

  unsigned long long h_idx = 0;
  double h_e = 0.0;
  Counters h_cnt;
  CUDA_TRY(cudaMemcpyAsync(&h_idx, p->d_arg_idx, sizeof(h_idx), cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaMemcpyAsync(&h_e, p->d_arg_e, sizeof(h_e), cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaMemcpyAsync(&h_cnt, p->d_counters, sizeof(h_cnt), cudaMemcpyDeviceToHost, p->stream));
  cudaError_t sync_err = cudaStreamSynchronize(p->stream);
  if (sync_err != cudaSuccess)
    return fail(OSA_ERR_CUDA, "annealing kernels failed: %s", cudaGetErrorString(sync_err));
  if (h_idx >= prm->num_tries) return fail(OSA_ERR_CUDA, "argmin returned an invalid index");

  if (best_energies)
    CUDA_TRY(cudaMemcpyAsync(best_energies, p->d_energy, prm->num_tries * sizeof(double),
                             cudaMemcpyDeviceToHost, p->stream));
  if (best_states_packed)
    CUDA_TRY(cudaMemcpyAsync(best_states_packed, p->d_states,
                             (size_t)prm->num_tries * p->nw * sizeof(uint32_t),
                             cudaMemcpyDeviceToHost, p->stream));
  if (trace_hash)
    CUDA_TRY(cudaMemcpyAsync(trace_hash, p->d_trace, prm->num_tries * sizeof(uint64_t),
                             cudaMemcpyDeviceToHost, p->stream));
  std::vector<uint32_t> win(p->nw);
  CUDA_TRY(cudaMemcpyAsync(win.data(), p->d_states + (size_t)h_idx * p->nw,
                           (size_t)p->nw * sizeof(uint32_t), cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  if (best_state)
    for (int i = 0; i < p->n; ++i) best_state[i] = (uint8_t)((win[i >> 5] >> (i & 31)) & 1u);
  if (best_energy) *best_energy = h_e;
  if (best_index) *best_index = prm->first_try + h_idx;

  if (stats) {
    memset(stats, 0, sizeof(*stats));
    const uint64_t per = (uint64_t)prm->num_iter * (uint64_t)prm->sweeps_per_beta *
                         (prm->mode == OSA_MODE_SEQUENTIAL_SWEEP ? (uint64_t)p->n : 1ull);
    stats->attempts = per * prm->num_tries;
    stats->accepts = h_cnt.accepts;
    stats->row_fetches = h_cnt.row_fetches;
    stats->init_row_fetches = h_cnt.init_row_fetches;
    stats->cyc_decide = h_cnt.cyc_decide;
    stats->cyc_apply = h_cnt.cyc_apply;
    stats->cyc_stage = h_cnt.cyc_stage;
    stats->cyc_init = h_cnt.cyc_init + h_cnt.pad;  // pad: only with OSA_WS_DEBUG=8
    cudaEventElapsedTime(&stats->ms_sweep, p->ev[0], p->ev[1]);
    cudaEventElapsedTime(&stats->ms_energy, p->ev[1], p->ev[2]);
    cudaEventElapsedTime(&stats->ms_reduce, p->ev[2], p->ev[3]);
    cudaEventElapsedTime(&stats->ms_total, p->ev[0], p->ev[3]);
    stats->kernel_id = kid;
    stats->traj_per_batch = info.traj_per_batch;
    stats->q_elem_bytes = f32 ? 4 : 8;
    stats->grid = info.grid;
    stats->launches = launches;
  }
  return OSA_OK;
}

int osa_pt_anneal(osa_problem *p, const double *betas, const osa_pt_params *prm,
                  double *best_energies, uint32_t *best_states_packed, uint8_t *best_state,
                  double *best_energy, uint64_t *best_index, osa_stats *stats) {
  if (!p || !betas || !prm) return fail(OSA_ERR_INVALID, "null argument");
  if (prm->num_groups < 1) return fail(OSA_ERR_INVALID, "num_groups must be >= 1");
  if (prm->num_replicas < 1 || prm->num_replicas > 4096)
    return fail(OSA_ERR_INVALID, "num_replicas must be in 1..4096");
  if (prm->num_rounds < 1) return fail(OSA_ERR_INVALID, "num_rounds must be >= 1");
  if (prm->sweeps_per_round < 1) return fail(OSA_ERR_INVALID, "sweeps_per_round must be >= 1");
  if (prm->flags != 0) return fail(OSA_ERR_INVALID, "unknown flags 0x%x", prm->flags);
  if (prm->accept_rule != OSA_ACCEPT_REFERENCE && prm->accept_rule != OSA_ACCEPT_BOLTZMANN)
    return fail(OSA_ERR_INVALID, "unknown accept rule %d", prm->accept_rule);
  if ((uint64_t)prm->num_rounds * (uint64_t)prm->sweeps_per_round >= (1ull << 32))
    return fail(OSA_ERR_INVALID, "num_rounds * sweeps_per_round must be < 2^32");
  const int M = prm->num_replicas;
  if (prm->num_groups > (1ull << 40) || (prm->first_group + prm->num_groups) > (1ull << 40))
    return fail(OSA_ERR_INVALID, "group ids must stay below 2^40");
  for (int j = 0; j < M; ++j) {
    if (!(betas[j] > 0.0) || !std::isfinite(betas[j]))
      return fail(OSA_ERR_INVALID, "betas[%d] = %g is not a positive finite number", j, betas[j]);
    if (j > 0 && !(betas[j] > betas[j - 1]))
      return fail(OSA_ERR_INVALID, "betas must be strictly increasing (betas[%d] <= betas[%d])", j,
                  j - 1);
  }
  const bool f32 = p->prec == OSA_SWEEP_F32;
  if (p->sparse || !dense_seq_supported(p->n, f32 ? 4 : 8))
    return fail(OSA_ERR_UNSUPPORTED,
                "parallel tempering runs on dense problems with n <= %d (this sweep precision)",
                f32 ? 8192 : 4096);
  const uint64_t tries = prm->num_groups * (uint64_t)M;
  const uint64_t first_try = prm->first_group * (uint64_t)M;

  DeviceGuard guard;
  int rc = select_device(p->device);
  if (rc) return rc;
  rc = ensure_workspace(p, tries, 1);
  if (rc) return rc;

  // ladder: threshold scale per rung (accept iff dE < tscale * (-ln u)) and the differences of the
  // inverse temperatures that weigh the exchanges
  std::vector<double> ts64(M), dinv(M, 0.0);
  std::vector<float> ts32(M);
  for (int j = 0; j < M; ++j) {
    ts64[j] = prm->accept_rule == OSA_ACCEPT_REFERENCE ? betas[j] : 1.0 / betas[j];
    ts32[j] = (float)ts64[j];
  }
  for (int j = 0; j + 1 < M; ++j) {
    const double bj = prm->accept_rule == OSA_ACCEPT_REFERENCE ? 1.0 / betas[j] : betas[j];
    const double bk = prm->accept_rule == OSA_ACCEPT_REFERENCE ? 1.0 / betas[j + 1] : betas[j + 1];
    dinv[j] = bj - bk;
  }
  std::vector<int32_t> rung(tries);
  for (uint64_t t = 0; t < tries; ++t) rung[t] = (int32_t)(t % (uint64_t)M);

  const size_t esz = f32 ? 4 : 8;
  const size_t words = (size_t)tries * p->nw;
  uint32_t *d_cur = nullptr, *d_keep = nullptr;
  double *d_ecur = nullptr, *d_beste = nullptr, *d_dinv = nullptr;
  char *d_fields = nullptr;  // [tries][ld] local fields carried from round to round
  void *d_ts_traj = nullptr, *d_ts_rung = nullptr;
  int32_t *d_temp_of_slot = nullptr, *d_slot_of_temp = nullptr;
  unsigned long long *d_swaps = nullptr;
  cudaStream_t st = p->stream;
  auto release = [&]() {
    dev_free(d_cur, st);
    dev_free(d_fields, st);
    dev_free(d_keep, st);
    dev_free(d_ecur, st);
    dev_free(d_beste, st);
    dev_free(d_dinv, st);
    dev_free((char *)d_ts_traj, st);
    dev_free((char *)d_ts_rung, st);
    dev_free(d_temp_of_slot, st);
    dev_free(d_slot_of_temp, st);
    dev_free(d_swaps, st);
  };
#define PT_TRY(expr)                                                                        \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      release();                                                                            \
      return fail(e_ == cudaErrorMemoryAllocation ? OSA_ERR_NOMEM : OSA_ERR_CUDA,           \
                  "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), "osa_api.cu", __LINE__); \
    }                                                                                       \
  } while (0)
  PT_TRY(dev_alloc(&d_cur, words * sizeof(uint32_t), st));
  PT_TRY(dev_alloc(&d_fields, (size_t)tries * p->ld * esz, st));
  PT_TRY(dev_alloc(&d_keep, words * sizeof(uint32_t), st));
  PT_TRY(dev_alloc(&d_ecur, tries * sizeof(double), st));
  PT_TRY(dev_alloc(&d_beste, tries * sizeof(double), st));
  PT_TRY(dev_alloc(&d_dinv, (size_t)M * sizeof(double), st));
  PT_TRY(dev_alloc((char **)&d_ts_traj, tries * esz, st));
  PT_TRY(dev_alloc((char **)&d_ts_rung, (size_t)M * esz, st));
  PT_TRY(dev_alloc(&d_temp_of_slot, tries * sizeof(int32_t), st));
  PT_TRY(dev_alloc(&d_slot_of_temp, tries * sizeof(int32_t), st));
  PT_TRY(dev_alloc(&d_swaps, sizeof(unsigned long long), st));

  PT_TRY(cudaMemcpyAsync(d_dinv, dinv.data(), (size_t)M * sizeof(double), cudaMemcpyHostToDevice, st));
  PT_TRY(cudaMemcpyAsync(d_ts_rung, f32 ? (const void *)ts32.data() : (const void *)ts64.data(),
                         (size_t)M * esz, cudaMemcpyHostToDevice, st));
  PT_TRY(cudaMemcpyAsync(d_temp_of_slot, rung.data(), tries * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  PT_TRY(cudaMemcpyAsync(d_slot_of_temp, rung.data(), tries * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  {
    // slot k starts on rung k
    std::vector<unsigned char> init(tries * esz);
    for (uint64_t t = 0; t < tries; ++t) {
      if (f32) ((float *)init.data())[t] = ts32[t % (uint64_t)M];
      else ((double *)init.data())[t] = ts64[t % (uint64_t)M];
    }
    PT_TRY(cudaMemcpyAsync(d_ts_traj, init.data(), tries * esz, cudaMemcpyHostToDevice, st));
    PT_TRY(cudaStreamSynchronize(st));  // `init` goes out of scope
  }
  {
    std::vector<double> inf(tries, std::numeric_limits<double>::infinity());
    PT_TRY(cudaMemcpyAsync(d_beste, inf.data(), tries * sizeof(double), cudaMemcpyHostToDevice, st));
    PT_TRY(cudaStreamSynchronize(st));
  }
  PT_TRY(cudaMemsetAsync(p->d_counters, 0, sizeof(Counters), st));
  PT_TRY(cudaMemsetAsync(d_swaps, 0, sizeof(unsigned long long), st));

  LaunchInfo info = {0, 0, 0, 0};
  int launches = 0;
  float ms_sweep = 0.f, ms_energy = 0.f;
  PT_TRY(cudaEventRecord(p->ev[0], st));
  PT_TRY(launch_pt_init_states(prm->seed, first_try, tries, p->n, p->nw, d_cur, st));
  ++launches;
  rc = exact_energies(p, d_cur, tries, d_ecur);
  if (rc) {
    release();
    return rc;
  }
  ++launches;
  for (int round = 0; round < prm->num_rounds; ++round) {
    auto run = [&](auto tag) -> cudaError_t {
      using T = decltype(tag);
      DenseParams<T> dp{};
      dp.qoff = (const T *)p->d_qoff;
      dp.diag = (const T *)p->d_diag;
      dp.tscale = nullptr;
      dp.ld = p->ld;
      dp.n = p->n;
      dp.num_iter = 1;
      dp.sweeps_per_beta = prm->sweeps_per_round;
      dp.mode = OSA_MODE_SEQUENTIAL_SWEEP;
      dp.seed = prm->seed;
      dp.first_try = first_try;
      dp.num_tries = tries;
      dp.best_rel = p->d_best_rel;
      dp.best_states = p->d_states;
      dp.nw = p->nw;
      dp.counters = p->d_counters;
      dp.init_states = d_cur;
      dp.final_states = d_cur;
      dp.tscale_traj = (const T *)d_ts_traj;
      dp.step_base = (uint32_t)round * (uint32_t)prm->sweeps_per_round;
      // the first round builds the fields from the spins, later rounds continue from what the
      // round before left (states stay in their slots; only temperatures move)
      dp.fields_in = round > 0 ? (const T *)d_fields : nullptr;
      dp.fields_out = (T *)d_fields;
      return launch_dense_seq_ws<T>(dp, st, &info);
    };
    PT_TRY(f32 ? run(float()) : run(double()));
    // best state of the round against the best kept so far (e_cur = energy the round started from)
    PT_TRY(launch_pt_track_best(d_ecur, p->d_best_rel, p->d_states, tries, p->nw, d_beste, d_keep, st));
    rc = exact_energies(p, d_cur, tries, d_ecur);
    if (rc) {
      release();
      return rc;
    }
    if (f32)
      PT_TRY(launch_pt_swap<float>(prm->seed, prm->first_group, prm->num_groups, M, (uint32_t)round,
                                   d_ecur, d_dinv, (const float *)d_ts_rung, d_temp_of_slot,
                                   d_slot_of_temp, (float *)d_ts_traj, d_swaps, st));
    else
      PT_TRY(launch_pt_swap<double>(prm->seed, prm->first_group, prm->num_groups, M, (uint32_t)round,
                                    d_ecur, d_dinv, (const double *)d_ts_rung, d_temp_of_slot,
                                    d_slot_of_temp, (double *)d_ts_traj, d_swaps, st));
    launches += 4;
  }
  PT_TRY(cudaEventRecord(p->ev[1], st));
  // exact energies of the kept best states, then the usual argmin (lowest id wins ties)
  rc = exact_energies(p, d_keep, tries, p->d_energy);
  if (rc) {
    release();
    return rc;
  }
  ++launches;
  PT_TRY(cudaEventRecord(p->ev[2], st));
  PT_TRY(launch_argmin(p->d_energy, tries, p->d_arg_idx, p->d_arg_e, st));
  ++launches;
  PT_TRY(cudaEventRecord(p->ev[3], st));

  unsigned long long h_idx = 0, h_swaps = 0;
  double h_e = 0.0;
  Counters h_cnt;
  PT_TRY(cudaMemcpyAsync(&h_idx, p->d_arg_idx, sizeof(h_idx), cudaMemcpyDeviceToHost, st));
  PT_TRY(cudaMemcpyAsync(&h_e, p->d_arg_e, sizeof(h_e), cudaMemcpyDeviceToHost, st));
  PT_TRY(cudaMemcpyAsync(&h_cnt, p->d_counters, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
  PT_TRY(cudaMemcpyAsync(&h_swaps, d_swaps, sizeof(h_swaps), cudaMemcpyDeviceToHost, st));
  {
    cudaError_t sync_err = cudaStreamSynchronize(st);
    if (sync_err != cudaSuccess) {
      release();
      return fail(OSA_ERR_CUDA, "parallel tempering kernels failed: %s", cudaGetErrorString(sync_err));
    }
  }
  if (h_idx >= tries) {
    release();
    return fail(OSA_ERR_CUDA, "argmin returned an invalid index");
  }
  if (best_energies)
    PT_TRY(cudaMemcpyAsync(best_energies, p->d_energy, tries * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (best_states_packed)
    PT_TRY(cudaMemcpyAsync(best_states_packed, d_keep, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  std::vector<uint32_t> win(p->nw);
  PT_TRY(cudaMemcpyAsync(win.data(), d_keep + (size_t)h_idx * p->nw, (size_t)p->nw * sizeof(uint32_t),
                         cudaMemcpyDeviceToHost, st));
  PT_TRY(cudaStreamSynchronize(st));
  if (best_state)
    for (int i = 0; i < p->n; ++i) best_state[i] = (uint8_t)((win[i >> 5] >> (i & 31)) & 1u);
  if (best_energy) *best_energy = h_e;
  if (best_index) *best_index = first_try + h_idx;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->attempts = (uint64_t)prm->num_rounds * (uint64_t)prm->sweeps_per_round * (uint64_t)p->n * tries;
    stats->accepts = h_cnt.accepts;
    stats->row_fetches = h_cnt.row_fetches;
    stats->init_row_fetches = h_cnt.init_row_fetches;
    stats->pt_swaps = h_swaps;
    cudaEventElapsedTime(&ms_sweep, p->ev[0], p->ev[1]);
    cudaEventElapsedTime(&ms_energy, p->ev[1], p->ev[2]);
    stats->ms_sweep = ms_sweep;
    stats->ms_energy = ms_energy;
    cudaEventElapsedTime(&stats->ms_reduce, p->ev[2], p->ev[3]);
    cudaEventElapsedTime(&stats->ms_total, p->ev[0], p->ev[3]);
    stats->kernel_id = KID_DENSE_SEQ;
    stats->traj_per_batch = info.traj_per_batch;
    stats->q_elem_bytes = f32 ? 4 : 8;
    stats->grid = info.grid;
    stats->launches = launches;
  }
  release();
#undef PT_TRY
  return OSA_OK;
}

// ---------------------------------------------------------------------------
// population annealing (osa_pa.cu): sweep -> exact energies -> resample, per temperature step
// ---------------------------------------------------------------------------
int osa_pa_anneal(osa_problem *p, const double *betas, const osa_pa_params *prm,
                  double *best_energies, uint32_t *best_states_packed, uint8_t *best_state,
                  double *best_energy, uint64_t *best_index, osa_stats *stats) {
  if (!p || !betas || !prm) return fail(OSA_ERR_INVALID, "null argument");
  if (prm->num_populations < 1) return fail(OSA_ERR_INVALID, "num_populations must be >= 1");
  if (prm->population_size < 1 || prm->population_size > (1 << 20))
    return fail(OSA_ERR_INVALID, "population_size must be in 1..2^20");
  if (prm->num_steps < 1) return fail(OSA_ERR_INVALID, "num_steps must be >= 1");
  if (prm->sweeps_per_step < 1) return fail(OSA_ERR_INVALID, "sweeps_per_step must be >= 1");
  if (prm->flags != 0) return fail(OSA_ERR_INVALID, "unknown flags 0x%x", prm->flags);
  if (prm->accept_rule != OSA_ACCEPT_REFERENCE && prm->accept_rule != OSA_ACCEPT_BOLTZMANN)
    return fail(OSA_ERR_INVALID, "unknown accept rule %d", prm->accept_rule);
  if ((uint64_t)prm->num_steps * (uint64_t)prm->sweeps_per_step >= (1ull << 32))
    return fail(OSA_ERR_INVALID, "num_steps * sweeps_per_step must be < 2^32");
  const int M = prm->population_size;
  if (prm->num_populations > (1ull << 31) / (uint64_t)M ||
      (prm->first_population + prm->num_populations) > (1ull << 40))
    return fail(OSA_ERR_INVALID, "too many populations (replicas must stay below 2^31, ids below 2^40)");
  for (int t = 0; t < prm->num_steps; ++t)
    if (!(betas[t] > 0.0) || !std::isfinite(betas[t]))
      return fail(OSA_ERR_INVALID, "betas[%d] = %g is not a positive finite number", t, betas[t]);
  const bool f32 = p->prec == OSA_SWEEP_F32;
  if (p->sparse || !dense_seq_supported(p->n, f32 ? 4 : 8))
    return fail(OSA_ERR_UNSUPPORTED,
                "population annealing runs on dense problems with n <= %d (this sweep precision)",
                f32 ? 8192 : 4096);
  const uint64_t tries = prm->num_populations * (uint64_t)M;
  const uint64_t first_try = prm->first_population * (uint64_t)M;

  DeviceGuard guard;
  int rc = select_device(p->device);
  if (rc) return rc;
  rc = ensure_workspace(p, tries, 1);
  if (rc) return rc;

  const size_t esz = f32 ? 4 : 8;
  const size_t words = (size_t)tries * p->nw;
  uint32_t *d_cur = nullptr, *d_nxt = nullptr, *d_keep = nullptr;
  double *d_ecur = nullptr, *d_enxt = nullptr, *d_beste = nullptr;
  unsigned long long *d_cum = nullptr, *d_replaced = nullptr;
  char *d_fields = nullptr, *d_fields_nxt = nullptr;  // [tries][ld] local fields, carried and resampled
  void *d_ts_traj = nullptr;
  cudaStream_t st = p->stream;
  auto release = [&]() {
    dev_free(d_cur, st);
    dev_free(d_fields, st);
    dev_free(d_fields_nxt, st);
    dev_free(d_nxt, st);
    dev_free(d_keep, st);
    dev_free(d_ecur, st);
    dev_free(d_enxt, st);
    dev_free(d_beste, st);
    dev_free(d_cum, st);
    dev_free(d_replaced, st);
    dev_free((char *)d_ts_traj, st);
  };
#define PA_TRY(expr)                                                                        \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      release();                                                                            \
      return fail(e_ == cudaErrorMemoryAllocation ? OSA_ERR_NOMEM : OSA_ERR_CUDA,           \
                  "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), "osa_api.cu", __LINE__); \
    }                                                                                       \
  } while (0)
  PA_TRY(dev_alloc(&d_cur, words * sizeof(uint32_t), st));
  PA_TRY(dev_alloc(&d_nxt, words * sizeof(uint32_t), st));
  PA_TRY(dev_alloc(&d_fields, (size_t)tries * p->ld * esz, st));
  PA_TRY(dev_alloc(&d_fields_nxt, (size_t)tries * p->ld * esz, st));
  PA_TRY(dev_alloc(&d_keep, words * sizeof(uint32_t), st));
  PA_TRY(dev_alloc(&d_ecur, tries * sizeof(double), st));
  PA_TRY(dev_alloc(&d_enxt, tries * sizeof(double), st));
  PA_TRY(dev_alloc(&d_beste, tries * sizeof(double), st));
  PA_TRY(dev_alloc(&d_cum, tries * sizeof(unsigned long long), st));
  PA_TRY(dev_alloc(&d_replaced, sizeof(unsigned long long), st));
  PA_TRY(dev_alloc((char **)&d_ts_traj, tries * esz, st));
  {
    std::vector<double> inf(tries, std::numeric_limits<double>::infinity());
    PA_TRY(cudaMemcpyAsync(d_beste, inf.data(), tries * sizeof(double), cudaMemcpyHostToDevice, st));
    PA_TRY(cudaStreamSynchronize(st));
  }
  PA_TRY(cudaMemsetAsync(p->d_counters, 0, sizeof(Counters), st));
  PA_TRY(cudaMemsetAsync(d_replaced, 0, sizeof(unsigned long long), st));

  LaunchInfo info = {0, 0, 0, 0};
  int launches = 0;
  float ms_sweep = 0.f, ms_energy = 0.f;
  PA_TRY(cudaEventRecord(p->ev[0], st));
  PA_TRY(launch_pt_init_states(prm->seed, first_try, tries, p->n, p->nw, d_cur, st));
  ++launches;
  rc = exact_energies(p, d_cur, tries, d_ecur);
  if (rc) {
    release();
    return rc;
  }
  ++launches;
  for (int step = 0; step < prm->num_steps; ++step) {
    // every replica of every population sweeps at betas[step]
    const double ts64 = prm->accept_rule == OSA_ACCEPT_REFERENCE ? betas[step] : 1.0 / betas[step];
    if (f32) PA_TRY(launch_pa_fill<float>((float *)d_ts_traj, (float)ts64, tries, st));
    else PA_TRY(launch_pa_fill<double>((double *)d_ts_traj, ts64, tries, st));
    auto run = [&](auto tag) -> cudaError_t {
      using T = decltype(tag);
      DenseParams<T> dp{};
      dp.qoff = (const T *)p->d_qoff;
      dp.diag = (const T *)p->d_diag;
      dp.tscale = nullptr;
      dp.ld = p->ld;
      dp.n = p->n;
      dp.num_iter = 1;
      dp.sweeps_per_beta = prm->sweeps_per_step;
      dp.mode = OSA_MODE_SEQUENTIAL_SWEEP;
      dp.seed = prm->seed;
      dp.first_try = first_try;
      dp.num_tries = tries;
      dp.best_rel = p->d_best_rel;
      dp.best_states = p->d_states;
      dp.nw = p->nw;
      dp.counters = p->d_counters;
      dp.init_states = d_cur;
      dp.final_states = d_cur;
      dp.tscale_traj = (const T *)d_ts_traj;
      dp.step_base = (uint32_t)step * (uint32_t)prm->sweeps_per_step;
      // the first step builds the fields from the spins; later steps continue from the fields of
      // the replica the slot was resampled from
      dp.fields_in = step > 0 ? (const T *)d_fields : nullptr;
      dp.fields_out = (T *)d_fields;
      return launch_dense_seq_ws<T>(dp, st, &info);
    };
    PA_TRY(f32 ? run(float()) : run(double()));
    // best state of the step against the best kept so far in this slot (e_cur = start energy)
    PA_TRY(launch_pt_track_best(d_ecur, p->d_best_rel, p->d_states, tries, p->nw, d_beste, d_keep, st));
    launches += 3;
    if (step + 1 == prm->num_steps) break;
    rc = exact_energies(p, d_cur, tries, d_ecur);
    if (rc) {
      release();
      return rc;
    }
    // weights for the move to the next temperature: exp(-(b' - b) E), b = inverse temperature
    const double b0 = prm->accept_rule == OSA_ACCEPT_REFERENCE ? 1.0 / betas[step] : betas[step];
    const double b1 =
        prm->accept_rule == OSA_ACCEPT_REFERENCE ? 1.0 / betas[step + 1] : betas[step + 1];
    PA_TRY(launch_pa_resample(prm->seed, prm->first_population, prm->num_populations, M,
                              (uint32_t)step, -(b1 - b0), d_cur, d_ecur, p->nw, d_cum, d_nxt, d_enxt,
                              nullptr, d_replaced, d_fields, d_fields_nxt, p->ld * esz, st));
    std::swap(d_cur, d_nxt);
    std::swap(d_ecur, d_enxt);
    std::swap(d_fields, d_fields_nxt);
    launches += 3;
  }
  PA_TRY(cudaEventRecord(p->ev[1], st));
  rc = exact_energies(p, d_keep, tries, p->d_energy);
  if (rc) {
    release();
    return rc;
  }
  ++launches;
  PA_TRY(cudaEventRecord(p->ev[2], st));
  PA_TRY(launch_argmin(p->d_energy, tries, p->d_arg_idx, p->d_arg_e, st));
  ++launches;
  PA_TRY(cudaEventRecord(p->ev[3], st));

  unsigned long long h_idx = 0, h_replaced = 0;
  double h_e = 0.0;
  Counters h_cnt;
  PA_TRY(cudaMemcpyAsync(&h_idx, p->d_arg_idx, sizeof(h_idx), cudaMemcpyDeviceToHost, st));
  PA_TRY(cudaMemcpyAsync(&h_e, p->d_arg_e, sizeof(h_e), cudaMemcpyDeviceToHost, st));
  PA_TRY(cudaMemcpyAsync(&h_cnt, p->d_counters, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
  PA_TRY(cudaMemcpyAsync(&h_replaced, d_replaced, sizeof(h_replaced), cudaMemcpyDeviceToHost, st));
  {
    cudaError_t sync_err = cudaStreamSynchronize(st);
    if (sync_err != cudaSuccess) {
      release();
      return fail(OSA_ERR_CUDA, "population annealing kernels failed: %s", cudaGetErrorString(sync_err));
    }
  }
  if (h_idx >= tries) {
    release();
    return fail(OSA_ERR_CUDA, "argmin returned an invalid index");
  }
  if (best_energies)
    PA_TRY(cudaMemcpyAsync(best_energies, p->d_energy, tries * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (best_states_packed)
    PA_TRY(cudaMemcpyAsync(best_states_packed, d_keep, words * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  std::vector<uint32_t> win(p->nw);
  PA_TRY(cudaMemcpyAsync(win.data(), d_keep + (size_t)h_idx * p->nw, (size_t)p->nw * sizeof(uint32_t),
                         cudaMemcpyDeviceToHost, st));
  PA_TRY(cudaStreamSynchronize(st));
  if (best_state)
    for (int i = 0; i < p->n; ++i) best_state[i] = (uint8_t)((win[i >> 5] >> (i & 31)) & 1u);
  if (best_energy) *best_energy = h_e;
  if (best_index) *best_index = first_try + h_idx;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->attempts = (uint64_t)prm->num_steps * (uint64_t)prm->sweeps_per_step * (uint64_t)p->n * tries;
    stats->accepts = h_cnt.accepts;
    stats->row_fetches = h_cnt.row_fetches;
    stats->init_row_fetches = h_cnt.init_row_fetches;
    stats->pt_swaps = h_replaced;
    cudaEventElapsedTime(&ms_sweep, p->ev[0], p->ev[1]);
    cudaEventElapsedTime(&ms_energy, p->ev[1], p->ev[2]);
    stats->ms_sweep = ms_sweep;
    stats->ms_energy = ms_energy;
    cudaEventElapsedTime(&stats->ms_reduce, p->ev[2], p->ev[3]);
    cudaEventElapsedTime(&stats->ms_total, p->ev[0], p->ev[3]);
    stats->kernel_id = KID_DENSE_SEQ;
    stats->traj_per_batch = info.traj_per_batch;
    stats->q_elem_bytes = f32 ? 4 : 8;
    stats->grid = info.grid;
    stats->launches = launches;
  }
  release();
#undef PA_TRY
  return OSA_OK;
}

int osa_energy_batch(osa_problem *p, const uint32_t *states_packed, uint64_t count, double *out) {
  if (!p || !states_packed || !out) return fail(OSA_ERR_INVALID, "null argument");
  if (count == 0) return OSA_OK;
  DeviceGuard guard;
  int rc = select_device(p->device);
  if (rc) return rc;
  rc = ensure_workspace(p, count, 1);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(p->d_states, states_packed, (size_t)count * p->nw * sizeof(uint32_t),
                           cudaMemcpyHostToDevice, p->stream));
  rc = exact_energies(p, p->d_states, count, p->d_energy);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(out, p->d_energy, count * sizeof(double), cudaMemcpyDeviceToHost,
                           p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return OSA_OK;
}

int osa_host_alloc_pinned(size_t bytes, void **out) {
  if (!out || bytes == 0) return fail(OSA_ERR_INVALID, "bad argument");
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(OSA_ERR_NO_DEVICE, "no CUDA device available");
  CUDA_TRY(cudaMallocHost(out, bytes));
  return OSA_OK;
}

int osa_host_free_pinned(void *ptr) {
  if (!ptr) return OSA_OK;
  CUDA_TRY(cudaFreeHost(ptr));
  return OSA_OK;
}

int osa_exhaustive_dense_f64(const double *qsym, int n, int device, uint8_t *best_state,
                             double *best_energy) {
  if (!qsym || !best_state || !best_energy) return fail(OSA_ERR_INVALID, "null argument");
  if (n < 1 || n > 40) return fail(OSA_ERR_UNSUPPORTED, "exhaustive search supports 1 <= n <= 40");
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j)
      if (!(qsym[(size_t)i * n + j] == qsym[(size_t)j * n + i]))
        return fail(OSA_ERR_INVALID, "Q is not symmetric (expected helpers::flatten_qubo layout)");
  DeviceGuard guard;
  int rc = select_device(device);
  if (rc) return rc;
  unsigned long long x = 0;
  double e = 0.0;
  std::string msg;
  cudaError_t err = exhaustive_search(qsym, n, &x, &e, &msg);
  if (err != cudaSuccess) return fail(OSA_ERR_CUDA, "%s", msg.c_str());
  for (int i = 0; i < n; ++i) best_state[i] = (uint8_t)((x >> i) & 1ull);
  *best_energy = e;
  return OSA_OK;
}

int osa_measure_read_bandwidth(int device, size_t bytes, int iters, double *gbs) {
  if (!gbs || bytes < 16 || iters < 1) return fail(OSA_ERR_INVALID, "bad argument");
  DeviceGuard guard;
  int rc = select_device(device);
  if (rc) return rc;
  uint4 *buf = nullptr;
  unsigned int *sink = nullptr;
  cudaEvent_t a = nullptr, b = nullptr;
  const size_t n_vec = bytes / 16;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaError_t e = cudaMalloc(&buf, n_vec * 16);
  if (e == cudaSuccess) e = cudaMalloc(&sink, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(buf, 1, n_vec * 16);
  if (e == cudaSuccess) e = cudaEventCreate(&a);
  if (e == cudaSuccess) e = cudaEventCreate(&b);
  float ms = 0.f;
  if (e == cudaSuccess) e = launch_read_bw(buf, n_vec, 2, sink, sms * 8, 0);  // warm L2
  if (e == cudaSuccess) e = cudaEventRecord(a, 0);
  if (e == cudaSuccess) e = launch_read_bw(buf, n_vec, iters, sink, sms * 8, 0);
  if (e == cudaSuccess) e = cudaEventRecord(b, 0);
  if (e == cudaSuccess) e = cudaEventSynchronize(b);
  if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, a, b);
  if (a) cudaEventDestroy(a);
  if (b) cudaEventDestroy(b);
  cudaFree(buf);
  cudaFree(sink);
  if (e != cudaSuccess) return fail(OSA_ERR_CUDA, "bandwidth probe failed: %s", cudaGetErrorString(e));
  *gbs = (double)n_vec * 16.0 * iters / (ms * 1e-3) / 1e9;
  return OSA_OK;
}

}  // extern "C"
