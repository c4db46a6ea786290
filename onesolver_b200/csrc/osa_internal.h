// osa_internal.h -- pieces of the C ABI implementation shared by osa_api.cu and osa_multi.cu:
// the problem handle, the per-thread error message and the device guard.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "osa_common.cuh"

struct osa_problem {
  int device = 0;
  int n = 0;
  int nw = 0;
  bool sparse = false;
  int prec = OSA_SWEEP_F64;
  // dense
  size_t ld = 0;          // leading dimension of qoff (sweep precision)
  size_t rows_pad = 0;    // rows allocated (multiple of 32)
  void *d_qoff = nullptr; // zero-diagonal symmetric copy, sweep precision
  void *d_diag = nullptr; // [ld] sweep precision
  size_t ld64 = 0;
  double *d_q64 = nullptr; // [rows_pad][ld64] original values incl. diagonal (exact energies)
  // csr
  int64_t nnz = 0;
  int32_t *d_rowptr = nullptr, *d_col = nullptr;
  void *d_val = nullptr;      // sweep precision
  double *d_val64 = nullptr;
  double *d_diag64 = nullptr;
  // grouped layout of the sequential sparse sweep (osa_sparse.cu): groups of four sites
  uint32_t *d_gbase = nullptr, *d_ginfo = nullptr;
  unsigned char *d_gent = nullptr;
  int stage_ok = 0, group = 4;
  // execution
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  // workspace (grow-only)
  size_t cap_tries = 0;
  double *d_best_rel = nullptr;
  double *d_energy = nullptr;
  uint32_t *d_states = nullptr;
  size_t cap_states_words = 0;
  uint32_t *d_xbest_ws = nullptr;
  size_t cap_ws_words = 0;
  char *d_fields = nullptr;  // [tries][ld] initial local fields of the warp-per-trajectory kernel
  size_t cap_fields_bytes = 0;
  unsigned long long *d_trace = nullptr;  // [cap_trace] flip-trace hashes (osa_anneal_traced)
  size_t cap_trace = 0;
  void *d_tscale = nullptr;
  size_t cap_tscale_bytes = 0;
  osa::Counters *d_counters = nullptr;
  unsigned long long *d_arg_idx = nullptr;
  double *d_arg_e = nullptr;
};

// Sets the calling thread's osa_last_error() message and returns `code`.
int osa_fail(int code, const char *fmt, ...);

// Pack {energy, GLOBAL trajectory id, packed state} of the trajectory that osa_anneal's argmin
// picked (left in p->d_arg_e / d_arg_idx / d_states by the call) into rec[16 + 4 * nw] on the
// problem's device and stream -- the record that osa_multi_anneal gathers over NCCL.
cudaError_t osa_pack_best(osa_problem *p, uint64_t first_try, unsigned char *d_rec);

// stream-ordered allocations from the library's private pool of the current device (osa_api.cu)
cudaError_t osa_pool_alloc(void **ptr, size_t bytes, cudaStream_t stream);
void osa_pool_free(void *ptr, cudaStream_t stream);

namespace osa {
// Every API call runs on the device of its problem and leaves the caller's current device as it
// found it (a host application with several GPUs keeps its own cudaSetDevice state).
class DeviceGuard {
 public:
  DeviceGuard() { ok_ = cudaGetDevice(&prev_) == cudaSuccess; }
  ~DeviceGuard() {
    if (ok_) cudaSetDevice(prev_);
  }
  DeviceGuard(const DeviceGuard &) = delete;
  DeviceGuard &operator=(const DeviceGuard &) = delete;

 private:
  int prev_ = 0;
  bool ok_ = false;
};
}  // namespace osa
