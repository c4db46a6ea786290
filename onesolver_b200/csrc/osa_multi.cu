// osa_multi.cu -- osa_multi_*: the annealing call sharded over the GPUs of one box.
//
// One process, one host thread and one stream per GPU.  Trajectories are independent (reference
// annealing.hpp:85-86), so device k of G runs the global trajectory ids of its shard with its own
// replica of Q and no data-path collective; the only exchange is ONE ncclAllGather of
// {best energy, global trajectory id, packed best state} per device (16 + 4*ceil(N/32) bytes) at
// the end of the call, after which the winner is min energy, then min id -- std::min_element
// (annealing.hpp:134) applied across the shards.  The random streams are keyed by global ids, so
// the result does not depend on the number of devices.
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library is loaded on first use, see nccl_api()

#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "osa_internal.h"

using namespace osa;

// The communicators of a device list are created once per process and shared by every osa_multi
// with that list (ncclCommInitAll costs hundreds of milliseconds; sa::anneal creates and destroys
// a problem per call).  Calls on handles that share communicators are serialised by `busy`.
struct CommSet {
  std::vector<ncclComm_t> comms;
  std::mutex busy;
};

struct osa_multi {
  std::vector<int> devices;
  std::vector<osa_problem *> problems;
  std::shared_ptr<CommSet> comm_set;
  std::vector<unsigned char *> d_rec;  // per device: this device's record
  std::vector<unsigned char *> d_all;  // per device: the gathered records of all devices
  int n = 0, nw = 0;
  size_t rec_bytes = 0;
};

namespace {

// NCCL is bound at run time (dlopen) instead of at link time: a process that also loads PyTorch
// must end up with ONE libnccl.so.2, and with a link-time dependency whichever of the two libraries
// is loaded first decides which copy that is (the system's older one breaks `import torch`).
// dlopen by soname returns the copy the process already has, if any.
struct NcclApi {
  decltype(&ncclCommInitAll) comm_init_all = nullptr;
  decltype(&ncclCommDestroy) comm_destroy = nullptr;
  decltype(&ncclAllGather) all_gather = nullptr;
  decltype(&ncclGetErrorString) error_string = nullptr;
  decltype(&ncclGetVersion) get_version = nullptr;
  std::string error;
};

const NcclApi &nccl_api() {
  static const NcclApi api = [] {
    NcclApi a;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) {
      a.error = dlerror();
      return a;
    }
    a.comm_init_all = reinterpret_cast<decltype(a.comm_init_all)>(dlsym(h, "ncclCommInitAll"));
    a.comm_destroy = reinterpret_cast<decltype(a.comm_destroy)>(dlsym(h, "ncclCommDestroy"));
    a.all_gather = reinterpret_cast<decltype(a.all_gather)>(dlsym(h, "ncclAllGather"));
    a.error_string = reinterpret_cast<decltype(a.error_string)>(dlsym(h, "ncclGetErrorString"));
    a.get_version = reinterpret_cast<decltype(a.get_version)>(dlsym(h, "ncclGetVersion"));
    if (!a.comm_init_all || !a.comm_destroy || !a.all_gather || !a.error_string)
      a.error = "libnccl.so.2 lacks ncclCommInitAll / ncclAllGather";
    return a;
  }();
  return api;
}

__global__ void k_pack_best(const double *arg_e, const unsigned long long *arg_idx,
                            const uint32_t *states, int nw, unsigned long long first_try, int valid,
                            unsigned char *rec) {
  double *e = reinterpret_cast<double *>(rec);
  unsigned long long *id = reinterpret_cast<unsigned long long *>(rec + 8);
  uint32_t *st = reinterpret_cast<uint32_t *>(rec + 16);
  if (!valid) {  // a device without trajectories never wins
    if (threadIdx.x == 0) {
      *e = INFINITY;
      *id = ~0ull;
    }
    for (int k = threadIdx.x; k < nw; k += blockDim.x) st[k] = 0u;
    return;
  }
  const unsigned long long idx = *arg_idx;
  if (threadIdx.x == 0) {
    *e = *arg_e;
    *id = first_try + idx;
  }
  for (int k = threadIdx.x; k < nw; k += blockDim.x) st[k] = states[(size_t)idx * nw + k];
}

// contiguous id range of device k of g: the remainder goes to the low devices
// (the rule of onesolver_b200/multi.py::shard)
void shard(uint64_t num_tries, int g, int k, uint64_t *first, uint64_t *count) {
  const uint64_t base = num_tries / (uint64_t)g, rem = num_tries % (uint64_t)g;
  *count = base + ((uint64_t)k < rem ? 1 : 0);
  *first = (uint64_t)k * base + ((uint64_t)k < rem ? (uint64_t)k : rem);
}

int resolve_devices(const int *devices, int num_devices, std::vector<int> *out) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return osa_fail(OSA_ERR_NO_DEVICE, "no CUDA device available: %s",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  out->clear();
  if (!devices || num_devices <= 0) {
    const int g = (num_devices > 0 && num_devices < count) ? num_devices : count;
    for (int k = 0; k < g; ++k) out->push_back(k);
    return OSA_OK;
  }
  for (int k = 0; k < num_devices; ++k) {
    if (devices[k] < 0 || devices[k] >= count)
      return osa_fail(OSA_ERR_INVALID, "device %d out of range [0, %d)", devices[k], count);
    for (int j = 0; j < k; ++j)
      if (devices[j] == devices[k]) return osa_fail(OSA_ERR_INVALID, "device %d listed twice", devices[k]);
    out->push_back(devices[k]);
  }
  return OSA_OK;
}

// run fn(k) for every device on its own host thread; the first failure (lowest k) is reported
template <typename F>
int for_each_device(int g, F fn) {
  std::vector<int> rc(g, OSA_OK);
  std::vector<std::string> msg(g);
  std::vector<std::thread> th;
  th.reserve(g);
  for (int k = 0; k < g; ++k)
    th.emplace_back([&, k] {
      rc[k] = fn(k);
      if (rc[k] != OSA_OK) msg[k] = osa_last_error();
    });
  for (auto &t : th) t.join();
  for (int k = 0; k < g; ++k)
    if (rc[k] != OSA_OK) return osa_fail(rc[k], "device slot %d: %s", k, msg[k].c_str());
  return OSA_OK;
}

template <typename Create>
int multi_create(const int *devices, int num_devices, int n, Create create, osa_multi **out) {
  if (!out) return osa_fail(OSA_ERR_INVALID, "null argument");
  *out = nullptr;
  osa_multi *m = new (std::nothrow) osa_multi();
  if (!m) return osa_fail(OSA_ERR_NOMEM, "out of host memory");
  int rc = resolve_devices(devices, num_devices, &m->devices);
  if (rc) {
    delete m;
    return rc;
  }
  const int g = (int)m->devices.size();
  m->n = n;
  m->nw = (n + 31) / 32;
  m->rec_bytes = 16 + (size_t)((m->nw + 1) / 2 * 2) * 4;  // multiple of 8 bytes
  m->problems.assign(g, nullptr);
  m->d_rec.assign(g, nullptr);
  m->d_all.assign(g, nullptr);
  DeviceGuard guard;
  // Q replicated: every device gets its own upload from the caller's buffer, in parallel
  rc = for_each_device(g, [&](int k) { return create(m->devices[k], &m->problems[k]); });
  if (rc == OSA_OK && g > 1) {
    const NcclApi &nccl = nccl_api();
    if (!nccl.error.empty()) {
      rc = osa_fail(OSA_ERR_UNSUPPORTED, "NCCL is not available: %s", nccl.error.c_str());
    } else {
      static std::mutex cache_mutex;
      static std::map<std::vector<int>, std::shared_ptr<CommSet>> cache;  // lives as long as the process
      std::lock_guard<std::mutex> lock(cache_mutex);
      auto it = cache.find(m->devices);
      if (it != cache.end()) {
        m->comm_set = it->second;
      } else {
        auto set = std::make_shared<CommSet>();
        set->comms.assign(g, nullptr);
        ncclResult_t nr = nccl.comm_init_all(set->comms.data(), g, m->devices.data());
        if (nr != ncclSuccess)
          rc = osa_fail(OSA_ERR_CUDA, "ncclCommInitAll over %d devices failed: %s", g, nccl.error_string(nr));
        else
          cache[m->devices] = m->comm_set = set;
      }
    }
  }
  for (int k = 0; k < g && rc == OSA_OK; ++k) {
    cudaError_t e = cudaSetDevice(m->devices[k]);
    cudaStream_t st = m->problems[k]->stream;
    if (e == cudaSuccess) e = osa_pool_alloc(reinterpret_cast<void **>(&m->d_rec[k]), m->rec_bytes, st);
    if (e == cudaSuccess) e = osa_pool_alloc(reinterpret_cast<void **>(&m->d_all[k]), m->rec_bytes * g, st);
    if (e != cudaSuccess) rc = osa_fail(OSA_ERR_CUDA, "gather buffers on device %d: %s", m->devices[k], cudaGetErrorString(e));
  }
  if (rc) {
    const std::string keep = osa_last_error();
    osa_multi_destroy(m);
    return osa_fail(rc, "%s", keep.c_str());
  }
  *out = m;
  return OSA_OK;
}

}  // namespace

cudaError_t osa_pack_best(osa_problem *p, uint64_t first_try, unsigned char *d_rec) {
  k_pack_best<<<1, 128, 0, p->stream>>>(p->d_arg_e, p->d_arg_idx, p->d_states, p->nw, first_try, 1, d_rec);
  return cudaGetLastError();
}

extern "C" {

int osa_multi_create_dense_f64(const double *qsym, int n, const int *devices, int num_devices,
                               int sweep_precision, osa_multi **out) {
  return multi_create(devices, num_devices, n, [&](int dev, osa_problem **p) {
    return osa_problem_create_dense_f64(qsym, n, dev, sweep_precision, p);
  }, out);
}

int osa_multi_create_dense_f32(const float *qsym, int n, const int *devices, int num_devices,
                               osa_multi **out) {
  return multi_create(devices, num_devices, n, [&](int dev, osa_problem **p) {
    return osa_problem_create_dense_f32(qsym, n, dev, p);
  }, out);
}

int osa_multi_create_csr_f64(const int32_t *rowptr, const int32_t *col, const double *val,
                             const double *diag, int n, const int *devices, int num_devices,
                             int sweep_precision, osa_multi **out) {
  return multi_create(devices, num_devices, n, [&](int dev, osa_problem **p) {
    return osa_problem_create_csr_f64(rowptr, col, val, diag, n, dev, sweep_precision, p);
  }, out);
}

int osa_multi_destroy(osa_multi *m) {
  if (!m) return OSA_OK;
  DeviceGuard guard;
  for (size_t k = 0; k < m->devices.size(); ++k) {
    cudaSetDevice(m->devices[k]);
    cudaStream_t st = (k < m->problems.size() && m->problems[k]) ? m->problems[k]->stream : nullptr;
    if (k < m->d_rec.size() && m->d_rec[k]) osa_pool_free(m->d_rec[k], st);
    if (k < m->d_all.size() && m->d_all[k]) osa_pool_free(m->d_all[k], st);
  }
  for (auto p : m->problems) osa_problem_destroy(p);
  delete m;
  return OSA_OK;
}

int osa_multi_devices(const osa_multi *m, int *num_devices, int *devices, int capacity) {
  if (!m || !num_devices) return osa_fail(OSA_ERR_INVALID, "null argument");
  *num_devices = (int)m->devices.size();
  if (devices)
    for (int k = 0; k < *num_devices && k < capacity; ++k) devices[k] = m->devices[k];
  return OSA_OK;
}

int osa_multi_problem(osa_multi *m, int slot, osa_problem **out) {
  if (!m || !out || slot < 0 || slot >= (int)m->problems.size())
    return osa_fail(OSA_ERR_INVALID, "bad device slot");
  *out = m->problems[slot];
  return OSA_OK;
}

int osa_multi_anneal(osa_multi *m, const double *beta_schedule, const osa_anneal_params *prm,
                     double *best_energies, uint32_t *best_states_packed, uint8_t *best_state,
                     double *best_energy, uint64_t *best_index, osa_stats *stats,
                     osa_stats *device_stats) {
  if (!m || !prm) return osa_fail(OSA_ERR_INVALID, "null argument");
  if (prm->num_tries < 1) return osa_fail(OSA_ERR_INVALID, "num_tries must be >= 1");
  const int g = (int)m->devices.size();
  std::vector<osa_stats> st(g);
  std::vector<std::vector<unsigned char>> host_all(g);
  memset(st.data(), 0, sizeof(osa_stats) * g);
  DeviceGuard guard;
  std::unique_lock<std::mutex> comm_lock;
  if (m->comm_set) comm_lock = std::unique_lock<std::mutex>(m->comm_set->busy);
  int rc = for_each_device(g, [&](int k) -> int {
    osa_problem *p = m->problems[k];
    uint64_t first = 0, count = 0;
    shard(prm->num_tries, g, k, &first, &count);
    cudaError_t e = cudaSetDevice(m->devices[k]);
    if (e != cudaSuccess) return osa_fail(OSA_ERR_CUDA, "cudaSetDevice(%d): %s", m->devices[k], cudaGetErrorString(e));
    if (count > 0) {
      osa_anneal_params local = *prm;
      local.first_try = prm->first_try + first;
      local.num_tries = count;
      int r = osa_anneal(p, beta_schedule, &local, best_energies ? best_energies + first : nullptr,
                         best_states_packed ? best_states_packed + (size_t)first * m->nw : nullptr,
                         nullptr, nullptr, nullptr, &st[k]);
      if (r != OSA_OK) return r;
      e = osa_pack_best(p, local.first_try, m->d_rec[k]);
    } else {
      k_pack_best<<<1, 128, 0, p->stream>>>(nullptr, nullptr, nullptr, m->nw, 0ull, 0, m->d_rec[k]);
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) return osa_fail(OSA_ERR_CUDA, "packing the best record: %s", cudaGetErrorString(e));
    // the one collective of the call: every device's {energy, id, state}
    if (g > 1) {
      const NcclApi &nccl = nccl_api();
      ncclResult_t nr = nccl.all_gather(m->d_rec[k], m->d_all[k], m->rec_bytes, ncclUint8, m->comm_set->comms[k], p->stream);
      if (nr != ncclSuccess) return osa_fail(OSA_ERR_CUDA, "ncclAllGather: %s", nccl.error_string(nr));
    } else {
      e = cudaMemcpyAsync(m->d_all[k], m->d_rec[k], m->rec_bytes, cudaMemcpyDeviceToDevice, p->stream);
      if (e != cudaSuccess) return osa_fail(OSA_ERR_CUDA, "record copy: %s", cudaGetErrorString(e));
    }
    host_all[k].resize(m->rec_bytes * g);
    e = cudaMemcpyAsync(host_all[k].data(), m->d_all[k], m->rec_bytes * g, cudaMemcpyDeviceToHost, p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    if (e != cudaSuccess) return osa_fail(OSA_ERR_CUDA, "gather on device %d: %s", m->devices[k], cudaGetErrorString(e));
    return OSA_OK;
  });
  if (rc) return rc;

  // min energy, then min id, over the gathered records (every device holds the same table)
  const unsigned char *tab = host_all[0].data();
  int win = -1;
  double we = 0.0;
  uint64_t wi = 0;
  for (int k = 0; k < g; ++k) {
    double e;
    uint64_t id;
    memcpy(&e, tab + (size_t)k * m->rec_bytes, 8);
    memcpy(&id, tab + (size_t)k * m->rec_bytes + 8, 8);
    if (id == ~0ull) continue;
    if (win < 0 || e < we || (e == we && id < wi)) {
      win = k;
      we = e;
      wi = id;
    }
  }
  if (win < 0) return osa_fail(OSA_ERR_CUDA, "no device produced a result");
  for (int k = 1; k < g; ++k)
    if (memcmp(host_all[k].data(), tab, m->rec_bytes * g) != 0)
      return osa_fail(OSA_ERR_CUDA, "devices disagree on the gathered records");
  if (best_state) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(tab + (size_t)win * m->rec_bytes + 16);
    for (int i = 0; i < m->n; ++i) best_state[i] = (uint8_t)((w[i >> 5] >> (i & 31)) & 1u);
  }
  if (best_energy) *best_energy = we;
  if (best_index) *best_index = wi;
  if (device_stats) memcpy(device_stats, st.data(), sizeof(osa_stats) * g);
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    for (int k = 0; k < g; ++k) {
      const osa_stats &s = st[k];
      stats->attempts += s.attempts;
      stats->accepts += s.accepts;
      stats->row_fetches += s.row_fetches;
      stats->init_row_fetches += s.init_row_fetches;
      stats->cyc_decide += s.cyc_decide;
      stats->cyc_apply += s.cyc_apply;
      stats->cyc_stage += s.cyc_stage;
      stats->cyc_init += s.cyc_init;
      stats->grid += s.grid;
      stats->launches += s.launches + 1;  // + the record kernel
      // times: the slowest device (they run concurrently)
      stats->ms_total = std::fmax(stats->ms_total, s.ms_total);
      stats->ms_sweep = std::fmax(stats->ms_sweep, s.ms_sweep);
      stats->ms_energy = std::fmax(stats->ms_energy, s.ms_energy);
      stats->ms_reduce = std::fmax(stats->ms_reduce, s.ms_reduce);
      if (s.kernel_id) {
        stats->kernel_id = s.kernel_id;
        stats->traj_per_batch = s.traj_per_batch;
        stats->q_elem_bytes = s.q_elem_bytes;
      }
    }
    stats->reserved = g;  // devices used
  }
  return OSA_OK;
}

}  // extern "C"
