// helpers/devices.hpp -- device selection for the solvers.
//
// Mirrors the reference's devices::construct_device_selector
// (/root/reference/include/helpers/devices.hpp:27-42): "cpu" / "gpu" / "host" select a
// device class, anything else throws std::invalid_argument("Unknown device type: X").
// The reference returns SYCL selectors and the caller builds a sycl::queue from them
// (app/one-solver-anneal.cpp:141-143); here the selectors are plain classes and
// devices::queue plays the role of sycl::queue:
//   "gpu"          -> CUDA device(s) through the C ABI in onesolver_b200.h (hand-written
//                     sm_100a kernels): ALL visible GPUs of the box unless the caller names
//                     one device or a list (sa::anneal shards the trajectories over them,
//                     osa_multi_anneal).  No silent fallback: if no CUDA device can be opened
//                     the queue constructor throws std::runtime_error.
//   "cpu" / "host" -> the host engine in simulated_annealing/host_engine.hpp (same
//                     algorithm and random streams as the CUDA kernels, single thread per
//                     trajectory batch) -- the counterpart of SYCL's host/CPU devices.
#ifndef ONESOLVER_B200_HELPERS_DEVICES_HPP_
#define ONESOLVER_B200_HELPERS_DEVICES_HPP_

#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "onesolver_b200.h"

namespace devices {

enum class device_kind { host, cpu, gpu };

class device_selector {
public:
  virtual ~device_selector() = default;
  virtual device_kind kind() const = 0;
};
class host_selector : public device_selector {
public:
  device_kind kind() const override { return device_kind::host; }
};
class cpu_selector : public device_selector {
public:
  device_kind kind() const override { return device_kind::cpu; }
};
class gpu_selector : public device_selector {
public:
  device_kind kind() const override { return device_kind::gpu; }
};

using device_selector_ptr = std::unique_ptr<device_selector>;

inline device_selector_ptr construct_device_selector(std::string device_type) {
  if (device_type == "cpu") return device_selector_ptr(new cpu_selector());
  if (device_type == "gpu") return device_selector_ptr(new gpu_selector());
  if (device_type == "host") return device_selector_ptr(new host_selector());
  std::ostringstream error_stream;
  error_stream << "Unknown device type: " << device_type;
  throw std::invalid_argument(error_stream.str());
}

// Stand-in for sycl::queue: where sa::anneal / exhaustive::solve run.
class queue {
public:
  // every visible CUDA device for "gpu"
  explicit queue(const device_selector &selector) : kind_(selector.kind()) { open({}); }
  // one CUDA device
  queue(const device_selector &selector, int cuda_device) : kind_(selector.kind()) {
    open({cuda_device});
  }
  // a list of CUDA devices (empty: all visible)
  queue(const device_selector &selector, const std::vector<int> &cuda_devices)
      : kind_(selector.kind()) {
    open(cuda_devices);
  }
  bool is_gpu() const { return kind_ == device_kind::gpu; }
  device_kind kind() const { return kind_; }
  int cuda_device() const { return cuda_devices_.empty() ? 0 : cuda_devices_[0]; }
  const std::vector<int> &cuda_devices() const { return cuda_devices_; }
  const std::string &device_name() const { return name_; }
  unsigned max_compute_units() const {
    const unsigned hw = std::thread::hardware_concurrency();
    return kind_ == device_kind::host ? 1u : (hw ? hw : 1u);
  }

private:
  device_kind kind_;
  std::vector<int> cuda_devices_;
  std::string name_;

  void open(std::vector<int> wanted) {
    if (kind_ != device_kind::gpu) {
      name_ = kind_ == device_kind::cpu ? "Host CPU (onesolver_b200 host engine, all cores)"
                                        : "Host (onesolver_b200 host engine)";
      return;
    }
    int count = 0;
    if (osa_device_count(&count) != OSA_OK || count <= 0) {
      throw std::runtime_error(std::string("No CUDA device could be initialized: ") +
                               osa_last_error());
    }
    if (wanted.empty())
      for (int d = 0; d < count; ++d) wanted.push_back(d);
    for (int d : wanted)
      if (d < 0 || d >= count) throw std::runtime_error("CUDA device index out of range");
    char name[256];
    if (osa_device_name(wanted[0], name, sizeof(name)) != OSA_OK) {
      throw std::runtime_error(std::string("Cannot query CUDA device: ") + osa_last_error());
    }
    name_ = name;
    if (wanted.size() > 1) name_ += " x" + std::to_string(wanted.size());
    cuda_devices_ = wanted;
  }
};

}  // namespace devices

#endif
