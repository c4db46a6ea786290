// app/cli_common.hpp -- the parts the three command-line programs share: required --input/--output,
// device-type validation, model loading, device opening and the outermost error handler.  The
// texts and exit codes are the reference CLIs' (/root/reference/app/one-solver-anneal.cpp:78-115,
// 129-151,171-176; one-solver-exhaustive.cpp:40-60,86-104): 0 success/help, -1 bad arguments or
// unreadable input, 1 exception.
#ifndef ONESOLVER_B200_APP_CLI_COMMON_HPP_
#define ONESOLVER_B200_APP_CLI_COMMON_HPP_

#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "cli_options.hpp"
#include "helpers/devices.hpp"
#include "model/qubo.hpp"

namespace cli {

constexpr int kExitOk = 0, kExitUsage = -1, kExitError = 1;

// thrown by the helpers below after they have printed their message
struct Exit {
  int code;
};

inline void add_io_options(Options &options) {
  options.add("help", false, "", "produce help message")
      .add("input", true, "", "input file")
      .add("output", true, "", "output file");
}

inline void usage_error(const std::string &message) {
  std::cerr << message << std::endl;
  throw Exit{kExitUsage};
}

// --help, then the two mandatory paths
inline void require_io(const Options &options, std::string &input, std::string &output) {
  if (options.count("help")) {
    std::cout << options.help() << std::endl;
    throw Exit{kExitOk};
  }
  if (!options.count("input")) usage_error("No input file provided.");
  if (!options.count("output")) usage_error("No output file provided.");
  input = options.str("input");
  output = options.str("output");
}

inline std::string checked_device_type(const Options &options) {
  const std::string type = options.str("device-type");
  if (type != "cpu" && type != "gpu" && type != "host") usage_error("Unknown device type: " + type);
  return type;
}

// one of a fixed set of words, e.g. --schedule-type; `what` completes "Unknown <what>: <value>"
inline std::string checked_choice(const Options &options, const std::string &flag,
                                  std::initializer_list<const char *> allowed,
                                  const std::string &what) {
  const std::string value = options.str(flag);
  for (const char *a : allowed)
    if (value == a) return value;
  usage_error("Unknown " + what + ": " + value);
  return value;
}

inline qubo::QUBOModel<int, double> read_model(const std::string &path) {
  std::ifstream file(path);
  if (!file) usage_error("can not open input file: " + path);
  return qubo::QUBOModel<int, double>::load(file);
}

// the reference prints the first line and then dereferences a null queue; we stop instead.
// gpu_devices: the CUDA devices of a "gpu" queue (empty: every visible one)
inline std::unique_ptr<devices::queue> open_device(const std::string &type,
                                                   const std::vector<int> &gpu_devices) {
  try {
    std::unique_ptr<devices::queue> q(
        new devices::queue(*devices::construct_device_selector(type), gpu_devices));
    std::cout << "Using device: " << q->device_name() << std::endl;
    return q;
  } catch (const std::runtime_error &e) {
    std::cerr << "No devices of given type could be initialized." << std::endl;
    std::cerr << "error: " << e.what() << "\n";
    throw Exit{kExitError};
  }
}
inline std::unique_ptr<devices::queue> open_device(const std::string &type, int gpu_index = 0) {
  return open_device(type, std::vector<int>{gpu_index});
}

// main() body wrapper: maps Exit and exceptions to the reference's exit codes
inline int run(const std::function<void()> &body) {
  try {
    body();
  } catch (const Exit &e) {
    return e.code;
  } catch (std::exception &e) {
    std::cerr << "error: " << e.what() << "\n";
    return kExitError;
  } catch (...) {
    std::cerr << "Exception of unknown type!\n";
  }
  return kExitOk;
}

}  // namespace cli

#endif
