// one-solver-exhaustive -- brute-force ground state of a small `.qubo` instance.
// Same flags, banner, messages and exit codes as the reference CLI
// (/root/reference/app/one-solver-exhaustive.cpp:23-105): --input --output --device-type.
#include <fstream>
#include <iostream>
#include <memory>

#include "cli_options.hpp"
#include "exhaustive/exhaustive.hpp"
#include "helpers/devices.hpp"
#include "model/qubo.hpp"
#include "model/solution.hpp"

int main(int argc, char *argv[]) {
  try {
    cli::Options options("Allowed options");
    options.add("help", false, "", "produce help message")
        .add("input", true, "", "input file")
        .add("output", true, "", "output file")
        .add("device-type", true, "host", "device type to use (cpu, gpu or host)");
    options.parse(argc, argv);

    if (options.count("help")) {
      std::cout << options.help() << std::endl;
      return 0;
    }
    if (!options.count("input")) {
      std::cerr << "No input file provided." << std::endl;
      return -1;
    }
    if (!options.count("output")) {
      std::cerr << "No output file provided." << std::endl;
      return -1;
    }
    const std::string input_file = options.str("input"), output_file = options.str("output");
    const std::string device_type = options.str("device-type");
    if (device_type != "cpu" && device_type != "gpu" && device_type != "host") {
      std::cerr << "Unknown device type: " << device_type << std::endl;
      return -1;
    }
    std::cout << "Reading input from: " << input_file << std::endl;
    std::cout << "Output will be saved to: " << output_file << std::endl;

    std::ifstream qubo_file(input_file);
    if (!qubo_file) {
      std::cerr << "can not open input file: " << input_file << std::endl;
      return -1;
    }
    auto instance = qubo::QUBOModel<int, double>::load(qubo_file);

    std::unique_ptr<devices::queue> q_ptr;
    try {
      q_ptr.reset(new devices::queue(*devices::construct_device_selector(device_type)));
    } catch (const std::runtime_error &e) {
      std::cerr << "No devices of given type could be initialized." << std::endl;
      std::cerr << "error: " << e.what() << "\n";
      return 1;
    }
    devices::queue &q = *q_ptr;
    std::cout << "Using device: " << q.device_name() << std::endl;

    auto solution = exhaustive::solve(q, instance);

    std::ofstream results_file(output_file);
    solution.save(results_file);
    results_file.close();
  } catch (std::exception &e) {
    std::cerr << "error: " << e.what() << "\n";
    return 1;
  } catch (...) {
    std::cerr << "Exception of unknown type!\n";
  }
  return 0;
}
