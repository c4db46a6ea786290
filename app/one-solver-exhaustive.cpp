// one-solver-exhaustive -- brute-force ground state of a small `.qubo` instance, on the host
// threads or (`--device-type gpu`) with the CUDA Gray-code search.  Flags, banner, messages and
// exit codes are those of the reference program (/root/reference/app/one-solver-exhaustive.cpp);
// the shared plumbing lives in cli_common.hpp.
#include "cli_common.hpp"
#include "exhaustive/exhaustive.hpp"
#include "model/solution.hpp"

int main(int argc, char *argv[]) {
  return cli::run([&] {
    cli::Options options("Allowed options");
    cli::add_io_options(options);
    options.add("device-type", true, "host", "device type to use (cpu, gpu or host)");
    options.parse(argc, argv);

    std::string input_file, output_file;
    cli::require_io(options, input_file, output_file);
    const std::string device_type = cli::checked_device_type(options);
    std::cout << "Reading input from: " << input_file << std::endl;
    std::cout << "Output will be saved to: " << output_file << std::endl;

    auto instance = cli::read_model(input_file);  // exhaustive::solve takes the model by reference
    const auto device = cli::open_device(device_type);
    const auto ground_state = exhaustive::solve(*device, instance);

    std::ofstream results(output_file);
    ground_state.save(results);
  });
}
