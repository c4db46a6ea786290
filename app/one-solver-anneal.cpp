// one-solver-anneal -- simulated annealing of a `.qubo` instance, `--device-type gpu` running
// on the B200 CUDA engine.
//
// Flag surface, defaults, banner lines, messages and exit codes follow the reference CLI
// (/root/reference/app/one-solver-anneal.cpp:41-179): --input --output --num-iter (100)
// --num-tries (100) --schedule-type (geometric) --beta-min (0.1) --beta-max (1.0)
// --device-type (host); exit 0 on success/help, -1 on bad arguments or an unreadable input,
// 1 on an exception.  Extra flags (all optional, defaults keep the reference behaviour):
// --mode random|sweep, --accept reference|boltzmann, --sweeps-per-beta, --seed,
// --precision f64|f32, --layout auto|dense|csr, --gpu-index, --stats; --algorithm sa|pt with
// --num-replicas: parallel tempering (GPU only) -- the beta range becomes a ladder of
// --num-replicas temperatures (built by the chosen schedule type), --num-iter counts rounds,
// --sweeps-per-beta the sweeps per round and --num-tries the independent ladders.
#include <fstream>
#include <iostream>
#include <memory>
#include <vector>

#include "cli_options.hpp"
#include "helpers/devices.hpp"
#include "model/qubo.hpp"
#include "model/solution.hpp"
#include "schedules.hpp"
#include "simulated_annealing/annealing.hpp"

int main(int argc, char *argv[]) {
  try {
    cli::Options options("Allowed options");
    options.add("help", false, "", "produce help message")
        .add("input", true, "", "input file")
        .add("output", true, "", "output file")
        .add("num-iter", true, "100", "number of iterations of the algorithm")
        .add("num-tries", true, "100", "number of trajectories to try")
        .add("schedule-type", true, "geometric", "type of beta schedule tu use, either linear or geometric")
        .add("beta-min", true, "0.1", "minimum value of beta in the annealing schedule (default 0.1)")
        .add("beta-max", true, "1", "maximum value of beta in the annealing schedule (default 1.0)")
        .add("device-type", true, "host", "device type to use (cpu, gpu or host)")
        .add("mode", true, "random", "site visiting order: random (reference) or sweep (sequential sweeps)")
        .add("accept", true, "reference", "acceptance rule: reference (exp(-dE/beta)) or boltzmann (exp(-beta*dE))")
        .add("sweeps-per-beta", true, "1", "attempts (random) or sweeps (sweep) per schedule step")
        .add("seed", true, "1234", "random seed")
        .add("precision", true, "f64", "sweep arithmetic on the gpu: f64 or f32")
        .add("layout", true, "auto", "gpu problem layout: auto, dense or csr")
        .add("gpu-index", true, "0", "CUDA device to use with --device-type gpu")
        .add("stats", false, "", "print engine statistics (gpu)")
        .add("algorithm", true, "sa", "sa (simulated annealing) or pt (parallel tempering, gpu)")
        .add("num-replicas", true, "12", "temperatures per ladder with --algorithm pt");
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
    const std::string schedule_type = options.str("schedule-type");
    const std::string device_type = options.str("device-type");
    const unsigned int num_iter = static_cast<unsigned int>(options.uint("num-iter"));
    const unsigned int num_tries = static_cast<unsigned int>(options.uint("num-tries"));
    const double beta_min = options.real("beta-min"), beta_max = options.real("beta-max");

    if (device_type != "cpu" && device_type != "gpu" && device_type != "host") {
      std::cerr << "Unknown device type: " << device_type << std::endl;
      return -1;
    }
    if (schedule_type != "linear" && schedule_type != "geometric") {
      std::cerr << "Unknown beta schedule: " << schedule_type << std::endl;
      return -1;
    }
    if (beta_max < 0 || beta_min < 0) {
      std::cerr << "Invalid schedule, both ends of beta range need to be positive" << std::endl;
      return -1;
    }
    if (beta_min >= beta_max) {
      std::cerr << "Invalid schedule, initial beta is not lesser than final beta" << std::endl;
      return -1;
    }
    sa::Options engine;
    const std::string mode = options.str("mode"), accept = options.str("accept");
    const std::string precision = options.str("precision"), layout = options.str("layout");
    if (mode != "random" && mode != "sweep") {
      std::cerr << "Unknown mode: " << mode << std::endl;
      return -1;
    }
    if (accept != "reference" && accept != "boltzmann") {
      std::cerr << "Unknown acceptance rule: " << accept << std::endl;
      return -1;
    }
    if (precision != "f64" && precision != "f32") {
      std::cerr << "Unknown precision: " << precision << std::endl;
      return -1;
    }
    if (layout != "auto" && layout != "dense" && layout != "csr") {
      std::cerr << "Unknown layout: " << layout << std::endl;
      return -1;
    }
    const std::string algorithm = options.str("algorithm");
    if (algorithm != "sa" && algorithm != "pt") {
      std::cerr << "Unknown algorithm: " << algorithm << std::endl;
      return -1;
    }
    const unsigned int num_replicas = static_cast<unsigned int>(options.uint("num-replicas"));
    if (algorithm == "pt" && num_replicas < 2) {
      std::cerr << "Parallel tempering needs at least two replicas" << std::endl;
      return -1;
    }
    engine.mode = mode == "sweep" ? OSA_MODE_SEQUENTIAL_SWEEP : OSA_MODE_RANDOM_SITE;
    engine.accept_rule = accept == "boltzmann" ? OSA_ACCEPT_BOLTZMANN : OSA_ACCEPT_REFERENCE;
    engine.sweep_precision = precision == "f32" ? OSA_SWEEP_F32 : OSA_SWEEP_F64;
    engine.layout = layout == "csr" ? sa::Layout::csr
                                    : (layout == "dense" ? sa::Layout::dense : sa::Layout::automatic);
    engine.seed = options.uint("seed");
    const int sweeps_per_beta = static_cast<int>(options.uint("sweeps-per-beta"));
    osa_stats stats{};
    if (options.count("stats")) engine.stats = &stats;

    std::cout << "Reading input from: " << input_file << std::endl;
    std::cout << "Output will be saved to: " << output_file << std::endl;
    std::cout << "Schedule type: " << schedule_type << std::endl;
    std::cout << "Beta range: [" << beta_min << ", " << beta_max << "]" << std::endl;
    std::cout << "Number of iterations: " << num_iter << std::endl;
    std::cout << "Number of tries: " << num_tries << std::endl;

    std::ifstream qubo_file(input_file);
    if (!qubo_file) {
      std::cerr << "can not open input file: " << input_file << std::endl;
      return -1;
    }
    auto instance = qubo::QUBOModel<int, double>::load(qubo_file);

    std::unique_ptr<devices::queue> q_ptr;
    try {
      q_ptr.reset(new devices::queue(*devices::construct_device_selector(device_type),
                                     static_cast<int>(options.uint("gpu-index"))));
    } catch (const std::runtime_error &e) {
      // the reference prints this and then dereferences a null queue; we stop here instead
      std::cerr << "No devices of given type could be initialized." << std::endl;
      std::cerr << "error: " << e.what() << "\n";
      return 1;
    }
    std::cout << "Using device: " << q_ptr->device_name() << std::endl;

    // simulated annealing: one beta per iteration; parallel tempering: one beta per replica
    const unsigned int ladder = algorithm == "pt" ? num_replicas : num_iter;
    std::vector<double> beta_schedule(ladder);
    if (schedule_type == "linear") {
      construct_linear_beta_schedule(beta_schedule, beta_min, beta_max, ladder);
    } else {
      construct_geometric_beta_schedule(beta_schedule, beta_min, beta_max, ladder);
    }

    auto solution =
        algorithm == "pt"
            ? sa::parallel_tempering(instance, *q_ptr, beta_schedule, static_cast<int>(num_iter),
                                     sweeps_per_beta, num_tries, engine)
            : sa::anneal(instance, *q_ptr, beta_schedule, static_cast<int>(num_iter), num_tries,
                         sweeps_per_beta, engine);

    std::ofstream results_file(output_file);
    solution.save(results_file);
    results_file.close();

    if (options.count("stats") && q_ptr->is_gpu()) {
      std::cout << "Kernel: " << osa_kernel_name(stats.kernel_id) << ", attempts " << stats.attempts
                << ", accepts " << stats.accepts << ", row fetches " << stats.row_fetches
                << ", device ms " << stats.ms_total << " (sweep " << stats.ms_sweep << ", energy "
                << stats.ms_energy << ")" << std::endl;
      if (algorithm == "pt") std::cout << "Replica exchanges accepted: " << stats.pt_swaps << std::endl;
    }
  } catch (std::exception &e) {
    std::cerr << "error: " << e.what() << "\n";
    return 1;
  } catch (...) {
    std::cerr << "Exception of unknown type!\n";
  }
  return 0;
}
