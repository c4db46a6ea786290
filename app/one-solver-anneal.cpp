// one-solver-anneal -- simulated annealing of a `.qubo` instance, `--device-type gpu` running
// on the B200 CUDA engine.
//
// Flag surface, defaults, banner lines, messages and exit codes follow the reference CLI
// (/root/reference/app/one-solver-anneal.cpp:41-179): --input --output --num-iter (100)
// --num-tries (100) --schedule-type (geometric) --beta-min (0.1) --beta-max (1.0)
// --device-type (host); exit 0 on success/help, -1 on bad arguments or an unreadable input,
// 1 on an exception.  Extra flags (all optional, defaults keep the reference behaviour):
// --mode random|sweep, --accept reference|boltzmann, --sweeps-per-beta, --seed,
// --precision f64|f32, --layout auto|dense|csr, --stats; --gpu-index K (one CUDA device) or
// --num-gpus G (devices 0..G-1) -- without either, --device-type gpu uses every visible GPU and
// shards the trajectories over them (osa_multi_anneal); --algorithm sa|pt|pa with
// --num-replicas: parallel tempering (GPU only) -- the beta range becomes a ladder of
// --num-replicas temperatures (built by the chosen schedule type), --num-iter counts rounds,
// --sweeps-per-beta the sweeps per round and --num-tries the independent ladders; population
// annealing (GPU only) -- the schedule of --num-iter temperatures is walked by --num-tries
// populations of --num-replicas replicas, resampled between the temperatures.
#include <vector>

#include "cli_common.hpp"
#include "model/solution.hpp"
#include "schedules.hpp"
#include "simulated_annealing/annealing.hpp"

namespace {

struct Settings {
  std::string input_file, output_file, schedule_type, device_type, algorithm;
  unsigned int num_iter = 0, num_tries = 0, num_replicas = 0;
  int sweeps_per_beta = 1;
  std::vector<int> gpu_devices;  // empty: every visible GPU
  double beta_min = 0.0, beta_max = 0.0;
  bool print_stats = false;
  sa::Options engine;
};

void declare_options(cli::Options &options) {
  cli::add_io_options(options);
  options.add("num-iter", true, "100", "number of iterations of the algorithm")
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
      .add("gpu-index", true, "", "use this CUDA device only (default: every visible GPU)")
      .add("num-gpus", true, "", "use CUDA devices 0..N-1 (default: every visible GPU)")
      .add("stats", false, "", "print engine statistics (gpu)")
      .add("algorithm", true, "sa",
           "sa (simulated annealing), pt (parallel tempering, gpu) or pa (population annealing, gpu)")
      .add("num-replicas", true, "12",
           "temperatures per ladder with --algorithm pt, replicas per population with pa");
}

// validation order and messages of the reference (one-solver-anneal.cpp:78-115), then our extras
Settings read_settings(const cli::Options &options) {
  Settings s;
  cli::require_io(options, s.input_file, s.output_file);
  s.schedule_type = options.str("schedule-type");
  s.num_iter = static_cast<unsigned int>(options.uint("num-iter"));
  s.num_tries = static_cast<unsigned int>(options.uint("num-tries"));
  s.beta_min = options.real("beta-min");
  s.beta_max = options.real("beta-max");
  s.device_type = cli::checked_device_type(options);
  if (s.schedule_type != "linear" && s.schedule_type != "geometric")
    cli::usage_error("Unknown beta schedule: " + s.schedule_type);
  if (s.beta_max < 0 || s.beta_min < 0)
    cli::usage_error("Invalid schedule, both ends of beta range need to be positive");
  if (s.beta_min >= s.beta_max)
    cli::usage_error("Invalid schedule, initial beta is not lesser than final beta");

  const std::string mode = cli::checked_choice(options, "mode", {"random", "sweep"}, "mode");
  const std::string accept =
      cli::checked_choice(options, "accept", {"reference", "boltzmann"}, "acceptance rule");
  const std::string precision = cli::checked_choice(options, "precision", {"f64", "f32"}, "precision");
  const std::string layout = cli::checked_choice(options, "layout", {"auto", "dense", "csr"}, "layout");
  s.algorithm = cli::checked_choice(options, "algorithm", {"sa", "pt", "pa"}, "algorithm");
  s.num_replicas = static_cast<unsigned int>(options.uint("num-replicas"));
  if (s.algorithm == "pt" && s.num_replicas < 2)
    cli::usage_error("Parallel tempering needs at least two replicas");

  s.engine.mode = mode == "sweep" ? OSA_MODE_SEQUENTIAL_SWEEP : OSA_MODE_RANDOM_SITE;
  s.engine.accept_rule = accept == "boltzmann" ? OSA_ACCEPT_BOLTZMANN : OSA_ACCEPT_REFERENCE;
  s.engine.sweep_precision = precision == "f32" ? OSA_SWEEP_F32 : OSA_SWEEP_F64;
  s.engine.layout = layout == "csr" ? sa::Layout::csr
                                    : (layout == "dense" ? sa::Layout::dense : sa::Layout::automatic);
  s.engine.seed = options.uint("seed");
  s.sweeps_per_beta = static_cast<int>(options.uint("sweeps-per-beta"));
  if (options.count("gpu-index") && options.count("num-gpus"))
    cli::usage_error("Give either --gpu-index or --num-gpus, not both");
  if (options.count("gpu-index")) s.gpu_devices = {static_cast<int>(options.uint("gpu-index"))};
  if (options.count("num-gpus")) {
    const int g = static_cast<int>(options.uint("num-gpus"));
    if (g < 1) cli::usage_error("--num-gpus must be at least 1");
    for (int d = 0; d < g; ++d) s.gpu_devices.push_back(d);
  }
  if (s.algorithm == "pa" && s.num_replicas < 1)
    cli::usage_error("Population annealing needs at least one replica per population");
  if (s.algorithm != "sa" && s.gpu_devices.empty()) s.gpu_devices = {0};  // pt / pa: one device
  s.print_stats = options.count("stats");
  return s;
}

void print_banner(const Settings &s) {
  std::cout << "Reading input from: " << s.input_file << std::endl;
  std::cout << "Output will be saved to: " << s.output_file << std::endl;
  std::cout << "Schedule type: " << s.schedule_type << std::endl;
  std::cout << "Beta range: [" << s.beta_min << ", " << s.beta_max << "]" << std::endl;
  std::cout << "Number of iterations: " << s.num_iter << std::endl;
  std::cout << "Number of tries: " << s.num_tries << std::endl;
}

}  // namespace

int main(int argc, char *argv[]) {
  return cli::run([&] {
    cli::Options options("Allowed options");
    declare_options(options);
    options.parse(argc, argv);
    Settings s = read_settings(options);
    osa_stats stats{};
    if (s.print_stats) s.engine.stats = &stats;
    print_banner(s);

    const auto instance = cli::read_model(s.input_file);
    const auto device = cli::open_device(s.device_type, s.gpu_devices);

    // simulated annealing: one beta per iteration; parallel tempering: one beta per replica
    const bool tempering = s.algorithm == "pt";
    const unsigned int ladder = tempering ? s.num_replicas : s.num_iter;
    std::vector<double> beta_schedule(ladder);
    if (s.schedule_type == "linear") {
      construct_linear_beta_schedule(beta_schedule, s.beta_min, s.beta_max, ladder);
    } else {
      construct_geometric_beta_schedule(beta_schedule, s.beta_min, s.beta_max, ladder);
    }

    // --algorithm pa: the schedule is the temperature ladder of a population of --num-replicas
    // replicas, --num-tries independent populations, --sweeps-per-beta sweeps per temperature
    const qubo::Solution solution =
        tempering ? sa::parallel_tempering(instance, *device, beta_schedule,
                                           static_cast<int>(s.num_iter), s.sweeps_per_beta,
                                           s.num_tries, s.engine)
        : s.algorithm == "pa"
            ? sa::population_annealing(instance, *device, beta_schedule, s.sweeps_per_beta,
                                       s.num_replicas, s.num_tries, s.engine)
            : sa::anneal(instance, *device, beta_schedule, static_cast<int>(s.num_iter),
                         s.num_tries, s.sweeps_per_beta, s.engine);

    std::ofstream results_file(s.output_file);
    solution.save(results_file);
    results_file.close();

    if (s.print_stats && device->is_gpu()) {
      std::cout << "Kernel: " << osa_kernel_name(stats.kernel_id) << ", attempts " << stats.attempts
                << ", accepts " << stats.accepts << ", row fetches " << stats.row_fetches
                << ", device ms " << stats.ms_total << " (sweep " << stats.ms_sweep << ", energy "
                << stats.ms_energy << ")" << std::endl;
      if (stats.reserved > 1) std::cout << "Devices: " << stats.reserved << std::endl;
      if (tempering) std::cout << "Replica exchanges accepted: " << stats.pt_swaps << std::endl;
      if (s.algorithm == "pa") std::cout << "Replicas resampled: " << stats.pt_swaps << std::endl;
    }
  });
}
