// one-solver-sweep -- the parameter sweep of the reference's annealing benchmark in one process.
//
// The reference drives its benchmark with Snakemake
// (/root/reference/benchmarks/annealing/Snakefile:19-72): for every point of the grid
// beta_min x num_iter x num_tries x schedule it launches one-solver-anneal, appends the four
// parameters to the two CSV lines with sed (:71-72) and concatenates the 1200 files with pandas
// (:41-58).  This program walks the same grid in the same order inside one process -- the model is
// parsed once, every point is one sa::anneal call -- and writes the merged table directly:
//
//   0,1,...,N-1,energy,beta_min,num_iter,num_tries,schedule
//   b0,b1,...,bN-1,<energy>,<beta_min>,<num_iter>,<num_tries>,<schedule>
//
// Each row is exactly what `one-solver-anneal` writes for those parameters (same schedule
// builders, same engine, same seed), so a table produced here can be compared line by line with
// the per-point CLI runs.  Lists are comma separated; integer lists also accept first:last:step
// (last exclusive, like Python's range in Snakefile:14-17).
//
// Snakefile quirk: its compute rule never passes --beta-min to the CLI (:63-70), so all six
// "beta_min" groups of the reference's published CSVs were in fact run with the CLI default 0.1.
// --reference-quirk reproduces that; by default the labelled beta_min is the one that is used.
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "cli_options.hpp"
#include "helpers/devices.hpp"
#include "model/qubo.hpp"
#include "model/solution.hpp"
#include "schedules.hpp"
#include "simulated_annealing/annealing.hpp"

namespace {

std::vector<std::string> split_list(const std::string &text) {
  std::vector<std::string> out;
  std::stringstream ss(text);
  std::string item;
  while (std::getline(ss, item, ','))
    if (!item.empty()) out.push_back(item);
  return out;
}

unsigned long long parse_uint(const std::string &v, const std::string &flag) {
  char *end = nullptr;
  if (v.empty() || v[0] == '-' || v[0] == '+')
    throw std::runtime_error("the argument ('" + v + "') for option '--" + flag + "' is invalid");
  const unsigned long long r = std::strtoull(v.c_str(), &end, 10);
  if (*end != '\0')
    throw std::runtime_error("the argument ('" + v + "') for option '--" + flag + "' is invalid");
  return r;
}

// "100,200,300" or "100:1100:100"
std::vector<unsigned int> parse_uint_list(const std::string &text, const std::string &flag) {
  std::vector<unsigned int> out;
  if (text.find(':') != std::string::npos) {
    std::vector<std::string> parts;
    std::stringstream ss(text);
    std::string item;
    while (std::getline(ss, item, ':')) parts.push_back(item);
    if (parts.size() != 3)
      throw std::runtime_error("option '--" + flag + "' expects first:last:step");
    const auto first = parse_uint(parts[0], flag), last = parse_uint(parts[1], flag),
               step = parse_uint(parts[2], flag);
    if (step == 0) throw std::runtime_error("option '--" + flag + "' has a zero step");
    for (auto v = first; v < last; v += step) out.push_back(static_cast<unsigned int>(v));
  } else {
    for (const auto &item : split_list(text))
      out.push_back(static_cast<unsigned int>(parse_uint(item, flag)));
  }
  if (out.empty()) throw std::runtime_error("option '--" + flag + "' selects nothing");
  return out;
}

double parse_real(const std::string &v, const std::string &flag) {
  char *end = nullptr;
  const double r = std::strtod(v.c_str(), &end);
  if (v.empty() || *end != '\0')
    throw std::runtime_error("the argument ('" + v + "') for option '--" + flag + "' is invalid");
  return r;
}

}  // namespace

int main(int argc, char *argv[]) {
  try {
    cli::Options options("Allowed options");
    options.add("help", false, "", "produce help message")
        .add("input", true, "", "input file (.qubo)")
        .add("output", true, "", "output file (merged CSV)")
        .add("beta-min", true, "0.1,0.3,0.5,0.7,0.9,1.1", "list of initial betas (Snakefile BETA_MIN)")
        .add("beta-max", true, "10", "final beta (Snakefile BETA_MAX)")
        .add("num-iter", true, "100:1100:100", "list or first:last:step (Snakefile NUM_ITERATIONS)")
        .add("num-tries", true, "10:500:50", "list or first:last:step (Snakefile NUM_TRIES)")
        .add("schedule-type", true, "linear,geometric", "list of schedules (Snakefile SCHEDULES)")
        .add("reference-quirk", false, "", "run every point with beta_min 0.1, as the Snakefile does")
        .add("device-type", true, "host", "device type to use (cpu, gpu or host)")
        .add("mode", true, "random", "site visiting order: random (reference) or sweep")
        .add("accept", true, "reference", "acceptance rule: reference or boltzmann")
        .add("sweeps-per-beta", true, "1", "attempts (random) or sweeps (sweep) per schedule step")
        .add("seed", true, "1234", "random seed")
        .add("precision", true, "f64", "sweep arithmetic on the gpu: f64 or f32")
        .add("layout", true, "auto", "gpu problem layout: auto, dense or csr")
        .add("gpu-index", true, "0", "CUDA device to use with --device-type gpu")
        .add("quiet", false, "", "no progress lines");
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
    const std::string device_type = options.str("device-type");
    if (device_type != "cpu" && device_type != "gpu" && device_type != "host") {
      std::cerr << "Unknown device type: " << device_type << std::endl;
      return -1;
    }
    const auto beta_mins = split_list(options.str("beta-min"));
    const auto schedules = split_list(options.str("schedule-type"));
    const auto num_iters = parse_uint_list(options.str("num-iter"), "num-iter");
    const auto num_tries_list = parse_uint_list(options.str("num-tries"), "num-tries");
    const double beta_max = options.real("beta-max");
    if (beta_mins.empty() || schedules.empty()) {
      std::cerr << "Empty parameter grid." << std::endl;
      return -1;
    }
    for (const auto &s : schedules) {
      if (s != "linear" && s != "geometric") {
        std::cerr << "Unknown beta schedule: " << s << std::endl;
        return -1;
      }
    }
    const bool quirk = options.count("reference-quirk");
    for (const auto &b : beta_mins) {
      const double beta_min = quirk ? 0.1 : parse_real(b, "beta-min");
      if (beta_max < 0 || beta_min < 0) {
        std::cerr << "Invalid schedule, both ends of beta range need to be positive" << std::endl;
        return -1;
      }
      if (beta_min >= beta_max) {
        std::cerr << "Invalid schedule, initial beta is not lesser than final beta" << std::endl;
        return -1;
      }
    }
    for (unsigned int it : num_iters) {
      if (it < 2) {
        std::cerr << "Invalid grid, num-iter must be at least 2" << std::endl;
        return -1;
      }
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
    engine.mode = mode == "sweep" ? OSA_MODE_SEQUENTIAL_SWEEP : OSA_MODE_RANDOM_SITE;
    engine.accept_rule = accept == "boltzmann" ? OSA_ACCEPT_BOLTZMANN : OSA_ACCEPT_REFERENCE;
    engine.sweep_precision = precision == "f32" ? OSA_SWEEP_F32 : OSA_SWEEP_F64;
    engine.layout = layout == "csr" ? sa::Layout::csr
                                    : (layout == "dense" ? sa::Layout::dense : sa::Layout::automatic);
    engine.seed = options.uint("seed");
    const int sweeps_per_beta = static_cast<int>(options.uint("sweeps-per-beta"));
    const bool quiet = options.count("quiet");

    const std::string input_file = options.str("input"), output_file = options.str("output");
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
      std::cerr << "No devices of given type could be initialized." << std::endl;
      std::cerr << "error: " << e.what() << "\n";
      return 1;
    }
    const std::size_t points =
        beta_mins.size() * num_iters.size() * num_tries_list.size() * schedules.size();
    if (!quiet) {
      std::cout << "Reading input from: " << input_file << std::endl;
      std::cout << "Output will be saved to: " << output_file << std::endl;
      std::cout << "Using device: " << q_ptr->device_name() << std::endl;
      std::cout << "Grid points: " << points << std::endl;
    }

    std::ofstream table(output_file);
    if (!table) {
      std::cerr << "can not open output file: " << output_file << std::endl;
      return -1;
    }
    bool header_written = false;
    std::size_t done = 0;
    for (const auto &beta_min_text : beta_mins) {
      const double beta_min = quirk ? 0.1 : parse_real(beta_min_text, "beta-min");
      for (unsigned int num_iter : num_iters) {
        for (unsigned int num_tries : num_tries_list) {
          for (const auto &schedule_type : schedules) {
            std::vector<double> beta_schedule(num_iter);
            if (schedule_type == "linear") {
              construct_linear_beta_schedule(beta_schedule, beta_min, beta_max, num_iter);
            } else {
              construct_geometric_beta_schedule(beta_schedule, beta_min, beta_max, num_iter);
            }
            const auto solution = sa::anneal(instance, *q_ptr, beta_schedule,
                                             static_cast<int>(num_iter), num_tries,
                                             sweeps_per_beta, engine);
            // the two lines one-solver-anneal would have written for this point
            std::ostringstream two_lines;
            solution.save(two_lines);
            const std::string text = two_lines.str();
            const auto nl = text.find('\n');
            const std::string head = text.substr(0, nl);
            std::string row = text.substr(nl + 1);
            if (!row.empty() && row.back() == '\n') row.pop_back();
            if (!header_written) {
              table << head << ",beta_min,num_iter,num_tries,schedule\n";
              header_written = true;
            }
            table << row << ',' << beta_min_text << ',' << num_iter << ',' << num_tries << ','
                  << schedule_type << '\n';
            ++done;
            if (!quiet && (done % 100 == 0 || done == points))
              std::cout << "Finished " << done << " / " << points << std::endl;
          }
        }
      }
    }
    table.close();
  } catch (std::exception &e) {
    std::cerr << "error: " << e.what() << "\n";
    return 1;
  } catch (...) {
    std::cerr << "Exception of unknown type!\n";
  }
  return 0;
}
