// tests/cpp/host_tests.cpp -- host-layer unit tests without Boost.Test.
// Ports every assertion of the reference's suites: tests/qubo_test.cpp (3 cases),
// tests/qubo_helpers_test.cpp (1), tests/io_test.cpp (6 + 5 data cases), tests/devices_test.cpp (3),
// tests/exhaustive_test.cpp (5 instances, 1e-13), and adds schedule / host-engine / CSR checks.
// Usage: host_tests [--dump-parse FILE]   (the dump mode feeds tests/test_host_cpp.py's fuzzing)
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "exhaustive/exhaustive.hpp"
#include "helpers/devices.hpp"
#include "helpers/qubo_helpers.hpp"
#include "model/qubo.hpp"
#include "model/solution.hpp"
#include "schedules.hpp"
#include "simulated_annealing/annealing.hpp"

static int g_failed = 0, g_checks = 0;
#define CHECK(cond)                                                                     \
  do {                                                                                  \
    ++g_checks;                                                                         \
    if (!(cond)) {                                                                      \
      ++g_failed;                                                                       \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);            \
    }                                                                                   \
  } while (0)
#define CHECK_THROWS(expr, exc)                                                         \
  do {                                                                                  \
    ++g_checks;                                                                         \
    bool thrown = false;                                                                \
    try { (void)(expr); } catch (const exc &) { thrown = true; } catch (...) {}         \
    if (!thrown) {                                                                      \
      ++g_failed;                                                                       \
      std::fprintf(stderr, "FAILED %s:%d: %s did not throw %s\n", __FILE__, __LINE__,   \
                   #expr, #exc);                                                        \
    }                                                                                   \
  } while (0)

using Model = qubo::QUBOModel<int, double>;

static Model load_text(const std::string &text) {
  std::stringstream stream(text);
  stream.unsetf(std::ios::skipws);  // the reference's callers do this; must be harmless here
  return Model::load(stream);
}

// ---- tests/qubo_test.cpp ---------------------------------------------------------------
static void qubo_suite() {
  {
    qubo::LinearCoef<int, double> lin{{1, -0.4}, {2, 1.1}, {3, -1.0}, {4, -1.2}};
    qubo::QuadraticCoef<int, double> quad{{{1, 2}, 1.0}, {{1, 3}, 7.2}, {{1, 4}, 1.0},
                                          {{2, 3}, 2.0}, {{2, 4}, 1.9}, {{3, 4}, 3.0}};
    Model m(lin, quad);
    m.add_variable(1, 20);
    CHECK(m.get_variable(1) == 20);
    CHECK(m.get_variable(99) == 0);
  }
  {
    qubo::LinearCoef<int, double> lin{};
    qubo::QuadraticCoef<int, double> quad{};
    Model m(lin, quad);
    m.add_connection(std::make_pair(1, 2), 20);
    CHECK(m.get_connection(std::make_pair(1, 2)) == 20);
    CHECK(m.get_connection(std::make_pair(2, 1)) == 0);
  }
  {
    qubo::LinearCoef<int, int> lin{};
    qubo::QuadraticCoef<int, int> quad{};
    qubo::QUBOModel<int, int> m(lin, quad);
    m.add_variable(1, 10);
    m.add_connection(std::make_pair(1, 2), 20);
    CHECK(m.str() == "QUBO model 1--1:10 1--2:20");
    std::ostringstream os;
    os << m;
    CHECK(os.str() == "QUBO model 1--1:10 1--2:20");
  }
}

// ---- tests/qubo_helpers_test.cpp -------------------------------------------------------
static void flatten_suite() {
  qubo::LinearCoef<int, double> lin{{0, 0.5}, {1, -2.0}, {2, 1.0}, {4, -1.5}};
  qubo::QuadraticCoef<int, double> quad{{{0, 1}, 1.0}, {{0, 3}, 7.2},  {{1, 4}, -1.0},
                                        {{2, 3}, 2.0}, {{2, 4}, -1.5}, {{3, 4}, -3.5}};
  Model m(lin, quad);
  m.set_nodes(5);
  const std::vector<double> expected{0.5, 1.0,  0.0, 7.2, 0.0, 1.0, -2.0, 0.0, 0.0,  -1.0, 0.0,  0.0, 1.0,
                                     2.0, -1.5, 7.2, 0.0, 2.0, 0.0, -3.5, 0.0, -1.0, -1.5, -3.5, -1.5};
  CHECK(helpers::flatten_qubo(m) == expected);
  // a model holding both orientations gets their sum on both sides
  m.add_connection(std::make_pair(1, 0), 0.25);
  const auto both = helpers::flatten_qubo(m);
  CHECK(both[0 + 1 * 5] == 1.25 && both[1 + 0 * 5] == 1.25);
  // CSR view of the same model: symmetric, sorted, diagonal separate
  const auto csr = helpers::build_csr(m);
  CHECK(csr.rowptr.size() == 6 && csr.rowptr[5] == 12);
  CHECK(csr.diag[3] == 0.0 && csr.diag[4] == -1.5);
  bool ok = true;
  for (int i = 0; i < 5; ++i)
    for (int p = csr.rowptr[i]; p < csr.rowptr[i + 1]; ++p) {
      ok = ok && csr.val[p] == both[i * 5 + csr.col[p]];
      if (p > csr.rowptr[i]) ok = ok && csr.col[p] > csr.col[p - 1];
    }
  CHECK(ok);
}

// ---- tests/io_test.cpp -----------------------------------------------------------------
static void io_suite() {
  const std::string good = "c qubo Target MaxNodes NumNodes NumLinks\n"
                           "p qubo 0 100 2 2\n"
                           "c comment\n"
                           "0 0 -0.5\n"
                           "0 1 2.0\n"
                           "1 2 4\n"
                           "2 2 -0.7\n";
  {
    Model m = load_text(good);
    CHECK(m.get_nodes() == 3);  // max index + 1, not the header value
    CHECK(m.get_variable(0) == -0.5);
    CHECK(m.get_variable(2) == -0.7);
    CHECK(m.get_connection(std::pair(0, 1)) == 2.0);
    CHECK(m.get_connection(std::pair(1, 2)) == 4.0);
  }
  CHECK_THROWS(load_text("p qubo 0 100 2 2\n0 0 -0.5\n1 0 2.0\n1 2 4\n2 2 -0.7\n"),
               std::invalid_argument);
  const std::string malformed[] = {
      "p qubo 0 1 100 3 492\n0 0 -0.5\n0 1 2.0\n", "p qubo 0 100 3 492\n0 0 1 -0.5\n0 1 2.0\n",
      "p qubo 0 100 3 492\n0 0 -0.5\n0 2.0\n", "p qubo 0 100 492\n0 0 -0.5\n0 1 2.0\n",
      "p qubo 0 100 3 492\n0 0 -0.5\nunexpected string\n0 1 2.0\n"};
  for (const auto &text : malformed) CHECK_THROWS(load_text(text), std::invalid_argument);
  CHECK_THROWS(load_text("p qubo 0 100 3 3\n0 0 -0.5\n0 1 2.0\n1 2 4\n2 2 -0.7\n"),
               std::invalid_argument);
  {
    char state[] = {0, 1, 1, 0, 1};
    qubo::Solution solution(state, state + 5, -12.5);
    std::stringstream stream("");
    solution.save(stream);
    CHECK(stream.str() == "0,1,2,3,4,energy\n0,1,1,0,1,-12.5\n");
  }
  // grammar corners
  CHECK_THROWS(load_text(""), std::invalid_argument);
  CHECK_THROWS(load_text("c only a comment\n"), std::invalid_argument);
  CHECK_THROWS(load_text("0 0 1\n"), std::invalid_argument);                        // no header
  CHECK_THROWS(load_text("p qubo 0 4 1 0\n\n0 0 1\n"), std::invalid_argument);      // blank line
  CHECK_THROWS(load_text("p  qubo 0 4 1 0\n0 0 1\n"), std::invalid_argument);       // two spaces
  CHECK_THROWS(load_text("p qubo 0 4 1 0\n0\t0 1\n"), std::invalid_argument);       // tab after index
  CHECK_THROWS(load_text("0 0 1\np qubo 0 4 1 0\n"), std::invalid_argument);        // header late
  CHECK_THROWS(load_text("p qubo 0 1 2 0\n0 0 1\n1 1 1\n"), std::invalid_argument); // nLin > maxNodes
  CHECK(load_text("p qubo 0 4 1 0\n0 0 1").get_nodes() == 1);                       // no final newline
  CHECK(load_text("  p qubo 0 4 1 1 \r\n 0 0  1e0 \r\n0 3 -.5\t\n  ").get_nodes() == 4);
  CHECK(load_text("p qubo 0 4 2 0\n0 0 1\n0 0 7\n2 2 3\n").get_variable(0) == 1.0); // first duplicate wins
}

// ---- tests/devices_test.cpp ------------------------------------------------------------
static void devices_suite() {
  CHECK(dynamic_cast<devices::host_selector *>(devices::construct_device_selector("host").get()) != nullptr);
  CHECK(dynamic_cast<devices::cpu_selector *>(devices::construct_device_selector("cpu").get()) != nullptr);
  CHECK(dynamic_cast<devices::gpu_selector *>(devices::construct_device_selector("gpu").get()) != nullptr);
  bool message_ok = false;
  try {
    devices::construct_device_selector("fpga");
  } catch (const std::invalid_argument &e) {
    message_ok = std::string(e.what()) == "Unknown device type: fpga";
  }
  CHECK(message_ok);
}

// ---- tests/exhaustive_test.cpp ---------------------------------------------------------
static void exhaustive_suite(const std::string &examples_dir) {
  const std::pair<const char *, double> cases[] = {
      {"test1.qubo", -12.0}, {"test2.qubo", -1.2}, {"csp5.qubo", -22.0},
      {"csp7.qubo", -14.0},  {"csp13.qubo", -32.0}, {"simple.qubo", -2.0}};
  devices::queue q(*devices::construct_device_selector("cpu"));
  for (const auto &c : cases) {
    std::ifstream f(examples_dir + "/" + c.first);
    CHECK(static_cast<bool>(f));
    if (!f) continue;
    auto model = Model::load(f);
    const auto solution = exhaustive::solve(q, model);
    CHECK(std::fabs(solution.energy - c.second) < 1e-13);
    const auto flat = helpers::flatten_qubo(model);
    CHECK(std::fabs(sa::energy(flat, solution.state, static_cast<int>(model.get_nodes())) - c.second) < 1e-13);
  }
  std::ifstream f(examples_dir + "/dwave_doc.qubo");  // "p  qubo": the reference rejects this file
  CHECK_THROWS(Model::load(f), std::invalid_argument);
}

// ---- schedules (one-solver-anneal.cpp:23-39) and the host engine -----------------------
static void engine_suite(const std::string &examples_dir) {
  std::vector<double> lin(100), geo(100);
  construct_linear_beta_schedule(lin, 0.1, 1.0, 100);
  construct_geometric_beta_schedule(geo, 0.1, 1.0, 100);
  CHECK(lin[0] == 0.1 && lin[99] == 0.1 + 1.0);  // the reference quirk: ends at min + max
  CHECK(geo[0] == 0.1 && std::fabs(geo[99] - 1.0) < 1e-12);
  CHECK(geo[1] == 0.1 * std::pow(10.0, 1.0 / 99));

  std::ifstream f(examples_dir + "/test1.qubo");
  auto model = Model::load(f);
  devices::queue host(*devices::construct_device_selector("host"));
  auto s = sa::anneal(model, host, geo, 100, 100);  // BASELINE config 1
  CHECK(s.energy == -12.0);
  CHECK(s.state == (std::vector<char>{1, 1, 0, 1}));
  devices::queue cpu(*devices::construct_device_selector("cpu"));
  auto s2 = sa::anneal(model, cpu, geo, 100, 100);
  CHECK(s2.energy == s.energy && s2.state == s.state);  // thread count does not change results
}

static int dump_parse(const char *path) {
  std::ifstream f(path, std::ios::binary);
  try {
    auto m = Model::load(f);
    std::printf("OK %lu %zu %zu\n", m.get_nodes(), m.linear_terms().size(), m.quadratic_terms().size());
    for (const auto &t : m.linear_terms()) std::printf("L %d %.17g\n", t.first, t.second);
    for (const auto &t : m.quadratic_terms())
      std::printf("Q %d %d %.17g\n", t.first.first, t.first.second, t.second);
  } catch (const std::invalid_argument &e) {
    std::printf("REJECT %s\n", e.what());
  }
  return 0;
}

int main(int argc, char **argv) {
  if (argc == 3 && std::string(argv[1]) == "--dump-parse") return dump_parse(argv[2]);
  const std::string examples = argc > 1 ? argv[1] : "examples";
  qubo_suite();
  flatten_suite();
  io_suite();
  devices_suite();
  exhaustive_suite(examples);
  engine_suite(examples);
  std::printf("%d checks, %d failed\n", g_checks, g_failed);
  return g_failed == 0 ? 0 : 1;
}
