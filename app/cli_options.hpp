// app/cli_options.hpp -- minimal `--flag value` / `--flag=value` parser for the two CLIs.
// Stands in for boost::program_options as used by the reference
// (/root/reference/app/one-solver-anneal.cpp:55-76): long options only, typed values with
// defaults, unknown options and malformed values raise std::runtime_error (caught by main and
// reported as "error: ...", exit code 1, like the reference's catch-all at :171-173).
#ifndef ONESOLVER_B200_APP_CLI_OPTIONS_HPP_
#define ONESOLVER_B200_APP_CLI_OPTIONS_HPP_

#include <cstdlib>
#include <iomanip>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace cli {

struct OptionSpec {
  std::string name;
  bool takes_value;
  std::string default_text;  // "" = no default
  std::string help;
};

class Options {
public:
  explicit Options(std::string caption) : caption_(std::move(caption)) {}

  Options &add(const std::string &name, bool takes_value, const std::string &default_text,
               const std::string &help) {
    specs_.push_back({name, takes_value, default_text, help});
    if (!default_text.empty()) values_[name] = default_text;
    return *this;
  }

  void parse(int argc, char **argv) {
    for (int k = 1; k < argc; ++k) {
      std::string arg = argv[k];
      if (arg.rfind("--", 0) != 0) throw std::runtime_error("too many positional options have been specified on the command line");
      std::string name = arg.substr(2), value;
      bool has_value = false;
      const auto eq = name.find('=');
      if (eq != std::string::npos) {
        value = name.substr(eq + 1);
        name = name.substr(0, eq);
        has_value = true;
      }
      const OptionSpec *spec = find(name);
      if (!spec) throw std::runtime_error("unrecognised option '--" + name + "'");
      if (spec->takes_value) {
        if (!has_value) {
          if (k + 1 >= argc) throw std::runtime_error("the required argument for option '--" + name + "' is missing");
          value = argv[++k];
        }
        values_[name] = value;
      } else {
        if (has_value) throw std::runtime_error("option '--" + name + "' does not take any arguments");
        values_[name] = "1";
      }
      seen_[name] = true;
    }
  }

  bool count(const std::string &name) const { return seen_.count(name) != 0; }

  std::string str(const std::string &name) const {
    const auto it = values_.find(name);
    return it == values_.end() ? std::string() : it->second;
  }
  unsigned long long uint(const std::string &name) const {
    const std::string v = str(name);
    char *end = nullptr;
    if (v.empty() || v[0] == '-' || v[0] == '+') bad(name, v);
    const unsigned long long r = std::strtoull(v.c_str(), &end, 10);
    if (*end != '\0') bad(name, v);
    return r;
  }
  double real(const std::string &name) const {
    const std::string v = str(name);
    char *end = nullptr;
    const double r = std::strtod(v.c_str(), &end);
    if (v.empty() || *end != '\0') bad(name, v);
    return r;
  }

  std::string help() const {
    std::ostringstream os;
    os << caption_ << ":\n";
    for (const auto &s : specs_) {
      std::string left = "  --" + s.name;
      if (s.takes_value) left += s.default_text.empty() ? " arg" : " arg (=" + s.default_text + ")";
      os << std::left << std::setw(36) << left << " " << s.help << "\n";
    }
    return os.str();
  }

private:
  std::string caption_;
  std::vector<OptionSpec> specs_;
  std::map<std::string, std::string> values_;
  std::map<std::string, bool> seen_;

  const OptionSpec *find(const std::string &name) const {
    for (const auto &s : specs_)
      if (s.name == name) return &s;
    return nullptr;
  }
  [[noreturn]] static void bad(const std::string &name, const std::string &v) {
    throw std::runtime_error("the argument ('" + v + "') for option '--" + name + "' is invalid");
  }
};

}  // namespace cli

#endif
