// model/qubo.hpp -- QUBO model container and the `.qubo` (qbsolv-style) reader.
//
// Same public surface as the reference's model layer
// (/root/reference/include/model/qubo.hpp): qubo::LinearCoef / QuadraticCoef,
// qubo::QUBOModel<Node, Coef> (:53-240), qubo::QUBOBuilder (:282-379),
// qubo::parse_qubo(first, last) (:392-417) and QUBOModel::load(istream&) (:427-430).
// The reference parses with Boost.Spirit; this reader is a Boost-free hand-written
// scanner that accepts and rejects exactly the same inputs (the grammar is restated
// rule by rule below; tests/cpp/host_tests.cpp ports the reference's io_test.cpp cases
// and tests/test_host_cpp.py fuzzes it against oracle/qubo_format.py).
#ifndef ONESOLVER_B200_MODEL_QUBO_HPP_
#define ONESOLVER_B200_MODEL_QUBO_HPP_

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <charconv>
#include <cstdlib>
#include <cstring>
#include <istream>
#include <iterator>
#include <limits>
#include <ostream>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>

#include "helpers/hash.hpp"
#include "helpers/insert.hpp"

namespace qubo {

template <class NodeType, class CoefType>
using LinearCoef = std::unordered_map<NodeType, CoefType>;

template <class NodeType, class CoefType>
using QuadraticCoef =
    std::unordered_map<std::pair<NodeType, NodeType>, CoefType, helpers::hash_pair>;

// Sparse coefficient store: linear terms q_ii and couplings q_ij keyed by (i, j).
template <class NodeType, class CoefType>
class QUBOModel {
public:
  using Linear = LinearCoef<NodeType, CoefType>;
  using Quadratic = QuadraticCoef<NodeType, CoefType>;

  QUBOModel() = default;
  QUBOModel(const Linear &c_linear, const Quadratic &c_quadratic)
      : linear(c_linear), quadratic(c_quadratic) {}
  QUBOModel(Linear &&c_linear, Quadratic &&c_quadratic)
      : linear(std::move(c_linear)), quadratic(std::move(c_quadratic)) {}

  // "QUBO model" + " i--i:v " per linear term + "i--j:v" per coupling (tests/qubo_test.cpp:54)
  std::string str() const {
    std::string text = "QUBO model";
    for (const auto &term : linear) {
      const std::string node = std::to_string(term.first);
      text += " " + node + "--" + node + ":" + std::to_string(term.second) + " ";
    }
    for (const auto &term : quadratic) {
      text += std::to_string(term.first.first) + "--" + std::to_string(term.first.second) + ":" +
              std::to_string(term.second);
    }
    return text;
  }

  // insert-or-overwrite, like the reference (qubo.hpp:205-214)
  void add_variable(const NodeType &vi, const CoefType &hi) {
    helpers::insert_model(linear, vi, hi);
  }
  void add_connection(const std::pair<NodeType, NodeType> &connection, const CoefType &Ji) {
    helpers::insert_model(quadratic, connection, Ji);
  }

  // absent entries read as 0 (qubo.hpp:217-240)
  const CoefType get_variable(const NodeType &vi) const {
    const auto it = linear.find(vi);
    return it == linear.end() ? CoefType(0) : it->second;
  }
  const CoefType get_connection(const std::pair<NodeType, NodeType> &connection) const {
    const auto it = quadratic.find(connection);
    return it == quadratic.end() ? CoefType(0) : it->second;
  }

  void set_nodes(int number) { num_nodes = static_cast<unsigned long>(number); }
  unsigned long get_nodes() const { return num_nodes; }

  // read-only views used by the O(nnz) layout builders (helpers/qubo_helpers.hpp)
  const Linear &linear_terms() const { return linear; }
  const Quadratic &quadratic_terms() const { return quadratic; }

  static QUBOModel<int, double> load(std::istream &stream);

protected:
  Linear linear;
  Quadratic quadratic;
  unsigned long num_nodes = 0;
};

template <class NodeType, class CoefType>
std::ostream &operator<<(std::ostream &os, const QUBOModel<NodeType, CoefType> &model) {
  return os << model.str();
}

// Collects what the reader sees and validates it against the header line.
struct QUBOBuilder {
  // i == j: linear, i < j: coupling, i > j: error. The first value of a duplicate wins.
  void add_element(int i, int j, double coef) {
    if (i > j) {
      throw std::invalid_argument(
          "Incorrect file, encountered coefficient from lower triangle of QUBO matrix.");
    }
    if (i == j) {
      linear_c.emplace(i, coef);
    } else {
      quadratic_c.emplace(std::make_pair(i, j), coef);
    }
    largest_node = std::max(largest_node, j);  // i <= j here
  }
  // the header announces the term counts: size the maps once instead of rehashing on the way
  void set_num_quadratic(int value) {
    num_quadratic = value;
    if (value > 0) quadratic_c.reserve(static_cast<std::size_t>(value));
  }
  void set_num_linear(int value) {
    num_linear = value;
    if (value > 0) linear_c.reserve(static_cast<std::size_t>(value));
  }
  void set_max_nodes(int value) { max_nodes = value; }

  // checks and messages follow qubo.hpp:354-378, in the same order
  QUBOModel<int, double> build_qubo() {
    if (linear_c.empty() && quadratic_c.empty()) {
      throw std::invalid_argument("An empty input, no coefficients defined.");
    }
    if (num_quadratic == -1 || num_linear == -1 || max_nodes == -1) {
      throw std::invalid_argument("No header line or header line misformatted.");
    }
    if (num_linear > max_nodes) {
      throw std::invalid_argument("Number of linear terms is greater than num nodes.");
    }
    if (quadratic_c.size() != static_cast<std::size_t>(num_quadratic)) {
      throw std::invalid_argument("Number of quadratic terms is not equal to the declared one.");
    }
    if (linear_c.size() != static_cast<std::size_t>(num_linear)) {
      throw std::invalid_argument("Number of linear terms is not equal to the declared one.");
    }
    QUBOModel<int, double> model(std::move(linear_c), std::move(quadratic_c));
    model.set_nodes(largest_node + 1);  // N = largest index seen + 1
    return model;
  }

private:
  LinearCoef<int, double> linear_c;
  QuadraticCoef<int, double> quadratic_c;
  int num_quadratic = -1, num_linear = -1, max_nodes = -1;
  int largest_node = -1;
};

namespace detail {

// Scanner over the whole text.  Grammar (reference qubo.hpp:395-411), blanks = ' ' | '\t'
// skipped before every token:
//   file         := comment* header? (comment | coefficients)*          -- all input consumed
//   comment      := 'c' printable* (EOL | EOI)
//   header       := "p qubo" uint uint uint uint (EOL | EOI)            -- target maxNodes nLin nQuad
//   coefficients := uint ' ' uint ' ' real (EOL | EOI)                   -- literal space after each index
//   EOL          := "\r\n" | "\n" | "\r"
class QuboScanner {
public:
  QuboScanner(const std::string &text, QUBOBuilder &builder) : s(text), b(builder) {}

  bool run() {
    std::size_t pos = 0, next = 0;
    while (comment(pos, next) && next != pos) pos = next;
    if (header(pos, next)) pos = next;
    while (pos < s.size()) {
      if (!(comment(pos, next) || coefficients(pos, next)) || next == pos) break;
      pos = next;
    }
    return skip(pos) == s.size();
  }

private:
  const std::string &s;
  QUBOBuilder &b;

  std::size_t skip(std::size_t p) const {
    while (p < s.size() && (s[p] == ' ' || s[p] == '\t')) ++p;
    return p;
  }
  bool line_end(std::size_t p, std::size_t &out) const {
    p = skip(p);
    if (p == s.size()) {
      out = p;
      return true;
    }
    bool matched = false;
    if (p < s.size() && s[p] == '\r') { ++p; matched = true; }
    if (p < s.size() && s[p] == '\n') { ++p; matched = true; }
    out = p;
    return matched;
  }
  // unsigned decimal that fits in 32 bits
  bool uint(std::size_t p, std::size_t &out, unsigned long long &value) const {
    std::size_t q = p;
    value = 0;
    while (q < s.size() && s[q] >= '0' && s[q] <= '9') {
      value = value * 10 + static_cast<unsigned>(s[q] - '0');
      if (value > 0xFFFFFFFFull) return false;
      ++q;
    }
    out = q;
    return q != p;
  }
  // [+-] (digits [. digits*] | . digits) [e[+-]digits] | [+-] inf|infinity|nan[(...)]
  bool real(std::size_t p, std::size_t &out, double &value) const {
    std::size_t q = p;
    if (q < s.size() && (s[q] == '+' || s[q] == '-')) ++q;
    auto digits = [&](std::size_t &r) {
      const std::size_t r0 = r;
      while (r < s.size() && s[r] >= '0' && s[r] <= '9') ++r;
      return r != r0;
    };
    auto word = [&](std::size_t r, const char *w) {
      const std::size_t len = std::strlen(w);
      if (r + len > s.size()) return false;
      for (std::size_t k = 0; k < len; ++k)
        if (std::tolower(static_cast<unsigned char>(s[r + k])) != w[k]) return false;
      return true;
    };
    std::size_t end = q;
    bool numeric = true;
    if (digits(end)) {
      if (end < s.size() && s[end] == '.') {
        ++end;
        digits(end);
      }
    } else if (end < s.size() && s[end] == '.') {
      ++end;
      if (!digits(end)) return false;
    } else if (word(q, "infinity")) {
      end = q + 8;
      numeric = false;
    } else if (word(q, "inf")) {
      end = q + 3;
      numeric = false;
    } else if (word(q, "nan")) {
      end = q + 3;
      numeric = false;
      if (end < s.size() && s[end] == '(') {
        const std::size_t close = s.find(')', end);
        if (close != std::string::npos) end = close + 1;
      }
    } else {
      return false;
    }
    if (numeric && end < s.size() && (s[end] == 'e' || s[end] == 'E')) {
      std::size_t r = end + 1;
      if (r < s.size() && (s[r] == '+' || s[r] == '-')) ++r;
      if (digits(r)) end = r;  // otherwise the 'e' is left for the line-end check to reject
    }
    out = end;
    if (numeric) {
      // plain decimal: std::from_chars on the validated extent (correctly rounded like strtod,
      // no copy of the token); it takes '-' but not '+'
      const char *first = s.data() + p, *last = s.data() + end;
      if (*first == '+') ++first;
      const auto res = std::from_chars(first, last, value);
      if (res.ec == std::errc() && res.ptr == last) return true;
      // out of range etc.: strtod's answer (+-inf / denormal) below, as before
    }
    const std::string token = s.substr(p, end - p);
    char *stop = nullptr;
    value = std::strtod(token.c_str(), &stop);
    return stop != token.c_str();
  }

  bool comment(std::size_t p, std::size_t &out) const {
    p = skip(p);
    if (p >= s.size() || s[p] != 'c') return false;
    ++p;
    for (;;) {
      p = skip(p);
      if (p < s.size() && static_cast<unsigned char>(s[p]) >= 0x20 &&
          static_cast<unsigned char>(s[p]) <= 0x7E) {
        ++p;
      } else {
        break;
      }
    }
    return line_end(p, out);
  }
  bool header(std::size_t p, std::size_t &out) {
    p = skip(p);
    if (s.compare(p, 6, "p qubo") != 0) return false;
    p += 6;
    unsigned long long v[4];
    for (int k = 0; k < 4; ++k) {
      p = skip(p);
      std::size_t q;
      if (!uint(p, q, v[k])) return false;
      p = q;
    }
    if (!line_end(p, out)) return false;
    b.set_max_nodes(static_cast<int>(v[1]));
    b.set_num_linear(static_cast<int>(v[2]));
    b.set_num_quadratic(static_cast<int>(v[3]));
    return true;
  }
  bool coefficients(std::size_t p, std::size_t &out) {
    unsigned long long idx[2];
    for (int k = 0; k < 2; ++k) {
      p = skip(p);
      std::size_t q;
      if (!uint(p, q, idx[k]) || q >= s.size() || s[q] != ' ') return false;
      if (idx[k] > static_cast<unsigned long long>(std::numeric_limits<int>::max())) return false;
      p = q + 1;
    }
    p = skip(p);
    std::size_t q;
    double coef;
    if (!real(p, q, coef)) return false;
    // the reference fires its semantic action as soon as the triple is read
    b.add_element(static_cast<int>(idx[0]), static_cast<int>(idx[1]), coef);
    return line_end(q, out);
  }
};

}  // namespace detail

// Parse a whole `.qubo` text given as an iterator range of chars.
template <typename Iterator>
QUBOModel<int, double> parse_qubo(Iterator first, Iterator last) {
  const std::string text(first, last);
  QUBOBuilder builder;
  detail::QuboScanner scanner(text, builder);
  if (!scanner.run()) {
    throw std::invalid_argument("Parsing failed. Incorrect file format.");
  }
  return builder.build_qubo();
}

// Load from a stream (works whether or not the caller cleared std::ios::skipws).
template <class NodeType, class CoefType>
QUBOModel<int, double> QUBOModel<NodeType, CoefType>::load(std::istream &stream) {
  std::istreambuf_iterator<char> first(stream), last;
  return parse_qubo(first, last);
}

}  // namespace qubo

#endif
