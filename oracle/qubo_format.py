"""oracle/qubo_format.py -- TEST INFRASTRUCTURE: restatement of the reference's .qubo reader.

Follows /root/reference/include/model/qubo.hpp:392-417 (Boost.Spirit grammar
`*comment >> -header >> *(comment | coefficients)` with a blank skipper) and the
QUBOBuilder checks at :282-379.  Used to check the product's Boost-free C++ parser
(include/model/qubo.hpp) and to load fixture instances in the tests.
"""
import re

_BLANK = " \t"


class QuboFormatError(ValueError):
    pass


def _skip(text, pos):
    while pos < len(text) and text[pos] in _BLANK:
        pos += 1
    return pos


def _eol_or_eoi(text, pos):
    """qi::eol | qi::eoi after the pre-skip; returns the new position or None."""
    pos = _skip(text, pos)
    if pos == len(text):
        return pos
    matched = False
    if pos < len(text) and text[pos] == "\r":
        pos += 1
        matched = True
    if pos < len(text) and text[pos] == "\n":
        pos += 1
        matched = True
    return pos if matched else None


def _uint(text, pos):
    m = re.compile(r"[0-9]+").match(text, pos)
    if not m or int(m.group()) > 0xFFFFFFFF:
        return None, pos
    return int(m.group()), m.end()


_REAL = re.compile(r"[+-]?(?:(?:[0-9]+(?:\.[0-9]*)?|\.[0-9]+)(?:[eE][+-]?[0-9]+)?|"
                   r"(?:inf(?:inity)?|nan(?:\([^)]*\))?))", re.IGNORECASE)


def _comment(text, pos):
    p = _skip(text, pos)
    if p >= len(text) or text[p] != "c":
        return None
    p += 1
    while True:  # *(qi::print) under the blank skipper
        p = _skip(text, p)
        if p < len(text) and 0x20 <= ord(text[p]) <= 0x7E:
            p += 1
        else:
            break
    return _eol_or_eoi(text, p)


def _header(text, pos, hdr):
    p = _skip(text, pos)
    if not text.startswith("p qubo", p):
        return None
    p += 6
    vals = []
    for _ in range(4):
        p = _skip(text, p)
        v, p2 = _uint(text, p)
        if v is None:
            return None
        vals.append(v)
        p = p2
    end = _eol_or_eoi(text, p)
    if end is None:
        return None
    hdr.update(max_nodes=vals[1], num_linear=vals[2], num_quadratic=vals[3])
    return end


def _coefficients(text, pos, add):
    p = pos
    idx = []
    for _ in range(2):  # index = lexeme[uint_ >> " "]
        p = _skip(text, p)
        v, p2 = _uint(text, p)
        if v is None or p2 >= len(text) or text[p2] != " ":
            return None
        idx.append(v)
        p = p2 + 1
    p = _skip(text, p)
    m = _REAL.match(text, p)
    if not m:
        return None
    add(idx[0], idx[1], float(m.group()))  # semantic action fires before the eol check
    return _eol_or_eoi(text, m.end())


def parse_qubo(text):
    """Returns (num_nodes, linear {i: v}, quadratic {(i, j): v}); raises QuboFormatError."""
    linear, quadratic, used, hdr = {}, {}, set(), {}

    def add(i, j, coef):  # QUBOBuilder::add_element, qubo.hpp:304-319
        if i == j:
            linear.setdefault(i, coef)
        elif i < j:
            quadratic.setdefault((i, j), coef)
        else:
            raise QuboFormatError("Incorrect file, encountered coefficient from lower triangle "
                                  "of QUBO matrix.")
        used.add(i)
        used.add(j)

    pos = 0
    while True:
        nxt = _comment(text, pos)
        if nxt is None or nxt == pos:
            break
        pos = nxt
    nxt = _header(text, pos, hdr)
    if nxt is not None:
        pos = nxt
    while pos < len(text):
        nxt = _comment(text, pos)
        if nxt is None:
            nxt = _coefficients(text, pos, add)
        if nxt is None or nxt == pos:
            break
        pos = nxt
    if _skip(text, pos) != len(text):  # phrase_parse post-skip, then first != last
        raise QuboFormatError("Parsing failed. Incorrect file format.")
    # QUBOBuilder::build_qubo, qubo.hpp:354-378
    if not linear and not quadratic:
        raise QuboFormatError("An empty input, no coefficients defined.")
    if not hdr:
        raise QuboFormatError("No header line or header line misformatted.")
    if hdr["num_linear"] > hdr["max_nodes"]:
        raise QuboFormatError("Number of linear terms is greater than num nodes.")
    if len(quadratic) != hdr["num_quadratic"]:
        raise QuboFormatError("Number of quadratic terms is not equal to the declared one.")
    if len(linear) != hdr["num_linear"]:
        raise QuboFormatError("Number of linear terms is not equal to the declared one.")
    return max(used) + 1, linear, quadratic


def load_qubo(path):
    with open(path) as f:
        return parse_qubo(f.read())


def ising_to_qubo(text):
    """benchmarks/annealing/scripts/convert_qbsolv_to_coo.py:23-37 (dimod change_vartype):
    1-indexed Ising `i i h` / `i j J` with s = 2x - 1 -> (n, linear, quadratic, offset)."""
    h, J = {}, {}
    for line in text.splitlines():
        if not line.strip() or line.startswith("#"):
            continue
        a, b, v = line.split()
        a, b, v = int(a) - 1, int(b) - 1, float(v)
        if a == b:
            h[a] = h.get(a, 0.0) + v
        else:
            J[(min(a, b), max(a, b))] = v
    nodes = set(h) | {i for k in J for i in k}
    linear = {i: 2.0 * h.get(i, 0.0) for i in nodes}
    quadratic = {}
    offset = -sum(h.values())
    for (i, j), v in J.items():
        quadratic[(i, j)] = 4.0 * v
        linear[i] -= 2.0 * v
        linear[j] -= 2.0 * v
        offset += v
    return max(nodes) + 1, linear, quadratic, offset
