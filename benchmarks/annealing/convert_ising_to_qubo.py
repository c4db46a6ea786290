#!/usr/bin/env python
"""Ising (COO text) -> `.qubo` converter for the annealing benchmark instances.

Does what /root/reference/benchmarks/annealing/scripts/convert_qbsolv_to_coo.py:23-37 does with
dimod (BQM.from_coo(vartype=SPIN) -> change_vartype(BINARY) -> relabel 1-based to 0-based ->
to_qubo -> sorted `i j coef` lines under a `p qubo 0 N N couplers` header), without dimod:
with s = 2x - 1,

    h_i s_i       = 2 h_i x_i - h_i
    J_ij s_i s_j  = 4 J_ij x_i x_j - 2 J_ij x_i - 2 J_ij x_j + J_ij

so linear_i = 2 h_i - 2 sum_j J_ij, quadratic_ij = 4 J_ij and the constant sum(J) - sum(h) is
dropped (the reference's plot.py:31 zeroes it too).  Input lines: `i i h` / `i j J`, 1-based,
`#` comments and blank lines ignored.
"""
import argparse


def read_ising(text):
    """-> (h, J) with 0-based indices; J keyed (min, max); repeated entries accumulate."""
    h, coupling = {}, {}
    for line in text.splitlines():
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        a, b, v = line.split()
        a, b, v = int(a) - 1, int(b) - 1, float(v)
        if a < 0 or b < 0:
            raise ValueError("Ising indices are 1-based")
        if a == b:
            h[a] = h.get(a, 0.0) + v
        else:
            key = (min(a, b), max(a, b))
            coupling[key] = coupling.get(key, 0.0) + v
    return h, coupling


def ising_to_qubo(h, coupling):
    """-> (num_variables, linear {i: a_i}, quadratic {(i, j): b_ij, i < j}, offset)."""
    nodes = set(h) | {i for key in coupling for i in key}
    if not nodes:
        raise ValueError("empty Ising instance")
    linear = {i: 2.0 * h.get(i, 0.0) for i in nodes}
    quadratic = {}
    offset = -sum(h.values())
    for (i, j), v in coupling.items():
        quadratic[(i, j)] = 4.0 * v
        linear[i] -= 2.0 * v
        linear[j] -= 2.0 * v
        offset += v
    return len(nodes), linear, quadratic, offset


def qubo_text(num_variables, linear, quadratic):
    """The file the reference script writes: header, then all terms sorted by (i, j)."""
    terms = {(i, i): v for i, v in linear.items()}
    terms.update(quadratic)
    lines = [f"p qubo 0 {num_variables} {num_variables} {len(quadratic)}\n"]
    lines += [f"{i} {j} {coef}\n" for (i, j), coef in sorted(terms.items())]
    return "".join(lines)


def ising_energy(h, coupling, spins):
    """Energy of a +-1 assignment (list indexed by 0-based variable)."""
    e = sum(v * spins[i] for i, v in h.items())
    e += sum(v * spins[i] * spins[j] for (i, j), v in coupling.items())
    return e


def main():
    parser = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    parser.add_argument("input_file", help="Ising instance (COO text, 1-based)")
    parser.add_argument("output_file", help="path of the .qubo file to write")
    args = parser.parse_args()
    with open(args.input_file) as f:
        h, coupling = read_ising(f.read())
    n, linear, quadratic, _ = ising_to_qubo(h, coupling)
    with open(args.output_file, "w") as f:
        f.write(qubo_text(n, linear, quadratic))


if __name__ == "__main__":
    main()
