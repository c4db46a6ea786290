#!/usr/bin/env python
"""Distance of the sweep's best energies from the known ground state.

The analysis of /root/reference/benchmarks/annealing/plot.py:19-52 as a table: the ground state
comes from a `groundstates_TN.txt` line (`001.txt : <ising energy> s1 s2 ...`, spins +-1), its
QUBO energy is evaluated on the instance converted from the Ising original with the constant
dropped (plot.py:30-31 does the same through dimod), and for every (num_tries, schedule,
num_iter) the minimum energy of the merged CSV minus that ground energy is reported.

plot.py:33-35 feeds the +-1 spins into the BINARY model unconverted, which evaluates a different
state; here the spins are mapped x = (s + 1) / 2 first.  --reference-quirk reproduces plot.py.
"""
import argparse
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from convert_ising_to_qubo import ising_to_qubo, read_ising  # noqa: E402


def read_ground_state(path, instance_name):
    """-> (ising_energy, [spins]) of the line that starts with `instance_name`."""
    with open(path) as f:
        for line in f:
            name, _, rest = line.partition(":")
            if name.strip() == instance_name:
                fields = rest.split()
                return float(fields[0]), [int(s) for s in fields[1:]]
    raise KeyError(f"{instance_name} not found in {path}")


def qubo_energy(linear, quadratic, x):
    e = sum(v * x[i] for i, v in linear.items())
    e += sum(v * x[i] * x[j] for (i, j), v in quadratic.items())
    return e


def read_results(path):
    """Merged sweep CSV -> list of dict(energy, beta_min, num_iter, num_tries, schedule)."""
    rows = []
    with open(path) as f:
        reader = csv.reader(f)
        header = next(reader)
        col = {name: header.index(name) for name in
               ("energy", "beta_min", "num_iter", "num_tries", "schedule")}
        for r in reader:
            if not r or r[0] == header[0] and r[col["energy"]] == "energy":
                continue
            rows.append({"energy": float(r[col["energy"]]), "beta_min": r[col["beta_min"]],
                         "num_iter": int(r[col["num_iter"]]), "num_tries": int(r[col["num_tries"]]),
                         "schedule": r[col["schedule"]]})
    return rows


def distance_table(rows, ground_energy):
    """{(num_tries, schedule): {num_iter: min energy - ground}} (plot.py:41-45 groups the same)."""
    table = {}
    for r in rows:
        cell = table.setdefault((r["num_tries"], r["schedule"]), {})
        d = r["energy"] - ground_energy
        cell[r["num_iter"]] = min(cell.get(r["num_iter"], d), d)
    return table


def main():
    parser = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    parser.add_argument("--results", required=True, help="merged CSV written by one-solver-sweep")
    parser.add_argument("--ising", required=True, help="Ising original of the instance")
    parser.add_argument("--ground-states", required=True, help="groundstates_TN.txt")
    parser.add_argument("--instance", default="001.txt", help="line of the ground-state file")
    parser.add_argument("--num-tries", type=int, nargs="*", help="only these trajectory counts")
    parser.add_argument("--reference-quirk", action="store_true",
                        help="use the +-1 spins as if they were 0/1 values, like plot.py")
    args = parser.parse_args()

    with open(args.ising) as f:
        h, coupling = read_ising(f.read())
    _, linear, quadratic, offset = ising_to_qubo(h, coupling)
    e_ising, spins = read_ground_state(args.ground_states, args.instance)
    x = spins if args.reference_quirk else [(s + 1) // 2 for s in spins]
    ground = qubo_energy(linear, quadratic, x)
    print(f"ground state: Ising energy {e_ising}, QUBO energy {ground:.6f} "
          f"(QUBO + dropped constant = {ground + offset:.6f})")
    table = distance_table(read_results(args.results), ground)
    iters = sorted({it for cell in table.values() for it in cell})
    print("num_tries,schedule," + ",".join(str(it) for it in iters))
    for (tries, schedule) in sorted(table):
        if args.num_tries and tries not in args.num_tries:
            continue
        cell = table[(tries, schedule)]
        print(f"{tries},{schedule}," + ",".join(
            f"{cell[it]:.6g}" if it in cell else "" for it in iters))


if __name__ == "__main__":
    main()
