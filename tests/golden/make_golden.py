#!/usr/bin/env python
"""Regenerates tests/golden/* from the read-only reference checkout at /root/reference.

Only DATA is taken from the reference: the instance files its benchmarks ship, the
known-answer values embedded in its unit tests, and (state, energy) rows its CLI produced.
No reference source code is copied.  Run once in the build container; the outputs are
committed because /root/reference does not exist on the GPU box.
"""
import csv
import json
import os
import re
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def c_string_literals(block):
    """Concatenate the adjacent C string literals found in `block`."""
    parts = re.findall(r'"((?:[^"\\]|\\.)*)"', block)
    return "".join(bytes(p, "utf-8").decode("unicode_escape") for p in parts)


def main():
    golden = {}

    # ---- tests/exhaustive_test.cpp:12-142 : five instances with ground energies (1e-13)
    src = open(os.path.join(REF, "tests/exhaustive_test.cpp")).read()
    pairs = re.findall(r"std::make_pair<std::string, double>\((.*?),\s*(-?[0-9.]+)\)", src, re.S)
    golden["exhaustive_test"] = [{"qubo": c_string_literals(text), "energy": float(e)}
                                 for text, e in pairs]
    assert len(golden["exhaustive_test"]) == 5

    # ---- tests/io_test.cpp:12-37 : the valid file and the five malformed ones
    src = open(os.path.join(REF, "tests/io_test.cpp")).read()
    valid = re.search(r"qubo_file_contents\((.*?)\);", src, re.S).group(1)
    golden["io_valid"] = c_string_literals(valid)
    bad_block = re.search(r"grammatically_incorrect_file_contents\[\]\{(.*?)\};", src, re.S).group(1)
    golden["io_malformed"] = [c_string_literals(b) for b in re.findall(r"std::string\((.*?)\)", bad_block, re.S)]
    assert len(golden["io_malformed"]) == 5
    golden["io_lower_triangle"] = "p qubo 0 100 2 2\n0 0 -0.5\n1 0 2.0\n1 2 4\n2 2 -0.7\n"   # io_test.cpp:72-76
    golden["io_count_mismatch"] = "p qubo 0 100 3 3\n0 0 -0.5\n0 1 2.0\n1 2 4\n2 2 -0.7\n"  # io_test.cpp:95-99
    golden["io_solution_csv"] = {"state": [0, 1, 1, 0, 1], "energy": -12.5,
                                 "text": "0,1,2,3,4,energy\n0,1,1,0,1,-12.5\n"}             # io_test.cpp:108-118

    # ---- tests/qubo_helpers_test.cpp:14-29 : flatten layout
    golden["flatten"] = {
        "n": 5,
        "linear": {"0": 0.5, "1": -2.0, "2": 1.0, "4": -1.5},
        "quadratic": [[0, 1, 1.0], [0, 3, 7.2], [1, 4, -1.0], [2, 3, 2.0], [2, 4, -1.5], [3, 4, -3.5]],
        "expected": [0.5, 1.0, 0.0, 7.2, 0.0, 1.0, -2.0, 0.0, 0.0, -1.0, 0.0, 0.0, 1.0,
                     2.0, -1.5, 7.2, 0.0, 2.0, 0.0, -3.5, 0.0, -1.0, -1.5, -3.5, -1.5],
    }
    # ---- tests/qubo_test.cpp:54
    golden["str"] = "QUBO model 1--1:10 1--2:20"

    # ---- benchmarks/exhaustive_search/examples/*.qubo : instance files -> examples/
    ex_dir = os.path.join(ROOT, "examples")
    os.makedirs(ex_dir, exist_ok=True)
    for name in sorted(os.listdir(os.path.join(REF, "benchmarks/exhaustive_search/examples"))):
        shutil.copy(os.path.join(REF, "benchmarks/exhaustive_search/examples", name),
                    os.path.join(ex_dir, name))
    # brute-forced optima of the shipped examples
    golden["examples_ground"] = {
        "simple.qubo": {"energy": -2.0, "state": [0, 0, 1, 1]},
        "test1.qubo": {"energy": -12.0, "state": [1, 1, 0, 1]},
        "test2.qubo": {"energy": -1.2, "state": [1, 0, 1]},
        "csp5.qubo": {"energy": -22.0, "state": [0, 1, 1, 1, 0]},
        "csp7.qubo": {"energy": -14.0, "state": [1, 1, 1, 0, 0, 1, 0]},
        "csp13.qubo": {"energy": -32.0, "state": [1, 1, 1, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0]},
    }

    # ---- Chimera droplet instances 001 (qbsolv format + Ising original) and a sample of the
    #      (state, energy) rows the reference CLI produced for them
    for size in ("128", "512"):
        d = os.path.join(HERE, f"chimera{size}")
        os.makedirs(d, exist_ok=True)
        shutil.copy(os.path.join(REF, f"benchmarks/annealing/chimera_droplets_qbsolv/{size}power/001.txt"),
                    os.path.join(d, "001.qubo"))
        shutil.copy(os.path.join(REF, f"benchmarks/annealing/chimera_droplets/{size}power/001.txt"),
                    os.path.join(d, "001.ising.txt"))
        rows = []
        with open(os.path.join(REF, f"benchmarks/annealing/results/{size}/001.csv")) as f:
            rd = csv.reader(f)
            header = next(rd)
            n = int(size)
            all_rows = [r for r in rd if r and r[0] != "0" or len(r) > 1]
            for k, r in enumerate(all_rows):
                if r[:3] == header[:3]:
                    continue
                if k % 24 == 0:  # 1200 rows -> 50 samples
                    rows.append({"state": "".join(r[:n]), "energy": float(r[n]), "beta_min": r[n + 1],
                                 "num_iter": int(r[n + 2]), "num_tries": int(r[n + 3]),
                                 "schedule": r[n + 4]})
        with open(os.path.join(d, "reference_results_sample.json"), "w") as f:
            json.dump(rows, f)
        line = open(os.path.join(REF, f"benchmarks/annealing/results/{size}/groundstates_TN.txt")).readline()
        name, rest = line.split(":")
        vals = rest.split()
        with open(os.path.join(d, "groundstate_TN.json"), "w") as f:
            json.dump({"file": name.strip(), "ising_energy": float(vals[0]),
                       "spins": [int(v) for v in vals[1:]]}, f)

    for base, _, files in os.walk(ROOT):
        if base.startswith(os.path.join(ROOT, "examples")) or base.startswith(HERE):
            for name in files:
                os.chmod(os.path.join(base, name), 0o644)
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(golden, f, indent=1)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
