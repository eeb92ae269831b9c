#!/usr/bin/env python
"""Generate tests/golden/tutorial_cases.json from the reference's rendered tutorial.

`/root/reference/doc/tutorial.md` holds JVM-produced result maps: every
`;; =>` block that contains a `:results {:final-state … :circuit {:operations …}}` map is a
recorded run of the reference's ideal or hardware simulator together with the exact circuit
it executed.  This script extracts (circuit, recorded outputs) pairs so that the oracle and the
CUDA path can replay the circuits without a JVM.  It only runs in the build container
(`/root/reference` is not present on the GPU box); the JSON it writes is committed.

Usage: python tests/golden/make_tutorial_golden.py
"""
from __future__ import annotations

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import edn  # noqa: E402

TUTORIAL = "/root/reference/doc/tutorial.md"


def jsonable(v):
    if isinstance(v, dict):
        return {str(k): jsonable(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [jsonable(x) for x in v]
    if isinstance(v, edn.Tagged):
        return repr(v)
    return v


def find_results(v, out, path=()):
    """Collect every map that has :final-state (with :state-vector) and a :circuit, either as a
    sibling (ideal simulator: inside :results) or one level up (hardware simulator: job level)."""
    if isinstance(v, dict):
        fs = v.get(":final-state")
        circ = v.get(":circuit")
        if isinstance(fs, dict) and ":state-vector" in fs and isinstance(circ, dict) and ":operations" in circ:
            out.append((path, v))
        res = v.get(":results")
        if (isinstance(res, dict) and isinstance(res.get(":final-state"), dict)
                and ":circuit" not in res and isinstance(circ, dict) and ":operations" in circ):
            merged = dict(res)
            merged[":circuit"] = circ
            out.append((path + (":results",), merged))
        for k, x in v.items():
            find_results(x, out, path + (k,))
    elif isinstance(v, (list, tuple)):
        for i, x in enumerate(v):
            find_results(x, out, path + (i,))


def main():
    text = open(TUTORIAL, encoding="utf-8").read()
    line_starts = [0]
    for i, ch in enumerate(text):
        if ch == "\n":
            line_starts.append(i + 1)

    def line_of(pos):
        import bisect
        return bisect.bisect_right(line_starts, pos)

    cases = []
    energies = []
    pos = 0
    nblocks = nfail = 0
    while True:
        k = text.find(";; =>", pos)
        if k < 0:
            break
        pos = k + 5
        nblocks += 1
        try:
            val, end = edn.read_from(text, pos)
        except Exception:
            nfail += 1
            continue
        found = []
        find_results(val, found)
        for path, res in found:
            circ = res[":circuit"]
            entry = {
                "source": f"doc/tutorial.md:{line_of(k)}",
                "path": [str(p) for p in path],
                "num_qubits": circ.get(":num-qubits"),
                "name": circ.get(":name"),
                "operations": jsonable(circ[":operations"]),
                "final_state": jsonable(res[":final-state"][":state-vector"]),
            }
            for key in (":measurement-results", ":probability-results", ":hamiltonian-result",
                        ":expectation-results", ":amplitude-results", ":state-vector-result",
                        ":result-types", ":trajectory-count", ":density-matrix-trace"):
                if key in res:
                    entry[key[1:].replace("-", "_")] = jsonable(res[key])
            # noisy runs: keep first trajectories (states) and the density matrix corner only
            if ":trajectories" in res:
                tr = res[":trajectories"]
                entry["trajectories"] = jsonable([t[":state-vector"] for t in tr[:100] if isinstance(t, dict)])
            if ":density-matrix" in res:
                entry["density_matrix"] = jsonable(res[":density-matrix"])
            parent = val
            # job-level fields live one level above :results
            if isinstance(val, dict):
                node = val
                for p in path[:-1]:
                    node = node[p] if isinstance(node, dict) else node[p]
                if isinstance(node, dict):
                    for key in (":shots-executed", ":execution-time-ms", ":job-status"):
                        if key in node:
                            entry[key[1:].replace("-", "_")] = jsonable(node[key])
            cases.append(entry)
        # variational runs: final circuit + Hamiltonian + optimal energy (+ optimiser history)
        if isinstance(val, dict) and ":optimal-energy" in (val.get(":optimization") or val) \
                and isinstance(val.get(":circuit"), dict):
            opt = val.get(":optimization") or val
            ham = (val.get(":config") or {}).get(":hamiltonian") or val.get(":problem-hamiltonian")
            energies.append({
                "source": f"doc/tutorial.md:{line_of(k)}",
                "algorithm": val.get(":algorithm"),
                "ansatz_type": val.get(":ansatz-type"),
                "num_qubits": val[":circuit"].get(":num-qubits"),
                "operations": jsonable(val[":circuit"][":operations"]),
                "hamiltonian": jsonable(ham),
                "optimal_energy": opt.get(":optimal-energy"),
                "optimal_parameters": jsonable(opt.get(":optimal-parameters")),
                "initial_parameters": jsonable(opt.get(":initial-parameters")),
                "history": jsonable([{"parameters": h.get(":parameters"), "energy": h.get(":energy"),
                                      "gradients": h.get(":gradients")}
                                     for h in (opt.get(":history") or [])]),
                "measurement_distribution": jsonable((val.get(":problem-solutions") or {}).get(":measurement-distribution")),
            })
        pos = end

    # de-duplicate identical (ops, final_state) pairs that the tutorial prints more than once
    uniq, seen = [], set()
    for c in cases:
        key = json.dumps([c["operations"], c["final_state"], c.get("measurement_results")], sort_keys=True)
        if key in seen:
            continue
        seen.add(key)
        uniq.append(c)

    out = os.path.join(HERE, "tutorial_cases.json")
    with open(out, "w") as f:
        json.dump({"generator": "tests/golden/make_tutorial_golden.py",
                   "reference": "lsolbach/qclojure doc/tutorial.md (JVM-produced outputs)",
                   "cases": uniq, "energy_cases": energies}, f, indent=1)
    print(f"blocks={nblocks} unparsed={nfail} results={len(cases)} unique={len(uniq)} -> {out}")
    for c in uniq:
        print(" ", c["source"], c["name"], "n=", c["num_qubits"], "ops=", len(c["operations"]),
              "keys=", [k for k in c if k not in ("operations", "final_state", "source", "path", "name", "num_qubits")])
    for e in energies:
        print("  energy", e["source"], e["algorithm"], e["ansatz_type"], "E=", e["optimal_energy"], "hist=", len(e["history"]))


if __name__ == "__main__":
    main()
