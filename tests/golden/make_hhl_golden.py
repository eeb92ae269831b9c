#!/usr/bin/env python
"""Generate tests/golden/hhl_tutorial.json from the reference's rendered tutorial.

`/root/reference/doc/tutorial.md:6340-6520` records a JVM run of `(hhl/hhl-algorithm (sim/create-simulator) [[3 1] [1 2]] [7 5]
{:shots 10000})` and prints `:probability-results` with all 64 probabilities of the 6-qubit circuit that
`hhl/hhl-circuit matrix b 4 1` builds (application/algorithm/hhl.clj:625-714).  The circuit contains four `:cry` gates, so
this recorded output PINS the reference's controlled-gate convention (`apply-controlled-gate` applies the TRANSPOSED 2x2,
domain/gate.clj:473-483, i.e. CRY(theta) acts as controlled-RY(-theta)), which no reference test pins.
Runs only in the build container (the GPU box has no /root/reference); the JSON it writes is committed."""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
TUTORIAL = "/root/reference/doc/tutorial.md"


def main():
    t = open(TUTORIAL, encoding="utf-8").read()
    i = t.index("(get-in hhl-result [:execution-result :results :probability-results])")
    j = t.index(":all-probabilities", i)
    k = t.index("]", j)
    probs = [float(x) for x in re.findall(r"[-+]?\d+\.\d+(?:E-?\d+)?", t[j:k])]
    assert len(probs) == 64 and abs(sum(probs) - 1.0) < 1e-12
    line = t.count("\n", 0, i) + 1
    out = {"generator": "tests/golden/make_hhl_golden.py", "reference": f"doc/tutorial.md:{line}",
           "builder": "application/algorithm/hhl.clj:625-714 (hhl-circuit matrix b-vector 4 1)",
           "matrix": [[3, 1], [1, 2]], "vector": [7, 5], "precision_qubits": 4, "ancilla_qubits": 1,
           "all_probabilities": probs}
    with open(os.path.join(HERE, "hhl_tutorial.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote hhl_tutorial.json:", len(probs), "probabilities from tutorial.md line", line)


if __name__ == "__main__":
    main()
