"""Launches every streaming reduction once on a 28-qubit state (4 GiB) so that `ncu --metrics gpu__time_duration.sum,...`
can attribute a duration and DRAM traffic to each (roofline of the reductions)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qclojure_b200 import _lib as L, circuits as C  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
u = np.random.default_rng(1).random(256)
H = C.max_cut_hamiltonian(C.random_regular_graph(n, 3, seed=11), n)[:16] + [{"coefficient": 0.5, "pauli-string": "X" * 2 + "I" * (n - 2)}]
with L.StateVector(n) as sv:
    sv.apply_circuit(C.random_brickwork_circuit(n, 4))
    for _ in range(2):
        print("norm", sv.norm2())
        print("energy", sv.expect_hamiltonian(H))
        print("1q", sv.expect_1q(np.array([[0, 1], [1, 0]]), 3))
        print("shots", sv.sample(u)[:4])
        print("marg", sv.marginal_probabilities([0, 5, n - 1])[:2])
        sv.normalize()
        print("probs", sv.probabilities(0, 1 << 20)[:2])
