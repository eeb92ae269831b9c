"""Times the Pauli-expectation passes on a resident state with CUDA events on the library's stream (16 ZZ terms in one pass,
one X term, a mixed Hamiltonian) and prints GB/s of the 16 B per amplitude each pass must read.
Usage: python scripts/expect_probe.py [qubits]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qclojure_b200 import _lib as L, circuits as C  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
zz = C.max_cut_hamiltonian(C.random_regular_graph(n, 3, seed=11), n)[:16]
x1 = [{"coefficient": 0.5, "pauli-string": "X" * 2 + "I" * (n - 2)}]
xy = [{"coefficient": 0.25, "pauli-string": "XY" + "Z" * 3 + "I" * (n - 5)}, {"coefficient": 0.5, "pauli-string": "XY" + "I" * (n - 2)}]
with L.StateVector(n) as sv:
    sv.apply_circuit(C.random_brickwork_circuit(n, 2))
    for name, H, passes in (("16 ZZ terms, one pass", zz, 1), ("1 X term", x1, 1), ("2 XY.. terms, one pass", xy, 1)):
        sv.expect_hamiltonian(H)
        sv.synchronize()
        reps = 5
        sv.timer_start()
        for _ in range(reps):
            e = sv.expect_hamiltonian(H)
        ms = sv.timer_stop() / reps
        print(f"{name}: {ms:.3f} ms  {16.0 * (1 << n) * passes / (ms * 1e-3) / 1e9:.0f} GB/s  energy {e:.12f}")
