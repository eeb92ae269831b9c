"""Times the marginal histogram (measure-specific-qubits) on a resident state for a few choices of measured qubits and checks
each against the probabilities summed on the host at a size where that is cheap.
Usage: python scripts/marginal_probe.py [qubits]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qclojure_b200 import _lib as L, circuits as C  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
# parity at 22 qubits against NumPy
m = 22
with L.StateVector(m) as sv:
    sv.apply_circuit(C.random_brickwork_circuit(m, 4))
    probs = np.abs(sv.get_state()) ** 2
    for qs in ([0], [m - 1], [0, 5, m - 1], [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14], [m - 1, m - 2, m - 3, m - 7, m - 9], list(range(12, 22))):
        got = sv.marginal_probabilities(qs)
        idx = np.arange(1 << m)
        key = np.zeros(1 << m, dtype=np.int64)
        for i, q in enumerate(qs):        # reference order (state.clj:961-963): bit i of the outcome <-> i-th listed qubit
            key |= ((idx >> (m - 1 - q)) & 1) << i
        want = np.bincount(key, weights=probs, minlength=1 << len(qs))
        err = float(np.max(np.abs(got - want)))
        print(f"parity {m} q, qubits {qs}: max|err| {err:.2e}")
        assert err < 1e-12, err
with L.StateVector(n) as sv:
    sv.apply_circuit(C.random_brickwork_circuit(n, 2))
    for qs in ([0, 5, n - 1], [0], [n - 8, n - 9], list(range(12))):
        sv.marginal_probabilities(qs)
        sv.synchronize()
        reps = 5
        sv.timer_start()
        for _ in range(reps):
            sv.marginal_probabilities(qs)
        ms = sv.timer_stop() / reps
        print(f"marginal of qubits {qs if len(qs) < 6 else str(len(qs)) + ' qubits'}: {ms:.3f} ms  {16.0 * (1 << n) / (ms * 1e-3) / 1e9:.0f} GB/s")
