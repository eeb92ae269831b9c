"""Device time of the QFT circuit (application/algorithm/quantum_fourier_transform.clj:34-62) on a resident state with far phases on
and off (QCB_FAR_PHASE is read when a handle is created), plus a closed-form check: QFT of |0...0> is the uniform superposition.
Usage: python scripts/qft_probe.py [qubits ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qclojure_b200 import _lib as L, circuits as C, ops as OPS  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [26, 28, 30]:
    circ = C.quantum_fourier_transform_circuit(n)
    enc = OPS.encode_ops(circ["operations"])
    for far in ("0", "1"):
        os.environ["QCB_FAR_PHASE"] = far
        with L.StateVector(n) as sv:
            for _ in range(2):
                sv.set_zero(); sv.apply_ops(enc)
            sv.synchronize()
            sv.timer_start()
            reps = 3
            for _ in range(reps):
                sv.set_zero(); sv.apply_ops(enc)
            ms = sv.timer_stop() / reps
            st = sv.stats()
            amps = sv.get_amplitudes([0, 1, (1 << n) - 1, 12345 % (1 << n)])
            err = float(np.max(np.abs(amps - 2.0 ** (-n / 2))))
        print(f"QFT-{n} far phases {far}: {ms:8.2f} ms  sweeps {st['n_sweeps']} rounds {st['n_rounds']}  {len(circ['operations']) / ms * 1e3:7.0f} gates/s  "
              f"|amp - 2^(-n/2)| <= {err:.1e}")
