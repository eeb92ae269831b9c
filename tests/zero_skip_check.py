"""Hardware check of the EXPERIMENTAL zero-state support path (csrc/plan.h: Plan::support_in; sim.cu: QCB_ZERO_SKIP=1), which
the last session of round 2 could verify on the host emulator only.  Run on a B200:

    QCB_ZERO_SKIP=1 python tests/zero_skip_check.py            # parity with the oracle + timing against QCB_ZERO_SKIP=0

Compares every state with the oracle (NumPy oracle up to 16 qubits, C oracle at 24 and 28) for circuits applied straight after
qcb_set_zero, in several calls, with mid-circuit measurement, after qcb_set_state (the claim must be dropped), and prints the
device time of the 30-qubit benchmark circuit.  Exit code 0 = all comparisons within 1e-10."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle as CO                      # noqa: E402
from oracle import qc_oracle as O                      # noqa: E402
from qclojure_b200 import _lib as L                    # noqa: E402
from qclojure_b200 import circuits as C                # noqa: E402
from qclojure_b200 import ops as OPS                   # noqa: E402

TOL = 1e-10
bad = 0


def check(name, got, ref):
    global bad
    err = float(np.max(np.abs(got - ref)))
    print(f"{name}: max abs err {err:.2e}", flush=True)
    if not err <= TOL:
        bad += 1


def oracle(circ, init=None):
    n = circ["num-qubits"]
    if n >= 20:
        return CO.apply_circuit(circ, init)
    st = np.zeros(1 << n, dtype=np.complex128) if init is None else init.copy()
    if init is None:
        st[0] = 1
    for op in circ["operations"]:
        st = O.apply_gate_to_state(st, op)
    return st


print("QCB_ZERO_SKIP =", os.environ.get("QCB_ZERO_SKIP"))
for n, depth in ((5, 6), (12, 8), (13, 10), (16, 12), (20, 10), (24, 20), (28, 20)):
    circ = C.random_brickwork_circuit(n, depth)
    with L.StateVector(n) as sv:
        sv.set_zero(); sv.apply_circuit(circ)
        check(f"brickwork {n} q from |0>", sv.get_state(), oracle(circ))
        # the same circuit in two calls (the support is carried from one qcb_apply_ops to the next)
        ops = circ["operations"]
        sv.set_zero(); sv.apply_ops(OPS.encode_ops(ops[:len(ops) // 3])); sv.apply_ops(OPS.encode_ops(ops[len(ops) // 3:]))
        check(f"brickwork {n} q in two calls", sv.get_state(), oracle(circ))
        if n <= 20:
            # a dense state between set_zero and the circuit: the claim must be gone
            rng = np.random.default_rng(n)
            init = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
            init /= np.linalg.norm(init)
            sv.set_zero(); sv.set_state(init); sv.apply_circuit(circ)
            check(f"brickwork {n} q from a dense state", sv.get_state(), oracle(circ, init))
for n in (10, 14, 20):
    for name, circ in (("qft", C.quantum_fourier_transform_circuit(n)), ("ghz", C.ghz_state_circuit(n))):
        with L.StateVector(n) as sv:
            sv.set_zero(); sv.apply_circuit(circ)
            check(f"{name} {n} q", sv.get_state(), oracle(circ))
n = 30
enc = OPS.encode_ops(C.random_brickwork_circuit(n, 20)["operations"])
with L.StateVector(n) as sv:
    for _ in range(2):
        sv.set_zero(); sv.apply_ops(enc)
    sv.synchronize(); sv.timer_start()
    for _ in range(3):
        sv.set_zero(); sv.apply_ops(enc)
    ms = sv.timer_stop() / 3
    print(f"30 q depth-20 brickwork: {ms:.1f} ms per step, norm {sv.norm2():.15f}", flush=True)
print("FAILED" if bad else "ok")
sys.exit(1 if bad else 0)
