"""The C restatement (oracle/qc_oracle.c) must agree with the golden-pinned NumPy oracle."""
import math

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import qc_oracle as O
from qclojure_b200 import circuits as C


def _all_gates_circuit(n, rng):
    c = C.create_circuit(n)
    one = ["x", "y", "z", "h", "s", "s-dag", "t", "t-dag"]
    for _ in range(60):
        k = rng.integers(0, 9)
        qs = rng.permutation(n)[:3].tolist()
        a = float(rng.random() * 2 * math.pi)
        if k == 0:
            C.add_gate(c, one[rng.integers(0, len(one))], target=qs[0])
        elif k == 1:
            C.add_gate(c, ["rx", "ry", "rz", "phase"][rng.integers(0, 4)], target=qs[0], angle=a)
        elif k == 2:
            C.add_gate(c, ["cnot", "cz", "rydberg-cz"][rng.integers(0, 3)], control=qs[0], target=qs[1])
        elif k == 3:
            C.add_gate(c, ["crx", "cry", "crz", "rydberg-cphase"][rng.integers(0, 4)], control=qs[0], target=qs[1], angle=a)
        elif k == 4:
            C.add_gate(c, ["swap", "iswap"][rng.integers(0, 2)], qubit1=qs[0], qubit2=qs[1])
        elif k == 5:
            C.toffoli(c, qs[0], qs[1], qs[2])
        elif k == 6:
            C.fredkin(c, qs[0], qs[1], qs[2])
        elif k == 7:
            C.add_gate(c, "rydberg-blockade", qubit_indices=qs, angle=a)
        else:
            g = ["global-h", "global-x", "global-y", "global-z", "global-rx", "global-ry", "global-rz"][rng.integers(0, 7)]
            C.add_gate(c, g, angle=a)
    return c


@pytest.mark.parametrize("n", [3, 5, 8])
def test_c_oracle_matches_numpy_oracle_all_gates(n):
    rng = np.random.default_rng(100 + n)
    circ = _all_gates_circuit(n, rng)
    init = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    init /= np.linalg.norm(init)
    want = O.execute_circuit(circ, init)
    got = CO.apply_circuit(circ, init)
    assert np.max(np.abs(got - want)) <= 1e-12


def test_c_oracle_configs_qft_ghz_brickwork():
    for circ in (C.quantum_fourier_transform_circuit(10), C.ghz_state_circuit(10), C.random_brickwork_circuit(10, 6)):
        assert np.max(np.abs(CO.apply_circuit(circ) - O.execute_circuit(circ))) <= 1e-12


def test_c_oracle_sampling_and_expectation():
    circ = C.random_brickwork_circuit(10, 5)
    st = O.execute_circuit(circ)
    u = np.random.default_rng(1).random(500)
    assert np.array_equal(CO.sample(st, u), O.sample_outcomes(st, u))
    for ps in ("ZZIIXIYIIZ", "IIIIIIIIII", "XYZXYZXYZX", "YYIIIIIIII"):
        assert abs(CO.expect_pauli(st, ps) - O.pauli_string_expectation(ps, st)) <= 1e-12
    assert abs(CO.norm2(st) - 1.0) <= 1e-12


def test_c_oracle_dense_kron_reference_form():
    rng = np.random.default_rng(2)
    st = rng.standard_normal(64) + 1j * rng.standard_normal(64)
    U = O.rx_gate(0.3)
    for q in range(6):
        assert np.max(np.abs(CO.apply_1q_dense_kron(st, q, U) - O.apply_single_qubit_gate(st, U, q))) <= 1e-13


def test_c_oracle_rejects_unknown_gates_like_reference():
    c = C.add_gate(C.create_circuit(2), "cy", control=0, target=1)
    with pytest.raises(RuntimeError):
        CO.apply_circuit(c)
