"""Test infrastructure: an oracle-backed stand-in for `qclojure_b200._lib.StateVector` and for the P2 backend, so that the
host-side mirror logic (result extraction, job layer) can be exercised on a machine without a GPU.  It is only ever
installed by tests through monkeypatching; the product never imports it (`test_product_never_imports_the_oracle`)."""
import numpy as np

from oracle import qc_oracle as O


class FakeStateVector:
    def __init__(self, n_qubits, **_kw):
        self.n = int(n_qubits)
        self.local_count = 1 << self.n
        self.state = O.zero_state(self.n)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        pass

    def set_zero(self):
        self.state = O.zero_state(self.n)

    def set_state(self, amps):
        self.state = np.array(amps, dtype=np.complex128).reshape(-1)
        assert self.state.shape[0] == 1 << self.n

    def get_state(self, offset=0, count=None):
        return self.state[offset: None if count is None else offset + count].copy()

    def get_amplitudes(self, indices):
        return self.state[np.asarray(indices, dtype=np.int64)].copy()

    def apply_ops(self, ops):
        for op in ops:
            self.state = O.apply_gate_to_state(self.state, op)
        return self

    def probabilities(self, offset=0, count=None):
        return O.measurement_probabilities(self.state)[offset: None if count is None else offset + count]

    def sample(self, uniforms):
        return O.sample_outcomes(self.state, np.asarray(uniforms, dtype=np.float64))

    def measure_qubits(self, qubits, u):
        bits, self.state, probs = O.measure_specific_qubits(self.state, list(qubits), u)
        return bits, float(probs[sum(b << i for i, b in enumerate(bits))])

    def expect_1q(self, observable, target):
        return O.expectation_1q(self.state, np.asarray(observable, dtype=np.complex128).reshape(2, 2), int(target))

    def expect_pauli(self, pauli):
        return O.pauli_string_expectation(pauli, self.state)

    def expect_hamiltonian(self, hamiltonian):
        return O.hamiltonian_expectation(hamiltonian, self.state)

    def fidelity(self, reference):
        return O.state_fidelity(np.asarray(reference, dtype=np.complex128), self.state)


class FakeLinearAlgebra:
    """NumPy stand-in for the device products of `linalg.B200ComplexBackend`; decompositions use the real host code of
    libqcb200 (`host_only=True`), which needs no GPU."""

    def __init__(self):
        from qclojure_b200 import linalg
        self._host = linalg.B200ComplexBackend(host_only=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self._host.close()

    def inner_product(self, x, y):
        return np.vdot(x, y)

    def outer_product(self, x, y):
        return np.outer(x, np.conj(y))

    def matrix_vector_product(self, A, x):
        return np.asarray(A) @ np.asarray(x)

    def matrix_multiply(self, A, B):
        return np.asarray(A) @ np.asarray(B)

    def trace(self, A):
        return np.trace(A)

    def eigen_hermitian(self, A):
        return self._host.eigen_hermitian(A)
