"""Initial states: the state constructors of the reference, `src/org/soulspace/qclojure/domain/state.clj`,
restated for the Python mirror so that `:initial-state` options can be built the way QClojure callers build them.  A state
is `{"state-vector": complex128 array, "num-qubits": n}`; index i = sum b_q 2^(n-1-q), qubit 0 is the MSB (state.clj:114-162).
Caller-side helpers on small host vectors: the device entry points for the common cases are `qcb_set_zero`,
`qcb_set_basis` and `qcb_set_state` (no 2^n host vector is needed for |0...0> or a basis state).

The multi-qubit variants follow the reference CODE literally, including two quirks: `(minus-state n)`, `(plus-i-state n)`
and `(minus-i-state n)` for n > 1 are NOT tensor powers of the single-qubit state - the first amplitude is +a and all
the others are -a / +ia / -ia (state.clj:377-385, 413-421, 449-457)."""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np


def _state(vec, n) -> dict:
    return {"state-vector": np.asarray(vec, dtype=np.complex128), "num-qubits": int(n)}


def _vec(state) -> np.ndarray:
    return np.asarray(state["state-vector"] if isinstance(state, dict) else state, dtype=np.complex128).reshape(-1)


def multi_qubit_state(amplitudes) -> dict:
    """state.clj:221-249: num-qubits = max(1, log2int(count))."""
    a = np.asarray(amplitudes, dtype=np.complex128).reshape(-1)
    return _state(a, max(1, int(math.log2(a.shape[0])) if a.shape[0] > 0 else 1))


def zero_state(n: int = 1) -> dict:
    """state.clj:251-284."""
    v = np.zeros(1 << n, dtype=np.complex128)
    v[0] = 1.0
    return _state(v, n)


def one_state(n: int = 1) -> dict:
    """state.clj:286-313: |1...1>."""
    v = np.zeros(1 << n, dtype=np.complex128)
    v[-1] = 1.0
    return _state(v, n)


def _first_then(n: int, rest: complex) -> dict:
    size = 1 << n
    a = 1.0 / math.sqrt(size)
    v = np.full(size, rest * a, dtype=np.complex128)
    v[0] = a
    return _state(v, n)


def plus_state(n: int = 1) -> dict:
    """state.clj:315-348: all amplitudes 1/sqrt(2^n)."""
    return _first_then(n, 1.0)


def minus_state(n: int = 1) -> dict:
    """state.clj:350-385: [a, -a, -a, ...] (for n = 1 the usual |->)."""
    return _first_then(n, -1.0)


def plus_i_state(n: int = 1) -> dict:
    """state.clj:387-421: [a, ia, ia, ...]."""
    return _first_then(n, 1j)


def minus_i_state(n: int = 1) -> dict:
    """state.clj:423-457: [a, -ia, -ia, ...]."""
    return _first_then(n, -1j)


def bits_to_index(bits: Sequence[int]) -> int:
    """state.clj:114-136."""
    idx = 0
    for b in bits:
        idx = (idx << 1) | (int(b) & 1)
    return idx


def index_to_bits(index: int, n: int):
    """state.clj:138-162."""
    return [(index >> (n - 1 - q)) & 1 for q in range(n)]


def computational_basis_state(n: int, bits: Sequence[int]) -> dict:
    """state.clj:484-519."""
    if len(bits) != n or any(b not in (0, 1) for b in bits):
        raise ValueError("bits must be n values of 0 / 1")
    v = np.zeros(1 << n, dtype=np.complex128)
    v[bits_to_index(bits)] = 1.0
    return _state(v, n)


def normalize_state(state: dict, tolerance: float = 1e-12) -> dict:
    """state.clj:524-551: divide by the 2-norm only if it is positive and above the tolerance."""
    v = _vec(state)
    norm = float(np.sqrt(np.sum(np.abs(v) ** 2)))
    out = dict(state)
    out["state-vector"] = v * (1.0 / norm) if norm > 0 and norm > tolerance else v
    return out


def tensor_product(state1: dict, state2: dict) -> dict:
    """state.clj:553-592: state1's qubits become the more significant ones."""
    return _state(np.kron(_vec(state1), _vec(state2)), state1["num-qubits"] + state2["num-qubits"])
