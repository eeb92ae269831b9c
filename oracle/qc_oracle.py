"""CPU oracle: a NumPy restatement of the reference's state-vector hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `qclojure_b200/` imports this module; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it,
and only as the checker (never as the thing shipped or measured as the product).

The reference (lsolbach/qclojure) is pure Clojure and cannot run in this image (no JVM), so this
file restates the reference's *semantics* function by function, each citing the `file:line` it
follows (paths relative to /root/reference/src/org/soulspace/qclojure/).  The arithmetic is the
pairwise form of the reference's dense mat-vec: for a 2x2 gate the only non-zero terms of a row of
the Kronecker-expanded matrix are the two pair partners, added in ascending index order with the
naive complex multiply (ac-bd, ad+bc) — bit-identical to `fastmath` Vec2 `mult`/`add`.

Parity pinning: `tests/test_oracle_golden.py` checks this oracle against the JVM-produced outputs
recorded in the reference's `doc/tutorial.md` (fixtures in tests/golden/tutorial_cases.json) and
against the known-answer tests of the reference's own test-suite.

Conventions (domain/state.clj:114-162): qubit 0 is the MOST significant bit of the amplitude index,
i = sum_q b_q * 2^(n-1-q); bitstrings are printed MSB-first.  The one exception is SWAP/iSWAP, whose
`qubit1`/`qubit2` are bit positions counted from the LSB (domain/gate.clj:768-778, 823-833).

Circuits are plain dicts mirroring the reference's maps (domain/circuit.clj:30-37) with keyword
names spelled as strings without the colon:
  {"num-qubits": n, "operations": [{"operation-type": "h", "operation-params": {"target": 0}}, ...]}
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

TOLERANCE = 1e-12  # domain/math/complex_linear_algebra.clj:22 (*tolerance*)

# --------------------------------------------------------------------------------------
# bit / index utilities — domain/state.clj:114-162, 43-69
# --------------------------------------------------------------------------------------


def bits_to_index(bits: Sequence[int]) -> int:
    """state.clj:114-136 — MSB-first: [1 0 1] -> 5."""
    i = 0
    for b in bits:
        i = (i << 1) | (int(b) & 1)
    return i


def index_to_bits(index: int, n: int) -> List[int]:
    """state.clj:138-162."""
    return [(index >> (n - 1 - q)) & 1 for q in range(n)]


def basis_string(index: int, n: int) -> str:
    """state.clj:43-69 — MSB-first bitstring of a basis-state index."""
    return "".join(str(b) for b in index_to_bits(index, n))


# --------------------------------------------------------------------------------------
# states — domain/state.clj:251-284, 484-519, 524-551
# --------------------------------------------------------------------------------------


def zero_state(n: int) -> np.ndarray:
    s = np.zeros(1 << n, dtype=np.complex128)
    s[0] = 1.0
    return s


def computational_basis_state(n: int, bits: Sequence[int]) -> np.ndarray:
    s = np.zeros(1 << n, dtype=np.complex128)
    s[bits_to_index(bits)] = 1.0
    return s


def norm2(state: np.ndarray) -> float:
    """cla/norm2 — sqrt(sum |a|^2) (math/fastmath/complex_linear_algebra.clj norm2)."""
    return math.sqrt(float(np.sum(state.real * state.real + state.imag * state.imag)))


def normalize_state(state: np.ndarray) -> np.ndarray:
    """state.clj:544-551 — divide by the 2-norm only when norm > tolerance (1e-12)."""
    nrm = norm2(state)
    if nrm > 0 and nrm > TOLERANCE:
        return state * (1.0 / nrm)
    return state


def num_qubits_of(state: np.ndarray) -> int:
    n = int(round(math.log2(state.shape[0])))
    assert (1 << n) == state.shape[0]
    return n


# --------------------------------------------------------------------------------------
# gate matrices — domain/gate.clj:38-283
# --------------------------------------------------------------------------------------

_S2 = 1.0 / math.sqrt(2.0)
PAULI_I = np.array([[1, 0], [0, 1]], dtype=np.complex128)
PAULI_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
PAULI_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
PAULI_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
HADAMARD = np.array([[_S2, _S2], [_S2, -_S2]], dtype=np.complex128)  # gate.clj:96-107
S_GATE = np.array([[1, 0], [0, 1j]], dtype=np.complex128)
S_DAG_GATE = np.array([[1, 0], [0, -1j]], dtype=np.complex128)


def phase_gate(phi: float) -> np.ndarray:
    """gate.clj:109-137 — diag(1, cos phi + i sin phi)."""
    return np.array([[1, 0], [0, complex(math.cos(phi), math.sin(phi))]], dtype=np.complex128)


T_GATE = phase_gate(math.pi / 4)       # gate.clj t-gate  = (phase-gate (/ PI 4))
T_DAG_GATE = phase_gate(math.pi / -4)  # gate.clj t-dag-gate = (phase-gate (/ PI -4))


def rx_gate(theta: float) -> np.ndarray:
    """gate.clj:198-222 — [[c, -i s], [-i s, c]]."""
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[complex(c, 0), complex(0, -s)], [complex(0, -s), complex(c, 0)]], dtype=np.complex128)


def ry_gate(theta: float) -> np.ndarray:
    """gate.clj:224-250 — [[c, -s], [s, c]]."""
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, -s], [s, c]], dtype=np.complex128)


def rz_gate(theta: float) -> np.ndarray:
    """gate.clj:252-283 — diag(e^{-i theta/2}, e^{+i theta/2}) with cos/sin of theta/-2 and theta/2."""
    en = complex(math.cos(theta / -2), math.sin(theta / -2))
    ep = complex(math.cos(theta / 2), math.sin(theta / 2))
    return np.array([[en, 0], [0, ep]], dtype=np.complex128)


# --------------------------------------------------------------------------------------
# gate application — domain/gate.clj:323-486, 752-969, 1028-1253
# --------------------------------------------------------------------------------------


def _view(state: np.ndarray, n: int, q: int) -> np.ndarray:
    """View the state as [hi, 2, lo] with the middle axis = bit (n-1-q) (reference qubit q)."""
    return state.reshape(1 << q, 2, 1 << (n - 1 - q))


def apply_single_qubit_gate(state: np.ndarray, U: np.ndarray, q: int) -> np.ndarray:
    """gate.clj:346-353, 384-395 — psi' = (I x..x U_q x..x I) psi, qubit 0 leftmost.

    Pairwise form of the dense mat-vec: new0 = U00*a0 + U01*a1, new1 = U10*a0 + U11*a1.
    """
    n = num_qubits_of(state)
    assert 0 <= q < n
    v = _view(state, n, q)
    out = np.empty_like(v)
    a0, a1 = v[:, 0, :], v[:, 1, :]
    out[:, 0, :] = U[0, 0] * a0 + U[0, 1] * a1
    out[:, 1, :] = U[1, 0] * a0 + U[1, 1] * a1
    return out.reshape(-1)


def apply_controlled_gate(state: np.ndarray, control: int, target: int, U: np.ndarray) -> np.ndarray:
    """gate.clj:451-486 — scatter-add form.  NOTE the reference uses U10 for the |1>->|0>
    contribution and U01 for |0>->|1> (lines 473-483), i.e. it applies U^T on the target when the
    control bit is 1.  Reproduced literally: new0 = U00*a0 + U10*a1 ; new1 = U01*a0 + U11*a1."""
    n = num_qubits_of(state)
    assert control != target
    idx = np.arange(state.shape[0], dtype=np.int64)
    cb = (idx >> (n - 1 - control)) & 1
    tb = (idx >> (n - 1 - target)) & 1
    out = state.copy()
    sel0 = np.nonzero((cb == 1) & (tb == 0))[0]
    sel1 = sel0 | (1 << (n - 1 - target))
    a0, a1 = state[sel0], state[sel1]
    out[sel0] = a0 * U[0, 0] + a1 * U[1, 0]
    out[sel1] = a0 * U[0, 1] + a1 * U[1, 1]
    return out


def cnot(state, control, target):
    """gate.clj:627-639."""
    return apply_controlled_gate(state, control, target, PAULI_X)


def controlled_z(state, control, target):
    """gate.clj:712-720."""
    return apply_controlled_gate(state, control, target, PAULI_Z)


def controlled_y(state, control, target):
    """gate.clj:722-729 (through apply-controlled-gate, hence controlled-(Y^T) = controlled-(-Y))."""
    return apply_controlled_gate(state, control, target, PAULI_Y)


def controlled_rx(state, control, target, theta):
    """gate.clj:655-667."""
    return apply_controlled_gate(state, control, target, rx_gate(theta))


def controlled_ry(state, control, target, theta):
    """gate.clj:669-681 — acts as controlled-RY(-theta) because of the transpose."""
    return apply_controlled_gate(state, control, target, ry_gate(theta))


def controlled_rz(state, control, target, theta):
    """gate.clj:683-696."""
    return apply_controlled_gate(state, control, target, rz_gate(theta))


def swap_gate(state: np.ndarray, qubit1: int, qubit2: int, phase: complex = 1.0) -> np.ndarray:
    """gate.clj:752-779 — new[i] = old[j], j = i with index bits `qubit1`,`qubit2` exchanged,
    where the bit positions are counted from the LSB (bit-shift-right i qubit1), unlike every other
    gate.  iswap (gate.clj:781-834) multiplies the moved amplitudes by i."""
    assert qubit1 != qubit2
    idx = np.arange(state.shape[0], dtype=np.int64)
    b1 = (idx >> qubit1) & 1
    b2 = (idx >> qubit2) & 1
    j = (idx & ~((1 << qubit1) | (1 << qubit2))) | (b2 << qubit1) | (b1 << qubit2)
    out = state[j]
    if phase != 1.0:
        diff = b1 != b2
        out = out.copy()
        out[diff] = phase * out[diff]
    return out


def iswap_gate(state, qubit1, qubit2):
    return swap_gate(state, qubit1, qubit2, phase=1j)


def toffoli_gate(state, control1, control2, target):
    """gate.clj:869-900 — flip bit(n-1-target) where both control bits are 1."""
    n = num_qubits_of(state)
    idx = np.arange(state.shape[0], dtype=np.int64)
    both = (((idx >> (n - 1 - control1)) & 1) == 1) & (((idx >> (n - 1 - control2)) & 1) == 1)
    src = np.where(both, idx ^ (1 << (n - 1 - target)), idx)
    return state[src]


def fredkin_gate(state, control, target1, target2):
    """gate.clj:935-969 — swap bits target1/target2 (MSB-first) where the control bit is 1."""
    n = num_qubits_of(state)
    idx = np.arange(state.shape[0], dtype=np.int64)
    p1, p2 = n - 1 - target1, n - 1 - target2
    c = ((idx >> (n - 1 - control)) & 1) == 1
    b1 = (idx >> p1) & 1
    b2 = (idx >> p2) & 1
    swapped = (idx & ~((1 << p1) | (1 << p2))) | (b2 << p1) | (b1 << p2)
    src = np.where(c, swapped, idx)
    return state[src]


def rydberg_cphase_gate(state, control, target, phi):
    """gate.clj:1028-1053 — e^{i phi} where both bits are 1."""
    n = num_qubits_of(state)
    idx = np.arange(state.shape[0], dtype=np.int64)
    both = (((idx >> (n - 1 - control)) & 1) == 1) & (((idx >> (n - 1 - target)) & 1) == 1)
    out = state.copy()
    out[both] = out[both] * complex(math.cos(phi), math.sin(phi))
    return out


def rydberg_blockade_gate(state, qubit_indices, phi):
    """gate.clj:1055-1095 — e^{i phi} where exactly one of the listed qubits is 1."""
    n = num_qubits_of(state)
    idx = np.arange(state.shape[0], dtype=np.int64)
    ones = np.zeros_like(idx)
    for q in qubit_indices:
        ones += (idx >> (n - 1 - q)) & 1
    out = state.copy()
    sel = ones == 1
    out[sel] = out[sel] * complex(math.cos(phi), math.sin(phi))
    return out


def _global(state, U):
    """gate.clj:1115-1253 — the same 1q gate on every qubit, qubit 0 first."""
    n = num_qubits_of(state)
    for q in range(n):
        state = apply_single_qubit_gate(state, U, q)
    return state


# --------------------------------------------------------------------------------------
# circuit execution — domain/circuit.clj:933-1111, 1738-1792
# --------------------------------------------------------------------------------------

GATE_ALIASES = {  # domain/operation_registry.clj:387-407
    "not": "x", "bit-flip": "x", "phase-flip": "z", "id": "i", "cx": "cnot", "ccx": "toffoli",
    "ccnot": "toffoli", "cswap": "fredkin", "p": "phase", "u1": "phase", "sdg": "s-dag",
    "tdg": "t-dag", "phaseshift": "phase", "si": "s-dag", "ti": "t-dag",
}


class UnknownGate(Exception):
    """circuit.clj:1072 — (throw (ex-info "Unknown gate type" ...))."""


def _kw(x):
    return x[1:] if isinstance(x, str) and x.startswith(":") else x


def normalize_op(op: dict) -> Tuple[str, dict]:
    """Accept both "h" and ":h" spellings (the golden fixtures keep the colon)."""
    typ = _kw(op.get("operation-type", op.get(":operation-type")))
    params = op.get("operation-params", op.get(":operation-params")) or {}
    params = {_kw(k): v for k, v in params.items()}
    return typ, params


def apply_gate_to_state(state: np.ndarray, op: dict, *, superset: bool = False) -> np.ndarray:
    """circuit.clj:952-1072 — `case` on the alias-resolved :operation-type.

    `superset=True` additionally accepts `:i` and `:cy`, for which the reference has builders
    (circuit.clj:225, 480) but no executor branch (it throws "Unknown gate type")."""
    typ, p = normalize_op(op)
    g = GATE_ALIASES.get(typ, typ)
    target = p.get("target")
    t0 = target if target is not None else 0  # (or target 0), circuit.clj:977-984
    angle = p.get("angle")
    control = p.get("control")

    def need(*names):
        for nm in names:
            if p.get(nm) is None:
                raise ValueError(f"{g} requires {names}")

    if g == "x":
        return apply_single_qubit_gate(state, PAULI_X, t0)
    if g == "y":
        return apply_single_qubit_gate(state, PAULI_Y, t0)
    if g == "z":
        return apply_single_qubit_gate(state, PAULI_Z, t0)
    if g == "h":
        return apply_single_qubit_gate(state, HADAMARD, t0)
    if g == "s":
        return apply_single_qubit_gate(state, S_GATE, t0)
    if g == "s-dag":
        return apply_single_qubit_gate(state, S_DAG_GATE, t0)
    if g == "t":
        return apply_single_qubit_gate(state, T_GATE, t0)
    if g == "t-dag":
        return apply_single_qubit_gate(state, T_DAG_GATE, t0)
    if g == "rx":
        return apply_single_qubit_gate(state, rx_gate(angle), t0)
    if g == "ry":
        return apply_single_qubit_gate(state, ry_gate(angle), t0)
    if g == "rz":
        return apply_single_qubit_gate(state, rz_gate(angle), t0)
    if g == "phase":
        return apply_single_qubit_gate(state, phase_gate(angle), t0)
    if g == "cnot":
        need("control", "target")
        return cnot(state, control, target)
    if g in ("cz", "rydberg-cz"):
        need("control", "target")
        return controlled_z(state, control, target)
    if g == "crz":
        need("control", "target", "angle")
        return controlled_rz(state, control, target, angle)
    if g == "crx":
        need("control", "target", "angle")
        return controlled_rx(state, control, target, angle)
    if g == "cry":
        need("control", "target", "angle")
        return controlled_ry(state, control, target, angle)
    if g == "swap":
        need("qubit1", "qubit2")
        return swap_gate(state, p["qubit1"], p["qubit2"])
    if g == "iswap":
        need("qubit1", "qubit2")
        return iswap_gate(state, p["qubit1"], p["qubit2"])
    if g == "toffoli":
        need("control1", "control2", "target")
        return toffoli_gate(state, p["control1"], p["control2"], target)
    if g == "fredkin":
        need("control", "target1", "target2")
        return fredkin_gate(state, control, p["target1"], p["target2"])
    if g == "rydberg-cphase":
        need("control", "target", "angle")
        return rydberg_cphase_gate(state, control, target, angle)
    if g == "rydberg-blockade":
        need("qubit-indices", "angle")
        return rydberg_blockade_gate(state, p["qubit-indices"], angle)
    if g == "global-rx":
        return _global(state, rx_gate(angle))
    if g == "global-ry":
        return _global(state, ry_gate(angle))
    if g == "global-rz":
        return _global(state, rz_gate(angle))
    if g == "global-h":
        return _global(state, HADAMARD)
    if g == "global-x":  # gate.clj:1224-1225 — (global-rx-gate state PI): keeps the (-i)^n phase
        return _global(state, rx_gate(math.pi))
    if g == "global-y":
        return _global(state, ry_gate(math.pi))
    if g == "global-z":
        return _global(state, rz_gate(math.pi))
    if superset and g == "i":
        return state
    if superset and g == "cy":
        need("control", "target")
        return controlled_y(state, control, target)
    raise UnknownGate(g)


def measurement_probabilities(state: np.ndarray) -> np.ndarray:
    """state.clj:676-682, 1087-1092 — p_i = |a_i|^2 computed as (fc/abs a)^2 (hypot, then square)."""
    mag = np.abs(state)
    return mag * mag


def measure_state(state: np.ndarray, u: float) -> Tuple[int, float]:
    """state.clj:894-913 with the JVM's (rand total) replaced by total*u, u in [0,1).

    outcome = #{i : cum_i < r} clamped to N-1, cum = sequential left-to-right running sum."""
    probs = measurement_probabilities(state)
    total = _seq_sum(probs)
    if abs(total - 1.0) > 1e-8:
        raise ValueError(f"State is not properly normalized: {total}")
    cum = np.cumsum(probs)  # numpy cumsum is a sequential left-to-right sum, as `reductions +`
    r = total * u
    outcome = int(np.searchsorted(cum, r, side="left"))  # count of cum_i < r
    outcome = min(outcome, state.shape[0] - 1)
    return outcome, float(probs[outcome])


def _seq_sum(x: np.ndarray) -> float:
    """Sequential left-to-right sum (`reduce +`), not numpy's pairwise sum."""
    if x.shape[0] == 0:
        return 0.0
    return float(np.cumsum(x)[-1])


def sample_outcomes(state: np.ndarray, uniforms: Iterable[float]) -> np.ndarray:
    """result.clj:224-225 — (repeatedly shots #(state/measure-state final-state)); one draw per shot."""
    probs = measurement_probabilities(state)
    cum = np.cumsum(probs)
    total = float(cum[-1])
    if abs(total - 1.0) > 1e-8:
        raise ValueError(f"State is not properly normalized: {total}")
    u = np.asarray(list(uniforms) if not isinstance(uniforms, np.ndarray) else uniforms, dtype=np.float64)
    out = np.searchsorted(cum, total * u, side="left")
    return np.minimum(out, state.shape[0] - 1).astype(np.int64)


def sample_boundary_distance(state: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
    """Distance |cum_i - r| to the nearest cumulative boundary for each draw (used by the parity
    tests to excuse shots that sit on a boundary within tolerance, BASELINE.json north_star)."""
    probs = measurement_probabilities(state)
    cum = np.cumsum(probs)
    r = float(cum[-1]) * np.asarray(uniforms, dtype=np.float64)
    k = np.searchsorted(cum, r, side="left")
    lo = np.abs(cum[np.clip(k - 1, 0, len(cum) - 1)] - r)
    hi = np.abs(cum[np.clip(k, 0, len(cum) - 1)] - r)
    return np.minimum(lo, hi)


def measure_specific_qubits(state: np.ndarray, qubits: Sequence[int], u: float):
    """state.clj:946-1014.  Outcome keys are enumerated with bit i of outcome-idx <-> i-th listed
    qubit (lines 961-963); one draw r = total*u; first outcome (in enumeration order) with
    cum >= r; amplitudes inconsistent with the outcome are zeroed and the rest scaled by
    1/sqrt(p) (only if p > 0).  Enumeration order = insertion order (exact for <= 3 measured qubits,
    where Clojure keeps an array-map; parity unpinned beyond that)."""
    n = num_qubits_of(state)
    m = len(qubits)
    idx = np.arange(state.shape[0], dtype=np.int64)
    key = np.zeros_like(idx)
    for i, q in enumerate(qubits):
        key |= ((idx >> (n - 1 - q)) & 1) << i
    probs = measurement_probabilities(state)
    outcome_probs = [_seq_sum(probs[key == k]) for k in range(1 << m)]
    total = 0.0
    cum = []
    for pk in outcome_probs:
        total += pk
        cum.append(total)
    r = total * u
    sel = 0
    while sel < len(cum) and cum[sel] < r:
        sel += 1
    sel = min(sel, len(cum) - 1)
    p_sel = outcome_probs[sel]
    factor = 1.0 / math.sqrt(p_sel) if p_sel > 0 else 1.0
    out = np.where(key == sel, state, 0.0) * complex(factor, 0.0)
    outcome_bits = [(sel >> i) & 1 for i in range(m)]
    return outcome_bits, out, outcome_probs


def apply_operation_to_state(state, op, draws=None, *, superset=False):
    """circuit.clj:1106-1111 — :measure collapses (one uniform consumed), everything else is a gate."""
    typ, p = normalize_op(op)
    if typ == "measure":
        qs = p.get("measurement-qubits")
        if qs is None:
            raise ValueError("Measure requires measurement-qubits parameter")
        u = next(draws) if draws is not None else 0.0
        return measure_specific_qubits(state, list(qs), u)[1]
    return apply_gate_to_state(state, op, superset=superset)


def execute_circuit(circuit: dict, initial_state: Optional[np.ndarray] = None, draws=None, *, superset=False):
    """circuit.clj:1778-1792 — reduce apply-operation-to-state over (:operations circuit)."""
    n = circuit.get("num-qubits", circuit.get(":num-qubits"))
    ops = circuit.get("operations", circuit.get(":operations"))
    state = zero_state(n) if initial_state is None else np.array(initial_state, dtype=np.complex128)
    it = iter(draws) if draws is not None else None
    for op in ops:
        state = apply_operation_to_state(state, op, it, superset=superset)
    return state


# --------------------------------------------------------------------------------------
# observables — domain/observables.clj:216-251, 314-320 ; domain/hamiltonian.clj:91-114
# --------------------------------------------------------------------------------------


def apply_pauli_string(state: np.ndarray, pauli: str) -> np.ndarray:
    """P|psi> with string char k <-> qubit k (leftmost = qubit 0 = MSB), observables.clj:216-229."""
    n = num_qubits_of(state)
    assert len(pauli) == n, "pauli string length must equal the number of qubits"
    out = state
    mats = {"I": None, "X": PAULI_X, "Y": PAULI_Y, "Z": PAULI_Z}
    for q, ch in enumerate(pauli):
        m = mats[ch]
        if m is not None:
            out = apply_single_qubit_gate(out, m, q)
    return out


def pauli_string_expectation(pauli: str, state: np.ndarray) -> float:
    """hamiltonian.clj:91-94 + observables.clj:246-251 — Re <psi|P|psi> (inner product conjugates
    the first argument)."""
    return float(np.vdot(state, apply_pauli_string(state, pauli)).real)


def hamiltonian_expectation(hamiltonian: Sequence[dict], state: np.ndarray) -> float:
    """hamiltonian.clj:96-114 — sum_i c_i <P_i>, left-to-right."""
    e = 0.0
    for term in hamiltonian:
        c = term.get("coefficient", term.get(":coefficient"))
        ps = term.get("pauli-string", term.get(":pauli-string"))
        e += c * pauli_string_expectation(ps, state)
    return e


def expectation_1q(state: np.ndarray, obs: np.ndarray, target: int) -> float:
    """result.clj:266-288 — single-qubit observable expanded with identities on the other qubits,
    then observables.clj:246-251."""
    return float(np.vdot(state, apply_single_qubit_gate(state, np.asarray(obs, dtype=np.complex128), target)).real)


def variance_1q(state: np.ndarray, obs: np.ndarray, target: int) -> float:
    """observables.clj:314-320 — <O^2> - <O>^2."""
    obs = np.asarray(obs, dtype=np.complex128)
    e = expectation_1q(state, obs, target)
    e2 = expectation_1q(state, obs @ obs, target)
    return e2 - e * e


def state_fidelity(a: np.ndarray, b: np.ndarray) -> float:
    """state.clj:1176-1185 — |<a|b>|."""
    return float(abs(np.vdot(a, b)))


# --------------------------------------------------------------------------------------
# noise channels — domain/channel.clj:52-121, 162-244, 259-264 ; domain/noise.clj:65-202
# --------------------------------------------------------------------------------------


def depolarizing_kraus_operators(p: float) -> List[np.ndarray]:
    """channel.clj:40-60 — {sqrt(1-p) I, sqrt(p/3) X, sqrt(p/3) Y, sqrt(p/3) Z}."""
    a = math.sqrt(1.0 - p)
    b = math.sqrt(p / 3.0)
    return [complex(a, 0) * PAULI_I, complex(b, 0) * PAULI_X, complex(b, 0) * PAULI_Y, complex(b, 0) * PAULI_Z]


def amplitude_damping_kraus_operators(gamma: float) -> List[np.ndarray]:
    """channel.clj:62-79."""
    return [np.array([[1.0, 0], [0, math.sqrt(1.0 - gamma)]], dtype=np.complex128),
            np.array([[0, math.sqrt(gamma)], [0, 0]], dtype=np.complex128)]


def phase_damping_kraus_operators(gamma: float) -> List[np.ndarray]:
    """channel.clj:81-99."""
    return [np.array([[1.0, 0], [0, math.sqrt(1.0 - gamma)]], dtype=np.complex128),
            np.array([[0, 0], [0, math.sqrt(gamma)]], dtype=np.complex128)]


def coherent_error_kraus_operator(angle: float, axis: str) -> np.ndarray:
    """channel.clj:101-121 — note: x/y are REAL rotation matrices and z = diag(cos a, cos(-a))
    (not unitary; the state is renormalised after application)."""
    c, s = math.cos(angle / 2.0), math.sin(angle / 2.0)
    axis = _kw(axis)
    if axis == "x":
        return np.array([[c, -s], [s, c]], dtype=np.complex128)
    if axis == "y":
        return np.array([[c, s], [-s, c]], dtype=np.complex128)
    if axis == "z":
        return np.array([[math.cos(angle), 0], [0, math.cos(-angle)]], dtype=np.complex128)
    raise ValueError(axis)


def apply_single_qubit_kraus_operator(state: np.ndarray, K: np.ndarray, q: int) -> np.ndarray:
    """channel.clj:162-200 — psi' = K psi / ||K psi|| on MSB-indexed qubit q (normalize-state rule)."""
    return normalize_state(apply_single_qubit_gate(state, K, q))


def kraus_selection_probabilities(kraus_ops: Sequence[np.ndarray]) -> List[float]:
    """channel.clj:225-233 — p_k = max over matrix elements of re^2 + im^2 (state independent)."""
    return [float(np.max(K.real * K.real + K.imag * K.imag)) for K in kraus_ops]


def select_kraus_index(kraus_ops: Sequence[np.ndarray], u: float) -> int:
    """channel.clj:235-242 — first k with u < cumulative, else the last operator."""
    probs = kraus_selection_probabilities(kraus_ops)
    cum = 0.0
    for k, pk in enumerate(probs):
        cum += pk
        if u < cum or k >= len(kraus_ops) - 1:
            return k
    return len(kraus_ops) - 1


def apply_quantum_channel(state, kraus_ops, q, draws):
    """channel.clj:215-244 — single operator: apply directly (no draw); otherwise one (rand)."""
    if len(kraus_ops) == 1:
        return apply_single_qubit_kraus_operator(state, kraus_ops[0], q)
    u = next(draws)
    return apply_single_qubit_kraus_operator(state, kraus_ops[select_kraus_index(kraus_ops, u)], q)


def calculate_decoherence_params(t1: float, t2: float, gate_time: float) -> Dict[str, float]:
    """channel.clj:259-264 — gamma = 1 - exp(-(gate_time[ns]/1000)/T[us])."""
    gt = gate_time / 1000.0
    return {"gamma-1": 1.0 - math.exp(-(gt / t1)), "gamma-2": 1.0 - math.exp(-(gt / t2))}


def _nm_get(d: dict, key: str, default=None):
    if d is None:
        return default
    if key in d:
        return d[key]
    if ":" + key in d:
        return d[":" + key]
    return default


def gate_noise_channel(op: dict, noise_model: dict):
    """noise.clj:65-103 — returns (kraus_ops, target_qubit, needs_draw) or None.

    Noise is looked up by the op's (un-aliased) :operation-type and applied to (:target params),
    default qubit 0.  Amplitude/phase damping take gamma from T1/T2 + gate time when present."""
    typ, p = normalize_op(op)
    target = p.get("target")
    target = 0 if target is None else target
    gate_noise = _nm_get(noise_model, "gate-noise") or {}
    cfg = _nm_get(gate_noise, typ)
    if not cfg:
        return None
    ntype = _kw(_nm_get(cfg, "noise-type"))
    t1, t2, gt = _nm_get(cfg, "t1-time"), _nm_get(cfg, "t2-time"), _nm_get(cfg, "gate-time")
    strength = _nm_get(cfg, "noise-strength", 0.01)
    if ntype == "depolarizing":
        return depolarizing_kraus_operators(strength), target
    if ntype == "amplitude-damping":
        dec = calculate_decoherence_params(t1, t2 or t1, gt) if (t1 and gt) else {"gamma-1": strength, "gamma-2": 0}
        return amplitude_damping_kraus_operators(dec["gamma-1"]), target
    if ntype == "phase-damping":
        dec = calculate_decoherence_params(t1 or t2, t2, gt) if (t2 and gt) else {"gamma-1": 0, "gamma-2": strength}
        return phase_damping_kraus_operators(dec["gamma-2"]), target
    if ntype == "coherent":
        cc = _nm_get(cfg, "coherent-error") or {"rotation-angle": 0.01, "rotation-axis": "z"}
        return [coherent_error_kraus_operator(_nm_get(cc, "rotation-angle"), _nm_get(cc, "rotation-axis"))], target
    return None


def apply_gate_noise(state, op, noise_model, draws):
    """noise.clj:65-103."""
    ch = gate_noise_channel(op, noise_model)
    if ch is None:
        return state
    kraus_ops, target = ch
    return apply_quantum_channel(state, kraus_ops, target, draws)


def calculate_final_bitstring(clean: str, n: int, readout_cfg: dict, draws) -> str:
    """noise.clj:120-164 — qubit 0..n-1 in order, one (rand) each; flip prob = base x product of
    correlation factors from already-flipped qubits (nested map {src {dst f}} only)."""
    p01 = _nm_get(readout_cfg, "prob-0-to-1")
    p10 = _nm_get(readout_cfg, "prob-1-to-0")
    corr = _nm_get(readout_cfg, "correlated-errors")
    bits = list(clean)
    history: List[int] = []
    for q in range(n):
        orig = bits[q]
        base = p01 if orig == "0" else p10
        factor = 1.0
        if corr and history:
            for src in history:
                qc = corr.get(src, corr.get(str(src))) if isinstance(corr, dict) else None
                if isinstance(qc, dict):
                    f = qc.get(q, qc.get(str(q)))
                    if f is not None:
                        factor *= f
        eff = min(1.0, max(0.0, base * factor))
        if next(draws) < eff:
            bits[q] = "1" if orig == "0" else "0"
            history.append(q)
    return "".join(bits)


def apply_readout_noise(state, n, noise_model, draws) -> str:
    """noise.clj:193-202 — measure-state (one draw), MSB-first bitstring, then per-qubit flips only if
    :readout-error is configured."""
    outcome, _ = measure_state(state, next(draws))
    clean = basis_string(outcome, n)
    ro = _nm_get(noise_model, "readout-error")
    if ro:
        return calculate_final_bitstring(clean, n, ro, draws)
    return clean


def draws_per_shot(circuit: dict, noise_model: dict) -> int:
    """Number of uniforms one noisy shot consumes, in the order of SURVEY §8a row 17: circuit order
    (one per multi-Kraus noisy gate, one per :measure op), one for the final measurement, n for
    readout if :readout-error is present."""
    n = circuit.get("num-qubits", circuit.get(":num-qubits"))
    cnt = 0
    for op in circuit.get("operations", circuit.get(":operations")):
        typ, _ = normalize_op(op)
        if typ == "measure":
            cnt += 1
        ch = gate_noise_channel(op, noise_model)
        if ch is not None and len(ch[0]) > 1:
            cnt += 1
    cnt += 1
    if _nm_get(noise_model, "readout-error"):
        cnt += n
    return cnt


def execute_single_noisy_shot(circuit, noise_model, draws, initial_state=None, *, superset=False):
    """adapter/backend/hardware_simulator.clj:84-104 — gate, then its noise; final readout."""
    n = circuit.get("num-qubits", circuit.get(":num-qubits"))
    state = zero_state(n) if initial_state is None else np.array(initial_state, dtype=np.complex128)
    for op in circuit.get("operations", circuit.get(":operations")):
        state = apply_operation_to_state(state, op, draws, superset=superset)
        state = apply_gate_noise(state, op, noise_model, draws)
    bitstring = apply_readout_noise(state, n, noise_model, draws)
    return bitstring, state


def run_noisy(circuit, noise_model, uniforms: np.ndarray, max_trajectories: int = 100, *, superset=False, initial_state=None):
    """hardware_simulator.clj:120-185 — `shots` independent shots; counts keyed by bitstring; the first
    <= max_trajectories final states are kept.  uniforms: [shots, draws_per_shot] array.  initial_state:
    (or (:initial-state options) zero-state), hardware_simulator.clj:228-230."""
    counts: Dict[str, int] = {}
    trajectories = []
    last = None
    for row in np.asarray(uniforms):
        it = iter(row.tolist())
        bs, st = execute_single_noisy_shot(circuit, noise_model, it, initial_state, superset=superset)
        counts[bs] = counts.get(bs, 0) + 1
        if len(trajectories) < max_trajectories:
            trajectories.append(st)
        last = st
    return {"measurement-results": counts, "final-state": last, "trajectories": trajectories}


def trajectory_to_density_matrix(trajectories: Sequence[np.ndarray]) -> np.ndarray:
    """state.clj:750-792 — rho = sum_k (1/K) |psi_k><psi_k| (equal weights)."""
    k = len(trajectories)
    w = 1.0 / k
    wsum = w * k
    rho = None
    for psi in trajectories:
        proj = np.outer(psi, np.conj(psi)) * (w / wsum)
        rho = proj if rho is None else rho + proj
    return rho


# --------------------------------------------------------------------------------------
# reference-faithful O(4^n) single-qubit application (used only for the cpu_baseline "how the
# reference does it" timing, never for parity): gate.clj:346-353 + matrix-vector product.
# --------------------------------------------------------------------------------------


def apply_single_qubit_gate_dense(state: np.ndarray, U: np.ndarray, q: int) -> np.ndarray:
    n = num_qubits_of(state)
    full = np.array([[1.0 + 0j]])
    for i in range(n):
        full = np.kron(full, U if i == q else PAULI_I)
    return full @ state
