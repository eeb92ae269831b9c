"""Caller-side circuit builders (plain dict circuit maps) for the workloads in BASELINE.json.

These restate the *builders* of the reference so that the benchmark and the parity tests can feed
the backend the exact gate lists the reference would (paths relative to
/root/reference/src/org/soulspace/qclojure/).  They produce the reference's circuit map
(domain/circuit.clj:30-37) with keyword names spelled as plain strings:

    {"num-qubits": n, "name": ..., "operations": [{"operation-type": "h",
                                                    "operation-params": {"target": 0}}, ...]}

No arithmetic on state vectors happens here.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np


def create_circuit(n: int, name: str = "", description: str = "") -> dict:
    """domain/circuit.clj:47-70."""
    return {"num-qubits": int(n), "name": name, "description": description, "operations": []}


def add_gate(c: dict, typ: str, **params) -> dict:
    """domain/circuit.clj:112-160 (add-gate): appends {:operation-type :operation-params}."""
    c["operations"].append({"operation-type": typ,
                            "operation-params": {k.replace("_", "-"): v for k, v in params.items()}})
    return c


def h(c, t): return add_gate(c, "h", target=t)
def x(c, t): return add_gate(c, "x", target=t)
def y(c, t): return add_gate(c, "y", target=t)
def z(c, t): return add_gate(c, "z", target=t)
def s(c, t): return add_gate(c, "s", target=t)
def t_gate(c, t): return add_gate(c, "t", target=t)
def rx(c, t, a): return add_gate(c, "rx", target=t, angle=float(a))
def ry(c, t, a): return add_gate(c, "ry", target=t, angle=float(a))
def rz(c, t, a): return add_gate(c, "rz", target=t, angle=float(a))
def phase(c, t, a): return add_gate(c, "phase", target=t, angle=float(a))
def cnot(c, ctl, t): return add_gate(c, "cnot", control=ctl, target=t)
def cz(c, ctl, t): return add_gate(c, "cz", control=ctl, target=t)
def crz(c, ctl, t, a): return add_gate(c, "crz", control=ctl, target=t, angle=float(a))
def crx(c, ctl, t, a): return add_gate(c, "crx", control=ctl, target=t, angle=float(a))
def cry(c, ctl, t, a): return add_gate(c, "cry", control=ctl, target=t, angle=float(a))
def swap(c, q1, q2): return add_gate(c, "swap", qubit1=q1, qubit2=q2)
def iswap(c, q1, q2): return add_gate(c, "iswap", qubit1=q1, qubit2=q2)
def toffoli(c, c1, c2, t): return add_gate(c, "toffoli", control1=c1, control2=c2, target=t)
def fredkin(c, ctl, t1, t2): return add_gate(c, "fredkin", control=ctl, target1=t1, target2=t2)
def measure(c, qubits): return add_gate(c, "measure", measurement_qubits=list(qubits))


def bell_state_circuit() -> dict:
    """domain/circuit.clj:1651-1673."""
    c = create_circuit(2, "Bell State")
    return cnot(h(c, 0), 0, 1)


def ghz_state_circuit(n: int) -> dict:
    """domain/circuit.clj:1675-1703 — H on qubit 0 then CNOT(0, i) for i = 1..n-1."""
    assert n >= 2
    c = create_circuit(n, "GHZ State", f"Prepares {n}-qubit GHZ state")
    h(c, 0)
    for i in range(1, n):
        cnot(c, 0, i)
    return c


def quantum_fourier_transform_circuit(n: int) -> dict:
    """application/algorithm/quantum_fourier_transform.clj:34-62 — per qubit: H, then
    CRZ(pi/2^(k+1)) with control = qubit+k+1, target = qubit; finally swaps i <-> n-1-i."""
    c = create_circuit(n, "QFT", "Quantum Fourier Transform")
    for q in range(n):
        h(c, q)
        for k in range(n - q - 1):
            ctl = q + k + 1
            if ctl < n:
                crz(c, ctl, q, math.pi / math.pow(2, k + 1))
    for i in range(n // 2):
        j = n - 1 - i
        if i < j:
            swap(c, i, j)
    return c


def random_brickwork_circuit(n: int, depth: int = 20, seed: int | None = None) -> dict:
    """layer l = one 1q gate per qubit drawn uniformly from
    {H, RX(theta), RZ(theta)}, theta ~ U[0, 2pi); then 2q gates on pairs (q, q+1), q = l mod 2 (mod 2),
    each CNOT or CZ with probability 1/2, control = q.  Generator default_rng(1000 + n)."""
    rng = np.random.default_rng(1000 + n if seed is None else seed)
    c = create_circuit(n, "Random brickwork", f"depth {depth}")
    for layer in range(depth):
        kinds = rng.integers(0, 3, size=n)
        thetas = rng.random(n) * 2.0 * math.pi
        for q in range(n):
            if kinds[q] == 0:
                h(c, q)
            elif kinds[q] == 1:
                rx(c, q, thetas[q])
            else:
                rz(c, q, thetas[q])
        start = layer % 2
        pairs = list(range(start, n - 1, 2))
        coin = rng.integers(0, 2, size=len(pairs))
        for k, q in enumerate(pairs):
            if coin[k] == 0:
                cnot(c, q, q + 1)
            else:
                cz(c, q, q + 1)
    return c


def hardware_efficient_ansatz(n: int, parameters: Sequence[float], num_layers: int = 1,
                              entangling_gate: str = "cnot") -> dict:
    """domain/ansatz.clj:26-95 — per layer: RX,RY,RZ on every qubit (3 params each), then a linear
    chain of entanglers (q, q+1)."""
    c = create_circuit(n, "Hardware Efficient Ansatz")
    p = 0
    for _ in range(num_layers):
        for q in range(n):
            rx(c, q, parameters[p]); ry(c, q, parameters[p + 1]); rz(c, q, parameters[p + 2])
            p += 3
        for q in range(n - 1):
            if entangling_gate == "cnot":
                cnot(c, q, q + 1)
            elif entangling_gate == "cz":
                cz(c, q, q + 1)
            else:
                crz(c, q, q + 1, math.pi / 4)
    return c


def uccsd_inspired_ansatz(n: int, parameters: Sequence[float]) -> dict:
    """domain/ansatz.clj:97-141 — Hartree-Fock X on the lower half, then per excitation
    RY(i, a/2) CNOT(i,a) RY(a, a/2) CNOT(i,a) RY(i, -a/2)."""
    c = create_circuit(n, "UCCSD Inspired Ansatz")
    half = n // 2
    for q in range(half):
        x(c, q)
    for exc, angle in enumerate(parameters):
        i = exc % half
        a = half + (exc % half)
        ry(c, i, angle / 2); cnot(c, i, a); ry(c, a, angle / 2); cnot(c, i, a); ry(c, i, -(angle / 2))
    return c


def max_cut_hamiltonian(graph: Sequence[Sequence[float]], num_vertices: int) -> List[dict]:
    """application/algorithm/qaoa.clj:132-149 — per edge: w/2 * I...I  -  w/2 * Z_i Z_j."""
    terms = []
    for (i, j, w) in graph:
        coeff = w / 2.0
        zz = ["I"] * num_vertices
        zz[int(i)] = "Z"; zz[int(j)] = "Z"
        terms.append({"coefficient": coeff, "pauli-string": "I" * num_vertices})
        terms.append({"coefficient": -coeff, "pauli-string": "".join(zz)})
    return terms


def standard_mixer_hamiltonian(n: int) -> List[dict]:
    """application/algorithm/qaoa.clj:409-416."""
    return [{"coefficient": 1.0, "pauli-string": "I" * i + "X" + "I" * (n - 1 - i)} for i in range(n)]


def hamiltonian_evolution_circuit(c: dict, hamiltonian: Sequence[dict], t: float) -> dict:
    """application/algorithm/qaoa.clj:467-528 — angle = 2 t coeff; single-Pauli terms -> RX/RY/RZ;
    all-Z terms -> CNOT ladder + RZ on the last qubit; identity and mixed multi-qubit terms skipped."""
    for term in hamiltonian:
        coeff, ps = term["coefficient"], term["pauli-string"]
        angle = 2.0 * t * coeff
        pos = [(k, ch) for k, ch in enumerate(ps) if ch != "I"]
        if not pos:
            continue
        if len(pos) == 1:
            k, ch = pos[0]
            {"X": rx, "Y": ry, "Z": rz}[ch](c, k, angle)
        elif all(ch == "Z" for _, ch in pos):
            qs = [k for k, _ in pos]
            if len(qs) == 2:
                cnot(c, qs[0], qs[1]); rz(c, qs[1], angle); cnot(c, qs[0], qs[1])
            else:
                tgt, ctrls = qs[-1], qs[:-1]
                for ctl in ctrls:
                    cnot(c, ctl, tgt)
                rz(c, tgt, angle)
                for ctl in reversed(ctrls):
                    cnot(c, ctl, tgt)
    return c


def qaoa_ansatz_circuit(problem_h, mixer_h, parameters: Sequence[float], n: int) -> dict:
    """application/algorithm/qaoa.clj:530-564 — H on all qubits, then per (gamma, beta) pair the problem
    evolution followed by the mixer evolution."""
    c = create_circuit(n, "QAOA Ansatz")
    for q in range(n):
        h(c, q)
    for k in range(0, len(parameters), 2):
        gamma, beta = parameters[k], parameters[k + 1]
        hamiltonian_evolution_circuit(c, problem_h, gamma)
        hamiltonian_evolution_circuit(c, mixer_h, beta)
    return c


def hhl_circuit(matrix: Sequence[Sequence[float]], b_vector: Sequence[float], precision_qubits: int = 4,
                ancilla_qubits: int = 1) -> dict:
    """application/algorithm/hhl.clj:625-714 (`hhl-circuit`): vector qubit(s), precision register, ancilla.  RY state
    preparation for a 2-element b, H on the precision register, CRZ(0.1 * A[0][0] * 2^k) from precision qubit k onto the
    vector qubit, RY(pi/8) on the ancilla, CRY(pi/(2+k)) from precision qubit k onto the ancilla, H on the precision
    register in reverse order."""
    n = len(matrix)
    vq = max(1, int(math.ceil(math.log(n) / math.log(2))))
    total = vq + precision_qubits + ancilla_qubits
    c = create_circuit(total, "Working HHL Algorithm", f"Functional HHL for {n}×{n} matrix")
    vec = list(range(vq))
    prec = list(range(vq, vq + precision_qubits))
    anc = list(range(vq + precision_qubits, total))
    if n == 2:
        b1, b2 = b_vector
        norm = math.sqrt(b1 * b1 + b2 * b2)
        nb1, nb2 = b1 / norm, b2 / norm
        ry(c, vec[0], 2 * math.atan2(abs(nb2), abs(nb1)) if abs(nb2) > 1e-10 else 0.0)
    for q in prec:
        h(c, q)
    for k, pq in enumerate(prec):
        crz(c, pq, vec[0], 0.1 * matrix[0][0] * math.pow(2, k))
    ry(c, anc[0], math.pi / 8)
    for k, pq in enumerate(prec):
        cry(c, pq, anc[0], math.pi / (2 + k))
    for q in reversed(prec):
        h(c, q)
    return c


def random_regular_graph(n: int, degree: int = 3, seed: int = 11) -> List[List[float]]:
    """3-regular random graph for the QAOA sweep of SURVEY §8(d) config 5 (pairing model with
    rejection; default_rng(seed))."""
    rng = np.random.default_rng(seed)
    assert (n * degree) % 2 == 0
    while True:
        stubs = np.repeat(np.arange(n), degree)
        rng.shuffle(stubs)
        edges = set()
        ok = True
        for a, b in zip(stubs[::2], stubs[1::2]):
            a, b = int(min(a, b)), int(max(a, b))
            if a == b or (a, b) in edges:
                ok = False
                break
            edges.add((a, b))
        if ok:
            return [[a, b, 1.0] for a, b in sorted(edges)]


def grover_iterations(n: int, num_targets: int = 1) -> int:
    """application/algorithm/grover.clj:289-293 — max(1, floor(pi/4 * sqrt(N/M)))."""
    return max(1, int(math.floor(math.pi / 4.0 * math.sqrt((1 << n) / num_targets))))
