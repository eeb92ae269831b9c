"""OpenQASM 2.0 import / export of circuits, mirroring the reference's line-oriented converter
`src/org/soulspace/qclojure/application/format/qasm2.clj` (`circuit-to-qasm` :86-110, `gate-to-qasm-fn` :16-84, `qasm-to-gate`
:165-262, `qasm-to-circuit` :264-284), which `adapter/io/qasm.clj:12-18` binds to the `:qasm2` format.  Differences from the
QASM 3 converter that are the reference's own: no space after the comma between operands, every `:measure` op becomes a
comment and one `measure q -> c;` closes the program, lines are matched untrimmed and angles must be plain numbers
(`Double/parseDouble`, no `pi` expressions), measurements are not imported.  Host-side only."""
from __future__ import annotations

import re
from typing import Optional

from . import circuits as CB
from .qasm3 import _get, _name, _num

_ALIASES = {"sdg": "s-dag", "tdg": "t-dag", "id": "i"}


def _gate_to_qasm(op: dict, n: int) -> str:
    g = _name(_get(op, "operation-type"))
    p = {_name(k): v for k, v in (_get(op, "operation-params") or {}).items()}
    q = lambda k: f"q[{p.get(k)}]"            # noqa: E731
    a = lambda: _num(p.get("angle"))          # noqa: E731
    one = {"i": "id", "x": "x", "y": "y", "z": "z", "h": "h", "s": "s", "t": "t", "s-dag": "sdg", "t-dag": "tdg"}
    if g in one:
        return f"{one[g]} {q('target')};"
    if g == "phase":
        return f"p({a()}) {q('target')};"
    if g in ("rx", "ry", "rz"):
        return f"{g}({a()}) {q('target')};"
    if g in ("cnot", "cx"):
        return f"cx {q('control')},{q('target')};"
    if g in ("cz", "cy"):
        return f"{g} {q('control')},{q('target')};"
    if g in ("swap", "iswap"):
        return f"{g} {q('qubit1')},{q('qubit2')};"
    if g == "toffoli":
        return f"ccx {q('control1')},{q('control2')},{q('target')};"
    if g == "fredkin":
        return f"cswap {q('control')},{q('target1')},{q('target2')};"
    if g in ("crx", "cry", "crz"):
        return f"{g}({a()}) {q('control')},{q('target')};"
    glob = {"global-x": ("X", "x"), "global-y": ("Y", "y"), "global-z": ("Z", "z"), "global-h": ("Hadamard", "h")}
    if g in glob:
        nm, qs = glob[g]
        short = "H" if nm == "Hadamard" else nm
        return f"// Global {nm} gate - decomposed to individual {short} gates\n" + "\n".join(f"{qs} q[{i}];" for i in range(n))
    if g in ("global-rx", "global-ry", "global-rz"):
        r = g[-2:]
        return (f"// Global {r.upper()}({a()}) gate - decomposed to individual {r.upper()} gates\n"
                + "\n".join(f"{r}({a()}) q[{i}];" for i in range(n)))
    if g == "rydberg-cz":
        return f"// Rydberg CZ gate - decomposed to standard CZ\ncz {q('control')},{q('target')};"
    if g == "rydberg-cphase":
        return f"// Rydberg controlled phase gate - decomposed to CRZ\ncrz({a()}) {q('control')},{q('target')};"
    if g == "rydberg-blockade":
        return "// Rydberg blockade gate - cannot be expressed in QASM 2.0\n// Requires hardware-specific backend support"
    if g == "measure":
        return "// Measurement will be handled by final measure statement"
    return f"// Unknown gate: {g}"


def circuit_to_qasm(circuit: dict, result_specs: Optional[dict] = None) -> str:
    """qasm2.clj:86-110 (the result specs are accepted and ignored, like the reference)."""
    n = int(_get(circuit, "num-qubits"))
    header = f'OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[{n}];\ncreg c[{n}];\n\n'
    return header + "\n".join(_gate_to_qasm(op, n) for op in _get(circuit, "operations", [])) + "\nmeasure q -> c;"


_Q = r"q\[(\d+)\]"
_RULES = [
    (re.compile(rf"^(x|y|z|h|s|t|sdg|tdg|id)\s+{_Q}"), lambda c, m: CB.add_gate(c, _ALIASES.get(m[1], m[1]), target=int(m[2]))),
    (re.compile(rf"^c([xyz])\s+{_Q},\s*{_Q}"),
     lambda c, m: CB.add_gate(c, {"x": "cnot", "z": "cz", "y": "cy"}[m[1]], control=int(m[2]), target=int(m[3]))),
    (re.compile(rf"^swap\s+{_Q},\s*{_Q}"), lambda c, m: CB.swap(c, int(m[1]), int(m[2]))),
    (re.compile(rf"^iswap\s+{_Q},\s*{_Q}"), lambda c, m: CB.iswap(c, int(m[1]), int(m[2]))),
    (re.compile(rf"^ccx\s+{_Q},\s*{_Q},\s*{_Q}"), lambda c, m: CB.toffoli(c, int(m[1]), int(m[2]), int(m[3]))),
    (re.compile(rf"^cswap\s+{_Q},\s*{_Q},\s*{_Q}"), lambda c, m: CB.fredkin(c, int(m[1]), int(m[2]), int(m[3]))),
    (re.compile(rf"^cr([xyz])\((.+?)\)\s+{_Q},\s*{_Q}"),
     lambda c, m: CB.add_gate(c, "cr" + m[1], control=int(m[3]), target=int(m[4]), angle=float(m[2]))),
    (re.compile(rf"^p\((.+?)\)\s+{_Q}"), lambda c, m: CB.phase(c, int(m[2]), float(m[1]))),
    (re.compile(rf"^r([xyz])\((.+?)\)\s+{_Q}"), lambda c, m: CB.add_gate(c, "r" + m[1], target=int(m[3]), angle=float(m[2]))),
]


def qasm_to_gate(circuit: dict, line: str) -> dict:
    """qasm2.clj:165-262 (the line is NOT trimmed by the reference)."""
    for rx, fn in _RULES:
        m = rx.search(line)
        if m:
            return fn(circuit, m)
    return circuit


def qasm_to_circuit(qasm: str) -> dict:
    """qasm2.clj:264-284."""
    lines = qasm.splitlines()
    decl = next((ln for ln in lines if ln.startswith("qreg")), None)
    if decl is None:
        raise ValueError("no qreg declaration found")
    c = CB.create_circuit(int(re.search(r"\d+", decl).group(0)), "Converted Circuit")
    for ln in lines:
        qasm_to_gate(c, ln)
    return c


def export_quantum_circuit(circuit: dict, filename: str) -> str:
    with open(filename, "w") as f:
        f.write(circuit_to_qasm(circuit))
    return filename


def import_quantum_circuit(filename: str) -> dict:
    with open(filename) as f:
        return qasm_to_circuit(f.read())
