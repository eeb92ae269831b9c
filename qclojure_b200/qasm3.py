"""OpenQASM 3.0 import / export of circuits, mirroring the reference's line-oriented converter
`src/org/soulspace/qclojure/application/format/qasm3.clj` (`circuit-to-qasm` :321-365, `qasm-to-gate` :437-563,
`qasm-to-circuit` :565-597, `parse-qasm-expression` :391-441, result pragmas :43-131 / :367-389 / :443-470), which is what
`adapter/io/qasm.clj` binds to the `:qasm3` format.  It lets benchmark circuits be exchanged with a QClojure installation
(or any OpenQASM 3 tool) as text.  Host-side only: no arithmetic on states.

Gate names read from QASM go through the reference's alias table (`operation_registry.clj:387-407`): cx -> cnot,
sdg -> s-dag, tdg -> t-dag, id -> i.  Like the reference, unknown lines are skipped silently and register names other
than `q` / `c` are not recognised.
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Optional, Tuple

from . import circuits as CB
from .io import _fmt_double

_ALIASES = {"cx": "cnot", "sdg": "s-dag", "tdg": "t-dag", "id": "i", "ccx": "toffoli", "cswap": "fredkin", "p": "phase"}
_PRAGMA_TARGET = {"qclojure": "qclojure", "braket": "braket"}


def _kw(x):
    return x[1:] if isinstance(x, str) and x.startswith(":") else x


def _get(d, key, default=None):
    if not isinstance(d, dict):
        return default
    return d.get(key, d.get(":" + key, default))


def _num(x) -> str:
    """How Clojure's `str` prints the number (doubles in Java's shortest round-trip form)."""
    if isinstance(x, bool):
        return "true" if x else "false"
    if isinstance(x, int):
        return str(x)
    return _fmt_double(float(x))


def _name(x) -> str:
    return str(_kw(x))


# ------------------------------------------------------------------ emission
def emit_qasm_pragmas(options: Optional[dict]) -> Optional[str]:
    """qasm3.clj:43-131."""
    specs = _get(options, "result-specs")
    tgt = _PRAGMA_TARGET[_name(_get(options, "target", "qclojure"))]
    if not specs:
        return None
    lines: List[str] = []

    def per_observable(kind, spec, with_shots=False):
        obs = _get(spec, "observables", []) or []
        targets = _get(spec, "targets", []) or []
        for i, o in enumerate(obs):
            s = f"#pragma {tgt} result {kind} observable={_name(o)}"
            if with_shots:
                s += f" shots={_get(spec, 'shots', 1000)}"
            if i < len(targets):
                s += f" target={targets[i]}"
            lines.append(s)

    for key, spec in specs.items():
        typ = _name(key)
        if typ == "measurements":
            s = f"#pragma {tgt} result measurement shots={_get(spec, 'shots', 1000)}"
            if _get(spec, "qubits"):
                s += " qubits=" + ",".join(str(q) for q in _get(spec, "qubits"))
            lines.append(s)
        elif typ == "expectation":
            per_observable("expectation", spec)
        elif typ == "variance":
            per_observable("variance", spec)
        elif typ == "probability":
            s = f"#pragma {tgt} result probability"
            if _get(spec, "targets"):
                s += " targets=" + ",".join(str(t) for t in _get(spec, "targets"))
            if _get(spec, "states"):
                s += " states=" + ",".join(str(t) for t in _get(spec, "states"))
            lines.append(s)
        elif typ == "amplitude":
            lines.append(f"#pragma {tgt} result amplitude states=" + ",".join(str(t) for t in (_get(spec, "states", []) or [])))
        elif typ == "sample":
            per_observable("sample", spec, with_shots=True)
        elif typ in ("state-vector", "density-matrix", "fidelity"):
            lines.append(f"#pragma {tgt} result {typ.replace('-', '_')} // Simulation-only result")
        else:
            lines.append(f"// Unknown result type: {typ}")
    if not lines:
        return None
    return "\n// Result extraction specifications\n" + "\n".join(lines) + "\n"


def _gate_to_qasm(op: dict, n: int, braket: bool) -> str:
    """qasm3.clj:133-245."""
    g = _name(_get(op, "operation-type"))
    p = {_name(k): v for k, v in (_get(op, "operation-params") or {}).items()}
    q = lambda k: f"q[{p.get(k)}]"            # noqa: E731
    a = lambda: _num(p.get("angle"))          # noqa: E731
    one = {"i": "id", "x": "x", "y": "y", "z": "z", "h": "h", "s": "s", "t": "t",
           "s-dag": "si" if braket else "sdg", "t-dag": "ti" if braket else "tdg"}
    if g in one:
        return f"{one[g]} {q('target')};"
    if g in ("rx", "ry", "rz"):
        return f"{g}({a()}) {q('target')};"
    if g == "phase":
        return f"{'phaseshift' if braket else 'p'}({a()}) {q('target')};"
    if g in ("cnot", "cx"):
        return f"{'cnot' if braket and g == 'cnot' else 'cx'} {q('control')}, {q('target')};"
    if g in ("cz", "cy"):
        return f"{g} {q('control')}, {q('target')};"
    if g in ("crx", "cry", "crz"):
        return f"{g}({a()}) {q('control')}, {q('target')};"
    if g == "controlled":
        return f"ctrl @ {p.get('gate')} {q('control')}, {q('target')};"
    if g in ("swap", "iswap"):
        return f"{g} {q('qubit1')}, {q('qubit2')};"
    if g == "toffoli":
        return f"{'ccnot' if braket else 'ccx'} {q('control1')}, {q('control2')}, {q('target')};"
    if g == "fredkin":
        return f"cswap {q('control')}, {q('target1')}, {q('target2')};"
    glob = {"global-x": ("X", "x"), "global-y": ("Y", "y"), "global-z": ("Z", "z"), "global-h": ("Hadamard", "h")}
    if g in glob:
        nm, qs = glob[g]
        return f"// Global {nm} gate - apply {nm[0] if nm != 'Hadamard' else 'H'} to all qubits\n" + \
            "\n".join(f"{qs} q[{i}];" for i in range(n))
    if g in ("global-rx", "global-ry", "global-rz"):
        r = g[-2:]
        return f"// Global {r.upper()}({a()}) gate - apply {r.upper()} to all qubits\n" + \
            "\n".join(f"{r}({a()}) q[{i}];" for i in range(n))
    if g == "rydberg-cz":
        return f"// Rydberg CZ gate - decomposed to standard CZ\ncz {q('control')}, {q('target')};"
    if g == "rydberg-cphase":
        return f"// Rydberg controlled phase gate - decomposed to CRZ\ncrz({a()}) {q('control')}, {q('target')};"
    if g == "rydberg-blockade":
        return ("// Rydberg blockade gate - hardware specific\n// Cannot be directly expressed in standard QASM\n"
                "// Requires hardware-specific backend support")
    if g == "measure":
        return "\n".join(f"c[{m}] = measure q[{m}];" for m in p.get("measurement-qubits", []))
    return f"// Unknown gate: {g}"


def circuit_to_qasm(circuit: dict, options: Optional[dict] = None) -> str:
    """qasm3.clj:321-365."""
    options = options or {}
    braket = _name(_get(options, "target", "qclojure")) == "braket"
    n = int(_get(circuit, "num-qubits"))
    header = "OPENQASM 3.0;\n" + ("" if braket else 'include "stdgates.inc";\n\n') + f"qubit[{n}] q;\nbit[{n}] c;\n"
    gates = "\n".join(_gate_to_qasm(op, n, braket) for op in _get(circuit, "operations", []))
    return header + (emit_qasm_pragmas(options) or "") + "\n" + gates


# ------------------------------------------------------------------ parsing
def parse_qasm_expression(expr: str) -> float:
    """qasm3.clj:391-441: numbers, pi, -pi, pi/n, n*pi, pi*n, a/b."""
    e = expr.strip()
    num = r"\d+(\.\d+)?"
    if re.fullmatch(rf"-?{num}", e):
        return float(e)
    if e == "pi":
        return math.pi
    if e == "-pi":
        return -math.pi
    m = re.fullmatch(rf"(-?)pi/({num})", e)
    if m:
        return (-math.pi if m.group(1) else math.pi) / float(m.group(2))
    m = re.fullmatch(rf"(-?{num})\*pi", e)
    if m:
        return float(m.group(1)) * math.pi
    m = re.fullmatch(rf"(-?)pi\*({num})", e)
    if m:
        # the reference's `^pi\*(.+)$` capture fails on a leading minus (re-find returns nil -> NPE); the sign is honoured here
        return (-math.pi if m.group(1) else math.pi) * float(m.group(2))
    m = re.fullmatch(rf"(-?{num})/({num})", e)
    if m:
        return float(m.group(1)) / float(m.group(3))
    try:
        return float(e)
    except ValueError:
        raise ValueError(f"Unsupported QASM expression: {e}. Supported: numbers, pi, pi/n, n*pi, pi*n, fractions") from None


def parse_result_pragma(line: str) -> Optional[Tuple[str, dict]]:
    """qasm3.clj:367-389."""
    if not line.startswith("#pragma qclojure result"):
        return None
    parts = line.split()
    if len(parts) < 4:
        return None
    params: Dict[str, object] = {}
    for pair in parts[4:]:
        if "=" not in pair:
            if pair.startswith("//"):
                break
            continue
        k, v = pair.split("=", 1)
        if re.fullmatch(r"\d+", v):
            params[k] = int(v)
        elif "," in v:
            params[k] = [s.strip() for s in v.split(",")]
        else:
            params[k] = v
    return parts[3], params


def collect_result_specs_from_qasm(lines) -> dict:
    """qasm3.clj:443-470."""
    specs: Dict[str, dict] = {}
    for ln in lines:
        pr = parse_result_pragma(ln.strip())
        if not pr:
            continue
        typ, params = pr
        if typ in specs and typ in ("expectation", "variance", "sample"):
            specs[typ].setdefault("observables", []).append(params.get("observable"))
            specs[typ].setdefault("targets", []).append(params.get("target"))
        else:
            specs[typ] = params
    return specs


_Q = r"q\[(\d+)\]"
_RULES = [
    (re.compile(rf"^(x|y|z|h|s|t|sdg|tdg|id)\s+{_Q}"), lambda c, m: CB.add_gate(c, _ALIASES.get(m[1], m[1]), target=int(m[2]))),
    (re.compile(rf"^(cx|cz|cy)\s+{_Q},\s*{_Q}"), lambda c, m: CB.add_gate(c, _ALIASES.get(m[1], m[1]), control=int(m[2]), target=int(m[3]))),
    (re.compile(rf"^swap\s+{_Q},\s*{_Q}"), lambda c, m: CB.swap(c, int(m[1]), int(m[2]))),
    (re.compile(rf"^iswap\s+{_Q},\s*{_Q}"), lambda c, m: CB.iswap(c, int(m[1]), int(m[2]))),
    (re.compile(rf"^ccx\s+{_Q},\s*{_Q},\s*{_Q}"), lambda c, m: CB.toffoli(c, int(m[1]), int(m[2]), int(m[3]))),
    (re.compile(rf"^cswap\s+{_Q},\s*{_Q},\s*{_Q}"), lambda c, m: CB.fredkin(c, int(m[1]), int(m[2]), int(m[3]))),
    (re.compile(rf"^cr([xyz])\((.+?)\)\s+{_Q},\s*{_Q}"),
     lambda c, m: CB.add_gate(c, "cr" + m[1], control=int(m[3]), target=int(m[4]), angle=parse_qasm_expression(m[2]))),
    (re.compile(rf"^p\((.+?)\)\s+{_Q}"), lambda c, m: CB.phase(c, int(m[2]), parse_qasm_expression(m[1]))),
    (re.compile(rf"^r([xyz])\((.+?)\)\s+{_Q}"),
     lambda c, m: CB.add_gate(c, "r" + m[1], target=int(m[3]), angle=parse_qasm_expression(m[2]))),
    (re.compile(rf"^c\[(\d+)\]\s*=\s*measure\s+{_Q}"), lambda c, m: CB.measure(c, [int(m[2])])),
]


def qasm_to_gate(circuit: dict, line: str) -> dict:
    """qasm3.clj:472-563: first matching rule wins, everything else (comments, declarations) is skipped."""
    line = line.strip()
    for rx, fn in _RULES:
        m = rx.search(line)
        if m:
            return fn(circuit, m)
    return circuit


def qasm_to_circuit(qasm: str) -> dict:
    """qasm3.clj:565-597: circuit map with the parsed pragmas under "result-specs"."""
    lines = qasm.splitlines()
    decl = next((ln for ln in lines if ln.strip().startswith("qubit[")), None)
    if decl is None:
        raise ValueError("no qubit[n] declaration found")
    c = CB.create_circuit(int(re.search(r"\d+", decl).group(0)), "Converted Circuit")
    for ln in lines:
        qasm_to_gate(c, ln)
    c["result-specs"] = collect_result_specs_from_qasm(lines)
    return c


# ------------------------------------------------------------------ adapter/io/qasm.clj
def export_quantum_circuit(circuit: dict, filename: str, options: Optional[dict] = None) -> str:
    with open(filename, "w") as f:
        f.write(circuit_to_qasm(circuit, options))
    return filename                      # util/io save-file returns the file name (doc/tutorial.md prints it)


def import_quantum_circuit(filename: str) -> dict:
    with open(filename) as f:
        return qasm_to_circuit(f.read())
