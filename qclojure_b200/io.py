"""Wire / on-disk formats of the reference: quantum states and circuits as EDN or JSON, exactly as
`src/org/soulspace/qclojure/adapter/io.clj:33-89` serialises them and `adapter/io/edn.clj`, `adapter/io/json.clj` write
them, so that states, circuits and golden vectors can be exchanged with a real QClojure installation.

  state   {:state-vector [{:real r :imag i} ...] :num-qubits n :metadata {...} :format-version "1.0"}     (io.clj:33-51)
  circuit {:operations [{:operation-type :h :operation-params {:target 0}} ...] :num-qubits n :name s :description s
           :metadata {...} :format-version "1.0"}                                                       (io.clj:68-89)

Python side: maps are dicts whose keys are the keyword names without the colon ("state-vector"), keyword VALUES are
`Keyword` (a str subclass, so `op["operation-type"] == "h"` holds), state vectors are complex128 arrays.  The JSON form is
what `clojure.data.json/write-str` produces (keywords -> plain strings); like the reference's JSON import
(`io/json.clj:22-25`, `:key-fn keyword`) reading it back leaves `:operation-type` a string, which every consumer here
accepts.  Host-side only: no arithmetic, no GPU."""
from __future__ import annotations

import json
import re
from typing import Any, Dict, List

import numpy as np

FORMAT_VERSION = "1.0"


class Keyword(str):
    """An EDN keyword value (`:h`); compares equal to its name without the colon."""
    __slots__ = ()

    def __repr__(self):
        return ":" + str.__str__(self)


# ------------------------------------------------------------------ EDN (the subset pr-str emits for these maps)
_TOKEN = re.compile(r"""[\s,]*(~@|[\[\]{}()]|#\{|"(?:\\.|[^\\"])*"|;[^\n]*|[^\s\[\]{}()"`,;]+)""")
_INT = re.compile(r"^[+-]?\d+N?$")
_FLOAT = re.compile(r"^[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)M?$")
_RATIO = re.compile(r"^[+-]?\d+/\d+$")
_ESC = {"n": "\n", "t": "\t", "r": "\r", '"': '"', "\\": "\\"}


def _tokens(text: str):
    pos = 0
    while True:
        m = _TOKEN.match(text, pos)
        if not m or m.end() == pos:
            break
        pos = m.end()
        tok = m.group(1)
        if tok and not tok.startswith(";"):
            yield tok
    if text[pos:].strip(" \t\r\n,"):
        raise ValueError(f"EDN: cannot tokenise near {text[pos:pos + 20]!r}")


def _atom(tok: str):
    if tok.startswith('"'):
        return re.sub(r"\\(.)", lambda m: _ESC.get(m.group(1), m.group(1)), tok[1:-1])
    if tok.startswith(":"):
        return Keyword(tok[1:])
    if tok == "nil":
        return None
    if tok == "true":
        return True
    if tok == "false":
        return False
    if _INT.match(tok):
        return int(tok.rstrip("N"))
    if _FLOAT.match(tok):
        return float(tok.rstrip("M"))
    if _RATIO.match(tok):                      # Clojure ratio literal, e.g. an angle written as 1/2
        num, den = tok.split("/")
        return int(num) / int(den)
    if tok.startswith("\\") and len(tok) > 1:  # character literal
        return {"\\newline": "\n", "\\space": " ", "\\tab": "\t"}.get(tok, tok[1:])
    if tok in ("##Inf", "##-Inf", "##NaN"):
        return {"##Inf": float("inf"), "##-Inf": float("-inf"), "##NaN": float("nan")}[tok]
    return tok                      # symbol


def _read(it, tok):
    if tok == "[" or tok == "(":
        close = "]" if tok == "[" else ")"
        out = []
        for t in it:
            if t == close:
                return out
            out.append(_read(it, t))
        raise ValueError("EDN: unterminated vector")
    if tok == "#{":
        out = []
        for t in it:
            if t == "}":
                return set(out)
            out.append(_read(it, t))
        raise ValueError("EDN: unterminated set")
    if tok == "{":
        items = []
        for t in it:
            if t == "}":
                if len(items) % 2:
                    raise ValueError("EDN: map with an odd number of forms")
                return {(str(k) if isinstance(k, Keyword) else k): v for k, v in zip(items[::2], items[1::2])}
            items.append(_read(it, t))
        raise ValueError("EDN: unterminated map")
    if tok in ("]", ")", "}"):
        raise ValueError(f"EDN: unexpected {tok}")
    return _atom(tok)


def read_edn(text: str):
    """Parse one EDN form (maps -> dict with keyword names as str keys, vectors/lists -> list, keywords -> Keyword)."""
    it = _tokens(text)
    try:
        first = next(it)
    except StopIteration:
        raise ValueError("EDN: empty input") from None
    return _read(it, first)


def _fmt_double(x: float) -> str:
    if x != x:
        return "##NaN"
    if x in (float("inf"), float("-inf")):
        return "##Inf" if x > 0 else "##-Inf"
    r = repr(float(x))               # shortest round-trip form, like Java's Double.toString up to exponent style
    if "e" in r:
        mant, exp = r.split("e")
        if "." not in mant:
            mant += ".0"
        return f"{mant}E{int(exp)}"
    return r


def write_edn(v: Any) -> str:
    """`pr-str` of the value: dict keys are written as keywords, Keyword values as keywords, other str as strings."""
    if v is None:
        return "nil"
    if v is True:
        return "true"
    if v is False:
        return "false"
    if isinstance(v, Keyword):
        return ":" + str(v)
    if isinstance(v, str):
        return '"' + v.replace("\\", "\\\\").replace('"', '\\"').replace("\n", "\\n") + '"'
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if isinstance(v, (float, np.floating)):
        return _fmt_double(float(v))
    if isinstance(v, dict):
        return "{" + ", ".join(f":{k} {write_edn(x)}" if isinstance(k, str) else f"{write_edn(k)} {write_edn(x)}" for k, x in v.items()) + "}"
    if isinstance(v, (set, frozenset)):
        return "#{" + " ".join(write_edn(x) for x in sorted(v, key=str)) + "}"
    if isinstance(v, (list, tuple, np.ndarray)):
        return "[" + " ".join(write_edn(x) for x in v) + "]"
    raise TypeError(f"cannot write {type(v).__name__} as EDN")


# ------------------------------------------------------------------ io.clj:10-89
def complex_to_map(z) -> Dict[str, float]:
    return {"real": float(np.real(z)), "imag": float(np.imag(z))}


def map_to_complex(m) -> complex:
    return complex(m["real"], m["imag"])


def _get(d, key, default=None):
    return d.get(key, d.get(":" + key, default))


def serialize_quantum_state(state: dict) -> dict:
    """io.clj:33-51.  `state` = {"state-vector": array-like of complex, "num-qubits": n, optional "metadata"}."""
    vec = np.asarray(_get(state, "state-vector"), dtype=np.complex128).reshape(-1)
    return {"state-vector": [complex_to_map(z) for z in vec], "num-qubits": int(_get(state, "num-qubits")),
            "metadata": _get(state, "metadata") or {}, "format-version": FORMAT_VERSION}


def deserialize_quantum_state(data: dict) -> dict:
    """io.clj:53-66."""
    vec = np.array([map_to_complex(m) for m in _get(data, "state-vector")], dtype=np.complex128)
    n = int(_get(data, "num-qubits"))
    if vec.shape[0] != 1 << n:
        raise ValueError(f"state vector of length {vec.shape[0]} does not match {n} qubits")
    return {"state-vector": vec, "num-qubits": n, "metadata": _get(data, "metadata")}


def _kw_op(op: dict) -> dict:
    params = _get(op, "operation-params") or {}
    return {"operation-type": Keyword(str(_get(op, "operation-type")).lstrip(":")),
            "operation-params": {str(k).lstrip(":"): v for k, v in params.items()}}


def serialize_quantum_circuit(circuit: dict) -> dict:
    """io.clj:68-89."""
    return {"operations": [_kw_op(op) for op in _get(circuit, "operations")], "num-qubits": int(_get(circuit, "num-qubits")),
            "name": _get(circuit, "name"), "description": _get(circuit, "description"),
            "metadata": _get(circuit, "metadata") or {}, "format-version": FORMAT_VERSION}


def deserialize_quantum_circuit(data: dict) -> dict:
    c = {"operations": [_kw_op(op) for op in _get(data, "operations")], "num-qubits": int(_get(data, "num-qubits"))}
    for k in ("name", "description", "metadata"):
        if _get(data, k) is not None:
            c[k] = _get(data, k)
    return c


def serialize_quantum_data(data: dict) -> dict:
    """io.clj:90-101: dispatch on the shape of the value."""
    if _get(data, "state-vector") is not None:
        return serialize_quantum_state(data)
    if _get(data, "operations") is not None:
        return serialize_quantum_circuit(data)
    raise ValueError("Unsupported quantum data type")


def deserialize_quantum_data(data: dict) -> dict:
    """io.clj:103-114."""
    if _get(data, "state-vector") is not None:
        return deserialize_quantum_state(data)
    if _get(data, "operations") is not None:
        return deserialize_quantum_circuit(data)
    raise ValueError("Unsupported quantum data format")


# ------------------------------------------------------------------ io/edn.clj, io/json.clj
def _fmt(fmt: str) -> str:
    f = str(fmt).lstrip(":").lower()
    if f not in ("edn", "json"):
        raise ValueError(f"unsupported format {fmt!r} (edn, json)")
    return f


def _dump(fmt: str, value: dict) -> str:
    if _fmt(fmt) == "edn":
        return write_edn(value)
    return json.dumps(value, separators=(",", ":"))


def _load(fmt: str, text: str) -> dict:
    return read_edn(text) if _fmt(fmt) == "edn" else json.loads(text)


def export_quantum_state(fmt: str, state: dict, filename: str):
    """Returns the file name like the reference's `save-file` - except for :edn, where the reference's method returns nil
    (`((io/serialize-quantum-state state) (qio/save-file ...))`, io/edn.clj:12-15; doc/tutorial.md prints `nil` there)."""
    with open(filename, "w") as f:
        f.write(_dump(fmt, serialize_quantum_state(state)))
    return None if _fmt(fmt) == "edn" else filename


def import_quantum_state(fmt: str, filename: str) -> dict:
    with open(filename) as f:
        return deserialize_quantum_state(_load(fmt, f.read()))


def export_quantum_circuit(fmt: str, circuit: dict, filename: str) -> str:
    f3 = str(fmt).lstrip(":").lower()
    if f3 in ("qasm2", "qasm3"):                             # adapter/io/qasm.clj:12-27
        from . import qasm2, qasm3
        return (qasm3 if f3 == "qasm3" else qasm2).export_quantum_circuit(circuit, filename)
    with open(filename, "w") as f:
        f.write(_dump(fmt, serialize_quantum_circuit(circuit)))
    return filename


def import_quantum_circuit(fmt: str, filename: str) -> dict:
    f3 = str(fmt).lstrip(":").lower()
    if f3 in ("qasm2", "qasm3"):
        from . import qasm2, qasm3
        return (qasm3 if f3 == "qasm3" else qasm2).import_quantum_circuit(filename)
    with open(filename) as f:
        return deserialize_quantum_circuit(_load(fmt, f.read()))


def export_quantum_data(fmt: str, data: dict, filename: str) -> str:
    with open(filename, "w") as f:
        f.write(_dump(fmt, serialize_quantum_data(data)))
    return filename


def import_quantum_data(fmt: str, filename: str) -> dict:
    with open(filename) as f:
        return deserialize_quantum_data(_load(fmt, f.read()))


def state_from_backend_result(result: dict) -> dict:
    """{:state-vector :num-qubits} out of a backend result's :final-state (backend.py / ideal_simulator.clj:84-96)."""
    fs = result["results"]["final-state"] if "results" in result else result["final-state"]
    return {"state-vector": np.asarray(fs["state-vector"]), "num-qubits": fs["num-qubits"]}


__all__: List[str] = [n for n in dir() if not n.startswith("_")]
