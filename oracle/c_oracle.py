"""ctypes binding of oracle/_build/libqcoracle.so (the C restatement, see qc_oracle.c).

TEST INFRASTRUCTURE ONLY — see oracle/qc_oracle.py.  The op encoder here is deliberately independent
of qclojure_b200/ops.py so that an encoding bug in the product shows up as a parity failure.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libqcoracle.so")


class Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("q", C.c_int32 * 3), ("n_mask", C.c_int32), ("_pad", C.c_int32),
                ("mask", C.c_uint64), ("angle", C.c_double), ("mat", C.c_double * 8), ("ext", C.c_void_p)]


KINDS = ["i", "x", "y", "z", "h", "s", "s-dag", "t", "t-dag", "rx", "ry", "rz", "phase",
         "cnot", "cz", "cy", "crx", "cry", "crz", "swap", "iswap", "toffoli", "fredkin",
         "rydberg-cz", "rydberg-cphase", "rydberg-blockade",
         "global-h", "global-x", "global-y", "global-z", "global-rx", "global-ry", "global-rz",
         "u1q", "cu1q", "u2q", "mcphase", "phase-oracle", "grover-diffusion"]
KIND = {k: i for i, k in enumerate(KINDS)}
ALIASES = {"not": "x", "bit-flip": "x", "phase-flip": "z", "id": "i", "cx": "cnot", "ccx": "toffoli",
           "ccnot": "toffoli", "cswap": "fredkin", "p": "phase", "u1": "phase", "sdg": "s-dag",
           "tdg": "t-dag", "phaseshift": "phase", "si": "s-dag", "ti": "t-dag"}


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "qc_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_apply_ops.restype = C.c_int
        _lib.orc_apply_ops.argtypes = [C.c_void_p, C.c_int, C.POINTER(Op), C.c_uint64]
        _lib.orc_probabilities.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.orc_sample.restype = C.c_int
        _lib.orc_sample.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        _lib.orc_norm2.restype = C.c_double
        _lib.orc_norm2.argtypes = [C.c_void_p, C.c_int]
        _lib.orc_expect_pauli.restype = C.c_double
        _lib.orc_expect_pauli.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        _lib.orc_apply_1q_dense_kron.restype = C.c_int
        _lib.orc_apply_1q_dense_kron.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_set_num_threads.argtypes = [C.c_int]
    return _lib


def _kw(x):
    return x[1:] if isinstance(x, str) and x.startswith(":") else x


def encode(circuit_or_ops, n: int):
    ops = circuit_or_ops.get("operations", circuit_or_ops.get(":operations")) if isinstance(circuit_or_ops, dict) else circuit_or_ops
    arr = (Op * len(ops))()
    for k, op in enumerate(ops):
        typ = _kw(op.get("operation-type", op.get(":operation-type")))
        typ = ALIASES.get(typ, typ)
        p = {_kw(a): b for a, b in (op.get("operation-params", op.get(":operation-params")) or {}).items()}
        o = arr[k]
        o.kind = KIND[typ]
        o.q[0] = o.q[1] = o.q[2] = -1
        if "angle" in p and p["angle"] is not None:
            o.angle = float(p["angle"])
        if typ in ("cnot", "cz", "cy", "crx", "cry", "crz", "rydberg-cz", "rydberg-cphase"):
            o.q[0], o.q[1] = p["control"], p["target"]
        elif typ in ("swap", "iswap"):
            o.q[0], o.q[1] = p["qubit1"], p["qubit2"]
        elif typ == "toffoli":
            o.q[0], o.q[1], o.q[2] = p["control1"], p["control2"], p["target"]
        elif typ == "fredkin":
            o.q[0], o.q[1], o.q[2] = p["control"], p["target1"], p["target2"]
        elif typ == "rydberg-blockade":
            m = 0
            for q in p["qubit-indices"]:
                m |= 1 << q
            o.mask, o.n_mask = m, len(p["qubit-indices"])
        elif typ.startswith("global-"):
            pass
        else:
            t = p.get("target")
            o.q[0] = 0 if t is None else t
    return arr


def apply_circuit(circuit: dict, state: np.ndarray | None = None) -> np.ndarray:
    n = circuit.get("num-qubits", circuit.get(":num-qubits"))
    if state is None:
        state = np.zeros(1 << n, dtype=np.complex128)
        state[0] = 1.0
    else:
        state = np.ascontiguousarray(state, dtype=np.complex128).copy()
    arr = encode(circuit, n)
    rc = lib().orc_apply_ops(state.ctypes.data, n, arr, len(arr))
    if rc != 0:
        raise RuntimeError(f"orc_apply_ops failed: {rc}")
    return state


def sample(state: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
    n = int(round(math.log2(state.shape[0])))
    u = np.ascontiguousarray(uniforms, dtype=np.float64)
    out = np.zeros(u.shape[0], dtype=np.uint64)
    cum = np.empty(state.shape[0], dtype=np.float64)
    rc = lib().orc_sample(state.ctypes.data, n, u.ctypes.data, u.shape[0], out.ctypes.data, cum.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"orc_sample failed: {rc}")
    return out.astype(np.int64)


def expect_pauli(state: np.ndarray, pauli: str) -> float:
    n = int(round(math.log2(state.shape[0])))
    return float(lib().orc_expect_pauli(state.ctypes.data, n, pauli.encode()))


def norm2(state: np.ndarray) -> float:
    n = int(round(math.log2(state.shape[0])))
    return float(lib().orc_norm2(state.ctypes.data, n))


def apply_1q_dense_kron(state: np.ndarray, target: int, mat: np.ndarray) -> np.ndarray:
    n = int(round(math.log2(state.shape[0])))
    st = np.ascontiguousarray(state, dtype=np.complex128).copy()
    m = np.ascontiguousarray(mat, dtype=np.complex128)
    rc = lib().orc_apply_1q_dense_kron(st.ctypes.data, n, target, m.ctypes.data)
    if rc != 0:
        raise MemoryError(f"orc_apply_1q_dense_kron: {rc}")
    return st


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> int:
    """Sets the OpenMP thread count of the C oracle (e.g. to os.cpu_count() under torchrun, which exports
    OMP_NUM_THREADS=1) and returns what the runtime reports afterwards."""
    lib().orc_set_num_threads(int(n))
    return num_threads()
