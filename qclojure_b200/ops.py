"""ctypes mirror of include/qcb200.h structs + the encoder from QClojure circuit maps to qcb_op[].

The encoder follows the reference's dispatcher `apply-gate-to-state`
(src/org/soulspace/qclojure/domain/circuit.clj:952-1072): alias resolution
(domain/operation_registry.clj:387-407), the parameter keys each gate reads, "missing :target
defaults to qubit 0" for the single-qubit gates (circuit.clj:965-984) and the error raised for
missing operands.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Sequence, Tuple

import numpy as np


class QcbConfig(C.Structure):
    _fields_ = [("n_qubits", C.c_int32), ("device", C.c_int32), ("fusion", C.c_int32),
                ("strict_parity", C.c_int32), ("tile_bits", C.c_int32), ("low_bits", C.c_int32),
                ("rank", C.c_int32), ("world_size", C.c_int32), ("nccl_unique_id", C.c_void_p),
                ("max_stage_cost", C.c_int32), ("max_stage_rounds", C.c_int32), ("dense_mma", C.c_int32), ("tile_mover", C.c_int32),
                ("n_gpus", C.c_int32), ("device_ids", C.c_int32 * 8), ("reserved", C.c_int32 * 3)]


class QcbOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("q", C.c_int32 * 3), ("n_mask", C.c_int32), ("_pad", C.c_int32),
                ("mask", C.c_uint64), ("angle", C.c_double), ("mat", C.c_double * 8), ("ext", C.c_void_p)]


class QcbNoiseEntry(C.Structure):
    _fields_ = [("op_kind", C.c_int32), ("n_kraus", C.c_int32), ("kraus", (C.c_double * 8) * 4)]


class QcbNoiseTable(C.Structure):
    _fields_ = [("entries", C.POINTER(QcbNoiseEntry)), ("n_entries", C.c_int32), ("has_readout", C.c_int32),
                ("prob_0_to_1", C.c_double), ("prob_1_to_0", C.c_double), ("correlation", C.POINTER(C.c_double))]


class QcbStats(C.Structure):
    _fields_ = [("n_ops", C.c_uint64), ("n_gates_lowered", C.c_uint64), ("n_sweeps", C.c_uint64),
                ("n_rounds", C.c_uint64), ("n_kernel_launches", C.c_uint64), ("n_exchanges", C.c_uint64),
                ("bytes_exchanged", C.c_uint64), ("algorithmic_bytes", C.c_double), ("unfused_bytes", C.c_double),
                ("gpu_ms", C.c_double), ("exchange_ms", C.c_double)]


class QcbJobRequest(C.Structure):
    _fields_ = [("ops", C.POINTER(QcbOp)), ("n_ops", C.c_uint64),
                ("initial_state", C.POINTER(C.c_double)), ("initial_count", C.c_uint64),
                ("uniforms", C.POINTER(C.c_double)), ("n_shots", C.c_uint64),
                ("ham_coeffs", C.POINTER(C.c_double)), ("ham_strings", C.POINTER(C.c_char_p)), ("n_terms", C.c_uint64),
                ("want_probabilities", C.c_int32), ("want_state", C.c_int32)]


class QcbJobResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("execution_time_ms", C.c_double),
                ("n_shots", C.c_uint64), ("outcomes", C.POINTER(C.c_uint64)),
                ("energy", C.c_double), ("has_energy", C.c_int32),
                ("probabilities", C.POINTER(C.c_double)), ("prob_capacity", C.c_uint64),
                ("state", C.POINTER(C.c_double)), ("state_capacity", C.c_uint64),
                ("error_message", C.c_char * 256)]


# enum qcb_op_kind, in header order
KIND_NAMES = ["i", "x", "y", "z", "h", "s", "s-dag", "t", "t-dag", "rx", "ry", "rz", "phase",
              "cnot", "cz", "cy", "crx", "cry", "crz", "swap", "iswap", "toffoli", "fredkin",
              "rydberg-cz", "rydberg-cphase", "rydberg-blockade",
              "global-h", "global-x", "global-y", "global-z", "global-rx", "global-ry", "global-rz",
              "u1q", "cu1q", "u2q", "mcphase", "phase-oracle", "grover-diffusion", "measure"]
KIND = {name: i for i, name in enumerate(KIND_NAMES)}

# domain/operation_registry.clj:387-407
GATE_ALIASES = {"not": "x", "bit-flip": "x", "phase-flip": "z", "id": "i", "cx": "cnot", "ccx": "toffoli",
                "ccnot": "toffoli", "cswap": "fredkin", "p": "phase", "u1": "phase", "sdg": "s-dag",
                "tdg": "t-dag", "phaseshift": "phase", "si": "s-dag", "ti": "t-dag"}

_ONE_QUBIT = {"i", "x", "y", "z", "h", "s", "s-dag", "t", "t-dag", "rx", "ry", "rz", "phase"}
_ANGLE_1Q = {"rx", "ry", "rz", "phase"}
_CTRL = {"cnot", "cz", "cy", "rydberg-cz"}
_CTRL_ANGLE = {"crx", "cry", "crz", "rydberg-cphase"}
_GLOBAL_ANGLE = {"global-rx", "global-ry", "global-rz"}
_GLOBAL = {"global-h", "global-x", "global-y", "global-z"}


class GateError(ValueError):
    """Mirrors the ex-info thrown by apply-gate-to-state for malformed or unknown gates."""


def _kw(x):
    return x[1:] if isinstance(x, str) and x.startswith(":") else x


def normalize_op(op: dict) -> Tuple[str, dict]:
    typ = op.get("operation-type")
    if typ is None:
        typ = op.get(":operation-type")
    if type(typ) is not str or typ[:1] == ":":
        typ = _kw(typ)
    params = op.get("operation-params")
    if params is None:
        params = op.get(":operation-params") or {}
    for k in params:                      # keys spelled without the colon (the common case) are used as they are
        if type(k) is str and k[:1] == ":":
            return typ, {_kw(k): v for k, v in params.items()}
    return typ, params


def circuit_ops(circuit: dict) -> list:
    return circuit.get("operations", circuit.get(":operations"))


def circuit_num_qubits(circuit: dict) -> int:
    return int(circuit.get("num-qubits", circuit.get(":num-qubits")))


def split_at_measurements(ops: Sequence[dict]) -> List[Tuple[str, object]]:
    """Split an op list into runs of unitary gates and `:measure` ops (circuit.clj:1106-1111)."""
    out: List[Tuple[str, object]] = []
    run: List[dict] = []
    for op in ops:
        typ, p = normalize_op(op)
        if typ == "measure":
            if run:
                out.append(("gates", run))
                run = []
            qs = p.get("measurement-qubits")
            if qs is None:
                raise GateError("Measure requires measurement-qubits parameter")
            out.append(("measure", list(qs)))
        else:
            run.append(op)
    if run:
        out.append(("gates", run))
    return out


# numpy view of qcb_op (same layout as the ctypes struct): columns are filled in bulk, which is what makes encoding a
# 1000-gate circuit cost a few hundred microseconds instead of milliseconds
OP_DTYPE = np.dtype({"names": ["kind", "q", "n_mask", "mask", "angle", "mat", "ext"],
                     "formats": [np.int32, (np.int32, 3), np.int32, np.uint64, np.float64, (np.float64, 8), np.uint64],
                     "offsets": [QcbOp.kind.offset, QcbOp.q.offset, QcbOp.n_mask.offset, QcbOp.mask.offset, QcbOp.angle.offset,
                                 QcbOp.mat.offset, QcbOp.ext.offset],
                     "itemsize": C.sizeof(QcbOp)})
_K1 = {g: KIND[g] for g in _ONE_QUBIT}
_K2 = {g: KIND[g] for g in _CTRL}
_K2A = {g: KIND[g] for g in _CTRL_ANGLE}


def encode_ops(ops: Iterable[dict]):
    """QClojure gate maps -> (ctypes array of qcb_op, count, keepalive list)."""
    ops = list(ops)
    n = len(ops)
    arr = (QcbOp * max(1, n))()
    keep = []
    kinds, q0, q1, q2, angles = [0] * n, [-1] * n, [-1] * n, [-1] * n, [0.0] * n
    slow = []                                    # (index, gate name, params): kinds that carry more than qubits + angle
    for k, op in enumerate(ops):
        typ, p = normalize_op(op)
        g = GATE_ALIASES.get(typ, typ)
        get = p.get
        if g in _K1:
            kinds[k] = _K1[g]
            t = get("target")
            q0[k] = 0 if t is None else int(t)            # (or target 0), circuit.clj:977-984
            if g in _ANGLE_1Q:
                a = get("angle")
                if a is None:
                    raise GateError(f"{g} requires angle")
                angles[k] = float(a)
        elif g in _K2:
            c, t = get("control"), get("target")
            if c is None or t is None:
                raise GateError(f"{g} requires control, target")
            kinds[k], q0[k], q1[k] = _K2[g], int(c), int(t)
        elif g in _K2A:
            c, t, a = get("control"), get("target"), get("angle")
            if c is None or t is None or a is None:
                raise GateError(f"{g} requires control, target, angle")
            kinds[k], q0[k], q1[k], angles[k] = _K2A[g], int(c), int(t), float(a)
        elif g not in KIND:
            raise GateError(f"Unknown gate type {g}")
        else:
            kinds[k] = KIND[g]
            slow.append((k, g, p))
    if n:
        view = np.frombuffer(arr, dtype=OP_DTYPE, count=n)
        view["kind"] = kinds
        qv = view["q"]
        qv[:, 0] = q0
        qv[:, 1] = q1
        qv[:, 2] = q2
        view["angle"] = angles
    for k, g, p in slow:
        o = arr[k]
        angle = p.get("angle")

        def need(*names):
            for nm in names:
                if p.get(nm) is None:
                    raise GateError(f"{g} requires {', '.join(names)}")

        if g in ("swap", "iswap"):
            need("qubit1", "qubit2")
            o.q[0], o.q[1] = int(p["qubit1"]), int(p["qubit2"])
        elif g == "toffoli":
            need("control1", "control2", "target")
            o.q[0], o.q[1], o.q[2] = int(p["control1"]), int(p["control2"]), int(p["target"])
        elif g == "fredkin":
            need("control", "target1", "target2")
            o.q[0], o.q[1], o.q[2] = int(p["control"]), int(p["target1"]), int(p["target2"])
        elif g in ("rydberg-blockade", "mcphase"):
            need("qubit-indices", "angle")
            m = 0
            for q in p["qubit-indices"]:
                m |= 1 << int(q)
            o.mask, o.n_mask, o.angle = m, len(p["qubit-indices"]), float(angle)
        elif g in _GLOBAL_ANGLE:
            need("angle")
            o.angle = float(angle)
        elif g in _GLOBAL:
            pass
        elif g == "u1q":
            need("target", "matrix")
            o.q[0] = int(p["target"])
            _put_mat(o, p["matrix"])
        elif g == "cu1q":
            need("control", "target", "matrix")
            o.q[0], o.q[1] = int(p["control"]), int(p["target"])
            _put_mat(o, p["matrix"])
        elif g == "u2q":
            need("qubit1", "qubit2", "matrix")
            o.q[0], o.q[1] = int(p["qubit1"]), int(p["qubit2"])
            m = np.ascontiguousarray(np.asarray(p["matrix"], dtype=np.complex128).reshape(4, 4))
            keep.append(m)
            o.ext = m.ctypes.data
        elif g == "phase-oracle":
            need("index")
            o.mask = int(p["index"])
        elif g == "grover-diffusion":
            pass
        elif g == "measure":
            need("measurement-qubits")
            qs = np.ascontiguousarray(np.asarray(p["measurement-qubits"], dtype=np.int32))
            keep.append(qs)
            o.ext = qs.ctypes.data
            o.n_mask = int(qs.shape[0])
            o.angle = float(p.get("uniform", 0.0))
    return arr, n, keep


def _put_mat(o: QcbOp, mat) -> None:
    m = np.asarray(mat, dtype=np.complex128).reshape(4)
    for i in range(4):
        o.mat[2 * i] = float(m[i].real)
        o.mat[2 * i + 1] = float(m[i].imag)


def make_config(n_qubits: int, *, device: int = -1, fusion: int = 1, strict_parity: int = 1, tile_bits: int = 0,
                low_bits: int = 0, rank: int = 0, world_size: int = 1, nccl_unique_id=None,
                max_stage_cost: int = 0, max_stage_rounds: int = 0, dense_mma: int = 0, tile_mover: int = 0,
                n_gpus: int = 0, device_ids=None) -> QcbConfig:
    cfg = QcbConfig()
    cfg.n_gpus = n_gpus
    for i in range(8):
        cfg.device_ids[i] = device_ids[i] if device_ids is not None and i < len(device_ids) else -1
    cfg.n_qubits, cfg.device, cfg.fusion, cfg.strict_parity = n_qubits, device, fusion, strict_parity
    cfg.tile_bits, cfg.low_bits, cfg.rank, cfg.world_size = tile_bits, low_bits, rank, world_size
    cfg.nccl_unique_id = nccl_unique_id
    cfg.max_stage_cost = max_stage_cost
    cfg.max_stage_rounds = max_stage_rounds
    cfg.dense_mma = dense_mma
    cfg.tile_mover = tile_mover
    return cfg
