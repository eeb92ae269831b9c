"""Noise-model -> qcb_noise_table conversion (host side of the hardware-simulator path).

Restates the parameter handling of the reference's `apply-gate-noise`
(src/org/soulspace/qclojure/domain/noise.clj:65-103) and the Kraus generators of
domain/channel.clj:52-121, 259-264: the table handed to `qcb_run_noisy` holds, per gate kind, the Kraus
matrices with their coefficients.  Selection among several operators (max |coeff|^2 rule,
channel.clj:225-242) and application + renormalisation happen inside the library.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import ops as OPS


def _kw(x):
    return x[1:] if isinstance(x, str) and x.startswith(":") else x


def _get(d, key, default=None):
    if d is None:
        return default
    if key in d:
        return d[key]
    if ":" + key in d:
        return d[":" + key]
    return default


_I = np.eye(2, dtype=np.complex128)
_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)


def depolarizing_kraus(p: float) -> List[np.ndarray]:
    """channel.clj:40-60."""
    a, b = math.sqrt(1.0 - p), math.sqrt(p / 3.0)
    return [a * _I, b * _X, b * _Y, b * _Z]


def amplitude_damping_kraus(gamma: float) -> List[np.ndarray]:
    """channel.clj:62-79."""
    return [np.array([[1.0, 0], [0, math.sqrt(1.0 - gamma)]], dtype=np.complex128),
            np.array([[0, math.sqrt(gamma)], [0, 0]], dtype=np.complex128)]


def phase_damping_kraus(gamma: float) -> List[np.ndarray]:
    """channel.clj:81-99."""
    return [np.array([[1.0, 0], [0, math.sqrt(1.0 - gamma)]], dtype=np.complex128),
            np.array([[0, 0], [0, math.sqrt(gamma)]], dtype=np.complex128)]


def coherent_error_kraus(angle: float, axis: str) -> np.ndarray:
    """channel.clj:101-121 (x/y: real rotation matrices; z: diag(cos a, cos(-a)))."""
    c, s = math.cos(angle / 2.0), math.sin(angle / 2.0)
    axis = _kw(axis)
    if axis == "x":
        return np.array([[c, -s], [s, c]], dtype=np.complex128)
    if axis == "y":
        return np.array([[c, s], [-s, c]], dtype=np.complex128)
    if axis == "z":
        return np.array([[math.cos(angle), 0], [0, math.cos(-angle)]], dtype=np.complex128)
    raise ValueError(f"unknown rotation axis {axis}")


def decoherence_params(t1: float, t2: float, gate_time: float) -> Dict[str, float]:
    """channel.clj:259-264 — T in microseconds, gate time in nanoseconds."""
    gt = gate_time / 1000.0
    return {"gamma-1": 1.0 - math.exp(-(gt / t1)), "gamma-2": 1.0 - math.exp(-(gt / t2))}


def kraus_for_gate_noise(cfg: dict) -> Optional[List[np.ndarray]]:
    """noise.clj:72-101 for one {:noise-type ...} entry."""
    ntype = _kw(_get(cfg, "noise-type"))
    t1, t2, gt = _get(cfg, "t1-time"), _get(cfg, "t2-time"), _get(cfg, "gate-time")
    strength = _get(cfg, "noise-strength", 0.01)
    if ntype == "depolarizing":
        return depolarizing_kraus(strength)
    if ntype == "amplitude-damping":
        g = decoherence_params(t1, t2 or t1, gt)["gamma-1"] if (t1 and gt) else strength
        return amplitude_damping_kraus(g)
    if ntype == "phase-damping":
        g = decoherence_params(t1 or t2, t2, gt)["gamma-2"] if (t2 and gt) else strength
        return phase_damping_kraus(g)
    if ntype == "coherent":
        cc = _get(cfg, "coherent-error") or {"rotation-angle": 0.01, "rotation-axis": "z"}
        return [coherent_error_kraus(_get(cc, "rotation-angle"), _get(cc, "rotation-axis"))]
    return None


def build_noise_table(noise_model: dict, n_qubits: int) -> Tuple[Optional[OPS.QcbNoiseTable], list]:
    """Returns (table, keepalive).  table is None for an empty noise model."""
    gate_noise = _get(noise_model, "gate-noise") or {}
    entries = []
    for gate, cfg in gate_noise.items():
        name = _kw(gate)
        if name not in OPS.KIND:          # noise is looked up by the un-aliased :operation-type (noise.clj:69-71)
            continue
        ks = kraus_for_gate_noise(cfg)
        if not ks:
            continue
        e = OPS.QcbNoiseEntry()
        e.op_kind, e.n_kraus = OPS.KIND[name], len(ks)
        for k, K in enumerate(ks):
            flat = np.asarray(K, dtype=np.complex128).reshape(4)
            for i in range(4):
                e.kraus[k][2 * i] = float(flat[i].real)
                e.kraus[k][2 * i + 1] = float(flat[i].imag)
        entries.append(e)
    ro = _get(noise_model, "readout-error")
    if not entries and not ro:
        return None, []
    arr = (OPS.QcbNoiseEntry * max(1, len(entries)))(*entries)
    t = OPS.QcbNoiseTable()
    t.entries, t.n_entries = arr, len(entries)
    keep = [arr]
    if ro:
        t.has_readout = 1
        t.prob_0_to_1 = float(_get(ro, "prob-0-to-1"))
        t.prob_1_to_0 = float(_get(ro, "prob-1-to-0"))
        corr = _get(ro, "correlated-errors")
        # only the nested form {src {dst factor}} has an effect in the reference (noise.clj:136-145)
        if isinstance(corr, dict) and any(isinstance(v, dict) for v in corr.values()):
            m = np.ones((n_qubits, n_qubits), dtype=np.float64)
            for src, row in corr.items():
                if not isinstance(row, dict):
                    continue
                for dst, f in row.items():
                    si, di = int(src), int(dst)
                    if 0 <= si < n_qubits and 0 <= di < n_qubits:
                        m[si, di] = float(f)
            keep.append(m)
            t.correlation = m.ctypes.data_as(C.POINTER(C.c_double))
    return t, keep
