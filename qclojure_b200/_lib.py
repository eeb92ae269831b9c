"""ctypes loader of libqcb200.so (built in-tree by qclojure_b200/csrc/Makefile).

There is no fallback: if the shared library is missing, `load()` raises; if it is present but no
CUDA device is usable, `qcb_create` fails with QCB_ERR_CUDA and `StateVector(...)` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

from . import ops as OPS

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QCB_LIB") or os.path.join(HERE, "lib", "libqcb200.so")   # QCB_LIB: profiling build (make PROFILE=1)
CSRC = os.path.join(HERE, "csrc")

QCB_OK = 0
ERR_NAMES = {-1: "QCB_ERR_INVALID", -2: "QCB_ERR_UNSUPPORTED", -3: "QCB_ERR_CUDA", -4: "QCB_ERR_NOMEM",
             -5: "QCB_ERR_NCCL", -6: "QCB_ERR_STATE", -7: "QCB_ERR_NOTFOUND"}
JOB_STATUS = {0: "queued", 1: "running", 2: "completed", 3: "failed", 4: "cancelled", 5: "not-found"}


class QcbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {message}")
        self.code = code
        self.message = message


def build(force: bool = False) -> str:
    """Compile libqcb200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in ("kernels.cu", "sim.cu", "plan.cpp", "la_host.cpp", "kernels.h", "plan.h",
                                             "tile_core.h", "la_host.h")]
    srcs.append(os.path.join(os.path.dirname(HERE), "include", "qcb200.h"))
    stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", CSRC, "-s"] + (["-B"] if force else []))
    return LIB_PATH


_lib = None

_P = C.POINTER
_PROTOS = {
    "qcb_abi_version": (C.c_int32, []),
    "qcb_last_error": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "qcb_device_count": (C.c_int32, [_P(C.c_int32)]),
    "qcb_nccl_unique_id": (C.c_int32, [C.c_void_p]),
    "qcb_config_default": (C.c_int32, [_P(OPS.QcbConfig)]),
    "qcb_create": (C.c_int32, [_P(OPS.QcbConfig), _P(C.c_void_p)]),
    "qcb_destroy": (C.c_int32, [C.c_void_p]),
    "qcb_synchronize": (C.c_int32, [C.c_void_p]),
    "qcb_set_zero": (C.c_int32, [C.c_void_p]),
    "qcb_set_basis": (C.c_int32, [C.c_void_p, C.c_uint64]),
    "qcb_set_state": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "qcb_get_state": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "qcb_get_amplitudes": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "qcb_normalize": (C.c_int32, [C.c_void_p]),
    "qcb_state_dev_ptr": (C.c_int32, [C.c_void_p, _P(C.c_void_p), _P(C.c_uint64)]),
    "qcb_apply_ops": (C.c_int32, [C.c_void_p, _P(OPS.QcbOp), C.c_uint64]),
    "qcb_norm2": (C.c_int32, [C.c_void_p, _P(C.c_double)]),
    "qcb_probabilities": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "qcb_sample": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "qcb_measure_qubits": (C.c_int32, [C.c_void_p, _P(C.c_int32), C.c_int32, C.c_double, _P(C.c_int32), _P(C.c_double)]),
    "qcb_marginal_probabilities": (C.c_int32, [C.c_void_p, _P(C.c_int32), C.c_int32, C.c_void_p]),
    "qcb_expect_pauli": (C.c_int32, [C.c_void_p, C.c_char_p, _P(C.c_double)]),
    "qcb_expect_hamiltonian": (C.c_int32, [C.c_void_p, C.c_void_p, _P(C.c_char_p), C.c_uint64, _P(C.c_double), C.c_void_p]),
    "qcb_expect_1q": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, _P(C.c_double)]),
    "qcb_fidelity": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, _P(C.c_double)]),
    "qcb_apply_kraus_1q": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32]),
    "qcb_noisy_set_initial_state": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "qcb_noisy_draws_per_shot": (C.c_int32, [C.c_void_p, _P(OPS.QcbOp), C.c_uint64, _P(OPS.QcbNoiseTable), _P(C.c_uint64)]),
    "qcb_run_noisy": (C.c_int32, [C.c_void_p, _P(OPS.QcbOp), C.c_uint64, _P(OPS.QcbNoiseTable), C.c_void_p, C.c_uint64,
                                  C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]),
    "qcb_get_stats": (C.c_int32, [C.c_void_p, _P(OPS.QcbStats)]),
    "qcb_timer_start": (C.c_int32, [C.c_void_p]),
    "qcb_timer_stop": (C.c_int32, [C.c_void_p, _P(C.c_double)]),
    "qcb_plan_create": (C.c_int32, [_P(OPS.QcbConfig), _P(OPS.QcbOp), C.c_uint64, _P(C.c_void_p)]),
    "qcb_plan_create_replayed": (C.c_int32, [_P(OPS.QcbConfig), _P(OPS.QcbOp), _P(OPS.QcbOp), C.c_uint64, _P(C.c_void_p)]),
    "qcb_plan_destroy": (C.c_int32, [C.c_void_p]),
    "qcb_plan_serialize": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, _P(C.c_uint64)]),
    "qcb_plan_summary": (C.c_int32, [C.c_void_p, _P(C.c_uint64), _P(C.c_uint64), _P(C.c_uint64)]),
    "qcb_submit": (C.c_int32, [C.c_void_p, _P(OPS.QcbJobRequest), _P(C.c_uint64)]),
    "qcb_job_status": (C.c_int32, [C.c_void_p, C.c_uint64, _P(C.c_int32)]),
    "qcb_job_result_get": (C.c_int32, [C.c_void_p, C.c_uint64, _P(OPS.QcbJobResult)]),
    "qcb_cancel": (C.c_int32, [C.c_void_p, C.c_uint64, _P(C.c_int32)]),
    "qcb_job_release": (C.c_int32, [C.c_void_p, C.c_uint64]),
    "qcb_queue_status": (C.c_int32, [C.c_void_p, _P(C.c_uint64), _P(C.c_uint64), _P(C.c_uint64)]),
    "qcb_la_matvec": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "qcb_la_matmul": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]),
    "qcb_la_kron": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "qcb_la_inner": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "qcb_la_outer": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "qcb_la_trace": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "qcb_la_norm2": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, _P(C.c_double)]),
    "qcb_la_axpby": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    # host-side protocol methods (la_host.cpp): handle may be NULL
    "qcb_la_hadamard": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "qcb_la_transpose": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32, C.c_void_p]),
    "qcb_la_solve": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "qcb_la_inverse": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "qcb_la_is_hermitian": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, _P(C.c_int32)]),
    "qcb_la_is_diagonal": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, _P(C.c_int32)]),
    "qcb_la_is_unitary": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, _P(C.c_int32)]),
    "qcb_la_is_positive_semidefinite": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, _P(C.c_int32)]),
    "qcb_la_eigen_hermitian": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "qcb_la_eigen_general": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "qcb_la_svd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qcb_la_lu": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qcb_la_qr": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "qcb_la_cholesky": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "qcb_la_matrix_exp": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "qcb_la_matrix_log": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "qcb_la_matrix_sqrt": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "qcb_la_spectral_norm": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, _P(C.c_double)]),
    "qcb_la_condition_number": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, _P(C.c_double)]),
}
EXPORTED_SYMBOLS = tuple(_PROTOS)


def load():
    """Load libqcb200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(the CUDA library is the product; there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def last_error(handle) -> str:
    buf = C.create_string_buffer(1024)
    load().qcb_last_error(handle, buf, len(buf))
    return buf.value.decode(errors="replace")


def check(rc: int, handle=None):
    if rc != QCB_OK:
        raise QcbError(rc, last_error(handle))


def device_count() -> int:
    n = C.c_int32(0)
    rc = load().qcb_device_count(C.byref(n))
    return int(n.value) if rc == QCB_OK else 0


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(load().qcb_nccl_unique_id(buf))
    return buf.raw


def _c128(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.complex128)


class StateVector:
    """One n-qubit fp64 complex state vector resident in HBM: on one GPU, as one rank's slice of an SPMD job
    (world_size > 1, one process per GPU), or - n_gpus > 1 - the WHOLE state sharded over the GPUs of this process behind
    one handle (offsets, counts and results are then global).

    Thin 1:1 wrapper over the C ABI; qubit numbering is the reference's (qubit 0 = MSB)."""

    def __init__(self, n_qubits: int, *, device: int = -1, fusion: int = 1, strict_parity: int = 1, tile_bits: int = 0,
                 low_bits: int = 0, rank: int = 0, world_size: int = 1, nccl_id: Optional[bytes] = None, max_stage_cost: int = 0,
                 max_stage_rounds: int = 0, dense_mma: int = 0, tile_mover: int = 0, n_gpus: int = 0,
                 device_ids: Optional[Sequence[int]] = None):
        self._lib = load()
        if dense_mma == 0 and os.environ.get("QCB_DENSE_MMA"):
            dense_mma = int(os.environ["QCB_DENSE_MMA"])     # 1 = tensor-core rounds (default), 2 = interpreter only
        if tile_mover == 0 and os.environ.get("QCB_TILE_MOVER"):
            tile_mover = int(os.environ["QCB_TILE_MOVER"])   # 1 = cp.async mover (default), 2 = TMA mover
        self.n = int(n_qubits)
        self.rank, self.world_size = rank, world_size
        self._idbuf = C.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
        cfg = OPS.make_config(self.n, device=device, fusion=fusion, strict_parity=strict_parity, tile_bits=tile_bits,
                              low_bits=low_bits, rank=rank, world_size=world_size,
                              nccl_unique_id=C.cast(self._idbuf, C.c_void_p) if self._idbuf is not None else None,
                              max_stage_cost=max_stage_cost, max_stage_rounds=max_stage_rounds, dense_mma=dense_mma,
                              tile_mover=tile_mover, n_gpus=n_gpus, device_ids=device_ids)
        self.n_gpus = n_gpus
        h = C.c_void_p()
        rc = self._lib.qcb_create(C.byref(cfg), C.byref(h))
        if rc != QCB_OK:
            raise QcbError(rc, last_error(None))
        self._h = h
        p = world_size.bit_length() - 1
        self.local_count = 1 << (self.n - p)      # amplitudes this handle addresses (a multi-GPU handle: all of them)

    # -- lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self._lib.qcb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        check(rc, self._h)

    def synchronize(self):
        self._ck(self._lib.qcb_synchronize(self._h))

    # -- state
    def set_zero(self):
        self._ck(self._lib.qcb_set_zero(self._h))

    def set_basis(self, index: int):
        self._ck(self._lib.qcb_set_basis(self._h, index))

    def set_state(self, amps):
        a = _c128(amps)
        self._ck(self._lib.qcb_set_state(self._h, a.ctypes.data, a.shape[0]))

    def get_state(self, offset: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.local_count - offset if count is None else count
        out = np.empty(count, dtype=np.complex128)
        self._ck(self._lib.qcb_get_state(self._h, offset, count, out.ctypes.data))
        return out

    def get_amplitudes(self, indices: Sequence[int]) -> np.ndarray:
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        out = np.empty(idx.shape[0], dtype=np.complex128)
        self._ck(self._lib.qcb_get_amplitudes(self._h, idx.ctypes.data, idx.shape[0], out.ctypes.data))
        return out

    def normalize(self):
        self._ck(self._lib.qcb_normalize(self._h))

    def dev_ptr(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self._lib.qcb_state_dev_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    # -- gates
    def apply_ops(self, ops):
        """ops: list of QClojure gate maps, or a pre-encoded (array, count, keepalive) triple."""
        if isinstance(ops, tuple):
            arr, cnt, keep = ops
        else:
            arr, cnt, keep = OPS.encode_ops(ops)
        self._ck(self._lib.qcb_apply_ops(self._h, arr, cnt))
        return self

    def apply_circuit(self, circuit: dict):
        return self.apply_ops(OPS.circuit_ops(circuit))

    # -- measurement
    def norm2(self) -> float:
        v = C.c_double()
        self._ck(self._lib.qcb_norm2(self._h, C.byref(v)))
        return float(v.value)

    def probabilities(self, offset: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.local_count - offset if count is None else count
        out = np.empty(count, dtype=np.float64)
        self._ck(self._lib.qcb_probabilities(self._h, offset, count, out.ctypes.data))
        return out

    def sample(self, uniforms) -> np.ndarray:
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        out = np.zeros(u.shape[0], dtype=np.uint64)
        self._ck(self._lib.qcb_sample(self._h, u.ctypes.data, u.shape[0], out.ctypes.data))
        return out.astype(np.int64)

    def measure_qubits(self, qubits: Sequence[int], u: float):
        q = (C.c_int32 * len(qubits))(*[int(x) for x in qubits])
        bits = (C.c_int32 * len(qubits))()
        p = C.c_double()
        self._ck(self._lib.qcb_measure_qubits(self._h, q, len(qubits), float(u), bits, C.byref(p)))
        return list(bits), float(p.value)

    def marginal_probabilities(self, qubits: Sequence[int]) -> np.ndarray:
        q = (C.c_int32 * len(qubits))(*[int(x) for x in qubits])
        out = np.empty(1 << len(qubits), dtype=np.float64)
        self._ck(self._lib.qcb_marginal_probabilities(self._h, q, len(qubits), out.ctypes.data))
        return out

    # -- expectation
    def expect_pauli(self, pauli: str) -> float:
        v = C.c_double()
        self._ck(self._lib.qcb_expect_pauli(self._h, pauli.encode(), C.byref(v)))
        return float(v.value)

    def expect_hamiltonian(self, hamiltonian, return_terms: bool = False):
        coeffs = np.ascontiguousarray([t.get("coefficient", t.get(":coefficient")) for t in hamiltonian], dtype=np.float64)
        strs = [t.get("pauli-string", t.get(":pauli-string")).encode() for t in hamiltonian]
        arr = (C.c_char_p * max(1, len(strs)))(*strs)
        e = C.c_double()
        terms = np.empty(len(strs), dtype=np.float64)
        self._ck(self._lib.qcb_expect_hamiltonian(self._h, coeffs.ctypes.data, arr, len(strs), C.byref(e), terms.ctypes.data))
        return (float(e.value), terms) if return_terms else float(e.value)

    def expect_1q(self, observable, target: int) -> float:
        m = _c128(np.asarray(observable).reshape(4))
        v = C.c_double()
        self._ck(self._lib.qcb_expect_1q(self._h, m.ctypes.data, int(target), C.byref(v)))
        return float(v.value)

    def fidelity(self, reference) -> float:
        a = _c128(reference)
        v = C.c_double()
        self._ck(self._lib.qcb_fidelity(self._h, a.ctypes.data, a.shape[0], C.byref(v)))
        return float(v.value)

    # -- noise
    def apply_kraus_1q(self, matrix, target: int):
        m = _c128(np.asarray(matrix).reshape(4))
        self._ck(self._lib.qcb_apply_kraus_1q(self._h, m.ctypes.data, int(target)))

    def noisy_draws_per_shot(self, ops_enc, noise_table) -> int:
        arr, cnt, _ = ops_enc
        out = C.c_uint64()
        self._ck(self._lib.qcb_noisy_draws_per_shot(self._h, arr, cnt, C.byref(noise_table) if noise_table is not None else None, C.byref(out)))
        return int(out.value)

    def noisy_set_initial_state(self, amps=None):
        """Initial state of the trajectories of the following run_noisy calls (None = |0...0>)."""
        if amps is None:
            self._ck(self._lib.qcb_noisy_set_initial_state(self._h, None, 0))
        else:
            a = _c128(amps)
            self._ck(self._lib.qcb_noisy_set_initial_state(self._h, a.ctypes.data, a.shape[0]))

    def run_noisy(self, ops_enc, noise_table, uniforms: np.ndarray, max_trajectories: int = 0):
        arr, cnt, _ = ops_enc
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        shots, dps = u.shape
        outcomes = np.zeros(shots, dtype=np.uint64)
        ntraj = min(shots, max_trajectories)
        traj = np.empty((ntraj, self.local_count), dtype=np.complex128) if ntraj else None
        self._ck(self._lib.qcb_run_noisy(self._h, arr, cnt, C.byref(noise_table) if noise_table is not None else None,
                                         u.ctypes.data, dps, shots, outcomes.ctypes.data,
                                         traj.ctypes.data if traj is not None else None, ntraj))
        return outcomes.astype(np.int64), traj

    def timer_start(self):
        self._ck(self._lib.qcb_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._ck(self._lib.qcb_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def stats(self) -> dict:
        s = OPS.QcbStats()
        self._ck(self._lib.qcb_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in OPS.QcbStats._fields_}


def plan_summary(n_qubits: int, ops, **cfgkw) -> dict:
    """Host-only: what the scheduler does with an op list (no GPU needed)."""
    lib = load()
    cfg = OPS.make_config(n_qubits, **cfgkw)
    arr, cnt, keep = OPS.encode_ops(ops)
    p = C.c_void_p()
    rc = lib.qcb_plan_create(C.byref(cfg), arr, cnt, C.byref(p))
    if rc != QCB_OK:
        raise QcbError(rc, last_error(None))
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.qcb_plan_summary(p, C.byref(a), C.byref(b), C.byref(c))
    nw = C.c_uint64()
    lib.qcb_plan_serialize(p, None, 0, C.byref(nw))
    words = np.empty(int(nw.value), dtype=np.uint64)
    lib.qcb_plan_serialize(p, words.ctypes.data_as(C.c_void_p), nw.value, C.byref(nw))
    lib.qcb_plan_destroy(p)
    # passes over the shared tile (csrc/plan.h: program layout): a paired pass (round kind 3) carries two dense rounds
    sweeps = passes = pairs = 0
    per_sweep = []                                       # (passes, paired passes) of every tile sweep
    pos = 4
    for _ in range(int(words[1])):
        kind = int(words[pos]); pos += 2
        if kind == 3:                                    # S_GROVER: marked indices follow the header
            pos += int(words[pos - 1]) & 0xff
        if kind != 0:
            continue
        nr = int(words[pos + 3])
        sweeps += 1; passes += nr
        np_ = sum(1 for r in range(nr) if int(words[pos + 48 + 40 * r + 17]) == 3)
        pairs += np_
        per_sweep.append((nr, np_))
        pos += int(words[pos + 40])
    return {"stages": int(a.value), "rounds": int(b.value), "exchanges": int(c.value), "program_words": int(nw.value),
            "tile_sweeps": sweeps, "passes": passes, "paired_passes": pairs, "per_sweep": per_sweep}
