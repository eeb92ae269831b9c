"""Host-side mirror of the reference's backend protocol on top of libqcb200.so.

`B200Simulator` mirrors `LocalQuantumSimulator` (src/org/soulspace/qclojure/adapter/backend/
ideal_simulator.clj:100-176) and `B200HardwareSimulator` mirrors `QuantumHardwareSimulator`
(adapter/backend/hardware_simulator.clj:281-391): the eight `QuantumBackend` protocol methods
(application/backend.clj:72-112) with the same names (kebab-case -> snake_case), argument meaning,
result-map keys and error behaviour, so that the parity tests read like the reference's own tests.
Keyword keys are plain strings without the colon ("job-status", "measurement-results", ...).

Differences that are deliberate and documented in DESIGN.md:
  * randomness: the reference draws from Math/random with no seeding hook; here draws come from
    `options["uniforms"]` (explicit array) or `options["seed"]`/config seed through NumPy's PCG64, so a run
    is reproducible and can be compared shot by shot with the oracle;
  * `:final-state` / `:measurement-probabilities` are returned as NumPy arrays up to `max_state_qubits`
    (default 26) and as a lazy device handle above it (a 2^30-entry Clojure vector is not representable);
  * small jobs (<= 20 qubits) complete inside `submit_circuit`, so the first status poll already sees
    "completed" (the reference's blocking helper sleeps 100 ms between polls, backend.clj:255).
"""
from __future__ import annotations

import itertools
import threading
import time
from typing import Dict, List, Optional

import numpy as np

from . import _lib as L
from . import noise as NZ
from . import ops as OPS
from . import results as RS

_job_counter = itertools.count(1)
_jobs_lock = threading.Lock()
# process-global job table shared by all simulator instances, like the reference's atom
# (ideal_simulator.clj:56-57)
_JOBS: Dict[str, dict] = {}

NATIVE_GATES = sorted(["i", "x", "y", "z", "h", "s", "s-dag", "t", "t-dag", "rx", "ry", "rz", "phase", "cnot", "cz", "cy",
                       "swap", "crx", "cry", "crz", "toffoli", "fredkin"])     # operation_registry.clj:380-383


def _kw(x):
    return x[1:] if isinstance(x, str) and x.startswith(":") else x


def _opt(d: Optional[dict], key: str, default=None):
    if not d:
        return default
    if key in d:
        return d[key]
    if ":" + key in d:
        return d[":" + key]
    return default


def circuit_metadata(circuit: dict) -> dict:
    """circuit-depth / operation-count / gate-count (domain/circuit.clj:1370-1440; greedy layering)."""
    ops = OPS.circuit_ops(circuit)
    n = OPS.circuit_num_qubits(circuit)
    layers = [0] * n
    gates = 0
    for op in ops:
        typ, p = OPS.normalize_op(op)
        qs = []
        for k in ("target", "control", "control1", "control2", "target1", "target2", "qubit1", "qubit2"):
            if p.get(k) is not None:
                qs.append(int(p[k]))
        for k in ("measurement-qubits", "qubit-indices"):
            if p.get(k) is not None:
                qs += [int(q) for q in p[k]]
        if typ.startswith("global-"):
            qs = list(range(n))
        if typ != "measure":
            gates += 1
        if not qs:
            continue
        lvl = max(layers[q] for q in qs) + 1
        for q in qs:
            layers[q] = lvl
    return {"circuit-depth": max(layers) if ops else 0, "circuit-operation-count": len(ops), "circuit-gate-count": gates}


class StaleStateHandle(RuntimeError):
    """The device state behind a DeviceStateHandle has been reused by a later job (or released)."""


class DeviceStateHandle:
    """Lazy view of a final state that is too large to materialise as a host vector.  The view is valid until the backend
    runs its next job on the same state vector (HBM holds one such state); afterwards every access raises
    `StaleStateHandle` instead of silently returning the later job's amplitudes."""

    def __init__(self, sv: L.StateVector):
        self._sv = sv
        self.num_qubits = sv.n
        self._generation = getattr(sv, "generation", 0)

    def _live(self) -> L.StateVector:
        if getattr(self._sv, "_h", None) is None or getattr(self._sv, "generation", 0) != self._generation:
            raise StaleStateHandle("the device state of this result has been reused by a later job; results of jobs above "
                                   "max-state-qubits must be read before the next job runs")
        return self._sv

    def slice(self, offset: int, count: int) -> np.ndarray:
        return self._live().get_state(offset, count)

    def amplitudes(self, indices) -> np.ndarray:
        return self._live().get_amplitudes(indices)

    def probabilities(self, offset: int, count: int) -> np.ndarray:
        return self._live().probabilities(offset, count)


class _DrawSource:
    """One stream of uniform draws per job: consecutive, non-overlapping slices of `options["uniforms"]`, or one generator
    seeded from `options["seed"]` (else the backend's), so that mid-circuit :measure draws, the shot draws, :sample draws and
    the noisy trajectory matrix never reuse the same numbers inside a job."""

    def __init__(self, options, fallback_rng):
        u = _opt(options, "uniforms")
        self._u = None if u is None else np.asarray(u, dtype=np.float64).reshape(-1)
        self._pos = 0
        seed = _opt(options, "seed")
        self._rng = np.random.default_rng(seed) if seed is not None else fallback_rng

    def take(self, shape):
        cnt = int(np.prod(shape))
        if self._u is not None:
            if self._pos + cnt > self._u.size:
                raise ValueError(f"not enough uniforms supplied: {self._u.size} given, {self._pos + cnt} needed so far")
            out = self._u[self._pos:self._pos + cnt].reshape(shape)
            self._pos += cnt
            return out
        return self._rng.random(shape)


class _BackendBase:
    backend_type = "simulator"

    def __init__(self, config: Optional[dict] = None):
        self.config = dict(config or {})
        self._svs: Dict[int, L.StateVector] = {}
        self._lock = threading.RLock()
        self._rng = np.random.default_rng(_opt(self.config, "seed"))
        self.max_state_qubits = int(_opt(self.config, "max-state-qubits", 26))
        self._workers: List[threading.Thread] = []

    # -- handles are cached per qubit count (one state vector resident in HBM each)
    def _sv(self, n: int) -> L.StateVector:
        """The state vector a job runs on.  config "n-gpus" (2, 4, 8; optional "device-ids") shards states of more than
        "multi-gpu-min-qubits" (default 31) qubits over the GPUs of this process behind ONE handle."""
        sv = self._svs.get(n)
        if sv is None:
            for old in list(self._svs):            # keep HBM for the one in use
                self._svs.pop(old).close()
            n_gpus = int(_opt(self.config, "n-gpus", 0) or 0)
            if n_gpus > 1 and n < int(_opt(self.config, "multi-gpu-min-qubits", 31)):
                n_gpus = 0
            sv = L.StateVector(n, device=int(_opt(self.config, "device", -1)),
                               fusion=int(_opt(self.config, "fusion", 1)),
                               strict_parity=int(_opt(self.config, "strict-parity", 1)),
                               n_gpus=n_gpus, device_ids=_opt(self.config, "device-ids"))
            sv.generation = 0
            self._svs[n] = sv
        sv.generation = getattr(sv, "generation", 0) + 1     # invalidates the lazy handles of earlier jobs' results
        return sv

    def close(self):
        for t in self._workers:
            t.join()
        for sv in self._svs.values():
            sv.close()
        self._svs.clear()

    # -- QuantumBackend protocol (application/backend.clj:72-112)
    def available(self) -> bool:
        return L.device_count() > 0

    def job_status(self, job_id: str) -> str:
        with _jobs_lock:
            job = _JOBS.get(job_id)
        return job["status"] if job else "not-found"

    def job_result(self, job_id: str) -> dict:
        with _jobs_lock:
            job = _JOBS.get(job_id)
        if not job:
            return {"job-id": job_id, "job-status": "not-found", "error-message": "Job not found"}
        if job["status"] == "completed":
            return dict(job["result"], **{"job-id": job_id})
        out = {"job-id": job_id, "job-status": job["status"], "error-message": "Job not completed"}
        if job.get("result") and job["result"].get("error-message"):
            out["failure-message"] = job["result"]["error-message"]        # superset: the reference drops it
            out["exception-type"] = job["result"].get("exception-type")
        return out

    def cancel_job(self, job_id: str) -> str:
        with _jobs_lock:
            job = _JOBS.get(job_id)
            if not job:
                return "not-found"
            if job["status"] in ("queued", "running"):
                job["status"] = "cancelled"
                job["completed-at"] = time.time()
                return "cancelled"
            return "cannot-cancel"

    def queue_status(self) -> dict:
        with _jobs_lock:
            jobs = list(_JOBS.values())
        cnt = lambda s: sum(1 for j in jobs if j["status"] == s)   # noqa: E731
        return {"total-jobs": len(jobs), "queued": cnt("queued"), "running": cnt("running"), "completed": cnt("completed"),
                "backend-load": 0.0, "estimated-wait-time": 0}

    def submit_circuit(self, circuit: dict, options: Optional[dict] = None) -> str:
        options = options or {}
        job_id = f"{self._job_prefix}_{next(_job_counter)}_{int(time.time() * 1000)}"
        job = {"job-id": job_id, "circuit": circuit, "options": options, "status": "queued", "result": None,
               "created-at": time.time(), "completed-at": None}
        with _jobs_lock:
            _JOBS[job_id] = job

        def work():
            with _jobs_lock:
                if job["status"] == "cancelled":
                    return
                job["status"] = "running"
            result = self._execute_guarded(circuit, options)
            with _jobs_lock:
                if job["status"] != "cancelled":
                    job["status"] = result["job-status"]
                    job["result"] = result
                    job["completed-at"] = time.time()

        if OPS.circuit_num_qubits(circuit) <= 20:
            work()
        else:
            t = threading.Thread(target=work, daemon=True)
            self._workers.append(t)
            t.start()
        return job_id

    def _execute_guarded(self, circuit, options) -> dict:
        # never throw out of the worker (ideal_simulator.clj:93-96, hardware_simulator.clj:187-191)
        t0 = time.time()
        try:
            with self._lock:
                options = dict(options or {})
                options["__draws"] = _DrawSource(options, self._rng)
                res = self._execute(circuit, options)
            res["execution-time-ms"] = int((time.time() - t0) * 1000)
            return res
        except Exception as e:      # noqa: BLE001
            return {"job-status": "failed", "error-message": str(e), "exception-type": type(e).__name__}

    def _uniforms(self, options, shape):
        src = options.get("__draws") if isinstance(options, dict) else None
        if src is None:
            src = _DrawSource(options, self._rng)
            if isinstance(options, dict):
                options["__draws"] = src
        return src.take(shape)


# =============================================================================== ideal simulator
class B200Simulator(_BackendBase):
    """Mirror of LocalQuantumSimulator (ideal_simulator.clj:100-176) running on the B200."""

    _job_prefix = "sim_job"

    def backend_info(self) -> dict:
        return {"backend-type": "simulator", "backend-name": "B200 State-Vector Simulator",
                "description": "B200-native fp64 state-vector simulator (libqcb200.so) behind the QuantumBackend protocol",
                "backend-config": self.config, "max-qubits": int(_opt(self.config, "max-qubits", 33)),
                "capabilities": {"quantum-backend"}, "device": self.device(), "version": "0.1.0"}

    def device(self) -> dict:
        return {"id": "b200-simulator", "name": "B200 Ideal Quantum Simulator", "provider": "qclojure_b200",
                "platform": "Local", "technology": "simulator", "num-qubits": 33, "topology": "all-to-all",
                "connectivity": "full", "native-gates": NATIVE_GATES, "virtual-gates": [],
                "supported-operations": NATIVE_GATES, "measurement-basis": "any", "noise-model": {},
                "performance": {"gate-fidelity": 1.0, "readout-fidelity": 1.0}}

    # circuit/execute-circuit + result/extract-results (domain/circuit.clj:1778-1792, domain/result.clj:535-639)
    def _execute(self, circuit: dict, options: dict) -> dict:
        n = OPS.circuit_num_qubits(circuit)
        ops = OPS.circuit_ops(circuit)
        specs = _opt(options, "result-specs") or {}
        sv = self._sv(n)
        init = _opt(options, "initial-state")
        if init is not None:
            vec = init.get("state-vector", init.get(":state-vector")) if isinstance(init, dict) else init
            sv.set_state(np.asarray(vec, dtype=np.complex128))
        else:
            sv.set_zero()
        # :measure ops consume one draw each (state.clj:981)
        segs = OPS.split_at_measurements(ops)
        n_meas = sum(1 for k, _ in segs if k == "measure")
        mdraws = iter(self._uniforms(options, (n_meas,)).tolist()) if n_meas else iter(())
        for kind, payload in segs:
            if kind == "gates":
                sv.apply_ops(payload)
            else:
                sv.measure_qubits(payload, next(mdraws))
        results: dict = {"result-types": sorted(_kw(k) for k in specs), "circuit": circuit,
                         "circuit-metadata": circuit_metadata(circuit)}
        results["final-state"] = ({"state-vector": sv.get_state(), "num-qubits": n} if n <= self.max_state_qubits
                                  else DeviceStateHandle(sv))
        results.update(RS.extract_results(sv, specs, lambda shape: self._uniforms(options, shape),
                                          max_state_qubits=self.max_state_qubits, state_handle=DeviceStateHandle(sv)))
        return {"job-status": "completed", "results": results}


def create_simulator(config: Optional[dict] = None) -> B200Simulator:
    """Mirror of `create-simulator` (ideal_simulator.clj:181-195)."""
    return B200Simulator(config)


# =============================================================================== hardware (noisy) simulator
class B200HardwareSimulator(_BackendBase):
    """Mirror of QuantumHardwareSimulator (hardware_simulator.clj:281-391): per-shot trajectories with gate
    noise (Kraus channels) and readout noise from a device noise profile."""

    _job_prefix = "job"
    backend_type = "hardware-simulator"

    def __init__(self, device: Optional[dict] = None, config: Optional[dict] = None):
        super().__init__(config)
        self._device = device or {"id": "b200-hardware-simulator", "noise-model": {}}
        # the device catalogue (hardware_simulator.clj:44-52 loads resources/simulator-devices.edn): a list of device maps
        # under config "devices", or the path of such an EDN file under "devices-file"
        self._devices: List[dict] = list(_opt(self.config, "devices") or [])
        if _opt(self.config, "devices-file"):
            self._devices += load_device_catalog(_opt(self.config, "devices-file"))
        if device is not None and all(_opt(d, "id") != _opt(device, "id") for d in self._devices):
            self._devices.append(device)                     # create-hardware-simulator adds it (:407-411)

    def backend_info(self) -> dict:
        # hardware_simulator.clj:284-290: :backend-type :backend-name :devices :device :config :capabilities #{:multi-device}
        return {"backend-type": "hardware-simulator", "backend-name": "B200 Noisy Quantum Hardware Simulator",
                "backend-config": self.config, "config": self.config, "max-qubits": int(_opt(self.config, "max-qubits", 26)),
                "capabilities": {"quantum-backend", "multi-device"}, "device": self._device, "devices": self.devices(),
                "version": "0.1.0"}

    def queue_status(self) -> dict:
        """hardware_simulator.clj:372-381 reports :total-jobs :active-jobs :completed-jobs (the ideal simulator's keys are kept
        as well)."""
        qs = super().queue_status()
        qs["active-jobs"] = qs["queued"] + qs["running"]
        qs["completed-jobs"] = qs["completed"]
        return qs

    def device(self) -> dict:
        return self._device

    # MultiDeviceBackend (application/backend.clj:117-131, hardware_simulator.clj:382-391)
    def devices(self) -> List[dict]:
        return list(self._devices)

    def select_device(self, device):
        """A device map, or the id of a catalogue entry.  (The reference's keyword branch looks up `:device` instead of
        the id and selects nil, hardware_simulator.clj:387-389; here the id is resolved, unknown ids raise.)"""
        if not isinstance(device, dict):
            want = _kw(str(device))
            found = [d for d in self._devices if _kw(str(_opt(d, "id"))) == want]
            if not found:
                raise KeyError(f"unknown device {device!r}")
            device = found[0]
        self._device = device
        return device

    def _execute(self, circuit: dict, options: dict) -> dict:
        n = OPS.circuit_num_qubits(circuit)
        ops = OPS.circuit_ops(circuit)
        shots = int(_opt(options, "shots", 1024))                      # hardware_simulator.clj:123
        max_traj = int(_opt(options, "max-trajectories", 100))
        noise_model = _opt(self._device, "noise-model") or {}          # options :noise-model is ignored (:228)
        specs = _opt(options, "result-specs")
        sv = self._sv(n)
        table, keep = NZ.build_noise_table(noise_model, n)
        enc = OPS.encode_ops(ops)
        dps = sv.noisy_draws_per_shot(enc, table)
        init = _opt(options, "initial-state")                          # (or (:initial-state options) zero-state), :128-131
        if init is not None:
            vec = init.get("state-vector", init.get(":state-vector")) if isinstance(init, dict) else init
            sv.noisy_set_initial_state(np.asarray(vec, dtype=np.complex128))
        else:
            sv.noisy_set_initial_state(None)
        u = self._uniforms(options, (shots, dps))
        outcomes, traj = sv.run_noisy(enc, table, u, max_trajectories=max_traj if n <= 20 else 0)
        counts: Dict[str, int] = {}
        for o in outcomes.tolist():
            bs = format(o, f"0{n}b")
            counts[bs] = counts.get(bs, 0) + 1
        results: dict = {"measurement-results": counts,
                         "final-state": {"state-vector": sv.get_state(), "num-qubits": n} if n <= self.max_state_qubits else DeviceStateHandle(sv)}
        if traj is not None and len(traj):
            results["trajectories"] = [{"state-vector": t, "num-qubits": n} for t in traj]
            results["trajectory-count"] = len(traj)
            results["trajectory-weights"] = [1.0 / len(traj)] * len(traj)
            if n <= 10:       # rho = mean projector (state.clj:781-786); 4^n entries, small n only
                rho = sum(np.outer(t, np.conj(t)) for t in traj) / len(traj)
                results["density-matrix"] = rho
                results["density-matrix-trace"] = float(np.trace(rho).real)
        if specs:                                                      # result.clj:642-804
            results["shots-executed"] = shots
            dev = int(_opt(self.config, "device", -1))
            results = RS.extract_noisy_results(results, specs, n, lambda: L.StateVector(n, device=dev),
                                               lambda shape: self._uniforms(options, shape))
            results.pop("shots-executed", None)
        return {"job-status": "completed", "circuit": circuit, "circuit-metadata": circuit_metadata(circuit),
                "shots-executed": shots, "results": results}


def load_device_catalog(path: str) -> List[dict]:
    """Reads a device catalogue in the reference's format (`resources/simulator-devices.edn`: a vector of device maps with
    `:id :num-qubits :native-gates :coupling :noise-model {:gate-noise {...} :readout-error {...}} ...`)."""
    from . import io as QIO
    with open(path) as f:
        data = QIO.read_edn(f.read())
    if isinstance(data, dict):
        data = list(data.values())
    return [d for d in data if isinstance(d, dict)]


def create_hardware_simulator(device: Optional[dict] = None, config: Optional[dict] = None) -> B200HardwareSimulator:
    """Mirror of `create-hardware-simulator` (hardware_simulator.clj:396-414)."""
    return B200HardwareSimulator(device, config)


# =============================================================================== blocking helper
def execute_circuit(backend: _BackendBase, circuit: dict, options: Optional[dict] = None, *, poll_s: float = 0.1,
                    max_polls: int = 600) -> dict:
    """Mirror of `backend/execute-circuit` (application/backend.clj:209-256): submit, poll every 100 ms up to
    600 times, return the result map (status first checked immediately, so small jobs return without sleeping)."""
    if not backend.available():
        return {"job-status": "failed", "error-message": "Backend is not available"}
    job_id = backend.submit_circuit(circuit, options or {})
    for _ in range(max_polls):
        st = backend.job_status(job_id)
        if st in ("completed", "failed", "cancelled"):
            return backend.job_result(job_id)
        time.sleep(poll_s)
    return {"job-status": "failed", "job-id": job_id, "error-message": "Job timed out"}
