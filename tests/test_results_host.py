"""Result extraction, host logic on CPU: `qclojure_b200/results.py` against the oracle with an
oracle-backed stand-in for the device state (tests/fake_sv.py).  The same extractors run against the real device state
in tests/test_gpu_parity.py::test_result_extraction_*.  Reference: src/org/soulspace/qclojure/domain/result.clj."""
import math
import os

import numpy as np
import pytest

from oracle import qc_oracle as O
from qclojure_b200 import circuits as C
from qclojure_b200 import results as RS
from tests.fake_sv import FakeLinearAlgebra, FakeStateVector

TOL = 1e-10


@pytest.fixture(autouse=True)
def _host_la(monkeypatch):
    monkeypatch.setattr(RS, "_la", FakeLinearAlgebra)


def _sv(circ, init=None):
    sv = FakeStateVector(circ["num-qubits"])
    if init is not None:
        sv.set_state(init)
    sv.apply_ops(circ["operations"])
    return sv


def _uniforms(seed):
    rng = np.random.default_rng(seed)
    return lambda shape: rng.random(shape)


def test_bits_and_labels():
    # test/.../domain/state_test.clj:197-227
    assert RS.bits_to_index([1, 0, 1]) == 5 and RS.bits_to_index([0, 0, 0]) == 0 and RS.bits_to_index([1, 1]) == 3
    assert RS.basis_labels(2) == ["|00⟩", "|01⟩", "|10⟩", "|11⟩"]


def test_hamiltonian_grouping():
    # test/.../domain/hamiltonian_test.clj (grouping): commuting terms share a group, bases by Pauli alphabet
    H = [{"coefficient": 1.0, "pauli-string": "ZZ"}, {"coefficient": 0.5, "pauli-string": "XX"},
         {"coefficient": 0.2, "pauli-string": "ZI"}, {"coefficient": 0.1, "pauli-string": "XI"},
         {"coefficient": 0.3, "pauli-string": "II"}, {"coefficient": 0.3, "pauli-string": "XY"}]
    groups = RS.group_commuting_terms(H)
    assert [[t["pauli-string"] for t in g] for g in groups] == [["ZZ", "XX", "ZI", "II", "XY"], ["XI"]]
    assert sum(len(g) for g in groups) == len(H)
    bases = RS.group_pauli_terms_by_measurement_basis(H)
    assert {k: [t["pauli-string"] for t in v] for k, v in bases.items()} == \
        {"z": ["ZZ", "ZI"], "x": ["XX", "XI"], "identity": ["II"], "mixed": ["XY"]}


def test_ideal_extraction_all_result_types():
    circ = C.ghz_state_circuit(3)
    C.ry(circ, 1, 0.7)
    sv = _sv(circ)
    psi = sv.get_state()
    H = [{"coefficient": 0.5, "pauli-string": "ZZI"}, {"coefficient": -0.25, "pauli-string": "XXX"}]
    ZZ = np.kron(np.kron(O.PAULI_Z, O.PAULI_Z), np.eye(2))
    specs = {"measurements": {"shots": 50}, "expectation": {"observables": [O.PAULI_Z, O.PAULI_X], "targets": [0, 1]},
             "variance": {"observables": [O.PAULI_Z], "targets": [2]}, "hamiltonian": H,
             "probabilities": {"targets": [[1, 1, 1], 0, [0, 1, 0]]}, "amplitudes": {"basis-states": [0, 7]},
             "state-vector": True, "density-matrix": True, "fidelity": {"references": [O.zero_state(3), psi]},
             "sample": {"observables": [O.PAULI_Z, ZZ], "shots": 40, "targets": [1, None]}}
    u = np.random.default_rng(3).random(50)
    draws = iter([u, np.random.default_rng(4).random((2, 40))])
    r = RS.extract_results(sv, specs, lambda shape: next(draws))
    m = r["measurement-results"]
    assert m["measurement-outcomes"] == O.sample_outcomes(psi, u).tolist() and m["shot-count"] == 50
    assert m["source"] == "ideal-simulation" and m["measurement-qubits"] == [0, 1, 2]
    assert sum(m["frequencies"].values()) == 50 and abs(sum(m["empirical-probabilities"].values()) - 1) < 1e-12
    assert np.allclose(m["measurement-probabilities"], np.abs(psi) ** 2, atol=TOL)
    e = r["expectation-results"]
    assert abs(e[0]["expectation-value"] - O.expectation_1q(psi, O.PAULI_Z, 0)) <= TOL and e[0]["target-qubits"] == [0]
    assert abs(e[1]["expectation-value"] - O.expectation_1q(psi, O.PAULI_X, 1)) <= TOL
    v = r["variance-results"][0]
    assert abs(v["variance-value"] - O.variance_1q(psi, O.PAULI_Z, 2)) <= TOL
    assert abs(v["standard-deviation"] - math.sqrt(v["variance-value"])) <= TOL
    h = r["hamiltonian-result"]
    assert abs(h["energy-expectation"] - O.hamiltonian_expectation(H, psi)) <= TOL
    assert h["hamiltonian"] is H and len(h["measurement-groups"]) == 1 and set(h["measurement-bases"]) == {"z", "x"}
    p = r["probability-results"]
    assert abs(p["probability-outcomes"][(1, 1, 1)] - abs(psi[7]) ** 2) <= TOL
    assert abs(p["probability-outcomes"][0] - abs(psi[0]) ** 2) <= TOL
    assert abs(p["probability-outcomes"][(0, 1, 0)] - abs(psi[2]) ** 2) <= TOL
    a = r["amplitude-results"]
    assert a["basis-states"] == [0, 7] and abs(a["amplitude-values"][7] - psi[7]) <= TOL
    s = r["state-vector-result"]
    assert np.array_equal(s["state-vector"], psi) and s["num-qubits"] == 3 and s["basis-labels"][5] == "|101⟩"
    d = r["density-matrix-result"]
    assert np.allclose(d["density-matrix"], np.outer(psi, psi.conj()), atol=TOL) and d["trace-valid"] is True
    f = r["fidelity-results"]["fidelities"]
    assert abs(f["reference-0"] - abs(psi[0])) <= TOL and abs(f["reference-1"] - 1.0) <= TOL
    sa = r["sample-results"]
    assert sa[0]["shot-count"] == 40 and set(sa[0]["frequencies"]) <= {-1.0, 1.0} and len(sa[0]["sample-outcomes"]) == 40
    assert set(sa[1]["frequencies"]) <= {-1.0, 1.0}          # ZZ: two doubly degenerate eigenvalues, probabilities summed


def test_all_probabilities_without_targets_and_full_register_observable():
    circ = C.bell_state_circuit()
    sv = _sv(circ)
    r = RS.extract_results(sv, {"probabilities": {"qubits": [0, 1]},
                                "expectation": {"observables": [np.kron(O.PAULI_Z, O.PAULI_Z)]}}, _uniforms(0))
    assert np.allclose(r["probability-results"]["all-probabilities"], [0.5, 0, 0, 0.5], atol=TOL)
    assert r["probability-results"]["probability-outcomes"][3] == pytest.approx(0.5, abs=TOL)
    # test/.../domain/hamiltonian_test.clj:93-96: <ZZ> = 1 on the Bell state
    assert r["expectation-results"][0]["expectation-value"] == pytest.approx(1.0, abs=TOL)
    assert r["expectation-results"][0]["target-qubits"] is None
    with pytest.raises(ValueError):
        RS.observable_expectation(sv, np.eye(8), None)


def test_observable_measurement_probabilities_and_sampling_rule():
    # observables.clj:325-355 doc example: pauli-z on |+> -> {1.0 0.5, -1.0 0.5}; eigenvalues ascending
    sv = FakeStateVector(1)
    sv.set_state(np.array([1, 1]) / math.sqrt(2))
    mp = RS.observable_measurement_probabilities(sv, O.PAULI_Z)
    assert list(mp) == pytest.approx([-1.0, 1.0]) and list(mp.values()) == pytest.approx([0.5, 0.5])
    # result.clj:476-486: first eigenvalue whose cumulative probability exceeds the draw, else the last
    assert RS.sample_eigenvalues({-1.0: 0.25, 1.0: 0.75}, [0.0, 0.2499, 0.25, 0.99, 1.0]) == [-1.0, -1.0, 1.0, 1.0, 1.0]
    # qubit 1 of |01> measured in Z is -1 with certainty
    sv2 = FakeStateVector(2)
    sv2.set_state(np.array([0, 1, 0, 0]))
    assert RS.observable_measurement_probabilities(sv2, O.PAULI_Z, 1) == pytest.approx({-1.0: 1.0, 1.0: 0.0})
    assert RS.observable_measurement_probabilities(sv2, O.PAULI_Z, 0) == pytest.approx({-1.0: 0.0, 1.0: 1.0})


def test_noisy_extraction_matches_density_matrix_formulas():
    """result.clj:642-804 on trajectories produced by the oracle's noisy shot loop: every Tr(rho O) quantity equals the
    reference's density-matrix formula, representative-state quantities come from sqrt(diag rho)."""
    import json
    import os
    prof = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "device_profiles.json")))
    nm = [d for d in prof["devices"] if d["id"] == ":ibm-lagos"][0]["noise_model"]
    n = 3
    circ = C.ghz_state_circuit(n)
    C.rx(circ, 2, 0.4)
    shots = 60
    u = np.random.default_rng(9).random((shots, O.draws_per_shot(circ, nm)))
    raw = O.run_noisy(circ, nm, u, max_trajectories=100)
    traj = raw["trajectories"]
    rho = O.trajectory_to_density_matrix(traj)
    base = {"measurement-results": raw["measurement-results"], "final-state": {"state-vector": raw["final-state"], "num-qubits": n},
            "trajectories": [{"state-vector": t, "num-qubits": n} for t in traj], "trajectory-count": len(traj),
            "density-matrix": rho, "density-matrix-trace": float(np.trace(rho).real), "shots-executed": shots}
    H = [{"coefficient": 1.0, "pauli-string": "ZZI"}, {"coefficient": 0.5, "pauli-string": "IXX"}]
    Hm = np.kron(np.kron(O.PAULI_Z, O.PAULI_Z), np.eye(2)) + 0.5 * np.kron(np.eye(2), np.kron(O.PAULI_X, O.PAULI_X))
    Z0 = np.kron(O.PAULI_Z, np.eye(4))
    specs = {"measurements": {"measurement-qubits": [0, 1, 2]}, "hamiltonian": {"hamiltonian": H},
             "expectation": {"observables": [O.PAULI_Z], "target-qubits": [0]},
             "variance": {"observables": [O.PAULI_Z], "target-qubits": [0]},
             "probability": {"target-states": [[0, 0, 0], 7]}, "amplitude": {"basis-states": [0, 7]},
             "state-vector": True, "density-matrix": True, "fidelity": {"reference-states": [O.zero_state(n)]},
             "sample": {"observables": [O.PAULI_Z], "shots": 30, "target-qubits": [2]}, "bogus": 1}
    r = RS.extract_noisy_results(base, specs, n, lambda: FakeStateVector(n), _uniforms(1))
    m = r["measurement-results"]
    assert m["source"] == "noisy-simulation" and m["shot-count"] == shots and m["frequencies"] == raw["measurement-results"]
    assert m["measurement-probabilities"] == m["empirical-probabilities"]
    assert r["hamiltonian-result"]["energy-expectation"] == pytest.approx(np.trace(rho @ Hm).real, abs=TOL)
    assert r["hamiltonian-result"]["source"] == "density-matrix"
    assert r["expectation-results"][0] == pytest.approx(np.trace(rho @ Z0).real, abs=TOL)
    var = np.trace(rho @ Z0 @ Z0).real - np.trace(rho @ Z0).real ** 2
    assert r["variance-results"][0]["variance-value"] == pytest.approx(var, abs=TOL)
    pops = np.real(np.diagonal(rho))
    assert r["probability-results"]["probability-outcomes"][(0, 0, 0)] == pytest.approx(pops[0], abs=TOL)
    assert r["probability-results"]["probability-outcomes"][7] == pytest.approx(pops[7], abs=TOL)
    assert r["amplitude-results"]["amplitude-values"][7] == pytest.approx(math.sqrt(pops[7]), abs=TOL)
    assert np.allclose(r["state-vector-result"]["state-vector"], np.sqrt(pops), atol=TOL)
    assert r["state-vector-result"]["source"] == "density-matrix-diagonal"
    assert r["density-matrix-result"]["from-trajectories"] is True and r["density-matrix-result"]["trajectory-count"] == len(traj)
    assert r["fidelity-results"]["fidelities"]["reference-0"] == pytest.approx(math.sqrt(pops[0]), abs=TOL)
    assert r["sample-results"][0]["shot-count"] == 30
    assert "bogus-error" not in r and r["result-types"] == sorted(specs)
    # the bare ideal spelling of :hamiltonian (variational_algorithm.clj:345) is accepted too
    r2 = RS.extract_noisy_results(base, {"hamiltonian": H}, n, lambda: FakeStateVector(n), _uniforms(1))
    assert r2["hamiltonian-result"]["energy-expectation"] == pytest.approx(np.trace(rho @ Hm).real, abs=TOL)
    # an extractor that fails records <type>-error instead of failing the job (result.clj:797-800)
    r3 = RS.extract_noisy_results(base, {"expectation": {"observables": [np.eye(3)], "target-qubits": [0]}}, n,
                                  lambda: FakeStateVector(n), _uniforms(1))
    assert "expectation-error" in r3 and "expectation-results" not in r3


# ------------------------------------------------------------------ the job layer on top (backend.py), device faked
@pytest.fixture
def fake_device(monkeypatch):
    from qclojure_b200 import _lib as L
    monkeypatch.setattr(L, "StateVector", FakeStateVector)
    monkeypatch.setattr(L, "device_count", lambda: 1)


def test_ideal_backend_job_with_every_result_spec(fake_device):
    """ideal_simulator.clj:100-176 + result.clj:535-639 through the protocol methods."""
    from qclojure_b200 import backend as B
    sim = B.create_simulator({"seed": 5})
    circ = C.ghz_state_circuit(4)
    C.measure(circ, [3])                                  # mid-circuit :measure consumes one draw (state.clj:981)
    C.h(circ, 0)
    u = np.random.default_rng(2).random(1 + 64)
    specs = {"measurements": {"shots": 64}, "amplitudes": {"basis-states": [0, 15]},
             "hamiltonian": [{"coefficient": 1.0, "pauli-string": "ZZII"}], "state-vector": True}
    res = B.execute_circuit(sim, circ, {"result-specs": specs, "uniforms": u})
    assert res["job-status"] == "completed" and res["job-id"].startswith("sim_job")
    want = O.execute_circuit(circ, draws=iter(u[:1].tolist()))
    r = res["results"]
    assert np.max(np.abs(r["final-state"]["state-vector"] - want)) <= TOL
    assert r["result-types"] == sorted(specs)
    assert r["hamiltonian-result"]["energy-expectation"] == pytest.approx(O.pauli_string_expectation("ZZII", want), abs=TOL)
    assert r["circuit-metadata"]["circuit-operation-count"] == len(circ["operations"])
    assert r["measurement-results"]["shot-count"] == 64
    # a failing circuit never throws out of the worker (ideal_simulator.clj:93-96)
    bad = C.add_gate(C.create_circuit(2), "nonsense", target=0)
    out = B.execute_circuit(sim, bad, {})
    assert out["job-status"] == "failed" and out["error-message"] == "Job not completed" and out["failure-message"]
    sim.close()


def test_edn_circuit_to_backend(fake_device, tmp_path):
    """Wire format -> backend: a circuit exported as EDN, imported, executed; the final state exported as JSON."""
    from qclojure_b200 import backend as B
    from qclojure_b200 import io as QIO
    circ = C.quantum_fourier_transform_circuit(3)
    QIO.export_quantum_circuit("edn", circ, str(tmp_path / "qft.edn"))
    back = QIO.import_quantum_circuit("edn", str(tmp_path / "qft.edn"))
    sim = B.create_simulator()
    res = B.execute_circuit(sim, back, {})
    st = QIO.state_from_backend_result(res)
    # doc/tutorial.md:6077-6105: QFT-3 on |000> -> eight amplitudes 0.3535533905932737
    assert np.allclose(st["state-vector"], 0.3535533905932737, atol=TOL)
    QIO.export_quantum_state("json", st, str(tmp_path / "s.json"))
    assert np.array_equal(QIO.import_quantum_state("json", str(tmp_path / "s.json"))["state-vector"], st["state-vector"])
    sim.close()


def test_device_catalogue_in_the_reference_format(fake_device, tmp_path):
    """MultiDeviceBackend (application/backend.clj:117-131): the catalogue is an EDN vector of device maps like
    resources/simulator-devices.edn; the noise model of a selected entry feeds the C-ABI noise table unchanged."""
    import json
    import os
    from qclojure_b200 import backend as B
    from qclojure_b200 import io as QIO
    from qclojure_b200 import noise as NZ
    prof = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "device_profiles.json")))["devices"]
    prof = [d for d in prof if ":gate-noise" in d["noise_model"]]

    def unkw(x):      # the golden fixture spells keywords with the colon; rebuild the EDN text a QClojure resource holds
        if isinstance(x, dict):
            return {(k[1:] if isinstance(k, str) and k.startswith(":") else k): unkw(v) for k, v in x.items()}
        if isinstance(x, list):
            return [unkw(v) for v in x]
        return QIO.Keyword(x[1:]) if isinstance(x, str) and x.startswith(":") else x
    cat = [{"id": unkw(d["id"]), "name": d["id"][1:], "num-qubits": 7, "noise-model": unkw(d["noise_model"])} for d in prof[:4]]
    f = tmp_path / "devices.edn"
    f.write_text(QIO.write_edn(cat))
    sim = B.create_hardware_simulator(config={"devices-file": str(f)})
    ids = [str(d["id"]) for d in sim.devices()]
    assert ids == [d["id"][1:] for d in prof[:4]]
    dev = sim.select_device(ids[1])
    assert sim.device() is dev and sim.backend_info()["device"] is dev
    with pytest.raises(KeyError):
        sim.select_device("no-such-device")
    # the parsed noise model builds the same C-ABI table as the fixture's own spelling
    a, _ka = NZ.build_noise_table(dev["noise-model"], 7)
    b, _kb = NZ.build_noise_table(prof[1]["noise_model"], 7)
    assert a.n_entries == b.n_entries and a.has_readout == b.has_readout
    assert a.prob_0_to_1 == b.prob_0_to_1 and a.prob_1_to_0 == b.prob_1_to_0
    for k in range(a.n_entries):
        assert bytes(a.entries[k]) == bytes(b.entries[k])
    own = {"id": "mine", "noise-model": {}}
    sim2 = B.create_hardware_simulator(own, {"devices": cat})
    assert len(sim2.devices()) == 5 and sim2.device() is own
    sim.close(); sim2.close()


@pytest.mark.skipif(not os.path.exists("/root/reference/resources/simulator-devices.edn"), reason="reference tree not present")
def test_reference_device_resource_parses():
    from qclojure_b200 import backend as B
    cat = B.load_device_catalog("/root/reference/resources/simulator-devices.edn")
    assert len(cat) >= 10 and all("id" in d and "noise-model" in d for d in cat)
    lagos = [d for d in cat if d["id"] == "ibm-lagos"][0]
    assert lagos["noise-model"]["readout-error"]["prob-0-to-1"] == 0.013       # simulator-devices.edn:532-560


def test_observables_test_clj_known_answers():
    """test/.../domain/observables_test.clj:62-142: expectation values, variances and eigenvalue measurement probabilities of
    the Pauli observables on |0>, |1>, |+>, through the extractors (device faked by the oracle)."""
    from qclojure_b200 import states as S
    R = 1 / math.sqrt(2)
    ident = np.eye(2)

    def sv_of(st):
        sv = FakeStateVector(1)
        sv.set_state(st["state-vector"])
        return sv
    z0, z1, plus = sv_of(S.zero_state()), sv_of(S.one_state()), sv_of(S.plus_state())
    exp = RS.observable_expectation
    assert exp(z0, O.PAULI_Z, None) == pytest.approx(1.0) and exp(z1, O.PAULI_Z, None) == pytest.approx(-1.0)
    assert exp(plus, O.PAULI_Z, None) == pytest.approx(0.0, abs=TOL) and exp(plus, O.PAULI_X, None) == pytest.approx(1.0)
    assert exp(z0, O.PAULI_X, None) == pytest.approx(0.0, abs=TOL) and exp(z1, ident, None) == pytest.approx(1.0)
    var = lambda sv, o: RS.extract_variance_results(sv, [o])[0]["variance-value"]      # noqa: E731
    assert var(z0, O.PAULI_Z) == pytest.approx(0.0, abs=TOL) and var(z1, O.PAULI_Z) == pytest.approx(0.0, abs=TOL)
    assert var(plus, O.PAULI_Z) == pytest.approx(1.0) and var(plus, O.PAULI_X) == pytest.approx(0.0, abs=TOL)
    mp = RS.observable_measurement_probabilities
    assert mp(z0, O.PAULI_Z) == pytest.approx({-1.0: 0.0, 1.0: 1.0}) and mp(z1, O.PAULI_Z) == pytest.approx({-1.0: 1.0, 1.0: 0.0})
    px = mp(plus, O.PAULI_X)
    assert sum(px.values()) == pytest.approx(1.0) and sum(v for v in px.values() if v > 0.9) == pytest.approx(1.0)
    assert px[max(px)] == pytest.approx(1.0)                                            # the +1 eigenvalue of X on |+>
    assert R * R == pytest.approx(mp(plus, O.PAULI_Z)[1.0])


def test_hardware_simulator_protocol_surface(fake_device):
    """test/.../adapter/backend/hardware_simulator_test.clj:21-33, 76-102: protocol surface without running a noisy job."""
    from qclojure_b200 import backend as B
    sim = B.create_hardware_simulator({"id": "dev", "noise-model": {}})
    info = sim.backend_info()
    assert info["backend-type"] == "hardware-simulator" and "multi-device" in info["capabilities"]
    assert info["device"]["id"] == "dev" and [d["id"] for d in info["devices"]] == ["dev"] and sim.available()
    qs = sim.queue_status()
    assert {"total-jobs", "active-jobs", "completed-jobs"} <= set(qs)
    assert sim.job_status("fake-job-id") == "not-found" and sim.job_result("fake-job-id")["job-status"] == "not-found"
    sim.close()
