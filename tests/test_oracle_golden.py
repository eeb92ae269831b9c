"""Pin the CPU oracle against the reference's own recorded outputs and known-answer tests.

Sources: JVM-produced result maps in the reference's doc/tutorial.md (fixtures in
tests/golden/tutorial_cases.json, made by tests/golden/make_tutorial_golden.py) and the
deterministic known-answer tests of the reference's clojure.test suite (cited per test).
Tolerance 1e-10 is the reference's own `approx=` bar (src/.../util/test.clj:13).
"""
import json
import math
import os

import numpy as np
import pytest

from oracle import qc_oracle as O
from qclojure_b200 import circuits as C

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-10


def _cases():
    with open(os.path.join(GOLDEN, "tutorial_cases.json")) as f:
        return json.load(f)


def _sv(lst):
    return np.array([complex(a, b) for a, b in lst])


IDEAL = [c for c in _cases()["cases"] if "trajectories" not in c]
NOISY = [c for c in _cases()["cases"] if "trajectories" in c]
ENERGY = _cases()["energy_cases"]


@pytest.mark.parametrize("case", IDEAL, ids=[c["source"] for c in IDEAL])
def test_tutorial_ideal_final_state(case):
    circ = {"num-qubits": case["num_qubits"], "operations": case["operations"]}
    # mid-circuit :measure ops in the recorded runs have deterministic outcomes (p = 1); any draw
    # strictly inside (0, 1) selects them
    got = O.execute_circuit(circ, draws=iter([0.5] * len(case["operations"])))
    want = _sv(case["final_state"])
    assert np.max(np.abs(got - want)) <= TOL
    mr = case.get("measurement_results")
    if mr:
        probs = np.array(mr[":measurement-probabilities"])
        assert np.max(np.abs(O.measurement_probabilities(got) - probs)) <= TOL
        # recorded outcomes must lie in the support of the distribution
        for o in mr[":measurement-outcomes"]:
            assert probs[o] > 0
    pr = case.get("probability_results")
    if pr and ":all-probabilities" in pr:
        assert np.max(np.abs(O.measurement_probabilities(got) - np.array(pr[":all-probabilities"]))) <= TOL


def test_tutorial_grover_probabilities_bit_exact():
    """doc/tutorial.md:5905-5960 — the pairwise restatement reproduces the JVM's probabilities
    bit for bit (SURVEY §8c 'verified during this survey')."""
    case = [c for c in IDEAL if c["name"] == "Grover Search"][0]
    got = O.execute_circuit({"num-qubits": 3, "operations": case["operations"]})
    want = np.array(case["measurement_results"][":measurement-probabilities"])
    assert np.array_equal(O.measurement_probabilities(got), want)
    assert want[5] == 0.9453124999999959


@pytest.mark.parametrize("case", NOISY, ids=[c["source"] for c in NOISY])
def test_tutorial_noisy_trajectories_are_reachable(case):
    """Hardware-simulator runs (doc/tutorial.md:1261, 1925, 3210): randomness is unseeded in the
    reference, so check what is deterministic: every recorded trajectory state is a normalised state
    whose support is GHZ-like up to the Pauli errors of the profile, rho = mean projector and its trace."""
    trajs = [_sv(t) for t in case["trajectories"]]
    assert len(trajs) == case["trajectory_count"]
    for t in trajs:
        assert abs(np.sum(np.abs(t) ** 2) - 1.0) <= 1e-10
    rho = O.trajectory_to_density_matrix(trajs)
    want = np.array([[complex(a, b) for a, b in row] for row in case["density_matrix"]])
    assert np.max(np.abs(rho - want)) <= 1e-10
    assert abs(np.trace(rho).real - case["density_matrix_trace"]) <= 1e-10
    assert sum(case["measurement_results"].values()) == case["shots_executed"]
    # every recorded trajectory must be one of the states the oracle's noisy path can reach under
    # some Kraus selection (enumerate all selections by steering the per-gate draw)
    n = case["num_qubits"]
    circ = {"num-qubits": n, "operations": case["operations"]}
    profile = {"doc/tutorial.md:1261": ":ibm-lagos", "doc/tutorial.md:1925": ":ibm-lagos",
               "doc/tutorial.md:3210": ":ionq-forte"}[case["source"]]
    with open(os.path.join(GOLDEN, "device_profiles.json")) as f:
        nm = [d for d in json.load(f)["devices"] if d["id"] == profile][0]["noise_model"]
    chans = [O.gate_noise_channel(op, nm) for op in case["operations"]]
    multi = [ch for ch in chans if ch is not None and len(ch[0]) > 1]
    import itertools
    reachable = []
    for choice in itertools.product(*[range(len(ch[0])) for ch in multi]):
        draws = []
        for ch, k in zip(multi, choice):
            pr = O.kraus_selection_probabilities(ch[0])
            draws.append(sum(pr[:k]) + 0.5 * pr[k])
        st = O.zero_state(n)
        it = iter(draws)
        for op in case["operations"]:
            st = O.apply_operation_to_state(st, op)
            st = O.apply_gate_noise(st, op, nm, it)
        reachable.append(st)
    R = np.array(reachable)
    for t in trajs:
        assert np.min(np.max(np.abs(R - t[None, :]), axis=1)) <= 1e-10


@pytest.mark.parametrize("case", ENERGY, ids=[c["source"] for c in ENERGY])
def test_tutorial_variational_energies(case):
    """VQE / QAOA blocks (doc/tutorial.md:6693, 6983, 7364): final circuit + Hamiltonian ->
    recorded optimal energy; optimiser history (parameters -> energy) replayed through the restated
    ansatz builders."""
    n = case["num_qubits"]
    H = case["hamiltonian"]
    state = O.execute_circuit({"num-qubits": n, "operations": case["operations"]})
    e = O.hamiltonian_expectation(H, state)
    assert abs(e - case["optimal_energy"]) <= TOL
    at = case["ansatz_type"]
    for hrec in case["history"]:
        p = hrec["parameters"]
        if at == ":hardware-efficient":
            circ = C.hardware_efficient_ansatz(n, p, num_layers=1)
        elif at == ":uccsd":
            circ = C.uccsd_inspired_ansatz(n, p)
        else:
            Hn = [{"coefficient": t[":coefficient"], "pauli-string": t[":pauli-string"]} for t in H]
            circ = C.qaoa_ansatz_circuit(Hn, C.standard_mixer_hamiltonian(n), p, n)
        st = O.execute_circuit(circ)
        assert abs(O.hamiltonian_expectation(H, st) - hrec["energy"]) <= TOL


# ---------------------------------------------------------------- reference unit tests (known answers)

def test_bits_and_index_conventions():
    """test/.../domain/state_test.clj:197-227, measurement_test.clj:55-70."""
    assert O.bits_to_index([1, 0, 1]) == 5
    assert O.index_to_bits(5, 3) == [1, 0, 1]
    assert O.bits_to_index([0, 0]) == 0 and O.bits_to_index([1, 1]) == 3
    assert O.basis_string(2, 2) == "10"
    s = O.computational_basis_state(2, [1, 0])
    assert O.measure_state(s, 0.3)[0] == 2


def test_bell_state_amplitudes():
    """test/.../domain/circuit_test.clj:410-427."""
    st = O.execute_circuit(C.bell_state_circuit())
    a = 1 / math.sqrt(2)
    assert np.max(np.abs(st - np.array([a, 0, 0, a]))) <= TOL


def test_gate_truth_tables():
    """test/.../domain/gate_test.clj:15-452 (truth tables and phases on 1-3 qubit basis states)."""
    z2 = O.zero_state(2)
    # X on qubit 0 of |00> -> |10> (index 2): qubit 0 is the MSB
    assert O.apply_gate_to_state(z2, {"operation-type": "x", "operation-params": {"target": 0}})[2] == 1
    # CNOT(0,1) on |10> -> |11>
    s10 = O.computational_basis_state(2, [1, 0])
    assert O.cnot(s10, 0, 1)[3] == 1
    # Y|0> = i|1>
    y0 = O.apply_single_qubit_gate(O.zero_state(1), O.PAULI_Y, 0)
    assert y0[1] == 1j
    # S|1> = i|1>, T|1> = e^{i pi/4}|1>
    one = O.computational_basis_state(1, [1])
    assert O.apply_single_qubit_gate(one, O.S_GATE, 0)[1] == 1j
    assert abs(O.apply_single_qubit_gate(one, O.T_GATE, 0)[1] - complex(math.cos(math.pi / 4), math.sin(math.pi / 4))) <= TOL
    # SWAP(0,1) |01> -> |10> ; iSWAP |01> -> i|10>  (gate_test.clj:209-230)
    s01 = O.computational_basis_state(2, [0, 1])
    assert O.swap_gate(s01, 0, 1)[2] == 1
    assert O.iswap_gate(s01, 0, 1)[2] == 1j
    # Toffoli |110> -> |111> ; Fredkin |101> -> |110>
    assert O.toffoli_gate(O.computational_basis_state(3, [1, 1, 0]), 0, 1, 2)[7] == 1
    assert O.fredkin_gate(O.computational_basis_state(3, [1, 0, 1]), 0, 1, 2)[6] == 1
    # CZ |11> -> -|11>
    assert O.controlled_z(O.computational_basis_state(2, [1, 1]), 0, 1)[3] == -1
    # unknown gate types :i and :cy throw at execution (circuit.clj:1072)
    with pytest.raises(O.UnknownGate):
        O.apply_gate_to_state(z2, {"operation-type": "i", "operation-params": {"target": 0}})
    with pytest.raises(O.UnknownGate):
        O.apply_gate_to_state(z2, {"operation-type": "cy", "operation-params": {"control": 0, "target": 1}})


def test_controlled_gate_applies_transpose():
    """gate.clj:473-483 — CRY(theta) acts as controlled-RY(-theta); CRX unaffected (symmetric)."""
    th = 0.7
    s10 = O.computational_basis_state(2, [1, 0])
    got = O.controlled_ry(s10, 0, 1, th)
    c, s = math.cos(th / 2), math.sin(th / 2)
    assert np.max(np.abs(got - np.array([0, 0, c, -s]))) <= TOL  # RY(-theta)|0> = c|0> - s|1>
    gx = O.controlled_rx(s10, 0, 1, th)
    assert np.max(np.abs(gx - np.array([0, 0, c, -1j * s]))) <= TOL


def test_swap_lsb_quirk():
    """gate.clj:768-778 — qubit1/qubit2 are LSB bit positions: on 3 qubits swap(0,1) exchanges the two
    LEAST significant index bits, i.e. reference qubits 2 and 1."""
    s = O.computational_basis_state(3, [0, 0, 1])  # index 1
    assert O.swap_gate(s, 0, 1)[2] == 1            # -> index 2 = |010>
    # symmetric pairs {k, n-1-k} coincide with the MSB convention (QFT's swaps)
    s = O.computational_basis_state(3, [1, 0, 0])  # index 4
    assert O.swap_gate(s, 0, 2)[1] == 1


def test_global_x_keeps_phase():
    """gate.clj:1224-1225 — global-x = RX(pi) on every qubit: |00> -> (-i)^2 |11> = -|11>."""
    st = O.apply_gate_to_state(O.zero_state(2), {"operation-type": "global-x", "operation-params": {}})
    assert abs(st[3] - (-1)) <= TOL


def test_pauli_and_hamiltonian_expectations():
    """test/.../domain/hamiltonian_test.clj:66-115, result_test.clj:47-104."""
    z2 = O.zero_state(2)
    assert O.pauli_string_expectation("ZZ", z2) == 1.0
    assert O.pauli_string_expectation("ZI", z2) == 1.0
    assert abs(O.pauli_string_expectation("XI", z2)) <= TOL
    bell = O.execute_circuit(C.bell_state_circuit())
    assert abs(O.pauli_string_expectation("ZZ", bell) - 1.0) <= TOL
    assert abs(O.pauli_string_expectation("XX", bell) - 1.0) <= TOL
    assert abs(O.pauli_string_expectation("YY", bell) + 1.0) <= TOL
    H = [{"coefficient": 0.5, "pauli-string": "ZZ"}, {"coefficient": -0.25, "pauli-string": "XX"}]
    assert abs(O.hamiltonian_expectation(H, bell) - 0.25) <= TOL
    plus = O.apply_single_qubit_gate(O.zero_state(1), O.HADAMARD, 0)
    assert abs(O.expectation_1q(plus, O.PAULI_X, 0) - 1.0) <= TOL
    assert abs(O.variance_1q(plus, O.PAULI_Z, 0) - 1.0) <= TOL


def test_kraus_operators_and_decoherence():
    """test/.../domain/channel_test.clj:46-145, 170-177, 213-229."""
    ks = O.depolarizing_kraus_operators(0.1)
    assert abs(ks[0][0, 0] - math.sqrt(0.9)) <= TOL and abs(ks[1][0, 1] - math.sqrt(0.1 / 3)) <= TOL
    tot = sum(k.conj().T @ k for k in ks)
    assert np.max(np.abs(tot - np.eye(2))) <= TOL
    ad = O.amplitude_damping_kraus_operators(0.3)
    assert abs(ad[0][1, 1] - math.sqrt(0.7)) <= TOL and abs(ad[1][0, 1] - math.sqrt(0.3)) <= TOL
    pdk = O.phase_damping_kraus_operators(0.2)
    assert abs(pdk[1][1, 1] - math.sqrt(0.2)) <= TOL
    d = O.calculate_decoherence_params(100.0, 50.0, 1000.0)  # 1 us gate
    assert abs(d["gamma-1"] - (1 - math.exp(-0.01))) <= TOL and abs(d["gamma-2"] - (1 - math.exp(-0.02))) <= TOL
    # X Kraus on qubit 1 of |00> -> |01>
    st = O.apply_single_qubit_kraus_operator(O.zero_state(2), O.PAULI_X, 1)
    assert abs(st[1] - 1) <= TOL
    # selection probabilities are max |coeff|^2: damping channels always pick K0
    assert O.kraus_selection_probabilities(ad)[0] == 1.0
    assert O.select_kraus_index(ad, 0.999999) == 0
    assert O.kraus_selection_probabilities(ks) == pytest.approx([0.9, 0.1 / 3, 0.1 / 3, 0.1 / 3], abs=1e-15)


def test_readout_bitstring():
    """test/.../domain/noise_test.clj:92-102 — basis [1 0] reads "10" without readout error."""
    s = O.computational_basis_state(2, [1, 0])
    assert O.apply_readout_noise(s, 2, {}, iter([0.5])) == "10"
    ro = {"readout-error": {"prob-0-to-1": 1.0, "prob-1-to-0": 0.0}}
    assert O.apply_readout_noise(s, 2, ro, iter([0.5, 0.5, 0.5])) == "11"


def test_ideal_simulator_supports():
    """test/.../adapter/backend/ideal_simulator_test.clj:101-185 — Bell {0,3}, GHZ {0,7}, X->1, X;CNOT->3."""
    u = np.random.default_rng(5).random(1000)
    bell = O.sample_outcomes(O.execute_circuit(C.bell_state_circuit()), u)
    assert set(bell.tolist()) == {0, 3} and 0.45 <= np.mean(bell == 0) <= 0.55
    ghz = O.sample_outcomes(O.execute_circuit(C.ghz_state_circuit(3)), u)
    assert set(ghz.tolist()) == {0, 7}
    c = C.x(C.create_circuit(1), 0)
    assert set(O.sample_outcomes(O.execute_circuit(c), u).tolist()) == {1}
    c = C.cnot(C.x(C.create_circuit(2), 0), 0, 1)
    assert set(O.sample_outcomes(O.execute_circuit(c), u).tolist()) == {3}


def test_qft3_matches_builder_and_tutorial():
    """doc/tutorial.md:6077-6105 — the QFT builder restatement emits the recorded 7-op list."""
    case = [c for c in IDEAL if c["name"] == "QFT"][0]
    built = C.quantum_fourier_transform_circuit(3)["operations"]
    rec = [O.normalize_op(op) for op in case["operations"]]
    assert len(built) == len(rec) == 7
    for b, (typ, p) in zip(built, rec):
        assert b["operation-type"] == typ
        for k, v in p.items():
            assert b["operation-params"][k] == pytest.approx(v, abs=1e-15)


def test_sampling_rule_edge_cases():
    """state.clj:905-908 — draw of exactly 0 returns index 0 even if p0 = 0; clamp to N-1."""
    s = O.computational_basis_state(2, [1, 1])
    assert O.measure_state(s, 0.0)[0] == 0
    assert O.measure_state(s, 0.999999)[0] == 3
    s = O.execute_circuit(C.bell_state_circuit())
    assert O.sample_outcomes(s, [0.0, 0.4999, 0.51, 0.999999]).tolist() == [0, 0, 3, 3]


def test_partial_measurement_collapse():
    """state.clj:946-1014 on a Bell state: measuring qubit 0 collapses both."""
    s = O.execute_circuit(C.bell_state_circuit())
    bits, col, probs = O.measure_specific_qubits(s, [0], 0.25)
    assert bits == [0] and abs(col[0] - 1) <= TOL and probs == pytest.approx([0.5, 0.5], abs=1e-12)
    bits, col, _ = O.measure_specific_qubits(s, [0], 0.75)
    assert bits == [1] and abs(col[3] - 1) <= TOL


def test_dense_reference_form_equals_pairwise():
    """gate.clj:346-395 dense Kronecker mat-vec == pairwise update (bit-exact on random states)."""
    rng = np.random.default_rng(3)
    st = rng.standard_normal(32) + 1j * rng.standard_normal(32)
    st /= np.linalg.norm(st)
    for q in range(5):
        a = O.apply_single_qubit_gate(st, O.rx_gate(0.37), q)
        b = O.apply_single_qubit_gate_dense(st, O.rx_gate(0.37), q)
        assert np.max(np.abs(a - b)) <= 1e-15


def test_hhl_tutorial_run_pins_the_transposed_controlled_gate():
    """doc/tutorial.md (HHL section): the JVM-recorded 64 probabilities of `hhl-circuit [[3 1] [1 2]] [7 5] 4 1` are
    reproduced to 1e-15 by the oracle's literal `apply-controlled-gate` (transposed 2x2, gate.clj:473-483) and are off by
    0.1 under the textbook CRY: the recorded run pins the convention SURVEY §8a row 5 lists as unpinned by the
    reference's tests."""
    import math
    from qclojure_b200 import circuits as CB
    with open(os.path.join(GOLDEN, "hhl_tutorial.json")) as f:
        g = json.load(f)
    circ = CB.hhl_circuit(g["matrix"], g["vector"], g["precision_qubits"], g["ancilla_qubits"])
    assert circ["num-qubits"] == 6 and [o["operation-type"] for o in circ["operations"]].count("cry") == 4
    want = np.array(g["all_probabilities"])
    got = O.measurement_probabilities(O.execute_circuit(circ))
    assert np.max(np.abs(got - want)) <= 1e-15
    st = O.zero_state(6)
    for op in circ["operations"]:
        if op["operation-type"] == "cry":
            p = op["operation-params"]
            st = O.apply_controlled_gate(st, p["control"], p["target"], O.ry_gate(p["angle"]).T)   # U^T of U^T = textbook U
        else:
            st = O.apply_gate_to_state(st, op)
    assert np.max(np.abs(O.measurement_probabilities(st) - want)) > 0.05
