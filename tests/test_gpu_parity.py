"""GPU parity tests: the CUDA path (through the C ABI of libqcb200.so) against the oracle on the same
seeded inputs.  Tolerances: amplitudes / probabilities / expectations 1e-10 absolute in fp64 (the
reference's own `approx=` bar, src/.../util/test.clj:13, and BASELINE.json north_star); shot outcomes
identical on identical uniform draws except at cumulative-probability boundaries within 1e-12.
"""
import json
import math
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import qc_oracle as O
from qclojure_b200 import _lib as L
from qclojure_b200 import circuits as C
from tests.test_oracle_c import _all_gates_circuit

pytestmark = pytest.mark.gpu
TOL = 1e-10
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _rand_state(n, seed):
    rng = np.random.default_rng(seed)
    s = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return s / np.linalg.norm(s)


def _run(n, ops, init=None, **kw):
    with L.StateVector(n, **kw) as sv:
        if init is not None:
            sv.set_state(init)
        sv.apply_ops(ops)
        return sv.get_state()


# ------------------------------------------------------------------ gates
@pytest.mark.parametrize("n,tile,low,fusion", [
    (1, 0, 0, 1), (2, 0, 0, 1), (3, 0, 0, 1), (4, 0, 0, 0), (5, 0, 0, 1), (8, 5, 2, 1), (10, 7, 4, 1),
    (12, 0, 0, 1), (13, 10, 4, 1), (14, 12, 4, 1), (14, 0, 0, 0), (16, 0, 0, 1), (16, 11, 5, 1), (17, 13, 6, 1)])
def test_all_gates_match_oracle(n, tile, low, fusion):
    rng = np.random.default_rng(7 * n + tile)
    if n >= 3:
        circ = _all_gates_circuit(n, rng)
    else:
        circ = C.create_circuit(n)
        for _ in range(20):
            q = int(rng.integers(0, n))
            C.add_gate(circ, ["h", "x", "y", "z", "s", "t"][rng.integers(0, 6)], target=q)
            C.rx(circ, q, rng.random()); C.rz(circ, q, rng.random())
            if n == 2:
                C.cnot(circ, q, 1 - q); C.crz(circ, 1 - q, q, 0.3); C.swap(circ, 0, 1)
    init = _rand_state(n, n)
    want = O.execute_circuit(circ, init)
    got = _run(n, circ["operations"], init, tile_bits=tile, low_bits=low, fusion=fusion)
    assert np.max(np.abs(got - want)) <= TOL


@pytest.mark.parametrize("target", list(range(0, 20, 1)))
def test_single_qubit_gate_every_target_20q(target):
    """dense 1q gate on every target 0..n-1 (high targets = strided pairs, low targets =
    in-tile pairs)."""
    n = 20
    init = _rand_state(n, 99)
    ops = [{"operation-type": "rx", "operation-params": {"target": target, "angle": 0.73}},
           {"operation-type": "h", "operation-params": {"target": target}}]
    want = CO.apply_circuit({"num-qubits": n, "operations": ops}, init)
    for fusion in (0, 1):
        got = _run(n, ops, init, fusion=fusion)
        assert np.max(np.abs(got - want)) <= TOL


def test_every_two_qubit_pair_12q():
    n = 12
    init = _rand_state(n, 5)
    for a in range(n):
        ops = []
        for b in range(n):
            if a != b:
                ops += [{"operation-type": "cnot", "operation-params": {"control": a, "target": b}},
                        {"operation-type": "crx", "operation-params": {"control": b, "target": a, "angle": 0.4}},
                        {"operation-type": "iswap", "operation-params": {"qubit1": a, "qubit2": b}},
                        {"operation-type": "cz", "operation-params": {"control": a, "target": b}}]
        want = O.execute_circuit({"num-qubits": n, "operations": ops}, init)
        got = _run(n, ops, init, tile_bits=8, low_bits=3)
        assert np.max(np.abs(got - want)) <= TOL


def test_strict_parity_flag_and_unknown_gates():
    n = 4
    circ = C.create_circuit(n)
    C.h(circ, 0); C.h(circ, 2); C.cry(circ, 0, 1, 0.7); C.swap(circ, 0, 1); C.iswap(circ, 0, 2)
    assert np.max(np.abs(_run(n, circ["operations"]) - O.execute_circuit(circ))) <= TOL
    st = O.zero_state(n)
    st = O.apply_single_qubit_gate(st, O.HADAMARD, 0)
    st = O.apply_single_qubit_gate(st, O.HADAMARD, 2)
    st = O.apply_controlled_gate(st, 0, 1, O.ry_gate(0.7).T)
    st = O.swap_gate(st, n - 1 - 0, n - 1 - 1)
    st = O.iswap_gate(st, n - 1 - 0, n - 1 - 2)
    assert np.max(np.abs(_run(n, circ["operations"], strict_parity=0) - st)) <= TOL
    bad = [{"operation-type": "h", "operation-params": {"target": 0}},
           {"operation-type": "cy", "operation-params": {"control": 0, "target": 1}}]
    with L.StateVector(2) as sv:
        with pytest.raises(L.QcbError) as ei:
            sv.apply_ops(bad)
        assert ei.value.code == -2 and "Unknown gate type" in ei.value.message
        # a failed op list leaves the state untouched (the reference fails the whole job)
        assert np.max(np.abs(sv.get_state() - O.zero_state(2))) == 0
    want = O.apply_controlled_gate(O.apply_single_qubit_gate(O.zero_state(2), O.HADAMARD, 0), 0, 1, O.PAULI_Y.T)
    assert np.max(np.abs(_run(2, bad, strict_parity=0) - want)) <= TOL


def test_generic_ops_and_grover_operators():
    n = 6
    rng = np.random.default_rng(3)
    init = _rand_state(n, 11)

    def runitary(d):
        q, _ = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))
        return q

    U1, U2, CU = runitary(2), runitary(4), runitary(2)
    ops = [{"operation-type": "u1q", "operation-params": {"target": 4, "matrix": U1}},
           {"operation-type": "cu1q", "operation-params": {"control": 1, "target": 3, "matrix": CU}},
           {"operation-type": "u2q", "operation-params": {"qubit1": 5, "qubit2": 2, "matrix": U2}},
           {"operation-type": "mcphase", "operation-params": {"qubit-indices": [0, 2, 5], "angle": 0.9}},
           {"operation-type": "phase-oracle", "operation-params": {"index": 37}},
           {"operation-type": "grover-diffusion", "operation-params": {}},
           {"operation-type": "h", "operation-params": {"target": 1}},
           {"operation-type": "grover-diffusion", "operation-params": {}}]
    st = O.apply_single_qubit_gate(init, U1, 4)
    st = O.apply_controlled_gate(st, 1, 3, CU.T)
    idx = np.arange(1 << n)
    out = np.zeros_like(st)
    ba, bb = n - 1 - 5, n - 1 - 2
    for i in idx:
        col = (((i >> ba) & 1) << 1) | ((i >> bb) & 1)
        base = i & ~((1 << ba) | (1 << bb))
        for row in range(4):
            out[base | (((row >> 1) & 1) << ba) | ((row & 1) << bb)] += U2[row, col] * st[i]
    st = out
    allone = np.ones(1 << n, dtype=bool)
    for q in (0, 2, 5):
        allone &= ((idx >> (n - 1 - q)) & 1) == 1
    st = np.where(allone, st * np.exp(1j * 0.9), st)
    st[37] *= -1
    st = 2 * np.mean(st) - st
    st = O.apply_single_qubit_gate(st, O.HADAMARD, 1)
    st = 2 * np.mean(st) - st
    for kw in ({}, {"tile_bits": 4, "low_bits": 2}, {"fusion": 0}):
        assert np.max(np.abs(_run(n, ops, init, **kw) - st)) <= TOL


def test_grover_full_iteration_count_12q_and_16q():
    """SURVEY §8d config 2 parity leg: full iteration count with the fused oracle+diffusion operators."""
    for n, target in ((12, 0xAAA), (16, 0x2AAA)):
        iters = C.grover_iterations(n)
        ops = [{"operation-type": "global-h", "operation-params": {}}]
        for _ in range(iters):
            ops += [{"operation-type": "phase-oracle", "operation-params": {"index": target}},
                    {"operation-type": "grover-diffusion", "operation-params": {}}]
        st = np.full(1 << n, 1.0 / math.sqrt(1 << n), dtype=np.complex128)
        for _ in range(iters):
            st[target] *= -1
            st = 2 * np.mean(st) - st
        got = _run(n, ops)
        assert np.max(np.abs(got - st)) <= TOL
        assert abs(got[target]) ** 2 > 0.99


@pytest.mark.parametrize("n,marked", [(3, [5]), (9, [1, 77, 300]), (14, [0x2AAA & 0x3FFF, 5]), (20, [0xABCDE]), (11, [])])
def test_fused_grover_pass_on_device(n, marked):
    """plan.h S_GROVER / kernels.cu k_grover_step: diffusion + the phase oracles after it + the sum for the next diffusion in
    one streaming pass; several marked states; a diffusion followed by ordinary gates falls back to the tile path."""
    its = 6
    ops = [{"operation-type": "global-h", "operation-params": {}}]
    st = np.full(1 << n, 1.0 / math.sqrt(1 << n), dtype=np.complex128)
    for it in range(its):
        ops += [{"operation-type": "phase-oracle", "operation-params": {"index": mk}} for mk in marked]
        ops += [{"operation-type": "grover-diffusion", "operation-params": {}}]
        for mk in marked:
            st[mk] *= -1
        st = 2 * np.mean(st) - st
        if it == 2:
            ops += [{"operation-type": "h", "operation-params": {"target": n - 1}}]
            st = O.apply_single_qubit_gate(st, O.HADAMARD, n - 1)
    with L.StateVector(n) as sv:
        sv.apply_ops(ops)
        got = sv.get_state()
        stats = sv.stats()
    assert np.max(np.abs(got - st)) <= TOL
    assert abs(np.linalg.norm(got) - 1.0) <= 1e-12


# ------------------------------------------------------------------ BASELINE configs (parity legs)
@pytest.mark.parametrize("n", [20])
def test_config1_qft_ghz_20q_with_shots(n):
    """configs[0]: QFT + GHZ on 20 qubits, 1024 shots (uniforms from default_rng(20261017))."""
    u = np.random.default_rng(20261017).random(1024)
    for circ, init_index in ((C.quantum_fourier_transform_circuit(n), 0), (C.quantum_fourier_transform_circuit(n), 0x5A5A5),
                             (C.ghz_state_circuit(n), 0)):
        init = np.zeros(1 << n, dtype=np.complex128)
        init[init_index] = 1.0
        want = CO.apply_circuit(circ, init)
        with L.StateVector(n) as sv:
            sv.set_basis(init_index)
            sv.apply_circuit(circ)
            got = sv.get_state()
            assert np.max(np.abs(got - want)) <= TOL
            assert abs(sv.norm2() - 1.0) <= TOL
            shots = sv.sample(u)
        ref = O.sample_outcomes(want, u)
        dist = O.sample_boundary_distance(want, u)
        bad = (shots != ref) & (dist > 1e-12)
        assert not bad.any(), f"{bad.sum()} shots differ away from cumulative boundaries"


@pytest.mark.parametrize("n", [20, 24])
def test_config3_brickwork_matches_c_oracle(n):
    circ = C.random_brickwork_circuit(n, 20)
    want = CO.apply_circuit(circ)
    with L.StateVector(n) as sv:
        sv.apply_circuit(circ)
        got = sv.get_state()
        st = sv.stats()
    assert np.max(np.abs(got - want)) <= TOL
    assert st["n_sweeps"] < len(circ["operations"]) / 4          # fusion really fused
    got_u = _run(n, circ["operations"], fusion=0) if n <= 20 else None
    if got_u is not None:
        assert np.max(np.abs(got_u - want)) <= TOL


def _inverse_ops(ops):
    """Exact inverse of a brickwork op list: reverse order, RX/RZ angles negated, H / CNOT / CZ are involutions."""
    inv = []
    for op in reversed(ops):
        t, p = op["operation-type"], dict(op["operation-params"])
        if t in ("rx", "rz"):
            p["angle"] = -p["angle"]
        else:
            assert t in ("h", "cnot", "cz"), t
        inv.append({"operation-type": t, "operation-params": p})
    return inv


def test_config3_full_size_properties_30q():
    """BASELINE.json's full size (30 qubits, 16 GiB state) through size-independent properties: the norm stays 1, 4096
    spot amplitudes are stable under re-execution, and circuit followed by its exact inverse returns |0...0> to 1e-10
    (890 + 890 gates through the fused executor; the C oracle would need minutes at this size)."""
    n = 30
    circ = C.random_brickwork_circuit(n, 20)
    ops = circ["operations"]
    idx = np.random.default_rng(30).integers(0, 1 << n, 4096)
    with L.StateVector(n) as sv:
        sv.apply_ops(ops)
        assert abs(sv.norm2() - 1.0) <= 1e-10
        spot = sv.get_amplitudes(idx)
        assert np.max(np.abs(spot)) < 1e-3 and np.any(np.abs(spot) > 0)          # a scrambled state, not a basis state
        sv.apply_ops(_inverse_ops(ops))
        assert abs(sv.norm2() - 1.0) <= 1e-10
        back = sv.get_amplitudes(np.concatenate([[0], idx]))
        assert abs(back[0] - 1.0) <= 1e-10
        assert np.max(np.abs(back[1:][idx != 0])) <= 1e-10
        sv.set_zero()
        sv.apply_ops(ops)
        assert np.array_equal(sv.get_amplitudes(idx), spot)                        # deterministic re-execution


@pytest.mark.parametrize("n", [28, 30])
def test_config3_full_size_spot_amplitudes_vs_c_oracle(n):
    """SURVEY 8d config 3 at full size against an INDEPENDENT implementation: the whole depth-20 brickwork circuit through the
    C oracle on the host cores (in-place pairwise updates, no fusion, no tiles, no tensor cores; 16 GiB of host state at 30
    qubits, a few minutes) and 4096 random spot amplitudes + the 64 largest-index amplitudes (byte offsets >= 2^32 from
    28 qubits on) compared with the fused CUDA path to 1e-10.  Unlike the circuit-then-inverse property above, a consistent
    qubit-mapping error or a wrong-but-unitary gate cannot cancel here."""
    import psutil
    if psutil.virtual_memory().available < (16 << n) * 1.3:
        pytest.skip("not enough host memory for the oracle state")
    circ = C.random_brickwork_circuit(n, 20)
    rng = np.random.default_rng(2800 + n)
    idx = np.concatenate([rng.integers(0, 1 << n, 4096), np.arange((1 << n) - 64, 1 << n)])
    want_full = CO.apply_circuit(circ)
    want = want_full[idx].copy()
    norm_want = CO.norm2(want_full)
    del want_full
    with L.StateVector(n) as sv:
        sv.apply_circuit(circ)
        got = sv.get_amplitudes(idx)
        norm_got = sv.norm2()
    assert np.max(np.abs(got - want)) <= TOL
    assert abs(norm_got - norm_want) <= 1e-10
    assert np.max(np.abs(want)) > 0


def test_tutorial_golden_cases_on_gpu():
    """The reference's own recorded runs (doc/tutorial.md) replayed through the CUDA path."""
    with open(os.path.join(GOLDEN, "tutorial_cases.json")) as f:
        data = json.load(f)
    for case in data["cases"]:
        if "trajectories" in case:
            continue
        n = case["num_qubits"]
        want = np.array([complex(a, b) for a, b in case["final_state"]])
        ops = []
        for op in case["operations"]:
            typ, p = O.normalize_op(op)
            if typ == "measure":
                p = dict(p, uniform=0.5)
            ops.append({"operation-type": typ, "operation-params": p})
        with L.StateVector(n) as sv:
            sv.apply_ops(ops)
            assert np.max(np.abs(sv.get_state() - want)) <= TOL, case["source"]
            mr = case.get("measurement_results")
            if mr:
                assert np.max(np.abs(sv.probabilities() - np.array(mr[":measurement-probabilities"]))) <= TOL
    for case in data["energy_cases"]:
        n = case["num_qubits"]
        with L.StateVector(n) as sv:
            sv.apply_ops(case["operations"])
            e = sv.expect_hamiltonian(case["hamiltonian"])
        assert abs(e - case["optimal_energy"]) <= TOL, case["source"]


# ------------------------------------------------------------------ measurement
def test_probabilities_norm_amplitudes():
    n = 14
    init = _rand_state(n, 21)
    with L.StateVector(n) as sv:
        sv.set_state(init)
        assert np.max(np.abs(sv.probabilities() - O.measurement_probabilities(init))) <= TOL
        assert np.max(np.abs(sv.probabilities(100, 50) - O.measurement_probabilities(init)[100:150])) <= TOL
        assert abs(sv.norm2() - 1.0) <= TOL
        idx = [0, 5, 77, (1 << n) - 1]
        assert np.max(np.abs(sv.get_amplitudes(idx) - init[idx])) == 0
        assert np.max(np.abs(sv.get_state(10, 7) - init[10:17])) == 0
        assert abs(sv.fidelity(init) - 1.0) <= TOL
        other = _rand_state(n, 22)
        assert abs(sv.fidelity(other) - O.state_fidelity(init, other)) <= TOL
        sv.set_state(init * 3.0)
        sv.normalize()
        assert np.max(np.abs(sv.get_state() - init)) <= TOL


@pytest.mark.parametrize("n", [1, 3, 11, 13, 18, 25])      # 25: 8192 chunk sums = two tiles of the chunk-sum scan (carry)
def test_sampling_rule(n):
    init = _rand_state(n, 30 + n)
    u = np.concatenate([[0.0, 0.999999999999], np.random.default_rng(n).random(2000)])
    with L.StateVector(n) as sv:
        sv.set_state(init)
        got = sv.sample(u)
    ref = O.sample_outcomes(init, u)
    dist = O.sample_boundary_distance(init, u)
    assert not ((got != ref) & (dist > 1e-12)).any()
    # sparse state: draw of exactly 0 returns index 0 even if p0 = 0; clamp at the top (state.clj:905-908)
    with L.StateVector(n) as sv:
        sv.set_basis((1 << n) - 1)
        assert sv.sample([0.0, 0.3, 0.999999]).tolist() == [0, (1 << n) - 1, (1 << n) - 1]
        sv.set_state(np.zeros(1 << n, dtype=np.complex128) + 0.0)
        with pytest.raises(L.QcbError) as ei:
            sv.sample([0.5])
        assert ei.value.code == -6        # "State is not properly normalized"


def test_partial_measurement_collapse():
    n = 10
    init = _rand_state(n, 44)
    for qubits, u in (([0], 0.3), ([3, 1], 0.8), ([9, 0, 5], 0.55), ([2, 4, 6, 8], 0.07)):
        bits, col, probs = O.measure_specific_qubits(init, qubits, u)
        with L.StateVector(n) as sv:
            sv.set_state(init)
            assert np.max(np.abs(sv.marginal_probabilities(qubits) - np.array(probs))) <= TOL
            gbits, gp = sv.measure_qubits(qubits, u)
            assert gbits == bits
            assert np.max(np.abs(sv.get_state() - col)) <= TOL
    # :measure op inside an op list (circuit.clj:1106-1111)
    circ = C.bell_state_circuit()
    ops = circ["operations"] + [{"operation-type": "measure", "operation-params": {"measurement-qubits": [0], "uniform": 0.75}}]
    got = _run(2, ops)
    assert abs(got[3] - 1) <= TOL


def test_direct_store_variant_matches_oracle(monkeypatch):
    """The optional direct-store variant of the tile kernel (QCB_DIRECT_STORE=1: the last round of a sweep writes to HBM from
    registers) against the C oracle."""
    monkeypatch.setenv("QCB_DIRECT_STORE", "1")
    for n in (14, 22):
        circ = C.random_brickwork_circuit(n, 12)
        want = CO.apply_circuit(circ)
        with L.StateVector(n) as sv:
            sv.apply_circuit(circ)
            assert np.max(np.abs(sv.get_state() - want)) <= TOL


def test_measure_many_qubits_two_pass_and_determinism():
    """:measure of more than 12 qubits (two histogram passes: the bits above 12 of the outcome enumeration first, then the
    low 12 restricted to that choice) against the oracle's single enumeration, and run-to-run bit-identity of the marginal
    histogram (no floating-point atomics: per-warp private bins, fixed addition order)."""
    n = 16
    init = _rand_state(n, 45)
    rng = np.random.default_rng(46)
    for m, u in ((13, 0.31), (14, 0.77), (16, 0.05), (16, 0.999)):
        qubits = [int(q) for q in rng.permutation(n)[:m]]
        bits, col, _probs = O.measure_specific_qubits(init, qubits, u)
        with L.StateVector(n) as sv:
            sv.set_state(init)
            gbits, gp = sv.measure_qubits(qubits, u)
            assert gbits == bits, (m, u)
            assert np.max(np.abs(sv.get_state() - col)) <= TOL
            assert abs(sv.norm2() - 1.0) <= TOL
    with L.StateVector(n) as sv:
        sv.set_state(init)
        for qubits in ([0, 5, 9], list(range(12)), [15, 3, 8, 1, 12, 7, 0]):
            a = sv.marginal_probabilities(qubits)
            b = sv.marginal_probabilities(qubits)
            assert np.array_equal(a, b)                                    # bit-identical, not just close
            assert np.max(np.abs(a - np.array(O.measure_specific_qubits(init, qubits, 0.5)[2]))) <= TOL
        with pytest.raises(L.QcbError):
            sv.marginal_probabilities(list(range(13)))                     # a 2^13-entry distribution is not offered


# ------------------------------------------------------------------ expectation values
def test_pauli_and_hamiltonian_expectations():
    n = 12
    init = _rand_state(n, 71)
    rng = np.random.default_rng(8)
    strings = ["I" * n, "Z" * n, "X" * n, "Y" * n, "ZIIIIIIIIIII", "IIIIIIIIIIIX", "XYZIXYZIXYZI", "IIYYIIIIIIZZ"]
    strings += ["".join(rng.choice(list("IXYZ"), size=n)) for _ in range(40)]
    coeffs = rng.standard_normal(len(strings))
    H = [{"coefficient": float(c), "pauli-string": s} for c, s in zip(coeffs, strings)]
    with L.StateVector(n) as sv:
        sv.set_state(init)
        for s in strings[:10]:
            assert abs(sv.expect_pauli(s) - O.pauli_string_expectation(s, init)) <= TOL, s
        e, terms = sv.expect_hamiltonian(H, return_terms=True)
        for s, t in zip(strings, terms):
            assert abs(t - O.pauli_string_expectation(s, init)) <= TOL, s
        assert abs(e - O.hamiltonian_expectation(H, init)) <= 1e-9
        for tq in (0, 5, 11):
            for obs in (O.PAULI_X, O.PAULI_Y, O.PAULI_Z, O.HADAMARD):
                assert abs(sv.expect_1q(obs, tq) - O.expectation_1q(init, obs, tq)) <= TOL
        with pytest.raises(L.QcbError):
            sv.expect_pauli("ZZ")


def test_qaoa_energy_sweep_12q():
    """configs[4] parity leg: MaxCut QAOA p=2 on a 3-regular graph, <H> per (gamma, beta) point."""
    n = 12
    graph = C.random_regular_graph(n, 3, seed=11)
    Hp, Hm = C.max_cut_hamiltonian(graph, n), C.standard_mixer_hamiltonian(n)
    with L.StateVector(n) as sv:
        for g in np.linspace(0, math.pi, 4):
            for b in np.linspace(0, math.pi, 4):
                circ = C.qaoa_ansatz_circuit(Hp, Hm, [g, b, 0.5 * g, 0.5 * b], n)
                sv.set_zero()
                sv.apply_circuit(circ)
                want = O.hamiltonian_expectation(Hp, CO.apply_circuit(circ))
                assert abs(sv.expect_hamiltonian(Hp) - want) <= TOL


# ------------------------------------------------------------------ noise
def _noise_table(nm, n):
    from qclojure_b200 import noise as NZ
    return NZ.build_noise_table(nm, n)


def test_kraus_application():
    n = 8
    init = _rand_state(n, 5)
    for K in (O.amplitude_damping_kraus_operators(0.3)[0], O.amplitude_damping_kraus_operators(0.3)[1],
              O.depolarizing_kraus_operators(0.1)[2], O.coherent_error_kraus_operator(0.2, "z")):
        for tq in (0, 3, 7):
            with L.StateVector(n) as sv:
                sv.set_state(init)
                sv.apply_kraus_1q(K, tq)
                assert np.max(np.abs(sv.get_state() - O.apply_single_qubit_kraus_operator(init, K, tq))) <= TOL


@pytest.mark.parametrize("profile,n,shots", [(":ibm-lagos", 3, 300), (":ionq-aria-1", 5, 200), (":ibm-lagos", 7, 100),
                                              (":ionq-forte", 12, 24), (":rigetti-aspen-m3", 6, 100)])
def test_noisy_trajectories_match_oracle(profile, n, shots):
    """configs[4] parity leg: identical bitstring counts vs the oracle on the same uniform draws, in the
    reference's draw order."""
    from qclojure_b200 import ops as OPS
    with open(os.path.join(GOLDEN, "device_profiles.json")) as f:
        nm = [d for d in json.load(f)["devices"] if d["id"] == profile][0]["noise_model"]
    circ = C.ghz_state_circuit(n)
    C.x(circ, n - 1); C.z(circ, 0); C.y(circ, 1)
    extra = C.random_brickwork_circuit(n, 2, seed=3)["operations"]
    circ["operations"] += extra
    dps = O.draws_per_shot(circ, nm)
    u = np.random.default_rng(7).random((shots, dps))
    want = O.run_noisy(circ, nm, u, max_trajectories=8)
    table, keep = _noise_table(nm, n)
    enc = OPS.encode_ops(circ["operations"])
    with L.StateVector(n) as sv:
        assert sv.noisy_draws_per_shot(enc, table) == dps
        outcomes, traj = sv.run_noisy(enc, table, u, max_trajectories=8)
        last = sv.get_state()
    counts = {}
    for o in outcomes:
        bs = O.basis_string(int(o), n)
        counts[bs] = counts.get(bs, 0) + 1
    assert counts == want["measurement-results"]
    for k in range(min(8, shots)):
        assert np.max(np.abs(traj[k] - want["trajectories"][k])) <= TOL
    assert np.max(np.abs(last - want["final-state"])) <= TOL


def test_noisy_trajectory_tree_with_mid_circuit_measure(monkeypatch):
    """The trajectory tree (shots share every common prefix of their Kraus choices AND mid-circuit measurement outcomes;
    device checkpoints at the splits) against the oracle's shot-by-shot loop on the same draws: counts, the first
    trajectories, the last shot's final state - for a circuit with two :measure ops - and against the library's own per-shot
    path (QCB_NOISY_TREE=0); plus the :initial-state option of the noisy path."""
    from qclojure_b200 import ops as OPS
    with open(os.path.join(GOLDEN, "device_profiles.json")) as f:
        nm = [d for d in json.load(f)["devices"] if d["id"] == ":ibm-lagos"][0]["noise_model"]
    n, shots = 6, 300
    circ = C.ghz_state_circuit(n)
    C.measure(circ, [2])
    circ["operations"] += C.random_brickwork_circuit(n, 2, seed=5)["operations"]
    C.measure(circ, [0, 4])
    C.h(circ, 3); C.cnot(circ, 3, 5)
    dps = O.draws_per_shot(circ, nm)
    u = np.random.default_rng(17).random((shots, dps))
    want = O.run_noisy(circ, nm, u, max_trajectories=16)
    table, keep = _noise_table(nm, n)
    enc = OPS.encode_ops(circ["operations"])

    def run():
        with L.StateVector(n) as sv:
            outcomes, traj = sv.run_noisy(enc, table, u, max_trajectories=16)
            return outcomes, traj, sv.get_state(), sv.stats()

    outcomes, traj, last, st_tree = run()
    counts = {}
    for o in outcomes:
        bs = O.basis_string(int(o), n)
        counts[bs] = counts.get(bs, 0) + 1
    assert counts == want["measurement-results"]
    for k in range(16):
        assert np.max(np.abs(traj[k] - want["trajectories"][k])) <= TOL
    assert np.max(np.abs(last - want["final-state"])) <= TOL
    monkeypatch.setenv("QCB_NOISY_TREE", "0")
    # the environment switch is read once per process: compare through a fresh interpreter
    import subprocess, sys, pickle, tempfile
    with tempfile.TemporaryDirectory() as d:
        np.save(os.path.join(d, "u.npy"), u)
        with open(os.path.join(d, "circ.pkl"), "wb") as f:
            pickle.dump((circ, nm, n), f)
        code = ("import sys, pickle, numpy as np; sys.path.insert(0, %r)\n"
                "from qclojure_b200 import _lib as L, ops as OPS, noise as NZ\n"
                "circ, nm, n = pickle.load(open(%r, 'rb')); u = np.load(%r)\n"
                "table, keep = NZ.build_noise_table(nm, n); enc = OPS.encode_ops(circ['operations'])\n"
                "sv = L.StateVector(n); o, t = sv.run_noisy(enc, table, u, max_trajectories=0); np.save(%r, o); print(sv.stats()['n_kernel_launches'])\n"
                % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(d, "circ.pkl"), os.path.join(d, "u.npy"),
                   os.path.join(d, "o.npy")))
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        per_shot = np.load(os.path.join(d, "o.npy"))
        launches_per_shot = int(out.stdout.strip().splitlines()[-1])
    assert np.array_equal(per_shot, outcomes)
    assert st_tree["n_kernel_launches"] < launches_per_shot          # shared prefixes: fewer kernels than one plan per shot
    # :initial-state of the noisy path (hardware_simulator.clj:228-230)
    init = _rand_state(n, 91)
    want_i = O.run_noisy(circ, nm, u[:40], max_trajectories=4, initial_state=init)
    with L.StateVector(n) as sv:
        sv.noisy_set_initial_state(init)
        o1, t1 = sv.run_noisy(enc, table, u[:40], max_trajectories=4)
        sv.noisy_set_initial_state(None)
        o2, _t2 = sv.run_noisy(enc, table, u[:40], max_trajectories=0)
    assert np.array_equal(o2, outcomes[:40])
    for k in range(4):
        assert np.max(np.abs(t1[k] - want_i["trajectories"][k])) <= TOL
    cnt_i = {}
    for o in o1:
        cnt_i[O.basis_string(int(o), n)] = cnt_i.get(O.basis_string(int(o), n), 0) + 1
    assert cnt_i == want_i["measurement-results"]


# ------------------------------------------------------------------ jobs and P2
def test_job_layer():
    from qclojure_b200 import backend as B
    sim = B.create_simulator()
    circ = C.ghz_state_circuit(5)
    u = np.random.default_rng(1).random(256)
    jid = sim.submit_circuit(circ, {"result-specs": {"measurements": {"shots": 256}}, "uniforms": u})
    assert isinstance(jid, str)
    assert sim.job_status(jid) == "completed"            # small jobs finish inside submit
    res = sim.job_result(jid)
    assert res["job-status"] == "completed"
    freq = res["results"]["measurement-results"]["frequencies"]
    assert set(freq) == {0, 31} and sum(freq.values()) == 256
    assert sim.job_status("nope") == "not-found"
    bad = C.add_gate(C.create_circuit(2), "cy", control=0, target=1)
    jid = sim.submit_circuit(bad, {})
    assert sim.job_status(jid) == "failed"
    res = sim.job_result(jid)
    # the reference's job-result reports "Job not completed" for anything but :completed
    # (ideal_simulator.clj:143-152); the real cause is kept under an extra key
    assert res["job-status"] == "failed" and res["error-message"] == "Job not completed"
    assert "Unknown gate type" in res["failure-message"]
    assert sim.cancel_job(jid) == "cannot-cancel" and sim.cancel_job("nope") == "not-found"
    qs = sim.queue_status()
    assert qs["completed"] >= 1 and qs["queued"] == 0
    # blocking helper (application/backend.clj:209-256) + variational-style hamiltonian spec
    from oracle import qc_oracle as OO
    H = [{"coefficient": 0.5, "pauli-string": "ZZIII"}, {"coefficient": -0.3, "pauli-string": "XXXXX"}]
    res = B.execute_circuit(sim, circ, {"result-specs": {"hamiltonian": H, "state-vector": True,
                                                         "expectation": {"observables": [OO.PAULI_Z], "targets": [2]}}})
    want = OO.execute_circuit(circ)
    assert abs(res["results"]["hamiltonian-result"]["energy-expectation"] - OO.hamiltonian_expectation(H, want)) <= TOL
    assert np.max(np.abs(res["results"]["final-state"]["state-vector"] - want)) <= TOL
    assert abs(res["results"]["expectation-results"][0]["expectation-value"] - OO.expectation_1q(want, OO.PAULI_Z, 2)) <= TOL
    sim.close()


def test_result_extraction_every_spec_on_device():
    """every result spec of result.clj:535-639 through the backend on the real device state."""
    from qclojure_b200 import backend as B
    sim = B.create_simulator()
    circ = C.ghz_state_circuit(6)
    C.ry(circ, 2, 0.9); C.rx(circ, 5, 0.3)
    psi = O.execute_circuit(circ)
    H = [{"coefficient": 0.5, "pauli-string": "ZZIIII"}, {"coefficient": -0.25, "pauli-string": "XXXXXX"}]
    full = np.kron(np.kron(O.PAULI_Z, O.PAULI_X), np.eye(16))
    specs = {"measurements": {"shots": 128}, "expectation": {"observables": [O.PAULI_Z, O.PAULI_X], "targets": [0, 2]},
             "variance": {"observables": [O.PAULI_Z], "targets": [2]}, "hamiltonian": H,
             "probabilities": {"targets": [[1, 1, 1, 1, 1, 1], 0]}, "amplitudes": {"basis-states": [0, 63]},
             "state-vector": True, "density-matrix": True, "fidelity": {"references": [O.zero_state(6), psi]},
             "sample": {"observables": [O.PAULI_Z], "shots": 64, "targets": [2]}}
    u = np.random.default_rng(3).random(128 + 64)
    res = B.execute_circuit(sim, circ, {"result-specs": specs, "uniforms": u})
    assert res["job-status"] == "completed", res
    r = res["results"]
    assert r["measurement-results"]["measurement-outcomes"] == O.sample_outcomes(psi, u[:128]).tolist()
    assert abs(r["expectation-results"][1]["expectation-value"] - O.expectation_1q(psi, O.PAULI_X, 2)) <= TOL
    assert abs(r["variance-results"][0]["variance-value"] - O.variance_1q(psi, O.PAULI_Z, 2)) <= TOL
    assert abs(r["hamiltonian-result"]["energy-expectation"] - O.hamiltonian_expectation(H, psi)) <= TOL
    assert abs(r["probability-results"]["probability-outcomes"][(1, 1, 1, 1, 1, 1)] - abs(psi[63]) ** 2) <= TOL
    assert abs(r["amplitude-results"]["amplitude-values"][63] - psi[63]) <= TOL
    assert np.max(np.abs(r["state-vector-result"]["state-vector"] - psi)) <= TOL
    assert np.max(np.abs(r["density-matrix-result"]["density-matrix"] - np.outer(psi, psi.conj()))) <= TOL
    assert r["density-matrix-result"]["trace-valid"]
    assert abs(r["fidelity-results"]["fidelities"]["reference-1"] - 1.0) <= TOL
    assert abs(r["fidelity-results"]["fidelities"]["reference-0"] - abs(psi[0])) <= TOL
    sa = r["sample-results"][0]
    from qclojure_b200 import results as RS
    p_up = (1 + O.expectation_1q(psi, O.PAULI_Z, 2)) / 2
    # one draw stream per job: the :sample spec continues where the measurement shots stopped (no draw is used twice)
    assert sa["sample-outcomes"] == RS.sample_eigenvalues({-1.0: 1 - p_up, 1.0: p_up}, u[128:192])
    # full-register dense observable through the P2 ops (result.clj:279-280)
    res = B.execute_circuit(sim, circ, {"result-specs": {"expectation": {"observables": [full]}}})
    assert abs(res["results"]["expectation-results"][0]["expectation-value"] - np.vdot(psi, full @ psi).real) <= TOL
    sim.close()


def test_hardware_simulator_backend_and_noisy_extraction():
    """hardware_simulator.clj:84-240 + result.clj:642-804 through the protocol methods: counts identical to the oracle on
    the same draws, Tr(rho O) quantities equal to the density-matrix formulas."""
    from qclojure_b200 import backend as B
    with open(os.path.join(GOLDEN, "device_profiles.json")) as f:
        nm = [d for d in json.load(f)["devices"] if d["id"] == ":ibm-lagos"][0]["noise_model"]
    n, shots = 4, 80
    circ = C.ghz_state_circuit(n)
    C.rx(circ, 3, 0.4)
    dps = O.draws_per_shot(circ, nm)
    u = np.random.default_rng(9).random((shots, dps))
    want = O.run_noisy(circ, nm, u, max_trajectories=100)
    rho = O.trajectory_to_density_matrix(want["trajectories"])
    H = [{"coefficient": 1.0, "pauli-string": "ZZII"}, {"coefficient": 0.5, "pauli-string": "IIXX"}]
    Hm = np.kron(np.kron(O.PAULI_Z, O.PAULI_Z), np.eye(4)) + 0.5 * np.kron(np.eye(4), np.kron(O.PAULI_X, O.PAULI_X))
    sim = B.create_hardware_simulator({"id": "ibm-lagos", "noise-model": nm})
    res = B.execute_circuit(sim, circ, {"shots": shots, "uniforms": u})
    assert res["job-status"] == "completed", res
    assert res["shots-executed"] == shots and res["results"]["measurement-results"] == want["measurement-results"]
    assert np.max(np.abs(res["results"]["density-matrix"] - rho)) <= TOL
    specs = {"measurements": {}, "hamiltonian": {"hamiltonian": H}, "probability": {"target-states": [0, 15]},
             "expectation": {"observables": [O.PAULI_Z], "target-qubits": [1]}, "density-matrix": True}
    res = B.execute_circuit(sim, circ, {"shots": shots, "uniforms": u, "result-specs": specs})
    r = res["results"]
    assert r["measurement-results"]["frequencies"] == want["measurement-results"]
    assert abs(r["hamiltonian-result"]["energy-expectation"] - np.trace(rho @ Hm).real) <= TOL
    Z1 = np.kron(np.kron(np.eye(2), O.PAULI_Z), np.eye(4))
    assert abs(r["expectation-results"][0] - np.trace(rho @ Z1).real) <= TOL
    assert abs(r["probability-results"]["probability-outcomes"][15] - rho[15, 15].real) <= TOL
    assert r["density-matrix-result"]["from-trajectories"] is True
    # the bare :hamiltonian spelling the variational objective sends (variational_algorithm.clj:345)
    res = B.execute_circuit(sim, circ, {"shots": shots, "uniforms": u, "result-specs": {"hamiltonian": H}})
    assert abs(res["results"]["hamiltonian-result"]["energy-expectation"] - np.trace(rho @ Hm).real) <= TOL
    sim.close()


def test_c_job_api():
    """qcb_submit / qcb_job_status / qcb_job_result_get / qcb_cancel / qcb_queue_status through ctypes."""
    import ctypes as CT
    import time
    from qclojure_b200 import ops as OPS
    for n in (6, 21):            # 6: completes inside submit; 21: goes through the worker thread
        circ = C.ghz_state_circuit(n)
        u = np.random.default_rng(2).random(64)
        with L.StateVector(n) as sv:
            lib, h = sv._lib, sv._h
            arr, cnt, keep = OPS.encode_ops(circ["operations"])
            req = OPS.QcbJobRequest()
            req.ops, req.n_ops = arr, cnt
            req.uniforms, req.n_shots = u.ctypes.data_as(CT.POINTER(CT.c_double)), 64
            strs = (CT.c_char_p * 1)(("Z" * 2 + "I" * (n - 2)).encode())
            co = (CT.c_double * 1)(2.0)
            req.ham_coeffs, req.ham_strings, req.n_terms = co, strs, 1
            jid = CT.c_uint64()
            L.check(lib.qcb_submit(h, CT.byref(req), CT.byref(jid)), h)
            st = CT.c_int32(-1)
            for _ in range(600):
                lib.qcb_job_status(h, jid.value, CT.byref(st))
                if st.value in (2, 3, 4):
                    break
                time.sleep(0.05)
            assert st.value == 2
            res = OPS.QcbJobResult()
            outs = np.zeros(64, dtype=np.uint64)
            res.outcomes, res.n_shots = outs.ctypes.data_as(CT.POINTER(CT.c_uint64)), 64
            L.check(lib.qcb_job_result_get(h, jid.value, CT.byref(res)), h)
            assert res.status == 2 and res.has_energy == 1 and abs(res.energy - 2.0) <= TOL
            assert set(outs.tolist()) <= {0, (1 << n) - 1}
            lib.qcb_job_status(h, 12345, CT.byref(st))
            assert st.value == 5
            q, r, c = CT.c_uint64(), CT.c_uint64(), CT.c_uint64()
            lib.qcb_queue_status(h, CT.byref(q), CT.byref(r), CT.byref(c))
            assert c.value == 1 and q.value == 0
            lib.qcb_cancel(h, jid.value, CT.byref(st))
            assert st.value == 2          # cannot cancel a finished job: status unchanged


def test_p2_linear_algebra_ops():
    rng = np.random.default_rng(0)

    def rc(*shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    from qclojure_b200 import linalg as LA
    with LA.B200ComplexBackend() as la:
        A, B, x, y = rc(6, 5), rc(5, 7), rc(5), rc(6)
        assert np.max(np.abs(la.matrix_multiply(A, B) - A @ B)) <= TOL
        assert np.max(np.abs(la.matrix_vector_product(A, x) - A @ x)) <= TOL
        assert np.max(np.abs(la.kronecker_product(A, B) - np.kron(A, B))) <= TOL
        assert abs(la.inner_product(y, y[::-1].copy()) - np.vdot(y, y[::-1])) <= TOL
        assert np.max(np.abs(la.outer_product(x, y) - np.outer(x, np.conj(y)))) <= TOL
        S = rc(6, 6)
        assert abs(la.trace(S) - np.trace(S)) <= TOL
        assert abs(la.norm2(x) - np.linalg.norm(x)) <= TOL
        assert np.max(np.abs(la.add(A, A) - 2 * A)) <= TOL
        assert np.max(np.abs(la.scale(A, 2 - 1j) - (2 - 1j) * A)) <= TOL


# ------------------------------------------------------------------ mover variants of the tile kernel
_MOVER_SCRIPT = r"""
import numpy as np
from oracle import qc_oracle as O
from qclojure_b200 import _lib as L, circuits as C
for n, depth in ((9, 6), (14, 6), (18, 8)):
    circ = C.random_brickwork_circuit(n, depth)
    want = O.execute_circuit(circ)
    with L.StateVector(n) as sv:
        sv.apply_circuit(circ)
        err = float(np.max(np.abs(sv.get_state() - want)))
    assert err <= 1e-10, (n, err)
print("mover ok")
"""


@pytest.mark.parametrize("env", [{"QCB_TILE_MOVER": "2"}, {"QCB_TILE_MOVER": "2", "QCB_NO_TMA": "1"}, {"QCB_CONSUMERS": "2x8"},
                                 {"QCB_CONSUMERS": "1x8", "QCB_TILE_BUFFERS": "2"}, {"QCB_PAIR_ROUNDS": "0"},
                                 {"QCB_PAIR_EFF_PCT": "100", "QCB_CONSUMERS": "2x8"}])
def test_tile_kernel_variants_match_oracle(env):
    """The TMA mover (tensor copies + hardware swizzle layout), the same layout moved by the LSU mover, the other consumer
    layouts and a shorter buffer ring give the same amplitudes (the knobs are read once per process, hence the subprocess)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", _MOVER_SCRIPT], cwd=root, env={**os.environ, **env}, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0 and "mover ok" in out.stdout, out.stderr[-2000:]


def test_paired_rounds_on_device_match_oracle():
    """Paired passes (round kind 3: the second dense block consumes the first block's D registers) against the oracle, with
    controls / diagonal operands as condition bits of either block, at sizes with one tile (12 q), many tiles (18 q) and
    tiles whose id bits are condition bits (22 q vs the C oracle)."""
    from tests.test_oracle_c import _all_gates_circuit
    for n in (12, 15, 18):
        rng = np.random.default_rng(40 + n)
        circ = _all_gates_circuit(n, rng)
        init = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        init /= np.linalg.norm(init)
        want = O.execute_circuit(circ, init)
        ps = L.plan_summary(n, circ["operations"])
        assert ps["paired_passes"] > 0
        with L.StateVector(n) as sv:
            sv.set_state(init)
            sv.apply_circuit(circ)
            assert float(np.max(np.abs(sv.get_state() - want))) <= 1e-10
    n = 22
    circ = C.random_brickwork_circuit(n, 12)
    ps = L.plan_summary(n, circ["operations"])
    assert ps["paired_passes"] > 0 and ps["rounds"] == ps["passes"] + ps["paired_passes"]
    want = CO.apply_circuit(circ)
    with L.StateVector(n) as sv:
        sv.apply_circuit(circ)
        assert float(np.max(np.abs(sv.get_state() - want))) <= 1e-10


def test_far_phases_qft_on_device():
    """QFT ladders with controls outside the tile ride as tile-constant row scalings of the A fragments (far phases): amplitudes
    against the oracle at 14 - 22 qubits (2 - 10 far bits), with far phases off for comparison, and the plan really uses them."""
    for n in (14, 18, 22):
        circ = C.quantum_fourier_transform_circuit(n)
        init = _rand_state(n, 70 + n)
        want = O.execute_circuit(circ, init)
        with L.StateVector(n) as sv:
            sv.set_state(init)
            sv.apply_circuit(circ)
            assert float(np.max(np.abs(sv.get_state() - want))) <= TOL
    # a mix of CRZ (both orientations), CZ and non-diagonal gates on the same qubits
    n = 20
    rng = np.random.default_rng(5)
    circ = C.create_circuit(n)
    for _ in range(300):
        a, b = (int(x) for x in rng.choice(n, 2, replace=False))
        k = int(rng.integers(0, 6))
        if k == 0: C.add_gate(circ, "h", target=a)
        elif k == 1: C.rx(circ, a, rng.random() * 6)
        elif k == 2: C.crz(circ, a, b, rng.random() * 6)
        elif k == 3: C.cz(circ, a, b)
        elif k == 4: C.cnot(circ, a, b)
        else: C.rz(circ, a, rng.random() * 6)
    init = _rand_state(n, 6)
    want = O.execute_circuit(circ, init)
    with L.StateVector(n) as sv:
        sv.set_state(init)
        sv.apply_circuit(circ)
        assert float(np.max(np.abs(sv.get_state() - want))) <= TOL


def test_swap_kernel_unit_all_partner_schedules():
    """k_swap_global (the in-place multi-qubit exchange over peer memory) for worlds of 2, 4, 8 ranks emulated as slices on
    ONE device, k = 1..3 exchanged qubits at arbitrary positions: tests/cuda/swap_kernel_test.cu (54 cases)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cuda", "swap_kernel_test")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(root, "tests", "cuda"), "-s", "swap_kernel_test"], timeout=600)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "swap kernel ok" in out.stdout, (out.stdout[-1000:], out.stderr[-1000:])


def test_multi_gpu_sharded_matches_oracle():
    """The sharded path (top log2 P qubits global, NCCL qubit exchanges) on every GPU of the box against the oracle:
    tests/multi_gpu_check.py under torchrun.  Skipped on a single-GPU box."""
    import subprocess
    import sys
    ngpu = L.device_count()
    world = 1 << (ngpu.bit_length() - 1)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "multi-gpu ok" in out.stdout, (out.stdout[-2000:], out.stderr[-3000:])


def test_single_process_multi_gpu_handle():
    """ONE handle (qcb_config.n_gpus) owning the whole sharded state, driven by one host process - the mode a JVM host uses
    (SURVEY 8b).  tests/group_check.py in a subprocess on every GPU of the box; skipped on a single-GPU box."""
    import subprocess
    import sys
    ngpu = L.device_count()
    world = 1 << (ngpu.bit_length() - 1)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "group_check.py"), str(world)], cwd=root, capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0 and "group ok" in out.stdout, (out.stdout[-2000:], out.stderr[-3000:])


def test_multi_gpu_handle_rejects_bad_configs():
    with pytest.raises(L.QcbError):
        L.StateVector(10, n_gpus=3)
    with pytest.raises(L.QcbError):
        L.StateVector(10, n_gpus=2, world_size=2)
    if L.device_count() < 8:
        with pytest.raises(L.QcbError):
            L.StateVector(12, n_gpus=8)


def test_config2_grover_26q_full_size_property():
    """BASELINE.json configs[1] at its full size (26 qubits, 1 GiB state): after k oracle + diffusion operators the marked
    amplitude is sin((2k+1) theta) and every other amplitude cos((2k+1) theta) / sqrt(N-1), theta = asin(1/sqrt(N)) - a
    size-independent closed form (the oracle would need minutes here).  64 iterations through the fused Grover pass."""
    n, k, target = 26, 64, 0x2AAAAAA
    N = 1 << n
    ops = [{"operation-type": "global-h", "operation-params": {}}]
    for _ in range(k):
        ops += [{"operation-type": "phase-oracle", "operation-params": {"index": target}},
                {"operation-type": "grover-diffusion", "operation-params": {}}]
    theta = math.asin(1.0 / math.sqrt(N))
    idx = np.random.default_rng(26).integers(0, N, 1024)
    idx = idx[idx != target]
    with L.StateVector(n) as sv:
        sv.apply_ops(ops)
        amps = sv.get_amplitudes(np.concatenate([[target], idx]))
        assert abs(sv.norm2() - 1.0) <= TOL
        assert sv.stats()["n_sweeps"] <= k + 8                     # one streaming pass per iteration
    assert abs(amps[0] - math.sin((2 * k + 1) * theta)) <= TOL
    assert np.max(np.abs(amps[1:] - math.cos((2 * k + 1) * theta) / math.sqrt(N - 1))) <= TOL


def test_hhl_tutorial_probabilities_on_gpu():
    """JVM-recorded HHL run of the reference's tutorial (tests/golden/hhl_tutorial.json): 64 probabilities of a circuit with
    four :cry gates - pins the transposed controlled gate on the CUDA path (strict_parity = 1)."""
    with open(os.path.join(GOLDEN, "hhl_tutorial.json")) as f:
        g = json.load(f)
    circ = C.hhl_circuit(g["matrix"], g["vector"], g["precision_qubits"], g["ancilla_qubits"])
    want = np.array(g["all_probabilities"])
    with L.StateVector(6) as sv:
        sv.apply_circuit(circ)
        assert np.max(np.abs(sv.probabilities() - want)) <= TOL
    with L.StateVector(6, strict_parity=0) as sv:
        sv.apply_circuit(circ)
        assert np.max(np.abs(sv.probabilities() - want)) > 0.05
