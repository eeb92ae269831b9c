"""Small end-to-end run of every kernel variant, meant to be executed under compute-sanitizer (memcheck / racecheck /
synccheck) on a GPU box:

    compute-sanitizer --tool racecheck python tests/sanitize_check.py

Sizes are tiny (<= 13 qubits) because the sanitizer slows kernels down by 10-100x; every result is still compared with the
oracle so that a tool-induced scheduling change that exposes a real race also shows up as a wrong amplitude."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import qc_oracle as O
    from qclojure_b200 import _lib as L
    from qclojure_b200 import circuits as C
    from tests.test_oracle_c import _all_gates_circuit

    def check(tag, n, ops, **kw):
        want = O.execute_circuit({"num-qubits": n, "operations": ops})
        with L.StateVector(n, **kw) as sv:
            sv.apply_ops(ops)
            got = sv.get_state()
        err = float(np.max(np.abs(got - want)))
        assert err <= 1e-10, (tag, err)
        print(f"ok {tag}: n={n} err={err:.1e}", flush=True)

    if "--single-tile" in sys.argv:
        # ONE tile per CTA and sweep (12 qubits = the tile): no buffer of the ring is ever reused, so the only accesses ordered
        # through mbarriers are mover-load -> first round and last round -> mover write-back; the round-to-round hand-over
        # inside a consumer group goes through bar.sync alone.  racecheck models bar.sync but not mbarrier phases
        # (tests/cuda/mbar_racecheck_probe.cu): here it must not report any store / load pair of the rounds themselves.
        check("single tile, three-product rounds", 12, C.random_brickwork_circuit(12, 10)["operations"])
        print("sanitize_check ok (single tile)", flush=True)
        return
    brick13 = C.random_brickwork_circuit(13, 6)["operations"]       # tile = 12 bits: specialised mover, 2 tiles per sweep
    brick12 = C.random_brickwork_circuit(12, 6)["operations"]
    check("three-product rounds, cp.async mover", 13, brick13)
    check("16x16 real rounds (legacy form)", 13, brick13, dense_mma=3)
    check("interpreter rounds only", 12, brick12, dense_mma=2)
    check("unfused (one gate per sweep)", 12, brick12[:40], fusion=0)
    check("TMA mover", 13, brick13, tile_mover=2)
    check("small tiles (generic mover)", 12, brick12, tile_bits=8, low_bits=3)
    check("every gate kind", 8, _all_gates_circuit(8, np.random.default_rng(8))["operations"])
    os.environ["QCB_DIRECT_STORE"] = "1"
    check("direct store", 13, brick13)
    os.environ.pop("QCB_DIRECT_STORE")
    # reductions, sampling, expectation, marginal, collapse, Grover pass
    n = 12
    init = np.random.default_rng(3).standard_normal(1 << n) + 1j * np.random.default_rng(4).standard_normal(1 << n)
    init /= np.linalg.norm(init)
    u = np.random.default_rng(5).random(64)
    H = C.max_cut_hamiltonian(C.random_regular_graph(n, 3, seed=11), n) + [{"coefficient": 0.3, "pauli-string": "XY" + "I" * (n - 2)}]
    with L.StateVector(n) as sv:
        sv.set_state(init)
        assert abs(sv.norm2() - 1.0) <= 1e-12
        assert abs(sv.expect_hamiltonian(H) - O.hamiltonian_expectation(H, init)) <= 1e-10
        ref = O.sample_outcomes(init, u)
        dist = O.sample_boundary_distance(init, u)
        assert not ((sv.sample(u) != ref) & (dist > 1e-12)).any()
        assert np.max(np.abs(sv.probabilities() - np.abs(init) ** 2)) <= 1e-12
        bits, col, probs = O.measure_specific_qubits(init, [1, 7, 3], 0.4)
        assert np.max(np.abs(sv.marginal_probabilities([1, 7, 3]) - np.array(probs))) <= 1e-10
        assert sv.measure_qubits([1, 7, 3], 0.4)[0] == bits
        assert np.max(np.abs(sv.get_state() - col)) <= 1e-10
        sv.normalize()
    # block-wise reduction kernels (slices of >= 2^11 amplitudes with six free index bits): marginal histogram, diagonal terms
    n2 = 14
    init2 = np.random.default_rng(6).standard_normal(1 << n2) + 1j * np.random.default_rng(7).standard_normal(1 << n2)
    init2 /= np.linalg.norm(init2)
    H2 = C.max_cut_hamiltonian(C.random_regular_graph(n2, 3, seed=11), n2)
    with L.StateVector(n2) as sv:
        sv.set_state(init2)
        assert abs(sv.expect_hamiltonian(H2) - O.hamiltonian_expectation(H2, init2)) <= 1e-10
        for qs in ([0, 13], [2], [13, 12, 11, 0]):
            _bits, _col, probs2 = O.measure_specific_qubits(init2, qs, 0.4)
            assert np.max(np.abs(sv.marginal_probabilities(qs) - np.array(probs2))) <= 1e-10
    ops = [{"operation-type": "global-h", "operation-params": {}}]
    for _ in range(3):
        ops += [{"operation-type": "phase-oracle", "operation-params": {"index": 77}}, {"operation-type": "grover-diffusion", "operation-params": {}}]
    with L.StateVector(n) as sv:
        sv.apply_ops(ops)
        amp = sv.get_amplitudes([77])[0]
    th = np.arcsin(2.0 ** (-n / 2))
    assert abs(amp - np.sin(7 * th)) <= 1e-10
    print("sanitize_check ok", flush=True)


if __name__ == "__main__":
    main()
