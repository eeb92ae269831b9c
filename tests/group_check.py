"""Single-process multi-GPU check: ONE handle created with qcb_config.n_gpus = N owns the whole sharded
state; one host thread per device inside the library, nothing SPMD on the caller's side - this is how a JVM host (one
`submit-circuit` caller, application/backend.clj:72-112) drives 2 / 4 / 8 GPUs.

    python tests/group_check.py [N]            (N defaults to the largest power of two <= visible GPUs)

Checks against the oracle: the whole state (global offsets), the norm, two Hamiltonian energies (one with X / Y factors on
the global qubits), shot outcomes on identical uniforms, spot amplitudes, a probability range that crosses slice borders, a
1-qubit observable on a GLOBAL qubit, a mid-circuit measurement of a global and a local qubit, :initial-state, and the job
API (qcb_submit through the Python mirror's backend with config "n-gpus")."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import c_oracle as CO
    from oracle import qc_oracle as O
    from qclojure_b200 import _lib as L
    from qclojure_b200 import backend as B
    from qclojure_b200 import circuits as C

    ngpu = L.device_count()
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << (ngpu.bit_length() - 1)
    assert N >= 2 and ngpu >= N, f"need {N} GPUs, {ngpu} visible"
    p = N.bit_length() - 1
    sizes = [tuple(int(v) for v in item.split(":")) for item in os.environ.get("QCB_MGC_LOCAL", "12:6,18:8,21:8").split(",")]
    worst = 0.0
    for n, depth in [(nl + p, d) for nl, d in sizes]:
        circ = C.random_brickwork_circuit(n, depth)
        want = CO.apply_circuit(circ) if n > 16 else O.execute_circuit(circ)
        u = np.random.default_rng(n).random(512)
        H = C.max_cut_hamiltonian(C.random_regular_graph(n, 3 if n % 2 == 0 else 4, seed=11), n)
        HX = C.standard_mixer_hamiltonian(n) + [
            {"coefficient": 0.7, "pauli-string": "XY" + "I" * (n - 3) + "Z"}, {"coefficient": -0.4, "pauli-string": "Y" * 3 + "I" * (n - 3)}]
        Y = np.array([[0, -1j], [1j, 0]])
        with L.StateVector(n, n_gpus=N) as sv:
            sv.apply_circuit(circ)
            stats = sv.stats()
            nrm = sv.norm2()
            energy = sv.expect_hamiltonian(H)
            energy_x = sv.expect_hamiltonian(HX)
            e1 = sv.expect_1q(Y, 0)                        # qubit 0 = the most global qubit
            e1l = sv.expect_1q(Y, n - 1)
            shots = sv.sample(u)
            got = sv.get_state()
            lc = 1 << (n - p)
            spot_idx = np.random.default_rng(n).integers(0, 1 << n, 257)
            spots = sv.get_amplitudes(spot_idx)
            probs = sv.probabilities(lc - 100, 300)        # crosses the border between slices 0 and 1
            part = sv.get_state(lc - 5, 10)
            # collapse a global and a local qubit, compare with the oracle's measure-specific-qubits on the same draw
            bits, pm = sv.measure_qubits([0, n - 2], 0.37)
            after = sv.get_state()
            # :initial-state round trip: a random state goes in, comes back, and a circuit on it matches the oracle
            rs = np.random.default_rng(5).standard_normal(1 << n) + 1j * np.random.default_rng(6).standard_normal(1 << n)
            rs /= np.linalg.norm(rs)
            sv.set_state(rs)
            back = sv.get_state()
            sv.apply_circuit(circ)
            got2 = sv.get_state()
        err = float(np.max(np.abs(got - want)))
        worst = max(worst, err)
        assert err <= 1e-10, f"n={n}: amplitude mismatch {err}"
        assert abs(nrm - 1.0) <= 1e-10
        assert abs(energy - O.hamiltonian_expectation(H, want)) <= 1e-9
        assert abs(energy_x - O.hamiltonian_expectation(HX, want)) <= 1e-9, "energy with X/Y on global qubits"
        want_e1 = float(np.real(np.vdot(want, O.apply_single_qubit_gate(want, Y, 0))))
        want_e1l = float(np.real(np.vdot(want, O.apply_single_qubit_gate(want, Y, n - 1))))
        assert abs(e1 - want_e1) <= 1e-10 and abs(e1l - want_e1l) <= 1e-10, (e1, want_e1, e1l, want_e1l)
        ref = O.sample_outcomes(want, u)
        dist_b = O.sample_boundary_distance(want, u)
        assert not ((shots != ref) & (dist_b > 1e-12)).any(), "shot outcomes differ"
        assert np.max(np.abs(spots - want[spot_idx])) <= 1e-10
        assert np.max(np.abs(probs - np.abs(want[lc - 100:lc + 200]) ** 2)) <= 1e-10
        assert np.max(np.abs(part - want[lc - 5:lc + 5])) <= 1e-10
        want_bits, want_after, _probs = O.measure_specific_qubits(want, [0, n - 2], 0.37)
        assert [int(b) for b in bits] == [int(b) for b in want_bits], (bits, want_bits)
        assert np.max(np.abs(after - want_after)) <= 1e-10
        assert np.array_equal(back, rs)
        assert np.max(np.abs(got2 - (CO.apply_circuit(circ, rs) if n > 16 else O.execute_circuit(circ, rs)))) <= 1e-10
        print(f"n={n} gpus={N}: max|err|={err:.2e} exchanges={stats['n_exchanges']} sweeps={stats['n_sweeps']} "
              f"exchange_ms={stats['exchange_ms']:.3f} gpu_ms={stats['gpu_ms']:.3f}", flush=True)
    # the backend mirror over one multi-GPU handle (job layer: one caller, one job id)
    n = 20 + p
    sim = B.create_simulator({"n-gpus": N, "multi-gpu-min-qubits": 0, "max-state-qubits": 26})
    circ = C.ghz_state_circuit(n)
    uu = np.random.default_rng(1).random(256)
    res = B.execute_circuit(sim, circ, {"result-specs": {"measurements": {"shots": 256}}, "uniforms": uu}, poll_s=0.01)
    assert res["job-status"] == "completed", res
    outs = res["results"]["measurement-results"]["measurement-outcomes"]
    assert set(int(o) for o in outs) <= {0, (1 << n) - 1}, "GHZ outcomes"
    fs = res["results"]["final-state"]["state-vector"]
    assert abs(abs(fs[0]) - 2 ** -0.5) <= 1e-12 and abs(abs(fs[-1]) - 2 ** -0.5) <= 1e-12
    sim.close()
    print(f"group ok: gpus={N} max|err|={worst:.2e}", flush=True)


if __name__ == "__main__":
    main()
