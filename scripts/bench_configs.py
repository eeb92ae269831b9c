"""Timings of the other BASELINE.json configs (they are parity-test cases, not bench lines): config 1 QFT + GHZ 20 q with
1024 shots, config 2 Grover 26 q with the fused oracle / diffusion operators, config 5 noisy trajectories (ibm-lagos
profile) and a QAOA energy sweep.  Wall clock around the public calls, after one warm-up.

`bench.py` folds `run(quick=True)` into its JSON line as `other_configs` (with the clocks seen meanwhile);
`python scripts/bench_configs.py` prints the full-size version."""
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def timed(f, reps=3):
    f()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    return (time.perf_counter() - t0) / reps


def run(device: int = 0, quick: bool = False) -> dict:
    from qclojure_b200 import _lib as L, backend as B, circuits as C, noise as NZ, ops as OPS
    peak = 6550.4
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
    except Exception:
        pass
    out = {}
    # ---- config 1
    sim = B.create_simulator({"device": device})
    u = np.random.default_rng(20261017).random(1024)
    for name, circ in (("qft20", C.quantum_fourier_transform_circuit(20)), ("ghz20", C.ghz_state_circuit(20))):
        opt = {"result-specs": {"measurements": {"shots": 1024}}, "uniforms": u}
        dt = timed(lambda: B.execute_circuit(sim, circ, opt, poll_s=0.0005, max_polls=100000))
        out[f"config1_{name}_1024shots_ms"] = 1e3 * dt
    sim.close()
    # the same circuit family at scale: QFT on 30 qubits (480 gates, every rotation controlled by a later qubit), device time of
    # set_zero + apply_ops on a resident state
    try:
        n30 = 30
        enc30 = OPS.encode_ops(C.quantum_fourier_transform_circuit(n30)["operations"])
        with L.StateVector(n30, device=device) as sv:
            for _ in range(2):
                sv.set_zero(); sv.apply_ops(enc30)
            sv.synchronize()
            sv.timer_start()
            for _ in range(3):
                sv.set_zero(); sv.apply_ops(enc30)
            ms30 = sv.timer_stop() / 3
            st30 = sv.stats()
            a0 = sv.get_amplitudes([0, (1 << n30) - 1])
        out["config1_qft30_device_ms"] = ms30
        out["config1_qft30_sweeps_rounds"] = [int(st30["n_sweeps"]), int(st30["n_rounds"])]
        out["config1_qft30_uniform_amp_err"] = float(np.max(np.abs(a0 - 2.0 ** (-n30 / 2))))
    except Exception as ex:      # noqa: BLE001
        out["config1_qft30_device_ms"] = f"skipped: {ex}"
    # ---- config 2: Grover 26 q, 64 iterations of oracle + diffusion
    n = 26
    ops = [{"operation-type": "global-h", "operation-params": {}}]
    for _ in range(64):
        ops += [{"operation-type": "phase-oracle", "operation-params": {"index": 0x2AAAAAA}},
                {"operation-type": "grover-diffusion", "operation-params": {}}]
    enc = OPS.encode_ops(ops)
    with L.StateVector(n, device=device) as sv:
        def run_g():
            sv.set_zero(); sv.apply_ops(enc); sv.synchronize()
        dt = timed(run_g)
        st = sv.stats()
        out["config2_grover26_ms_per_iteration"] = 1e3 * dt / 64
        out["config2_grover26_hbm_frac"] = (32.0 * (1 << n) / (dt / 64)) / 1e9 / peak      # one 32 B/amplitude pass per iteration
        out["config2_grover26_sweeps_per_iteration"] = st["n_sweeps"] / 64
        out["config2_grover26_full_6433_iterations_s_estimate"] = dt / 64 * C.grover_iterations(26)
        out["config2_grover26_p_target_after_64"] = float(abs(sv.get_amplitudes([0x2AAAAAA])[0]) ** 2)
    # ---- config 5: noisy trajectories
    with open(os.path.join(ROOT, "tests", "golden", "device_profiles.json")) as f:
        nm = [d for d in json.load(f)["devices"] if d["id"] == ":ibm-lagos"][0]["noise_model"]
    for n, shots in ((7, 1024), (12, 1024), (20, 256)):
        circ = C.ghz_state_circuit(n)
        circ["operations"] += C.random_brickwork_circuit(n, 4, seed=3)["operations"]
        table, keep = NZ.build_noise_table(nm, n)
        enc = OPS.encode_ops(circ["operations"])
        with L.StateVector(n, device=device) as sv:
            dps = sv.noisy_draws_per_shot(enc, table)
            uu = np.random.default_rng(7).random((shots, dps))
            dt = timed(lambda: sv.run_noisy(enc, table, uu, max_trajectories=0), reps=1)
        out[f"config5_noisy_ghz+brick4_{n}q_shots_per_s"] = shots / dt
    # ---- config 5: QAOA sweep at 20 qubits, p = 2 (21 x 21 grid; quick: 9 x 9)
    n = 20
    grid = 9 if quick else 21
    graph = C.random_regular_graph(n, 3, seed=11)
    Hp, Hm = C.max_cut_hamiltonian(graph, n), C.standard_mixer_hamiltonian(n)
    with L.StateVector(n, device=device) as sv:
        def sweep():
            for g in np.linspace(0, math.pi, grid):
                for b in np.linspace(0, math.pi, grid):
                    sv.set_zero()
                    sv.apply_circuit(C.qaoa_ansatz_circuit(Hp, Hm, [g, b, 0.5 * g, 0.5 * b], n))
                    sv.expect_hamiltonian(Hp)
        dt = timed(sweep, reps=1)
    out["config5_qaoa20_p2_energies_per_s"] = grid * grid / dt
    return out


if __name__ == "__main__":
    print(json.dumps(run(), indent=1))
