"""Where does a noisy run spend its time?  Trajectory tree vs the plain per-sequence grouping (QCB_NOISY_TREE=0 in a
child process), 20 qubits, GHZ + depth-4 brickwork, ibm-lagos profile: wall time, device time, sweeps, launches."""
import json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qclojure_b200 import _lib as L, circuits as C, noise as NZ, ops as OPS  # noqa: E402

def run(n, shots, measure):
    with open(os.path.join(ROOT, "tests", "golden", "device_profiles.json")) as f:
        nm = [d for d in json.load(f)["devices"] if d["id"] == ":ibm-lagos"][0]["noise_model"]
    circ = C.ghz_state_circuit(n)
    if measure:
        C.measure(circ, [1])
    circ["operations"] += C.random_brickwork_circuit(n, 4, seed=3)["operations"]
    table, keep = NZ.build_noise_table(nm, n)
    enc = OPS.encode_ops(circ["operations"])
    with L.StateVector(n) as sv:
        dps = sv.noisy_draws_per_shot(enc, table)
        uu = np.random.default_rng(7).random((shots, dps))
        sv.run_noisy(enc, table, uu, max_trajectories=0)
        t0 = time.perf_counter()
        sv.run_noisy(enc, table, uu, max_trajectories=0)
        dt = time.perf_counter() - t0
        st = sv.stats()
    return {"n": n, "shots": shots, "measure": measure, "tree": os.environ.get("QCB_NOISY_TREE", "1"), "shots_per_s": round(shots / dt),
            "wall_ms": round(1e3 * dt, 2), "gpu_ms": round(st["gpu_ms"], 2), "sweeps": st["n_sweeps"], "launches": st["n_kernel_launches"]}

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        print(json.dumps([run(12, 1024, False), run(20, 256, False), run(12, 256, True), run(20, 64, True)]))
    else:
        for tree in ("1", "0"):
            env = dict(os.environ, QCB_NOISY_TREE=tree)
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True, timeout=900)
            for r in json.loads(out.stdout.strip().splitlines()[-1]) if out.returncode == 0 else [out.stderr[-500:]]:
                print(r)
