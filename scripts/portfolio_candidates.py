"""Host only: the cost-model score of every member of the plan portfolio (csrc/plan.cpp: schedule) for the benchmark circuits,
the chosen one starred.  Members 1-5 are the base settings, 6-9 the same pairing settings with a stage-yield threshold of
35 %, 10-13 with 75 %.  Usage: python scripts/portfolio_candidates.py > profiles/<tag>_portfolio_candidates.txt"""
import os
import sys

os.environ["QCB_PORTFOLIO_DEBUG"] = "1"
sys.path.insert(0, ".")
from qclojure_b200 import _lib as L                    # noqa: E402
from qclojure_b200 import circuits as C                # noqa: E402

cases = [(n, 1) for n in range(24, 34)] + [(31, 2), (32, 4), (33, 8), (34, 2), (35, 4), (36, 8)]
for n, world in cases:
    ops = C.random_brickwork_circuit(n, 20)["operations"]
    sys.stderr.write(f"brickwork depth 20, {n} qubits on {world} GPU(s): ")
    sys.stderr.flush()
    ps = L.plan_summary(n, ops, rank=0, world_size=world)
    sys.stderr.write(f"    -> sweeps {ps['tile_sweeps']} passes {ps['passes']} paired {ps['paired_passes']} exchanges {ps['exchanges']}\n")
ops = C.quantum_fourier_transform_circuit(30)["operations"]
sys.stderr.write("QFT, 30 qubits on 1 GPU: ")
sys.stderr.flush()
ps = L.plan_summary(30, ops)
sys.stderr.write(f"    -> sweeps {ps['tile_sweeps']} passes {ps['passes']} paired {ps['paired_passes']}\n")
