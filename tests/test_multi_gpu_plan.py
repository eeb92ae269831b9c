"""Multi-GPU host logic on CPU: the qubit-remapping scheduler (global <-> local exchanges), rank-dependent
diagonal/control handling and the layout bookkeeping, checked by emulating every rank of a 2/4/8-GPU run in
one process (tests/emu) against the oracle.  The gloo world_size-2 test drives the same plan with real
torch.distributed send/recv between two processes."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import qc_oracle as O
from qclojure_b200 import circuits as C
from tests.emu import emu as E
from tests.test_oracle_c import _all_gates_circuit

TOL = 1e-10
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("kind", ["brick", "qft", "ghz", "all"])
def test_sharded_plan_matches_oracle(world, kind):
    n = 11
    rng = np.random.default_rng(world * 10 + len(kind))
    circ = {"brick": lambda: C.random_brickwork_circuit(n, 6), "qft": lambda: C.quantum_fourier_transform_circuit(n),
            "ghz": lambda: C.ghz_state_circuit(n), "all": lambda: _all_gates_circuit(n, rng)}[kind]()
    init = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    init /= np.linalg.norm(init)
    want = O.execute_circuit(circ, init)
    got, plans = E.run_world(n, circ["operations"], init, world=world, tile_bits=5, low_bits=2, return_plans=True)
    assert np.max(np.abs(got - want)) <= TOL
    # every rank must have planned the same stage sequence (SPMD)
    kinds = [[p.stage_kind(i) for i in range(p.num_stages)] for p in plans]
    assert all(k == kinds[0] for k in kinds)


def test_diagonal_and_control_on_global_qubits_need_no_exchange():
    """diagonal gates and controls on global qubits are decided by the rank id."""
    n, world = 10, 4
    circ = C.create_circuit(n)
    for q in range(2, n):
        C.h(circ, q)
    C.z(circ, 0); C.rz(circ, 1, 0.3); C.cz(circ, 0, 5); C.cnot(circ, 1, 7); C.crz(circ, 0, 1, 0.4); C.t_gate(circ, 0)
    C.toffoli(circ, 0, 1, 4); C.add_gate(circ, "rydberg-blockade", qubit_indices=[0, 1, 3], angle=0.2)
    got, plans = E.run_world(n, circ["operations"], world=world, tile_bits=5, low_bits=2, return_plans=True)
    assert np.max(np.abs(got - O.execute_circuit(circ))) <= TOL
    assert all(plans[0].stage_kind(i) != E.S_EXCHANGE for i in range(plans[0].num_stages))
    # a dense gate on a global qubit does need one
    C.h(circ, 0)
    got, plans = E.run_world(n, circ["operations"], world=world, tile_bits=5, low_bits=2, return_plans=True)
    assert np.max(np.abs(got - O.execute_circuit(circ))) <= TOL
    assert sum(plans[0].stage_kind(i) == E.S_EXCHANGE for i in range(plans[0].num_stages)) == 1


def test_exchange_count_is_small_for_brickwork():
    """The remap scheduler batches gates per global qubit: a depth-20 brickwork circuit on 8 ranks needs
    far fewer exchanges than it has gates on global qubits."""
    n, world = 33 + 3, 8
    circ = C.random_brickwork_circuit(n, 20)
    p = E.EmuPlan(n, circ["operations"], rank=0, world=world)
    ex = sum(p.stage_kind(i) == E.S_EXCHANGE for i in range(p.num_stages))
    global_gate_count = sum(1 for op in circ["operations"]
                            if op["operation-type"] in ("h", "rx") and op["operation-params"]["target"] < 3)
    assert ex <= global_gate_count + 6
    assert ex <= 70


@pytest.mark.parametrize("n,world,max_ex,max_sweeps", [(31, 2, 1, 19), (32, 4, 2, 20), (33, 8, 3, 24), (36, 8, 3, 22)])
def test_batched_exchanges_with_free_partners(n, world, max_ex, max_sweeps):
    """Weak-scaling benchmark plans: once the local part of the circuit is done, every global qubit
    is swapped against a qubit no pending gate targets any more, back to back - log2(P) exchanges for the whole circuit
    and no thin sweeps between them (was 2 / 4 / 6 / 6 exchanges and 19 / 21 / 27 / 31 sweeps)."""
    circ = C.random_brickwork_circuit(n, 20)
    p = E.EmuPlan(n, circ["operations"], rank=0, world=world)
    kinds = [p.stage_kind(i) for i in range(p.num_stages)]
    ex = kinds.count(E.S_EXCHANGE)
    assert ex <= max_ex and len(kinds) - ex <= max_sweeps
    # the exchanges come back to back, and their local partners are low-traffic bits with >= 64 KiB rows
    first = kinds.index(E.S_EXCHANGE)
    assert kinds[first:first + ex] == [E.S_EXCHANGE] * ex
    n_local = n - (world.bit_length() - 1)
    for i in range(first, first + ex):
        g, l = p.stage_exchange(i)
        assert g >= n_local and 12 <= l < n_local


@pytest.mark.parametrize("world", [4, 8])
def test_batched_exchanges_emulated_parity(world):
    """Deep brickwork at a size the emulator finishes: the local part completes, all global qubits come in at once, the
    rest runs - amplitudes against the oracle, layout restored through perm_out."""
    n = 13
    circ = C.random_brickwork_circuit(n, 12, seed=5)
    got, plans = E.run_world(n, circ["operations"], world=world, tile_bits=5, low_bits=2, return_plans=True)
    assert np.max(np.abs(got - O.execute_circuit(circ))) <= TOL
    kinds = [plans[0].stage_kind(i) for i in range(plans[0].num_stages)]
    assert kinds.count(E.S_EXCHANGE) >= 1


def test_gloo_two_process_exchange():
    """world_size-2 run over torch.distributed (gloo): each process emulates its rank's tile stages and
    exchanges halves with real send/recv; rank 0 compares the gathered state with the oracle."""
    script = os.path.join(ROOT, "tests", "gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, script, str(r), "2"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "PARITY OK" in outs[0]
