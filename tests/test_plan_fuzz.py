"""Randomised checks of the host side of the gate executor (lowering, scheduler, exchanges, program encoding, interpreters)
on the CPU: random circuits over the whole gate vocabulary of SURVEY.md §8a', random tile geometry, 1 / 2 / 4 / 8 emulated
ranks, against the oracle; and plan-trace replay against fresh scheduling.  Fixed seeds (a wider sweep of 3000 seeds was
run while developing; it found the tile-geometry hang and the tiny-slice exchange gap that `config_from` / the exchange
partner fallback now close; the last session of round 2 ran another 4000 + 2000 small cases, 1700 cases with 12 - 17 qubits and
production-size tiles incl. QFT-like ladders, 2800 replay-equals-fresh cases and 2000 zero-support cases: all green)."""
import math

import numpy as np
import pytest

from oracle import qc_oracle as O
from qclojure_b200 import circuits as C
from tests.emu import emu as E
from tests.test_plan_trace import _fresh, _reangle, _replayed

TOL = 1e-10
_ONE = ["h", "x", "y", "z", "s", "t", "s-dag", "t-dag"]


def random_circuit(n, rng, length, grover=False):
    circ = C.create_circuit(n)
    for _ in range(length):
        r = int(rng.integers(0, 15 if grover else 14))
        q = [int(x) for x in rng.permutation(n)[:3]]
        a = float(rng.uniform(0, 2 * math.pi))
        if r == 0:
            C.add_gate(circ, _ONE[rng.integers(0, len(_ONE))], target=q[0])
        elif r == 1:
            C.rx(circ, q[0], a)
        elif r == 2:
            C.ry(circ, q[0], a)
        elif r == 3:
            C.rz(circ, q[0], a)
        elif r == 4:
            C.phase(circ, q[0], a if rng.random() < 0.8 else math.pi)
        elif r == 5:
            C.cnot(circ, q[0], q[1])
        elif r == 6:
            C.cz(circ, q[0], q[1])
        elif r == 7:
            C.add_gate(circ, ["crx", "cry", "crz"][rng.integers(0, 3)], control=q[0], target=q[1], angle=a)
        elif r == 8:
            (C.swap if rng.random() < 0.5 else C.iswap)(circ, q[0], q[1])
        elif r == 9 and n >= 3:
            C.toffoli(circ, q[0], q[1], q[2])
        elif r == 10 and n >= 3:
            C.fredkin(circ, q[0], q[1], q[2])
        elif r == 11:
            C.add_gate(circ, "rydberg-cphase", control=q[0], target=q[1], angle=a)
        elif r == 12 and n >= 3:
            C.add_gate(circ, "rydberg-blockade", qubit_indices=q[:int(rng.integers(2, 4))], angle=a)
        elif r == 13:
            C.add_gate(circ, ["global-h", "global-rx", "global-rz", "global-x"][rng.integers(0, 4)], angle=a)
        elif r == 14:
            for _k in range(int(rng.integers(0, 3))):
                C.add_gate(circ, "phase-oracle", index=int(rng.integers(0, 1 << n)))
            C.add_gate(circ, "grover-diffusion")
    return circ


def _reference(circ, init):
    """Oracle semantics + the operator-level ops the oracle does not know."""
    st = init.copy()
    n = int(math.log2(st.shape[0]))
    idx = np.arange(st.shape[0])
    for op in circ["operations"]:
        t = op["operation-type"]
        if t == "phase-oracle":
            st[op["operation-params"]["index"]] *= -1
        elif t == "mcphase":
            sel = np.ones(st.shape[0], dtype=bool)
            for q in op["operation-params"]["qubit-indices"]:
                sel &= ((idx >> (n - 1 - q)) & 1) == 1
            st = np.where(sel, st * np.exp(1j * op["operation-params"]["angle"]), st)
        elif t == "grover-diffusion":
            st = 2 * np.mean(st) - st
        else:
            st = O.apply_gate_to_state(st, op)
    return st


def _case(seed, grover=False):
    rng = np.random.default_rng(seed)
    world = int(rng.choice([1, 1, 2, 4, 8]))
    p = world.bit_length() - 1
    n = int(rng.integers(max(3, p + 3), 12))
    nl = n - p
    tile = int(rng.integers(min(3, nl), min(nl, 9) + 1))
    low = int(rng.integers(1, max(2, min(tile, 4)) + 1))
    circ = random_circuit(n, rng, int(rng.integers(5, 70)), grover)
    init = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    init /= np.linalg.norm(init)
    return n, world, tile, low, circ, init, int(rng.choice([1, 2]))


@pytest.mark.parametrize("block", range(6))
def test_random_circuits_random_geometry_emulated_ranks(block):
    for seed in range(40 * block, 40 * block + 40):
        n, world, tile, low, circ, init, mma = _case(seed, grover=(seed % 3 == 0))
        got = E.run_world(n, circ["operations"], init, world=world, tile_bits=tile, low_bits=low, dense_mma=mma)
        err = float(np.max(np.abs(got - _reference(circ, init))))
        assert err <= TOL, f"seed {seed}: n={n} world={world} tile={tile} low={low} err={err}"


def test_wide_diagonal_gate_followed_by_gates_on_its_bits():
    """A diagonal gate on more than MAX_SLOT_BITS tile-local bits (phase oracle, multi-controlled phase) followed by a gate
    that targets one of its bits used to be deferred for ever ("scheduler made no progress")."""
    for n in range(4, 13):
        circ = C.create_circuit(n)
        C.add_gate(circ, "phase-oracle", index=(1 << n) - 2)
        C.h(circ, n - 1)
        st = E.run_world(n, circ["operations"])
        init = np.zeros(1 << n, dtype=complex); init[0] = 1
        assert np.max(np.abs(st - _reference(circ, init))) <= TOL
    # textbook Grover iteration with explicit gates: H^n, oracle, H^n X^n, multi-controlled Z, X^n H^n
    for n, world in ((5, 1), (9, 1), (11, 2), (12, 4)):
        circ = C.create_circuit(n)
        for q in range(n): C.h(circ, q)
        for _it in range(2):
            C.add_gate(circ, "phase-oracle", index=5)
            for q in range(n): C.h(circ, q)
            for q in range(n): C.x(circ, q)
            C.add_gate(circ, "mcphase", qubit_indices=list(range(n)), angle=math.pi)
            for q in range(n): C.x(circ, q)
            for q in range(n): C.h(circ, q)
        init = np.zeros(1 << n, dtype=complex); init[0] = 1
        st = E.run_world(n, circ["operations"], world=world)
        assert np.max(np.abs(st - _reference(circ, init))) <= TOL
    # random: wide mcphase / oracles sprinkled between ordinary gates, random geometry
    for seed in range(60):
        rng = np.random.default_rng(7000 + seed)
        n = int(rng.integers(5, 12))
        world = int(rng.choice([1, 1, 2, 4]))
        circ = C.create_circuit(n)
        for _ in range(int(rng.integers(6, 30))):
            r = rng.random()
            if r < 0.25:
                k = int(rng.integers(4, n + 1))
                C.add_gate(circ, "mcphase", qubit_indices=[int(x) for x in rng.permutation(n)[:k]], angle=float(rng.uniform(0, 6.28)))
            elif r < 0.4:
                C.add_gate(circ, "phase-oracle", index=int(rng.integers(0, 1 << n)))
            elif r < 0.7:
                C.add_gate(circ, ["h", "x", "t"][rng.integers(0, 3)], target=int(rng.integers(0, n)))
            else:
                a, b = [int(x) for x in rng.permutation(n)[:2]]
                C.cnot(circ, a, b)
        init = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        init /= np.linalg.norm(init)
        nl = n - (world.bit_length() - 1)
        tile = int(rng.integers(min(4, nl), min(nl, 9) + 1))
        got = E.run_world(n, circ["operations"], init, world=world, tile_bits=tile)
        err = float(np.max(np.abs(got - _reference(circ, init))))
        assert err <= TOL, f"seed {seed}: n={n} world={world} tile={tile} err={err}"


def test_degenerate_tile_geometry_is_sanitised():
    # tile_bits == low_bits < n_local used to leave no room for a gate's targets (scheduler never finished)
    n = 9
    circ = C.create_circuit(n)
    C.swap(circ, 0, 1); C.iswap(circ, 2, 8); C.fredkin(circ, 0, 3, 4); C.h(circ, 0); C.cnot(circ, 8, 0)
    for kw in ({"tile_bits": 3, "low_bits": 3}, {"tile_bits": 5, "low_bits": 4, "world": 8}, {"tile_bits": 1, "low_bits": 1}):
        got = E.run_world(n, circ["operations"], **kw)
        assert np.max(np.abs(got - O.execute_circuit(circ))) <= TOL
    # slices of 3 local qubits: exchange partners fall back to any local bit
    circ = C.random_brickwork_circuit(6, 6, seed=2)
    got = E.run_world(6, circ["operations"], world=8)
    assert np.max(np.abs(got - O.execute_circuit(circ))) <= TOL


@pytest.mark.parametrize("block", range(3))
def test_trace_replay_equals_fresh_schedule_on_random_circuits(block):
    for seed in range(1000 + 25 * block, 1000 + 25 * block + 25):
        n, world, tile, low, circ, _init, mma = _case(seed, grover=(seed % 2 == 0))
        ops = circ["operations"]
        new = _reangle(ops, seed)
        for rank in {0, world - 1}:
            kw = {"tile_bits": tile, "low_bits": low, "rank": rank, "world_size": world, "dense_mma": mma}
            assert np.array_equal(_replayed(n, ops, new, **kw), _fresh(n, new, **kw)), f"seed {seed} rank {rank}"


def test_empty_and_single_gate_programs():
    from qclojure_b200 import _lib as L
    assert L.plan_summary(5, []) == {"stages": 0, "rounds": 0, "exchanges": 0, "program_words": 4, "tile_sweeps": 0, "passes": 0,
                                     "paired_passes": 0, "per_sweep": []}
    st = E.run_world(5, [])
    assert st[0] == 1.0 and np.count_nonzero(st) == 1
    for n in (1, 2, 13):
        circ = C.h(C.create_circuit(n), n - 1)
        assert np.max(np.abs(E.run_world(n, circ["operations"]) - O.execute_circuit(circ))) <= TOL


@pytest.mark.parametrize("seed", range(8))
def test_wire_formats_round_trip_random_circuits(seed):
    """EDN / JSON / OpenQASM 3 carry a random circuit (every gate kind the formats know) without loss: the re-imported
    circuit encodes to the same qcb_op array."""
    from qclojure_b200 import io as QIO
    from qclojure_b200 import ops as OPS
    from qclojure_b200 import qasm3 as Q
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(3, 9))
    circ = random_circuit(n, rng, 60)
    a, na, _k = OPS.encode_ops(circ["operations"])
    for fmt in ("edn", "json"):
        back = QIO.deserialize_quantum_circuit(QIO._load(fmt, QIO._dump(fmt, QIO.serialize_quantum_circuit(circ))))
        b, nb, _k2 = OPS.encode_ops(back["operations"])
        assert na == nb and bytes(a) == bytes(b)
    # QASM 3 has no spelling for the neutral-atom gates (they are emitted as decompositions / comments): leave them out
    plain = C.create_circuit(n)
    plain["operations"] = [o for o in circ["operations"]
                           if not o["operation-type"].startswith(("global-", "rydberg-"))]
    a, na, _k = OPS.encode_ops(plain["operations"])
    back = Q.qasm_to_circuit(Q.circuit_to_qasm(plain))
    b, nb, _k2 = OPS.encode_ops(back["operations"])
    assert na == nb and bytes(a) == bytes(b)


def test_known_zero_state_support_skips_tiles_and_keeps_results():
    """EXPERIMENTAL path (csrc/plan.h: Plan::support_in, off unless QCB_ZERO_SKIP=1): a plan that is told its input is |0...0>
    visits only the tiles that can hold non-zero amplitudes - same results as the oracle on random circuits over the whole
    vocabulary (Grover operators, fusion off, sharded worlds), the first sweeps shrink, and the claim matters (a dense state
    run under it comes out wrong: the tiles really are skipped)."""
    for seed in range(9000, 9120):
        rng = np.random.default_rng(seed)
        world = int(rng.choice([1, 1, 2, 4]))
        p = world.bit_length() - 1
        n = int(rng.integers(max(4, p + 3), 13))
        nl = n - p
        tile = int(rng.integers(min(3, nl), min(nl, 9) + 1))
        low = int(rng.integers(1, max(2, min(tile, 4)) + 1))
        circ = random_circuit(n, rng, int(rng.integers(5, 70)), grover=(seed % 3 == 0))
        init = np.zeros(1 << n, dtype=complex); init[0] = 1.0
        got = E.run_world(n, circ["operations"], init, world=world, tile_bits=tile, low_bits=low, fusion=int(seed % 5 != 0), support=0)
        err = float(np.max(np.abs(got - _reference(circ, init))))
        assert err <= TOL, f"seed {seed}: n={n} world={world} tile={tile} low={low} err={err}"
    n = 14
    circ = C.random_brickwork_circuit(n, 8)
    known = E.EmuPlan(n, circ["operations"], tile_bits=6, low_bits=2, support=0)
    plain = E.EmuPlan(n, circ["operations"], tile_bits=6, low_bits=2)
    fk = [known.stage_info(i)["fraction"] for i in range(known.num_stages)]
    fp = [plain.stage_info(i)["fraction"] for i in range(plain.num_stages)]
    assert fk[0] == 2.0 ** -(n - 6) and fk[0] < fk[1] < fk[2] <= 1.0 and fp[0] == 1.0
    assert all(a <= b for a, b in zip(fk, fp))
    rng = np.random.default_rng(1)
    dense = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    dense /= np.linalg.norm(dense)
    wrong = E.run_world(n, circ["operations"], dense, tile_bits=6, low_bits=2, support=0)
    assert np.max(np.abs(wrong - _reference(circ, dense))) > 1e-3
