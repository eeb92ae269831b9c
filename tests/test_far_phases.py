"""Far phases (csrc/plan.cpp: classify_far / build_far_table; csrc/tile_core.h "far phases"): a diagonal two-bit gate (CRZ, CZ,
controlled phase) with one operand on a slot of the round and the other OUTSIDE the tile costs no condition bit - for a given
tile the far bit is a constant, so the product of all such gates of a round is a row scaling of the round's 8x8 block that
every lane applies to its A fragments.  This is what lets one round absorb the tail of a QFT ladder
(application/algorithm/quantum_fourier_transform.clj:34-62).  CPU checks through the host emulator against the oracle."""
import ctypes as CT

import numpy as np
import pytest

from oracle import qc_oracle as O
from qclojure_b200 import circuits as C
from tests.emu import emu as E
from tests.test_plan_trace import _fresh, _replayed, _reangle

TOL = 1e-10


def _far_entries(plan):
    """Total number of far-table entries over all passes of a plan (RoundDesc word [39])."""
    nw = E.lib().emu_program_words(plan.h, None, 0)
    buf = (CT.c_uint64 * nw)()
    E.lib().emu_program_words(plan.h, buf, nw)
    w = np.frombuffer(buf, dtype=np.uint64)
    pos, total = 4, 0
    for _ in range(int(w[1])):
        kind = int(w[pos]); pos += 2
        if kind == 3:
            pos += int(w[pos - 1]) & 0xff
        if kind != 0:
            continue
        st = w[pos:]
        for r in range(int(st[3])):
            v = int(st[48 + 40 * r + 39])
            total += sum((v >> (8 * t)) & 0xff for t in range(4))
        pos += int(st[40])
    return total


def _rand_state(n, seed):
    rng = np.random.default_rng(seed)
    s = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return s / np.linalg.norm(s)


@pytest.mark.parametrize("n,kw", [(13, dict(tile_bits=8, low_bits=3)), (14, dict(tile_bits=10)), (16, dict()), (15, dict(tile_bits=10, world=4)),
                                  (14, dict(tile_bits=10, world=2))])
def test_qft_with_far_phases_matches_oracle(n, kw):
    circ = C.quantum_fourier_transform_circuit(n)
    init = _rand_state(n, n)
    want = O.execute_circuit(circ, init)
    got, plans = E.run_world(n, circ["operations"], init, return_plans=True, **kw)
    assert np.max(np.abs(got - want)) <= TOL
    assert _far_entries(plans[0]) > 0


@pytest.mark.parametrize("seed", range(6))
def test_far_phases_random_diagonal_mixes(seed):
    """CRZ in both orientations, CZ, rydberg-cphase and RZ / phase gates between far and near qubits, interleaved with gates
    that act non-diagonally on the same slots (which must end the far phase's round on that slot)."""
    rng = np.random.default_rng(100 + seed)
    n = 14 + seed % 2
    circ = C.create_circuit(n)
    for _ in range(140):
        a, b = (int(x) for x in rng.choice(n, 2, replace=False))
        k = int(rng.integers(0, 8))
        if k == 0: C.add_gate(circ, "h", target=a)
        elif k == 1: C.rx(circ, a, rng.random() * 6)
        elif k == 2: C.crz(circ, a, b, rng.random() * 6)
        elif k == 3: C.crz(circ, b, a, rng.random() * 6)
        elif k == 4: C.cz(circ, a, b)
        elif k == 5: C.rz(circ, a, rng.random() * 6)
        elif k == 6: C.cnot(circ, a, b)
        else: C.add_gate(circ, "t", target=a)
    init = _rand_state(n, seed)
    want = O.execute_circuit(circ, init)
    got, plans = E.run_world(n, circ["operations"], init, return_plans=True, tile_bits=9)
    assert np.max(np.abs(got - want)) <= TOL
    assert _far_entries(plans[0]) > 0


def test_far_phases_cut_the_qft_plan_and_can_be_switched_off(monkeypatch):
    from qclojure_b200 import _lib as L
    ops = C.quantum_fourier_transform_circuit(26)["operations"]
    on = L.plan_summary(26, ops)
    monkeypatch.setenv("QCB_FAR_PHASE", "0")
    off = L.plan_summary(26, ops)
    assert on["passes"] <= 0.7 * off["passes"] and on["rounds"] < off["rounds"]
    monkeypatch.delenv("QCB_FAR_PHASE")
    circ = C.quantum_fourier_transform_circuit(14)
    a = E.run_world(14, circ["operations"], tile_bits=9)
    monkeypatch.setenv("QCB_FAR_PHASE", "0")
    b = E.run_world(14, circ["operations"], tile_bits=9)
    assert np.max(np.abs(a - b)) <= 1e-13


def test_trace_replay_reproduces_far_phases():
    ops = C.quantum_fourier_transform_circuit(18)["operations"]
    ops2 = _reangle(ops, 3)
    a = _replayed(18, ops, ops2)
    b = _fresh(18, ops2)
    assert a.shape == b.shape and np.array_equal(a, b)
