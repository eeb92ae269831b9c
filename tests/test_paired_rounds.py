"""Paired rounds (round kind 3; csrc/tile_core.h "paired rounds", csrc/plan.cpp: build_k3_pair_round): two dense 8x8 complex
blocks on disjoint slot triples in one pass over the shared tile, the second block fed from the first block's D registers.
CPU checks through the host emulator, which models the mma.m8n8k4 fragment layouts lane by lane (tests/emu/emu.cpp:
emu_k3x_round), against the oracle; the GPU parity tests exercise the same plans on the device."""
import ctypes as CT

import numpy as np
import pytest

from oracle import qc_oracle as O
from qclojure_b200 import circuits as C
from tests.emu import emu as E
from tests.test_oracle_c import _all_gates_circuit
from tests.test_plan_trace import _fresh, _replayed, _reangle

TOL = 1e-10


def _round_kinds(plan):
    """[(stage, [kind of every pass])] from the serialised program (csrc/plan.h: RoundDesc word [17])."""
    nw = E.lib().emu_program_words(plan.h, None, 0)
    buf = (CT.c_uint64 * nw)()
    E.lib().emu_program_words(plan.h, buf, nw)
    w = np.frombuffer(buf, dtype=np.uint64)
    pos, out = 4, []
    for s in range(int(w[1])):
        kind = int(w[pos]); pos += 2
        if kind == 3:                                   # S_GROVER: marked indices follow
            pos += int(w[pos - 1]) & 0xff
        if kind != 0:
            continue
        st = w[pos:]
        out.append((s, [int(st[48 + 40 * r + 17]) for r in range(int(st[3]))], st))
        pos += int(st[40])
    return out


def _rand_state(n, seed):
    rng = np.random.default_rng(seed)
    s = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return s / np.linalg.norm(s)


def test_benchmark_plan_pairs_rounds_and_stays_conflict_free():
    circ = C.random_brickwork_circuit(30, 20)
    p = E.EmuPlan(30, circ["operations"])
    kinds = _round_kinds(p)
    passes = sum(len(k) for _, k, _ in kinds)
    pairs = sum(k.count(3) for _, k, _ in kinds)
    assert pairs >= 15                                   # a third of the passes or more carry two rounds
    assert p.num_rounds == passes + pairs                # n_rounds counts dense blocks: a paired pass counts two
    assert max(p.max_conflict(s) for s, _, _ in kinds) == 1
    for _, ks, st in kinds:
        m = int(st[1])
        for r, kd in enumerate(ks):
            if kd != 3:
                continue
            rd = st[48 + 40 * r: 48 + 40 * (r + 1)]
            s1 = {int(rd[4 + j]) for j in range(3)}
            s2 = {int(rd[19 + j]) for j in range(3)}
            cond = {int(rd[30 + j]) for j in range(int(rd[29]))}
            assert len(s1) == 3 and len(s2) == 3 and not (s1 & s2) and not (cond & (s1 | s2))
            assert int(rd[18]) == m - 3 and len(cond) <= 4


@pytest.mark.parametrize("n,tile,low", [(10, 10, 4), (12, 12, 4), (13, 10, 4), (14, 12, 4), (15, 13, 4), (16, 11, 3)])
def test_paired_rounds_match_oracle_all_gates(n, tile, low):
    """Every gate kind of the vocabulary (controls / diagonal operands become condition bits of either block)."""
    rng = np.random.default_rng(100 + n)
    circ = _all_gates_circuit(n, rng)
    init = _rand_state(n, n)
    want = O.execute_circuit(circ, init)
    got, plans = E.run_world(n, circ["operations"], init, tile_bits=tile, low_bits=low, return_plans=True)
    assert np.max(np.abs(got - want)) <= TOL
    assert any(3 in k for _, k, _ in _round_kinds(plans[0]))


@pytest.mark.parametrize("seed", range(6))
def test_paired_rounds_match_oracle_random_circuits(seed):
    """Random mixes of dense 1q / 2q gates, controlled rotations and diagonal gates at 13-14 qubits, 4 warps per group as in
    the kernel's default layout (nthreads = 128) and 16."""
    rng = np.random.default_rng(seed)
    n = 13 + seed % 2
    circ = C.create_circuit(n)
    for _ in range(160):
        a, b, c = (int(x) for x in rng.choice(n, 3, replace=False))
        k = int(rng.integers(0, 8))
        if k == 0: C.add_gate(circ, "h", target=a)
        elif k == 1: C.rx(circ, a, rng.random() * 6)
        elif k == 2: C.ry(circ, a, rng.random() * 6)
        elif k == 3: C.cnot(circ, a, b)
        elif k == 4: C.crz(circ, a, b, rng.random() * 6)
        elif k == 5: C.cz(circ, a, b)
        elif k == 6: C.swap(circ, a, b)
        else: C.toffoli(circ, a, b, c)
    init = _rand_state(n, seed)
    want = O.execute_circuit(circ, init)
    for nthreads in (128, 512):
        got, plans = E.run_world(n, circ["operations"], init, return_plans=True, nthreads=nthreads)
        assert np.max(np.abs(got - want)) <= TOL
    assert any(3 in k for _, k, _ in _round_kinds(plans[0]))


def test_pairing_off_gives_the_same_state(monkeypatch):
    circ = C.random_brickwork_circuit(14, 10)
    init = _rand_state(14, 3)
    on, plans_on = E.run_world(14, circ["operations"], init, return_plans=True)
    monkeypatch.setenv("QCB_PAIR_ROUNDS", "0")
    off, plans_off = E.run_world(14, circ["operations"], init, return_plans=True)
    assert np.max(np.abs(on - off)) <= 1e-13
    assert any(3 in k for _, k, _ in _round_kinds(plans_on[0]))
    assert not any(3 in k for _, k, _ in _round_kinds(plans_off[0]))


def test_sharded_plan_with_paired_rounds_matches_oracle():
    circ = C.random_brickwork_circuit(15, 8)
    want = O.execute_circuit(circ)
    got, plans = E.run_world(15, circ["operations"], world=4, tile_bits=10, return_plans=True)
    assert np.max(np.abs(got - want)) <= TOL
    assert any(3 in k for pl in plans for _, k, _ in _round_kinds(pl))


def test_trace_replay_reproduces_paired_rounds():
    """A replayed plan (variational loop) must equal the freshly scheduled one word for word, pairs included."""
    ops = C.random_brickwork_circuit(16, 10)["operations"]
    ops2 = _reangle(ops, 5)
    a = _replayed(16, ops, ops2)
    b = _fresh(16, ops2)
    assert a.shape == b.shape and np.array_equal(a, b)


def _cost(ps):
    """csrc/plan.cpp: plan_cost_ms for single-GPU plans of tensor-core rounds only (every sweep at least its HBM time)."""
    return sum(max(1.52 + 2.0 * (nr - npair) + 3.39 * npair, 5.8) for nr, npair in ps["per_sweep"])


def test_plan_portfolio_picks_the_cheapest_member_and_replays_word_for_word(monkeypatch):
    """Large circuits (>= 24 local qubits, >= 128 gates) are scheduled under five budget settings (plus eight variants of the
    stage-yield threshold, test below); the plan the cost model prefers is the one that is built (csrc/plan.cpp: schedule).  It can never be worse than any member, a pinned knob switches
    the portfolio off, and a replayed plan of the same structure equals the fresh one word for word."""
    from qclojure_b200 import _lib as L
    n = 24
    ops = C.random_brickwork_circuit(n, 20)["operations"]
    chosen = L.plan_summary(n, ops)
    members = []
    for rounds, cost_q, eff, search, pairs in [(7, 7, 170, 1, 1), (7, 7, 150, 4, 1), (6, 6, 170, 1, 1), (8, 7, 160, 4, 1), (5, 6, 170, 1, 0)]:
        for k, v in (("QCB_PAIR_COST_Q", cost_q), ("QCB_PAIR_EFF_PCT", eff), ("QCB_PAIR_SEARCH", search), ("QCB_PAIR_ROUNDS", pairs)):
            monkeypatch.setenv(k, str(v))
        members.append(_cost(L.plan_summary(n, ops, max_stage_rounds=rounds)))
    for k in ("QCB_PAIR_COST_Q", "QCB_PAIR_EFF_PCT", "QCB_PAIR_SEARCH", "QCB_PAIR_ROUNDS"):
        monkeypatch.delenv(k)
    assert abs(_cost(chosen) - min(members)) < 1e-9
    assert len(set(round(m, 6) for m in members)) > 1            # the settings really differ on this circuit
    ops2 = _reangle(ops, 9)
    a = _replayed(n, ops, ops2)
    b = _fresh(n, ops2)
    assert a.shape == b.shape and np.array_equal(a, b)
    monkeypatch.setenv("QCB_PLAN_PORTFOLIO", "0")
    fixed = L.plan_summary(n, ops)
    assert abs(_cost(fixed) - members[0]) < 1e-9                  # portfolio off = the first (default) setting


_BASE_KNOBS = [(7, 7, 170, 1, 1), (7, 7, 150, 4, 1), (6, 6, 170, 1, 1), (8, 7, 160, 4, 1), (5, 6, 170, 1, 0)]


def _member_costs(monkeypatch, n, ops, yield_pct, knobs, **kw):
    from qclojure_b200 import _lib as L
    out = []
    monkeypatch.setenv("QCB_ROUND_YIELD_PCT", str(yield_pct))
    for rounds, cost_q, eff, search, pairs in knobs:
        for k, v in (("QCB_PAIR_COST_Q", cost_q), ("QCB_PAIR_EFF_PCT", eff), ("QCB_PAIR_SEARCH", search), ("QCB_PAIR_ROUNDS", pairs)):
            monkeypatch.setenv(k, str(v))
        out.append(_cost(L.plan_summary(n, ops, max_stage_rounds=rounds, **kw)))
    for k in ("QCB_PAIR_COST_Q", "QCB_PAIR_EFF_PCT", "QCB_PAIR_SEARCH", "QCB_PAIR_ROUNDS", "QCB_ROUND_YIELD_PCT"):
        monkeypatch.delenv(k)
    return out


@pytest.mark.parametrize("n", [29, 30, 32])
def test_plan_portfolio_yield_variants(monkeypatch, n):
    """The portfolio also schedules under stage-yield thresholds of 35 % and 75 % (the four pairing settings each); a variant
    is built only when the cost model puts it at least 2 % below the best of the five base settings.  On the benchmark
    circuits that happens at 29 and 32 qubits and not at 30 (the headline plan is the one every hardware number was taken with)."""
    from qclojure_b200 import _lib as L
    ops = C.random_brickwork_circuit(n, 20)["operations"]
    chosen = _cost(L.plan_summary(n, ops))
    base = min(_member_costs(monkeypatch, n, ops, 50, _BASE_KNOBS))
    var = min(min(_member_costs(monkeypatch, n, ops, y, _BASE_KNOBS[:4])) for y in (35, 75))
    expect = var if var < 0.98 * base else base
    assert abs(chosen - expect) < 1e-9
    assert (n == 30) == (abs(chosen - base) < 1e-9)
    ops2 = _reangle(ops, 3)
    assert np.array_equal(_replayed(n, ops, ops2), _fresh(n, ops2))
