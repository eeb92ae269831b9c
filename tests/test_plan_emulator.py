"""CPU checks of the host side of the gate executor: lowering, fusion scheduler, program encoding
and the op interpreters (shared header csrc/tile_core.h), run through the host emulator against the
oracle.  The emulator is test infrastructure (tests/emu/); the GPU parity tests are in test_gpu_*.py."""
import math

import numpy as np
import pytest

from oracle import qc_oracle as O
from qclojure_b200 import circuits as C
from tests.emu import emu as E
from tests.test_oracle_c import _all_gates_circuit

TOL = 1e-10


def _rand_state(n, seed):
    rng = np.random.default_rng(seed)
    s = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    return s / np.linalg.norm(s)


@pytest.mark.parametrize("n,tile,low,fusion", [
    (1, 0, 0, 1), (2, 0, 0, 1), (3, 0, 0, 1), (5, 0, 0, 1), (6, 4, 2, 1), (8, 5, 2, 1), (9, 6, 3, 0),
    (10, 7, 4, 1), (11, 8, 3, 1), (13, 10, 4, 1), (14, 12, 4, 1), (12, 12, 4, 0), (15, 13, 6, 1), (14, 13, 4, 0),
    (9, 6, 2, 1), (10, 9, 3, 1)])
@pytest.mark.parametrize("dense_mma", [1, 2])
def test_emulated_executor_matches_oracle_all_gates(n, tile, low, fusion, dense_mma):
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
    got = E.run_world(n, circ["operations"], init, tile_bits=tile, low_bits=low, fusion=fusion, dense_mma=dense_mma)
    assert np.max(np.abs(got - want)) <= TOL


@pytest.mark.parametrize("builder", ["qft", "ghz", "brick"])
def test_emulated_configs(builder):
    n = 12
    circ = {"qft": C.quantum_fourier_transform_circuit, "ghz": C.ghz_state_circuit,
            "brick": lambda k: C.random_brickwork_circuit(k, 8)}[builder](n)
    want = O.execute_circuit(circ)
    got, plans = E.run_world(n, circ["operations"], tile_bits=8, low_bits=3, return_plans=True)
    assert np.max(np.abs(got - want)) <= TOL
    # fusion must actually fuse: far fewer sweeps than gates
    assert plans[0].num_stages < plans[0].num_gates


def test_unfused_mode_is_one_sweep_per_gate_with_partial_sweeps():
    n = 12
    circ = C.create_circuit(n)
    C.h(circ, 0); C.cnot(circ, 0, 1); C.cz(circ, 2, 3); C.z(circ, 1); C.toffoli(circ, 0, 1, 2); C.rz(circ, 5, 0.2)
    got, plans = E.run_world(n, circ["operations"], tile_bits=6, low_bits=3, fusion=0, return_plans=True)
    assert np.max(np.abs(got - O.execute_circuit(circ))) <= TOL
    p = plans[0]
    assert p.num_stages == 6
    fr = [p.stage_info(i)["fraction"] for i in range(6)]
    # H: full sweep; CNOT: control-half only; CZ: quarter; Z: half; Toffoli: quarter; RZ: full
    assert fr == [1.0, 0.5, 0.25, 0.5, 0.25, 1.0]
    assert p.algorithmic_bytes() == pytest.approx(32 * 2 ** n * sum(fr))
    assert p.unfused_bytes() == pytest.approx(32 * 2 ** n * sum(fr))


def test_strict_parity_flag():
    n = 4
    circ = C.create_circuit(n)
    C.h(circ, 0); C.h(circ, 2); C.cry(circ, 0, 1, 0.7); C.swap(circ, 0, 1); C.iswap(circ, 0, 2)
    strict = E.run_world(n, circ["operations"], strict_parity=1)
    assert np.max(np.abs(strict - O.execute_circuit(circ))) <= TOL
    # textbook semantics: CRY applies RY(theta) (not the transpose), swap operands are qubit numbers
    st = O.zero_state(n)
    st = O.apply_single_qubit_gate(st, O.HADAMARD, 0)
    st = O.apply_single_qubit_gate(st, O.HADAMARD, 2)
    st = O.apply_controlled_gate(st, 0, 1, O.ry_gate(0.7).T)      # oracle applies U^T, so pass U^T to get U
    st = O.swap_gate(st, n - 1 - 0, n - 1 - 1)
    st = O.iswap_gate(st, n - 1 - 0, n - 1 - 2)
    text = E.run_world(n, circ["operations"], strict_parity=0)
    assert np.max(np.abs(text - st)) <= TOL
    # :i and :cy: rejected like the reference under strict parity, accepted otherwise
    bad = C.add_gate(C.create_circuit(2), "cy", control=0, target=1)
    with pytest.raises(E.EmuError) as ei:
        E.EmuPlan(2, bad["operations"], strict_parity=1)
    assert ei.value.code == -2
    ok = E.run_world(2, C.h(C.create_circuit(2), 0)["operations"] + bad["operations"], strict_parity=0)
    want = O.apply_controlled_gate(O.apply_single_qubit_gate(O.zero_state(2), O.HADAMARD, 0), 0, 1, O.PAULI_Y.T)
    assert np.max(np.abs(ok - want)) <= TOL


def test_generic_ops_u1q_cu1q_u2q_mcphase_oracle_diffusion():
    n = 6
    rng = np.random.default_rng(3)
    init = _rand_state(n, 11)

    def runitary(d):
        a = rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
        q, _ = np.linalg.qr(a)
        return q

    U1, U2, CU = runitary(2), runitary(4), runitary(2)
    ops = [{"operation-type": "u1q", "operation-params": {"target": 4, "matrix": U1}},
           {"operation-type": "cu1q", "operation-params": {"control": 1, "target": 3, "matrix": CU}},
           {"operation-type": "u2q", "operation-params": {"qubit1": 5, "qubit2": 2, "matrix": U2}},
           {"operation-type": "u2q", "operation-params": {"qubit1": 0, "qubit2": 4, "matrix": U2}},
           {"operation-type": "mcphase", "operation-params": {"qubit-indices": [0, 2, 5], "angle": 0.9}},
           {"operation-type": "phase-oracle", "operation-params": {"index": 37}},
           {"operation-type": "grover-diffusion", "operation-params": {}},
           {"operation-type": "h", "operation-params": {"target": 1}},
           {"operation-type": "grover-diffusion", "operation-params": {}}]
    # reference model with dense matrices
    def kron_on(mats):  # mats: dict qubit->2x2
        full = np.array([[1.0 + 0j]])
        for q in range(n):
            full = np.kron(full, mats.get(q, np.eye(2)))
        return full
    st = init.copy()
    st = kron_on({4: U1}) @ st
    st = O.apply_controlled_gate(st, 1, 3, CU.T)
    def apply_u2(st, qa, qb, U):
        idx = np.arange(1 << n)
        out = np.zeros_like(st)
        ba, bb = n - 1 - qa, n - 1 - qb
        for i in idx:
            col = (((i >> ba) & 1) << 1) | ((i >> bb) & 1)
            base = i & ~((1 << ba) | (1 << bb))
            for row in range(4):
                j = base | (((row >> 1) & 1) << ba) | ((row & 1) << bb)
                out[j] += U[row, col] * st[i]
        return out
    st = apply_u2(st, 5, 2, U2)
    st = apply_u2(st, 0, 4, U2)
    idx = np.arange(1 << n)
    allone = np.ones_like(idx, dtype=bool)
    for q in (0, 2, 5):
        allone &= ((idx >> (n - 1 - q)) & 1) == 1
    st = np.where(allone, st * np.exp(1j * 0.9), st)
    st[37] *= -1
    st = 2 * np.mean(st) - st
    st = O.apply_single_qubit_gate(st, O.HADAMARD, 1)
    st = 2 * np.mean(st) - st
    got = E.run_world(n, ops, init, tile_bits=4, low_bits=2)
    assert np.max(np.abs(got - st)) <= TOL


def test_grover_operator_matches_gate_level_circuit():
    """fused oracle+diffusion operators validated against the reference's
    gate-level Grover circuit (tests/golden/tutorial_cases.json, doc/tutorial.md:5832; 3 qubits, target 5)."""
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "tutorial_cases.json")) as f:
        case = [c for c in json.load(f)["cases"] if c["name"] == "Grover Search"][0]
    n = 3
    ops = [{"operation-type": "global-h", "operation-params": {}}]
    for _ in range(2):
        ops += [{"operation-type": "phase-oracle", "operation-params": {"index": 5}},
                {"operation-type": "grover-diffusion", "operation-params": {}}]
    got = E.run_world(n, ops)
    want_p = np.array(case["measurement_results"][":measurement-probabilities"])
    assert np.max(np.abs(np.abs(got) ** 2 - want_p)) <= TOL


def _grover_reference(n, iterations, marked, extra=None):
    st = np.full(1 << n, 1.0 / math.sqrt(1 << n), dtype=np.complex128)
    for it in range(iterations):
        for mk in marked:
            st[mk] *= -1
        st = 2 * np.mean(st) - st
        if extra is not None and it == extra[0]:
            st = O.apply_single_qubit_gate(st, O.HADAMARD, extra[1])
    return st


@pytest.mark.parametrize("n,world,marked", [(8, 1, [0xA5]), (10, 1, [3, 700, 1023]), (10, 4, [3, 700, 1023]), (12, 8, [0xABC]),
                                            (9, 2, [])])
def test_fused_grover_pass(n, world, marked):
    """Diffusion + following phase oracles (+ the sum for the next diffusion) run as ONE streaming stage (plan.h: S_GROVER,
    kernels.cu: k_grover_step): one S_SUM for the whole loop, one pass per iteration; marked states land on the rank that
    holds them."""
    its = 5
    ops = [{"operation-type": "global-h", "operation-params": {}}]
    for _ in range(its):
        ops += [{"operation-type": "phase-oracle", "operation-params": {"index": mk}} for mk in marked]
        ops += [{"operation-type": "grover-diffusion", "operation-params": {}}]
    got, plans = E.run_world(n, ops, world=world, tile_bits=min(6, n - 3), low_bits=2, return_plans=True)
    assert np.max(np.abs(got - _grover_reference(n, its, marked))) <= TOL
    kinds = [plans[0].stage_kind(i) for i in range(plans[0].num_stages)]
    assert kinds.count(E.S_SUM) == 1 and kinds.count(E.S_GROVER) == its
    if marked:
        # every marked state is flipped by exactly one rank
        for i in range(plans[0].num_stages):
            if kinds[i] == E.S_GROVER and i != len(kinds) - 1:
                assert sum(len(p.stage_grover(i)[0]) for p in plans) == len(marked)


def test_fused_grover_pass_mixed_with_gates_and_permuted_layout():
    """A diffusion followed by ordinary gates keeps the tile path (affine lead round) and reuses the sum a fused pass left
    on the device; after a qubit exchange the marked index is translated to the physical layout."""
    n, world = 10, 4
    marked = [0x155, 0x2AA]
    ops = [{"operation-type": "global-h", "operation-params": {}},
           {"operation-type": "h", "operation-params": {"target": 0}},          # non-diagonal on a global qubit: exchange
           {"operation-type": "h", "operation-params": {"target": 0}}]
    for it in range(4):
        ops += [{"operation-type": "phase-oracle", "operation-params": {"index": mk}} for mk in marked]
        ops += [{"operation-type": "grover-diffusion", "operation-params": {}}]
        if it == 1:
            ops += [{"operation-type": "h", "operation-params": {"target": 7}}]
    got, plans = E.run_world(n, ops, world=world, tile_bits=5, low_bits=2, return_plans=True)
    assert np.max(np.abs(got - _grover_reference(n, 4, marked, extra=(1, 7)))) <= TOL
    kinds = [plans[0].stage_kind(i) for i in range(plans[0].num_stages)]
    assert E.S_EXCHANGE in kinds and kinds.count(E.S_GROVER) >= 2
    # unfused mode never forms the fused pass
    got = E.run_world(n, ops, world=1, fusion=0)
    assert np.max(np.abs(got - _grover_reference(n, 4, marked, extra=(1, 7)))) <= TOL


def test_hhl_tutorial_probabilities_through_the_executor():
    """The JVM-recorded HHL run (tests/golden/hhl_tutorial.json, doc/tutorial.md HHL section) through lowering + scheduler +
    interpreters: matches under strict_parity = 1 (the reference's transposed controlled gate), differs by 0.1 under
    textbook semantics."""
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "hhl_tutorial.json")) as f:
        g = json.load(f)
    circ = C.hhl_circuit(g["matrix"], g["vector"], g["precision_qubits"], g["ancilla_qubits"])
    want = np.array(g["all_probabilities"])
    for kw in ({}, {"tile_bits": 4, "low_bits": 2}, {"world": 4}):
        got = E.run_world(6, circ["operations"], strict_parity=1, **kw)
        assert np.max(np.abs(np.abs(got) ** 2 - want)) <= TOL
    text = E.run_world(6, circ["operations"], strict_parity=0)
    assert np.max(np.abs(np.abs(text) ** 2 - want)) > 0.05


def test_bank_conflict_free_lane_mapping():
    """Every round of a 30-qubit brickwork plan must give conflict-free shared-memory access under the default tile
    layout (full XOR fold, tile_core.h: swz with c = 0; lane choice in plan.cpp).  The optional TMA-compatible layout
    (tile_mover = 2) is bound to the hardware 128-byte swizzle: tiles whose low bits are contiguous (long TMA runs) leave
    only index bits 3-5 for the swizzle, which costs up to 4-way conflicts - one of the reasons it is not the default
    (DESIGN.md section 5)."""
    circ = C.random_brickwork_circuit(30, 20)
    p = E.EmuPlan(30, circ["operations"])
    worst = [p.max_conflict(i) for i in range(p.num_stages) if p.stage_kind(i) == E.S_TILE]
    assert max(worst) == 1
    assert p.num_stages < p.num_gates / 4
    p = E.EmuPlan(30, circ["operations"], tile_mover=2)
    worst = [p.max_conflict(i) for i in range(p.num_stages) if p.stage_kind(i) == E.S_TILE]
    assert max(worst) <= 4


@pytest.mark.parametrize("n", [7, 12, 15])
def test_tma_layout_matches_oracle(n):
    """The scheduler + interpreter / tensor-core rounds under the TMA-compatible layout give the oracle's amplitudes."""
    circ = C.random_brickwork_circuit(n, 6)
    got = E.run_world(n, circ["operations"], tile_mover=2)
    want = O.execute_circuit(circ)
    assert np.max(np.abs(got - want)) <= TOL


def test_tile_layout_is_linear_bijection():
    """swz(., c) must be a GF(2)-linear bijection of the tile for every run length c (the kernel composes addresses by
    XOR) and must keep a run of 2^c amplitudes in consecutive 128-byte rows (one TMA box)."""
    for m in (5, 9, 12, 13):
        for c in range(0, min(m, 11) + 1):
            img = np.array([E.lib().emu_swz(i, c) for i in range(1 << m)], dtype=np.int64)
            assert sorted(img.tolist()) == list(range(1 << m))
            rng = np.random.default_rng(c)
            a, b = rng.integers(0, 1 << m, 64), rng.integers(0, 1 << m, 64)
            assert all(img[x ^ y] == img[x] ^ img[y] for x, y in zip(a, b))
            if c >= 3:
                for h in range(0, 1 << m, 1 << c):
                    rows = img[h:h + (1 << c)] >> 3
                    assert rows.min() == img[h] >> 3 and rows.max() - rows.min() == (1 << (c - 3)) - 1
                    assert img[h] % 8 == (img[h] >> 3) % 8          # chunk = 0 ^ (row address & 7): hardware 128B swizzle


def test_direct_store_last_round(monkeypatch):
    """QCB_DIRECT_STORE=1: the last three-product round of a sweep writes to global memory itself (stage flag bit 1; the
    emulator follows the kernel's address arithmetic: batch offset XOR lane offset) - same amplitudes as the oracle."""
    import ctypes as CT
    from oracle import c_oracle as CO
    monkeypatch.setenv("QCB_DIRECT_STORE", "1")
    for n in (13, 16):
        circ = C.random_brickwork_circuit(n, 8)
        p = E.EmuPlan(n, circ["operations"])
        nw = E.lib().emu_program_words(p.h, None, 0)
        buf = (CT.c_uint64 * nw)()
        E.lib().emu_program_words(p.h, buf, nw)
        w = np.frombuffer(buf, dtype=np.uint64)
        pos, flagged, tiles = 4, 0, 0
        for _s in range(int(w[1])):
            kind = int(w[pos]); pos += 2
            if kind != 0:
                continue
            tiles += 1
            flagged += (int(w[pos + 41]) >> 1) & 1
            pos += int(w[pos + 40])
        assert tiles > 0 and flagged == tiles
        got = E.run_world(n, circ["operations"])
        assert np.max(np.abs(got - CO.apply_circuit(circ))) <= 1e-10
