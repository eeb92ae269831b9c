"""Plan traces (csrc/plan.h: PlanTrace): a circuit whose STRUCTURE the handle has seen before is planned by replaying the
recorded scheduler decisions with the new angles (variational loops: the VQE / QAOA objective of the reference,
application/algorithm/variational_algorithm.clj:330-360, submits the same ansatz with new parameters every evaluation).
The replayed program must be word-for-word the program a fresh scheduling run of the same circuit produces.
Host-only (qcb_plan_* API), no GPU."""
import ctypes as CT
import math

import numpy as np
import pytest

from qclojure_b200 import _lib as L
from qclojure_b200 import circuits as C
from qclojure_b200 import ops as OPS


def _words(plan_ptr):
    lib = L.load()
    nw = CT.c_uint64()
    lib.qcb_plan_serialize(plan_ptr, None, 0, CT.byref(nw))
    buf = np.empty(nw.value, dtype=np.uint64)
    lib.qcb_plan_serialize(plan_ptr, buf.ctypes.data, nw.value, CT.byref(nw))
    return buf


def _fresh(n, ops, **kw):
    lib = L.load()
    cfg = OPS.make_config(n, **kw)
    arr, cnt, keep = OPS.encode_ops(ops)
    p = CT.c_void_p()
    assert lib.qcb_plan_create(CT.byref(cfg), arr, cnt, CT.byref(p)) == 0, L.last_error(None)
    w = _words(p)
    lib.qcb_plan_destroy(p)
    return w


def _replayed(n, ops_recorded, ops, **kw):
    lib = L.load()
    cfg = OPS.make_config(n, **kw)
    a, cnt, keep_a = OPS.encode_ops(ops_recorded)
    b, cnt_b, keep_b = OPS.encode_ops(ops)
    assert cnt == cnt_b
    p = CT.c_void_p()
    rc = lib.qcb_plan_create_replayed(CT.byref(cfg), a, b, cnt, CT.byref(p))
    if rc != 0:
        raise L.QcbError(rc, L.last_error(None))
    w = _words(p)
    lib.qcb_plan_destroy(p)
    return w


def _qaoa(n, params):
    graph = C.random_regular_graph(n, 3, seed=11)
    return C.qaoa_ansatz_circuit(C.max_cut_hamiltonian(graph, n), C.standard_mixer_hamiltonian(n), params, n)["operations"]


def _reangle(ops, seed):
    rng = np.random.default_rng(seed)
    out = []
    for op in ops:
        p = dict(op["operation-params"])
        if "angle" in p:
            p["angle"] = float(rng.uniform(0.05, 2 * math.pi - 0.05))
        out.append({"operation-type": op["operation-type"], "operation-params": p})
    return out


@pytest.mark.parametrize("n,kw", [(12, {}), (16, {}), (20, {}), (14, {"tile_bits": 8, "low_bits": 3}), (14, {"dense_mma": 2})])
def test_replayed_qaoa_program_equals_fresh_program(n, kw):
    rec = _qaoa(n, [0.3, 0.2, 0.15, 0.1])
    new = _qaoa(n, [1.1, 0.7, 2.2, 0.9])
    assert np.array_equal(_replayed(n, rec, new, **kw), _fresh(n, new, **kw))
    assert not np.array_equal(_fresh(n, rec, **kw), _fresh(n, new, **kw))      # the angles do change the program


@pytest.mark.parametrize("n,depth", [(12, 6), (18, 8), (24, 10), (30, 20)])
def test_replayed_brickwork_program_equals_fresh_program(n, depth):
    rec = C.random_brickwork_circuit(n, depth, seed=1000 + n)["operations"]
    new = _reangle(rec, 5)
    assert np.array_equal(_replayed(n, rec, new), _fresh(n, new))


def test_replay_with_every_gate_kind_grover_operators_and_global_qubits():
    from tests.test_oracle_c import _all_gates_circuit
    n = 10
    rec = _all_gates_circuit(n, np.random.default_rng(3))["operations"]
    rec = [o for o in rec if o["operation-type"] != "measure"]
    rec += [{"operation-type": "phase-oracle", "operation-params": {"index": 5}},
            {"operation-type": "grover-diffusion", "operation-params": {}},
            {"operation-type": "h", "operation-params": {"target": 3}},
            {"operation-type": "grover-diffusion", "operation-params": {}}]
    new = _reangle(rec, 9)
    for kw in ({}, {"tile_bits": 6, "low_bits": 2}, {"fusion": 0}):
        assert np.array_equal(_replayed(n, rec, new, **kw), _fresh(n, new, **kw))
    # sharded layout: exchanges are part of the trace
    n = 14
    rec = C.random_brickwork_circuit(n, 6, seed=2)["operations"]
    new = _reangle(rec, 4)
    for rank in (0, 3):
        kw = {"rank": rank, "world_size": 4, "tile_bits": 8, "low_bits": 3}
        assert np.array_equal(_replayed(n, rec, new, **kw), _fresh(n, new, **kw))


def test_structure_mismatch_is_refused():
    n = 8
    a = C.random_brickwork_circuit(n, 4, seed=1)["operations"]
    b = [dict(o) for o in a]
    b[5] = {"operation-type": "h", "operation-params": {"target": (a[5]["operation-params"].get("target", 0) + 1) % n}}
    with pytest.raises(L.QcbError, match="differ in structure"):
        _replayed(n, a, b)
    # a phase of exactly pi is a sign flip (cheaper device op, different scheduling class): part of the key
    c1 = [{"operation-type": "phase", "operation-params": {"target": 0, "angle": 0.3}}] * 20
    c2 = [{"operation-type": "phase", "operation-params": {"target": 0, "angle": math.pi}}] * 20
    assert np.array_equal(_replayed(n, c1, _reangle(c1, 1)), _fresh(n, _reangle(c1, 1)))
    try:
        w = _replayed(n, c1, c2)
        assert np.array_equal(w, _fresh(n, c2))
    except L.QcbError as e:
        assert "differ in structure" in str(e)
