"""OpenQASM 3 import/export.  Cases follow the reference's own tests
(`/root/reference/test/org/soulspace/qclojure/application/format/qasm3_test.clj`) and the sample in
`application/format/qasm3.clj` (comment block).  Host-side only."""
import math

import numpy as np
import pytest

from qclojure_b200 import circuits as C
from qclojure_b200 import io as QIO
from qclojure_b200 import qasm3 as Q


def _lines(text):
    return text.splitlines()


def test_header_and_basic_gates():
    # qasm3_test.clj: test-circuit-to-qasm-header / -basic-gates / -pauli-gates / -controlled-gates
    c = C.create_circuit(3, "Test")
    for g, t in (("x", 0), ("y", 1), ("z", 0), ("h", 1), ("s", 0), ("t", 1), ("s-dag", 0), ("t-dag", 1), ("i", 2)):
        C.add_gate(c, g, target=t)
    C.cnot(c, 0, 1); C.cz(c, 1, 2); C.add_gate(c, "cy", control=0, target=2)
    q = Q.circuit_to_qasm(c)
    assert q.startswith('OPENQASM 3.0;\ninclude "stdgates.inc";\n\nqubit[3] q;\nbit[3] c;\n')
    for ln in ("x q[0];", "y q[1];", "z q[0];", "h q[1];", "s q[0];", "t q[1];", "sdg q[0];", "tdg q[1];", "id q[2];",
               "cx q[0], q[1];", "cz q[1], q[2];", "cy q[0], q[2];"):
        assert ln in _lines(q)


def test_rotation_swap_three_qubit_and_measure_lines():
    # qasm3_test.clj: -rotation-gates / -phase-gate / -swap-gates / -three-qubit-gates / -measurements
    c = C.create_circuit(4)
    C.rx(c, 0, math.pi / 2); C.ry(c, 1, math.pi / 4); C.rz(c, 0, math.pi / 3); C.phase(c, 0, math.pi / 4)
    C.swap(c, 0, 1); C.iswap(c, 1, 2); C.toffoli(c, 0, 1, 2); C.fredkin(c, 3, 0, 2)
    C.crx(c, 1, 3, math.pi / 4); C.cry(c, 2, 0, math.pi / 3); C.crz(c, 0, 3, math.pi / 8)
    C.measure(c, [0, 1])
    q = Q.circuit_to_qasm(c)
    for ln in ("rx(1.5707963267948966) q[0];", "ry(0.7853981633974483) q[1];", "rz(1.0471975511965976) q[0];",
               "p(0.7853981633974483) q[0];", "swap q[0], q[1];", "iswap q[1], q[2];", "ccx q[0], q[1], q[2];",
               "cswap q[3], q[0], q[2];", "crx(0.7853981633974483) q[1], q[3];", "cry(1.0471975511965976) q[2], q[0];",
               "crz(0.39269908169872414) q[0], q[3];", "c[0] = measure q[0];", "c[1] = measure q[1];"):
        assert ln in _lines(q), ln


def test_braket_dialect_and_neutral_atom_gates():
    c = C.create_circuit(2)
    C.add_gate(c, "s-dag", target=0); C.add_gate(c, "t-dag", target=1); C.phase(c, 0, 0.5); C.cnot(c, 0, 1)
    q = Q.circuit_to_qasm(c, {"target": "braket"})
    assert "stdgates.inc" not in q
    for ln in ("si q[0];", "ti q[1];", "phaseshift(0.5) q[0];", "cnot q[0], q[1];"):
        assert ln in _lines(q)
    c = C.create_circuit(2)
    C.add_gate(c, "global-h"); C.add_gate(c, "global-rx", angle=0.25)
    C.add_gate(c, "rydberg-cz", control=0, target=1); C.add_gate(c, "rydberg-cphase", control=0, target=1, angle=0.5)
    C.add_gate(c, "rydberg-blockade", qubit_indices=[0, 1], angle=0.1); C.add_gate(c, "mystery", target=0)
    q = Q.circuit_to_qasm(c)
    for ln in ("// Global Hadamard gate - apply H to all qubits", "h q[0];", "h q[1];", "rx(0.25) q[1];",
               "// Global RX(0.25) gate - apply RX to all qubits", "cz q[0], q[1];", "crz(0.5) q[0], q[1];",
               "// Rydberg blockade gate - hardware specific", "// Unknown gate: mystery"):
        assert ln in _lines(q), ln


def test_round_trip_preserves_operations():
    # qasm3_test.clj round-trip tests: gate types, qubits and angles survive; aliases resolve (cx -> cnot, sdg -> s-dag)
    c = C.create_circuit(4, "rt")
    C.h(c, 0); C.cnot(c, 0, 1); C.add_gate(c, "s-dag", target=2); C.add_gate(c, "t-dag", target=3)
    C.rx(c, 0, 0.1); C.ry(c, 1, 1e-7); C.rz(c, 2, -2.5); C.phase(c, 3, math.pi / 6)
    C.crx(c, 0, 1, 0.3); C.cry(c, 1, 2, 0.4); C.crz(c, 2, 3, 0.5); C.cz(c, 0, 3); C.add_gate(c, "cy", control=1, target=0)
    C.swap(c, 0, 3); C.iswap(c, 1, 2); C.toffoli(c, 0, 1, 2); C.fredkin(c, 0, 1, 2); C.add_gate(c, "i", target=1)
    C.measure(c, [0]); C.measure(c, [3])
    back = Q.qasm_to_circuit(Q.circuit_to_qasm(c))
    assert back["num-qubits"] == 4 and back["name"] == "Converted Circuit" and back["result-specs"] == {}
    assert [(o["operation-type"], o["operation-params"]) for o in back["operations"]] == \
        [(o["operation-type"], o["operation-params"]) for o in c["operations"]]


def test_brickwork_benchmark_circuit_round_trip_is_exact():
    """The BASELINE.json configs[2] generator survives QASM text bit for bit (angles print in shortest round-trip form)."""
    c = C.random_brickwork_circuit(12, 6, seed=1012)
    back = Q.qasm_to_circuit(Q.circuit_to_qasm(c))
    assert len(back["operations"]) == len(c["operations"])
    for a, b in zip(c["operations"], back["operations"]):
        assert a["operation-type"] == b["operation-type"] and a["operation-params"] == b["operation-params"]


def test_sample_program_from_the_reference():
    # qasm3.clj comment block: "OPENQASM 3;" header, pi expressions, trailing comments
    src = '''OPENQASM 3;
include "stdgates.inc";

qubit[4] q;
bit[4] c;

rz(pi/3) q[0];
rz(pi/5) q[0];   // can be folded to rz(pi/3 + pi/5)
rx(pi/2) q[1];
rx(-pi/2) q[1];  // cancels
h q[2];
h q[2];          // cancels
cx q[0], q[1];
cx q[0], q[1];   // cancels'''
    c = Q.qasm_to_circuit(src)
    assert c["num-qubits"] == 4
    ops = [(o["operation-type"], o["operation-params"]) for o in c["operations"]]
    assert ops == [("rz", {"target": 0, "angle": math.pi / 3}), ("rz", {"target": 0, "angle": math.pi / 5}),
                   ("rx", {"target": 1, "angle": math.pi / 2}), ("rx", {"target": 1, "angle": -math.pi / 2}),
                   ("h", {"target": 2}), ("h", {"target": 2}),
                   ("cnot", {"control": 0, "target": 1}), ("cnot", {"control": 0, "target": 1})]


def test_expressions():
    # qasm3.clj:391-441
    P = Q.parse_qasm_expression
    assert P("0.5") == 0.5 and P("-2") == -2.0 and P("pi") == math.pi and P("-pi") == -math.pi
    assert P("pi/4") == math.pi / 4 and P("-pi/2") == -math.pi / 2 and P("2*pi") == 2 * math.pi
    assert P("pi*2") == math.pi * 2 and P("3/4") == 0.75 and P(" 1.5e-3 ") == 1.5e-3
    with pytest.raises(ValueError, match="Unsupported QASM expression"):
        P("sin(pi)")


def test_result_pragmas_emit_and_parse():
    # qasm3_test.clj result-spec tests; qasm3.clj:43-131, 367-389, 443-470
    c = C.bell_state_circuit()
    specs = {"measurements": {"shots": 500, "qubits": [0, 1]},
             "expectation": {"observables": ["pauli-z", "pauli-x"], "targets": [0, 1]},
             "variance": {"observables": ["pauli-x"], "targets": [1]},
             "probability": {"targets": [0, 1], "states": ["00", "11"]}, "amplitude": {"states": ["00", "11"]},
             "sample": {"observables": ["pauli-z"], "shots": 100, "targets": [0]}, "state-vector": True}
    q = Q.circuit_to_qasm(c, {"result-specs": specs})
    for ln in ("// Result extraction specifications", "#pragma qclojure result measurement shots=500 qubits=0,1",
               "#pragma qclojure result expectation observable=pauli-z target=0",
               "#pragma qclojure result expectation observable=pauli-x target=1",
               "#pragma qclojure result variance observable=pauli-x target=1",
               "#pragma qclojure result probability targets=0,1 states=00,11",
               "#pragma qclojure result amplitude states=00,11",
               "#pragma qclojure result sample observable=pauli-z shots=100 target=0",
               "#pragma qclojure result state_vector // Simulation-only result"):
        assert ln in _lines(q), ln
    back = Q.qasm_to_circuit(q)
    rs = back["result-specs"]
    assert rs["measurement"] == {"shots": 500, "qubits": ["0", "1"]}
    # the first pragma of a kind is kept as parsed, later ones are merged into :observables / :targets (qasm3.clj:456-466)
    assert rs["expectation"]["observable"] == "pauli-z" and rs["expectation"]["observables"] == ["pauli-x"]
    assert rs["expectation"]["targets"] == [1]
    assert rs["variance"] == {"observable": "pauli-x", "target": 1}
    assert rs["amplitude"] == {"states": ["00", "11"]} and rs["state_vector"] == {}
    assert [o["operation-type"] for o in back["operations"]] == ["h", "cnot"]
    assert "#pragma braket result" in Q.circuit_to_qasm(c, {"result-specs": {"state-vector": True}, "target": "braket"})


def test_file_import_export_and_encoder(tmp_path):
    # adapter/io/qasm.clj
    from qclojure_b200 import ops as OPS
    c = C.quantum_fourier_transform_circuit(5)
    f = tmp_path / "qft.qasm"
    assert Q.export_quantum_circuit(c, str(f)) == str(f)
    back = Q.import_quantum_circuit(str(f))
    a, na, _ = OPS.encode_ops(OPS.circuit_ops(c))
    b, nb, _ = OPS.encode_ops(OPS.circuit_ops(back))
    assert na == nb and bytes(a) == bytes(b)
    with pytest.raises(ValueError):
        Q.qasm_to_circuit("OPENQASM 3.0;\nh q[0];")
    # the multimethod dispatch of adapter/io.clj on :qasm3
    QIO.export_quantum_circuit(":qasm3", c, str(tmp_path / "m.qasm"))
    assert len(QIO.import_quantum_circuit("qasm3", str(tmp_path / "m.qasm"))["operations"]) == len(c["operations"])


# ------------------------------------------------------------------ the reference's own recorded output (doc/tutorial.md:755-940)
def _tutorial_io_circuit():
    c = C.create_circuit(3, "I/O Test Circuit", "A circuit with medium complexity")
    C.h(c, 0); C.cnot(c, 0, 1); C.t_gate(c, 1); C.cnot(c, 1, 2); C.measure(c, [0, 1, 2])
    return c


TUTORIAL_QASM3 = '''OPENQASM 3.0;
include "stdgates.inc";

qubit[3] q;
bit[3] c;

h q[0];
cx q[0], q[1];
t q[1];
cx q[1], q[2];
c[0] = measure q[0];
c[1] = measure q[1];
c[2] = measure q[2];'''

TUTORIAL_QASM2 = '''OPENQASM 2.0;
include "qelib1.inc";
qreg q[3];
creg c[3];

h q[0];
cx q[0],q[1];
t q[1];
cx q[1],q[2];
// Measurement will be handled by final measure statement
measure q -> c;'''


def test_tutorial_qasm_text_is_reproduced_exactly(tmp_path):
    """The QASM 2 / QASM 3 files the reference's tutorial prints for its I/O test circuit, the circuits it reads back and
    the return values of the export calls (doc/tutorial.md: 'OpenQASM Support')."""
    from qclojure_b200 import qasm2 as Q2
    c = _tutorial_io_circuit()
    assert Q.circuit_to_qasm(c) == TUTORIAL_QASM3
    assert Q2.circuit_to_qasm(c) == TUTORIAL_QASM2
    f3, f2 = str(tmp_path / "test-circuit-qasm3.qasm"), str(tmp_path / "test-circuit-qasm2.qasm")
    assert QIO.export_quantum_circuit(":qasm3", c, f3) == f3 and QIO.export_quantum_circuit(":qasm2", c, f2) == f2
    back3 = QIO.import_quantum_circuit(":qasm3", f3)
    assert [(str(o["operation-type"]), o["operation-params"]) for o in back3["operations"]] == [
        ("h", {"target": 0}), ("cnot", {"control": 0, "target": 1}), ("t", {"target": 1}), ("cnot", {"control": 1, "target": 2}),
        ("measure", {"measurement-qubits": [0]}), ("measure", {"measurement-qubits": [1]}), ("measure", {"measurement-qubits": [2]})]
    assert back3["num-qubits"] == 3 and back3["name"] == "Converted Circuit" and back3["result-specs"] == {}
    back2 = QIO.import_quantum_circuit(":qasm2", f2)
    assert [(str(o["operation-type"]), o["operation-params"]) for o in back2["operations"]] == [
        ("h", {"target": 0}), ("cnot", {"control": 0, "target": 1}), ("t", {"target": 1}), ("cnot", {"control": 1, "target": 2})]
    assert back2["num-qubits"] == 3 and back2["name"] == "Converted Circuit" and "result-specs" not in back2


def test_qasm2_round_trip_and_quirks():
    from qclojure_b200 import qasm2 as Q2
    c = C.create_circuit(4)
    C.h(c, 0); C.add_gate(c, "s-dag", target=1); C.rx(c, 2, 0.5); C.phase(c, 3, 1e-7); C.crz(c, 0, 1, 0.25); C.cz(c, 1, 2)
    C.add_gate(c, "cy", control=2, target=3); C.swap(c, 0, 3); C.iswap(c, 1, 2); C.toffoli(c, 0, 1, 2); C.fredkin(c, 3, 0, 1)
    text = Q2.circuit_to_qasm(c)
    for ln in ("sdg q[1];", "rx(0.5) q[2];", "p(1.0E-7) q[3];", "crz(0.25) q[0],q[1];", "cy q[2],q[3];", "swap q[0],q[3];",
               "ccx q[0],q[1],q[2];", "cswap q[3],q[0],q[1];", "measure q -> c;"):
        assert ln in text.splitlines(), ln
    back = Q2.qasm_to_circuit(text)
    assert [(o["operation-type"], o["operation-params"]) for o in back["operations"]] == \
        [(o["operation-type"], o["operation-params"]) for o in c["operations"]]
    # lines are matched untrimmed and angles are plain numbers (qasm2.clj:165-262)
    assert Q2.qasm_to_circuit("qreg q[1];\n  h q[0];")["operations"] == []
    with pytest.raises(ValueError):
        Q2.qasm_to_circuit("qreg q[1];\nrx(pi/2) q[0];")
    g = Q2.circuit_to_qasm(C.add_gate(C.create_circuit(2), "global-h"))
    assert "// Global Hadamard gate - decomposed to individual H gates" in g and "h q[1];" in g


def test_tutorial_state_files(tmp_path):
    """doc/tutorial.md 'EDN Support' / 'JSON Support': |+> written and read back; EDN export returns nil, JSON the file name."""
    st = {"state-vector": [0.7071067811865475, 0.7071067811865475], "num-qubits": 1}
    fe, fj = str(tmp_path / "plus-state.edn"), str(tmp_path / "plus-state.json")
    assert QIO.export_quantum_state(":edn", st, fe) is None and QIO.export_quantum_state(":json", st, fj) == fj
    for fmt, f in ((":edn", fe), (":json", fj)):
        back = QIO.import_quantum_state(fmt, f)
        assert back["num-qubits"] == 1 and back["metadata"] == {}
        assert back["state-vector"].tolist() == [0.7071067811865475 + 0j, 0.7071067811865475 + 0j]
    assert open(fe).read() == ('{:state-vector [{:real 0.7071067811865475, :imag 0.0} {:real 0.7071067811865475, :imag 0.0}], '
                               ':num-qubits 1, :metadata {}, :format-version "1.0"}')
