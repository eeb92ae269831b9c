"""Wire formats: EDN / JSON round trips of states and circuits in the reference's layout
(`/root/reference/src/org/soulspace/qclojure/adapter/io.clj:33-114`, `adapter/io/edn.clj`, `adapter/io/json.clj`).
Host-side only; no GPU."""
import json
import math

import numpy as np
import pytest

from qclojure_b200 import circuits as CB
from qclojure_b200 import io as QIO


def test_complex_map_round_trip():
    # io.clj:10-31
    m = QIO.complex_to_map(0.5 - 0.25j)
    assert m == {"real": 0.5, "imag": -0.25}
    assert QIO.map_to_complex(m) == 0.5 - 0.25j


def test_state_serialisation_layout():
    # io.clj:33-51: keys and "1.0" format version, amplitudes as {:real :imag} maps in index order
    vec = np.array([1, 0, 0, 1j]) / math.sqrt(2)
    ser = QIO.serialize_quantum_state({"state-vector": vec, "num-qubits": 2})
    assert list(ser) == ["state-vector", "num-qubits", "metadata", "format-version"]
    assert ser["format-version"] == "1.0" and ser["metadata"] == {} and ser["num-qubits"] == 2
    assert ser["state-vector"][0] == {"real": 0.7071067811865475, "imag": 0.0}
    assert ser["state-vector"][3] == {"real": 0.0, "imag": 0.7071067811865475}
    back = QIO.deserialize_quantum_state(ser)
    assert back["num-qubits"] == 2 and np.array_equal(back["state-vector"], vec)


@pytest.mark.parametrize("fmt", ["edn", "json", ":edn", ":json"])
def test_state_file_round_trip_is_bit_exact(tmp_path, fmt):
    rng = np.random.default_rng(5)
    vec = rng.standard_normal(32) + 1j * rng.standard_normal(32)
    vec /= np.linalg.norm(vec)
    vec[3] = 1e-300 + 2.5e17j              # exponent forms
    f = tmp_path / f"state.{fmt.lstrip(':')}"
    ret = QIO.export_quantum_state(fmt, {"state-vector": vec, "num-qubits": 5, "metadata": {"tag": "t"}}, str(f))
    assert ret == (None if fmt.lstrip(":") == "edn" else str(f))        # the reference's return values (io/edn.clj:12-15)
    back = QIO.import_quantum_state(fmt, str(f))
    assert back["num-qubits"] == 5 and back["metadata"] == {"tag": "t"}
    assert np.array_equal(back["state-vector"], vec)      # shortest round-trip doubles, no loss


def test_edn_text_is_what_pr_str_writes():
    c = CB.create_circuit(2, "Bell State")
    CB.cnot(CB.h(c, 0), 0, 1)
    CB.rz(c, 1, 0.25)
    text = QIO.write_edn(QIO.serialize_quantum_circuit(c))
    assert text.startswith('{:operations [{:operation-type :h, :operation-params {:target 0}} '
                           '{:operation-type :cnot, :operation-params {:control 0, :target 1}} '
                           '{:operation-type :rz, :operation-params {:target 1, :angle 0.25}}], :num-qubits 2, '
                           ':name "Bell State"')
    assert text.endswith(':format-version "1.0"}')


def test_edn_reader_accepts_clojure_output():
    # a circuit as a QClojure REPL prints it (doc/tutorial.md op lists use this layout), with commas, nil and a comment
    text = """
    {:operations [{:operation-type :h, :operation-params {:target 0}}
                  {:operation-type :crz, :operation-params {:control 1, :target 0, :angle 1.5707963267948966}} ; QFT step
                  {:operation-type :measure, :operation-params {:measurement-qubits [0 1]}}],
     :num-qubits 2, :name "x", :description nil, :metadata {}, :format-version "1.0"}"""
    c = QIO.deserialize_quantum_circuit(QIO.read_edn(text))
    assert c["num-qubits"] == 2 and len(c["operations"]) == 3
    assert c["operations"][1]["operation-type"] == "crz"
    assert repr(c["operations"][1]["operation-type"]) == ":crz"
    assert c["operations"][1]["operation-params"] == {"control": 1, "target": 0, "angle": math.pi / 2}
    assert c["operations"][2]["operation-params"]["measurement-qubits"] == [0, 1]
    assert "description" not in c


def test_edn_scalars_and_errors():
    assert QIO.read_edn("[1 -2 3N 1.5 -2.0E-5 1e3 4.5M ##Inf ##-Inf nil true false \"a\\\"b\" :k sym]")[:-1] == \
        [1, -2, 3, 1.5, -2.0e-5, 1000.0, 4.5, float("inf"), float("-inf"), None, True, False, 'a"b', "k"]
    assert math.isnan(QIO.read_edn("##NaN"))
    assert QIO.read_edn("#{1 2}") == {1, 2}
    assert QIO.read_edn("[1/2 -3/4 \\a \\newline :ns/kw]") == [0.5, -0.75, "a", "\n", "ns/kw"]
    assert QIO.write_edn(1e-7) == "1.0E-7" and QIO.write_edn(1e22) == "1.0E22" and QIO.write_edn(0.1) == "0.1"
    for bad in ("", "[1 2", "{:a}", "]"):
        with pytest.raises(ValueError):
            QIO.read_edn(bad)


@pytest.mark.parametrize("fmt", ["edn", "json"])
def test_circuit_file_round_trip(tmp_path, fmt):
    c = CB.quantum_fourier_transform_circuit(4)
    CB.measure(c, [0, 1, 2, 3])
    f = tmp_path / f"c.{fmt}"
    QIO.export_quantum_circuit(fmt, c, str(f))
    back = QIO.import_quantum_circuit(fmt, str(f))
    assert back["num-qubits"] == 4 and back["name"] == c["name"]
    assert [(o["operation-type"], o["operation-params"]) for o in back["operations"]] == \
        [(o["operation-type"], o["operation-params"]) for o in c["operations"]]
    if fmt == "json":      # clojure.data.json writes keywords as plain strings (io/json.clj)
        raw = json.loads(f.read_text())
        assert raw["operations"][0] == {"operation-type": "h", "operation-params": {"target": 0}}
        assert raw["format-version"] == "1.0"


def test_quantum_data_dispatch(tmp_path):
    # io.clj:90-114
    st = {"state-vector": np.array([1, 0], dtype=complex), "num-qubits": 1}
    c = CB.bell_state_circuit()
    assert "state-vector" in QIO.serialize_quantum_data(st)
    assert "operations" in QIO.serialize_quantum_data(c)
    with pytest.raises(ValueError, match="Unsupported quantum data type"):
        QIO.serialize_quantum_data({"foo": 1})
    with pytest.raises(ValueError, match="Unsupported quantum data format"):
        QIO.deserialize_quantum_data({"foo": 1})
    for fmt in ("edn", "json"):
        f = tmp_path / f"d.{fmt}"
        QIO.export_quantum_data(fmt, c, str(f))
        assert QIO.import_quantum_data(fmt, str(f))["num-qubits"] == 2
        QIO.export_quantum_data(fmt, st, str(f))
        assert np.array_equal(QIO.import_quantum_data(fmt, str(f))["state-vector"], st["state-vector"])
    with pytest.raises(ValueError, match="unsupported format"):
        QIO.export_quantum_data("xml", c, str(tmp_path / "x"))


def test_state_length_must_match_qubits():
    with pytest.raises(ValueError):
        QIO.deserialize_quantum_state({"state-vector": [{"real": 1.0, "imag": 0.0}] * 3, "num-qubits": 2})


def test_imported_circuit_feeds_the_op_encoder():
    """A circuit read back from EDN/JSON encodes to the same C-ABI op list as the original (no GPU needed)."""
    from qclojure_b200 import ops as OPS
    c = CB.random_brickwork_circuit(6, 3, seed=4)
    a, na, _keep = OPS.encode_ops(OPS.circuit_ops(c))
    for fmt in ("edn", "json"):
        text = QIO._dump(fmt, QIO.serialize_quantum_circuit(c))
        c2 = QIO.deserialize_quantum_circuit(QIO._load(fmt, text))
        b, nb, _keep2 = OPS.encode_ops(OPS.circuit_ops(c2))
        assert na == nb == len(c["operations"])
        assert bytes(a) == bytes(b)
