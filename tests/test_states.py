"""Initial-state constructors against the reference's documented values
(`/root/reference/src/org/soulspace/qclojure/domain/state.clj` docstrings and test/.../domain/state_test.clj)."""
import math

import numpy as np
import pytest

from qclojure_b200 import states as S

R = 1 / math.sqrt(2)


def test_single_qubit_states():
    assert np.allclose(S.zero_state()["state-vector"], [1, 0]) and np.allclose(S.one_state()["state-vector"], [0, 1])
    assert np.allclose(S.plus_state()["state-vector"], [R, R]) and np.allclose(S.minus_state()["state-vector"], [R, -R])
    assert np.allclose(S.plus_i_state()["state-vector"], [R, R * 1j]) and np.allclose(S.minus_i_state()["state-vector"], [R, -R * 1j])


def test_multi_qubit_states_follow_the_reference_code():
    assert S.zero_state(3)["num-qubits"] == 3 and S.zero_state(3)["state-vector"][0] == 1 and S.one_state(3)["state-vector"][7] == 1
    assert np.allclose(S.plus_state(3)["state-vector"], 1 / math.sqrt(8))
    a = 0.5
    # state.clj:377-385, 413-421, 449-457: first amplitude +a, all others -a / +ia / -ia (not a tensor power)
    assert np.allclose(S.minus_state(2)["state-vector"], [a, -a, -a, -a])
    assert np.allclose(S.plus_i_state(2)["state-vector"], [a, a * 1j, a * 1j, a * 1j])
    assert np.allclose(S.minus_i_state(2)["state-vector"], [a, -a * 1j, -a * 1j, -a * 1j])
    for f in (S.plus_state, S.minus_state, S.plus_i_state, S.minus_i_state):
        assert abs(np.linalg.norm(f(4)["state-vector"]) - 1) < 1e-14


def test_basis_states_and_bit_order():
    # state_test.clj:197-227 and the docstring (computational-basis-state 2 [1 1]) => |11>
    assert S.bits_to_index([1, 0, 1]) == 5 and S.index_to_bits(5, 3) == [1, 0, 1] and S.index_to_bits(1, 3) == [0, 0, 1]
    st = S.computational_basis_state(3, [1, 0, 0])
    assert st["state-vector"][4] == 1 and np.count_nonzero(st["state-vector"]) == 1
    with pytest.raises(ValueError):
        S.computational_basis_state(2, [1, 2])


def test_normalize_and_tensor_product():
    # docstrings: normalize [3 4] -> [0.6 0.8]; tensor |0> x |1> = |01> = [0 1 0 0]
    n = S.normalize_state(S.multi_qubit_state([3, 4]))
    assert np.allclose(n["state-vector"], [0.6, 0.8]) and n["num-qubits"] == 1
    z = S.normalize_state(S.multi_qubit_state([0, 0]))
    assert np.allclose(z["state-vector"], [0, 0])                      # norm below tolerance: left alone
    t = S.tensor_product(S.zero_state(), S.one_state())
    assert t["num-qubits"] == 2 and np.allclose(t["state-vector"], [0, 1, 0, 0])
    t3 = S.tensor_product(S.plus_state(), S.computational_basis_state(2, [1, 0]))
    assert np.allclose(t3["state-vector"], [0, 0, R, 0, 0, 0, R, 0])
