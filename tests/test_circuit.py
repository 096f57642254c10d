"""QuantumCircuit mirror (quaru_b200/circuit.py): the reference's builder panics (operation.rs:276-321), its four
reducer unit tests (operation.rs:600-651), and -- what the reference never checks -- that every reducer preserves
the circuit's action (verified against the CPU oracle)."""
import numpy as np
import pytest

from oracle.oracle import OracleRegister
from tests import circuits
from tests.util import assert_close, random_state


def _q():
    import quaru_b200
    return quaru_b200


def test_two_pair_of_same_gates_cancel_with_one_pair_in_middle():     # operation.rs:600-613
    q = _q(); op = q.operation
    c = q.QuantumCircuit()
    for g in (op.hadamard(0), op.cnot(0, 1), op.cnot(0, 1), op.hadamard(0)):
        c.add_operation(g)
    c.reduce_circuit_cancel_gates()
    assert len(c.get_operations()) == 0


def test_gates_with_same_arity_and_target_mul_together():             # operation.rs:614-625
    q = _q(); op = q.operation
    c = q.QuantumCircuit()
    c.add_operation(op.hadamard(0))
    c.add_operation(op.not_(0))
    c.reduce_circuit_gates_with_same_targets()
    assert len(c.get_operations()) == 1
    want = op.not_(0).matrix() @ op.hadamard(0).matrix()               # later . earlier
    assert np.allclose(c.get_operations()[0].matrix(), want)


def test_hadamard_twice_and_cnot_twice_cancel():                      # operation.rs:627-651
    q = _q(); op = q.operation
    for g in (op.hadamard(0), op.cnot(0, 1)):
        c = q.QuantumCircuit()
        c.add_operation(g)
        c.add_operation(g)
        c.reduce_circuit_cancel_gates()
        assert len(c.get_operations()) == 0


def test_non_involutory_real_gates_are_not_cancelled():
    """The reference would drop two adjacent equal real rotations (operation.rs:202); that changes the circuit."""
    q = _q()
    ry = q.Operation(np.array([[np.cos(0.3), -np.sin(0.3)], [np.sin(0.3), np.cos(0.3)]]), [0])
    c = q.QuantumCircuit()
    c.add_operation(ry)
    c.add_operation(ry)
    c.reduce_circuit_cancel_gates()
    assert len(c.get_operations()) == 2


def test_builder_panics():                                            # operation.rs:276-321
    q = _q(); op = q.operation
    c = q.QuantumCircuit()
    c.add_operation(op.hadamard(0))
    c.add_measurement(0)
    with pytest.raises(RuntimeError, match="Cannot add operation after measurement"):
        c.add_operation(op.not_(0))
    with pytest.raises(RuntimeError, match="Cannot add measurement after measurement"):
        c.add_measurement(0)
    with pytest.raises(RuntimeError, match="Condition must be a measurement"):
        c.add_conditional_operation(op.not_(1), [(2, True)])
    c.add_conditional_operation(op.not_(1), [(0, True)])
    with pytest.raises(RuntimeError, match="conditional operation"):
        c.add_operation(op.hadamard(2))
    c.reset_register()
    assert c.get_operations() == [] and c.get_measurement_targets() == [] and c.get_conditional_operations() == []


def _apply_all(ops, n, seed):
    ora = OracleRegister(random_state(n, seed), n)
    for o in ops:
        ora.apply((o.matrix(), o.targets()))
    return ora.state


@pytest.mark.parametrize("reducer", ["reduce_circuit_cancel_gates", "reduce_circuit_gates_with_same_targets",
                                     "reduce_non_overlapping_gates", "reduce_gates_with_one_off_size"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_reducers_preserve_the_circuit(reducer, seed):
    q = _q()
    n = 6
    rng = np.random.default_rng(seed)
    raw = circuits.random_circuit(n, 3, seed=seed) + [o for o in circuits.mixed_circuit(n, 40, seed) if len(o[1]) <= 2]
    # sprinkle exact repeats so that the cancelling / merging reducers have something to do
    ops = []
    for m, t in raw:
        ops.append(q.Operation(m, t))
        if rng.random() < 0.3:
            ops.append(q.Operation(m, t))
    c = q.QuantumCircuit()
    for o in ops:
        c.add_operation(o)
    before = _apply_all(c.get_operations(), n, 5)
    n_before = len(c.get_operations())
    if reducer == "reduce_gates_with_one_off_size":
        c.reduce_gates_with_one_off_size(2)
    else:
        getattr(c, reducer)()
    assert len(c.get_operations()) <= n_before
    assert_close(_apply_all(c.get_operations(), n, 5), before, reducer)
    if reducer != "reduce_circuit_cancel_gates":
        assert len(c.get_operations()) < n_before
