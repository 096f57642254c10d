"""GPU parity of Register.apply_circuit (/root/reference/src/register.rs:345-377): the circuit's operations in order, then its
measurements in insertion order, then every conditional operation whose conditions are all among the results -- random
QuantumCircuits (gates + measurements + conditional operations, with and without each reducer of operation.rs:114-269)
through the engine and through the oracle under one shared stream of uniform draws `u`.  States <= 1e-12, outcome lists equal."""
import numpy as np
import pytest

from oracle import gates as og
from oracle.oracle import OracleRegister
from tests.util import assert_close, random_state

pytestmark = pytest.mark.gpu


def _random_circuit_parts(n, n_ops, n_meas, n_cond, rng):
    measured = [int(x) for x in rng.permutation(n)[:n_meas]]
    free = [q for q in range(n) if q not in measured]
    ops = []
    for _ in range(n_ops):
        kind = int(rng.integers(6))
        qs = [int(x) for x in rng.permutation(n)]
        if kind == 0:
            ops.append(og.hadamard(qs[0]))
        elif kind == 1:
            ops.append(og.cnot(qs[0], qs[1]))
        elif kind == 2:
            ops.append(og.rz(float(rng.uniform(0, 6.28)), qs[0]))
        elif kind == 3:
            ops.append(og.cz(qs[0], qs[1]))
        elif kind == 4:
            m = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
            ops.append((np.linalg.qr(m)[0], [qs[0]]))
        else:
            ops.append(og.swap(qs[0], qs[1]))
    conds = []
    for _ in range(n_cond):
        k = int(rng.integers(1, min(3, n_meas) + 1))
        cond = [(int(t), bool(rng.integers(2))) for t in rng.permutation(measured)[:k]]
        tq = [int(x) for x in rng.permutation(free)]
        g = [og.not_(tq[0]), og.pauli_z(tq[0]), og.hadamard(tq[0]), og.cnot(tq[0], tq[1])][int(rng.integers(4))]
        conds.append((g, cond))
    return ops, measured, conds


@pytest.mark.parametrize("n,seed", [(3, 0), (5, 1), (8, 2), (12, 3), (14, 4), (16, 5)])
@pytest.mark.parametrize("reducer", [None, "cancel", "same_targets", "non_overlapping", "one_off"])
def test_apply_circuit_matches_the_oracle(engine, n, seed, reducer):
    rng = np.random.default_rng(1000 * n + seed)
    ops, measured, conds = _random_circuit_parts(n, 40, min(3, n - 2), 4, rng)
    s = random_state(n, 77 + seed)
    us = [float(u) for u in rng.random(len(measured))]

    circ = engine.QuantumCircuit()
    for g in ops:
        circ.add_operation(engine.Operation(g[0], g[1]))
    if reducer == "cancel":
        circ.reduce_circuit_cancel_gates()
    elif reducer == "same_targets":
        circ.reduce_circuit_gates_with_same_targets()
    elif reducer == "non_overlapping":
        circ.reduce_non_overlapping_gates(3)
    elif reducer == "one_off":
        circ.reduce_gates_with_one_off_size(2)
    for t in measured:
        circ.add_measurement(t)
    for g, cond in conds:
        circ.add_conditional_operation(engine.Operation(g[0], g[1]), cond)

    reg = engine.Register.new([False] * n)
    reg.set_state(s)
    got = reg.apply_circuit(circ, us)

    ora = OracleRegister(s.copy(), n)
    want = ora.apply_circuit(ops, measured, conds, us)       # the UNREDUCED gate list: reducers must not change the map
    assert got == want, "measurement results, in insertion order"
    assert_close(reg.state, ora.state, f"apply_circuit n={n} reducer={reducer}")


def test_apply_circuit_without_measurements_and_builder_panics(engine):
    n = 6
    circ = engine.QuantumCircuit()
    circ.add_operation(engine.operation.hadamard(0))
    circ.add_operation(engine.operation.cnot(0, 5))
    reg = engine.Register.new([False] * n)
    assert reg.apply_circuit(circ) == []                      # register.rs:352: no measurement targets -> empty Vec
    ora = OracleRegister.new([False] * n)
    ora.apply_circuit([og.hadamard(0), og.cnot(0, 5)], [], [], [])
    assert_close(reg.state, ora.state)
    circ.add_measurement(5)
    with pytest.raises(RuntimeError):                         # operation.rs:277-283
        circ.add_operation(engine.operation.hadamard(5))
    with pytest.raises(RuntimeError):                         # operation.rs:293-297
        circ.add_measurement(5)
    with pytest.raises(RuntimeError):                         # operation.rs:309-315: condition must be a measurement
        circ.add_conditional_operation(engine.operation.not_(1), [(2, True)])
