"""Pins the CPU oracle: every known-answer / property test the reference holds for the hot path
(ported in tests/reference_suite.py) must pass on oracle/quaru_oracle.c, and the C oracle must agree with its two
numpy twins and with the reference's constructor tests (src/operation.rs:565-594, src/math.rs:91-122)."""
import numpy as np
import pytest

from oracle import gates as og
from oracle import oracle_np
from oracle.oracle import OracleRegister
from tests import reference_suite as rs
from tests.util import random_matrix, random_state


@pytest.mark.parametrize("test", rs.ALL, ids=lambda t: t.__name__)
def test_reference_suite_on_oracle(test):
    test(rs.OracleBackend(), np.random.default_rng(12345))


def test_period_finder_on_oracle():
    """examples/shors.rs:382-421 (20 trials, >= 15 successes), slow N < 6 and fast N < 15."""
    rs.t_period_finder_working(rs.OracleBackend(), np.random.default_rng(7))


def test_u_gate_on_oracle():
    """examples/shors.rs:433-463, exhaustive (61 440 cases)."""
    rs.t_test_u_gate(rs.OracleBackend(), None)


def test_shors_config_periods_on_oracle():
    """BASELINE config C2: find_period_slow for (N=15, a=7) -> 12 qubits; the period 4 (or a divisor) must come out."""
    rng = np.random.default_rng(3)
    ok = sum(4 % rs.find_period_slow(rs.OracleBackend(), rng, 15, 7) == 0 for _ in range(5))
    assert ok >= 4


@pytest.mark.parametrize("n", range(1, 9))
def test_c_oracle_equals_numpy_twins(n):
    """quaru_oracle.c == tensor-contraction twin == literal restatement of try_apply (register.rs:193-250)."""
    rng = np.random.default_rng(n)
    for k in range(1, min(n, 4) + 1):
        for trial in range(4):
            t = [int(x) for x in rng.permutation(n)[:k]]
            m = random_matrix(k, 100 * n + 10 * k + trial)
            s = random_state(n, 7 * n + trial, normalize=False)
            r = OracleRegister(s.copy(), n)
            r.apply((m, t))
            assert np.abs(r.state - oracle_np.apply_np(s, n, m, t)).max() < 1e-12
            assert np.abs(r.state - oracle_np.apply_slow_np(s, n, m, t)).max() < 1e-12
            r2 = OracleRegister(s.copy(), n)
            r2.apply((m, t), faithful=True)                      # allocation-faithful variant: same bits
            assert (r2.state == r.state).all()
            r3 = OracleRegister(s.copy(), n)
            r3.apply((m, t), threads=3)                          # OpenMP variant: same bits
            assert (r3.state == r.state).all()


def test_oracle_measure_matches_numpy_twin():
    for n in [1, 3, 6, 9]:
        s = random_state(n, n)
        for t in range(n):
            for u in [0.0, 0.3, 0.9999]:
                r = OracleRegister(s.copy(), n)
                got = r.measure(t, u)
                want, st = oracle_np.measure_np(s, n, t, u)
                assert got == want and np.abs(r.state - st).max() < 1e-14


def test_oracle_apply_all_literal_equals_per_qubit():
    """register.rs:273-279 (kron n times, one mat-vec) vs the gate on each qubit (apply_all_test :211-219)."""
    for n in range(1, 9):
        s = random_state(n, 50 + n)
        m = random_matrix(1, n, unitary=True)
        a = OracleRegister(s.copy(), n).try_apply_all((m, [0]), literal=True).state
        b = OracleRegister(s.copy(), n).try_apply_all((m, [0]), literal=False).state
        assert np.abs(a - b).max() < 1e-13


# ---- src/operation.rs:565-594 on the restated constructors --------------------------------------------
def _all_ops():
    return [og.identity(0), og.hadamard(0), og.cnot(0, 1), og.swap(0, 1), og.phase(0), og.not_(0), og.pauli_y(0),
            og.pauli_z(0), og.cnx([0], 1), og.cnx([0, 1], 2), og.cnx([0, 1, 2], 3), og.cnx([0, 1, 2, 3], 4),
            og.cnx([0, 1, 2, 3, 4], 5)]


def test_sz_matches_and_unitary():
    for m, t in _all_ops():
        assert m.shape == (1 << len(t), 1 << len(t))
        assert np.abs(m @ m.conj().T - np.eye(m.shape[0])).max() < 1e-8


def test_toffoli2_equals_cnot():
    assert np.abs(og.cnx([0], 1)[0] - og.cnot(0, 1)[0]).max() < 1e-8
    assert og.cnx([0], 1)[1] == og.cnot(0, 1)[1] == [1, 0]


def test_limit_denominator_working():
    """src/math.rs:94-121 brute force, m, n, l < 30."""
    mx = 30
    for m in range(mx):
        for n in range(1, mx):
            for l in range(1, mx):
                r, s = og.limit_denominator(m, n, l)
                if m == 0:
                    assert (r, s) == (0, 1)
                    continue
                best = min(abs(r2 / s2 - m / n) for r2 in range(1, mx) for s2 in range(1, l + 1))
                assert best + 1e-15 >= abs(r / s - m / n)


def test_product_constructors_match_oracle_constructors():
    """quaru_b200/operation.py (product) vs oracle/gates.py (restatement): same matrices, same target order."""
    from quaru_b200 import operation as op
    pairs = [(op.identity(3), og.identity(3)), (op.hadamard(2), og.hadamard(2)), (op.cnot(4, 1), og.cnot(4, 1)),
             (op.swap(0, 5), og.swap(0, 5)), (op.phase(1), og.phase(1)), (op.not_(0), og.not_(0)),
             (op.pauli_y(2), og.pauli_y(2)), (op.pauli_z(2), og.pauli_z(2)), (op.cnx([3, 1, 0], 2), og.cnx([3, 1, 0], 2)),
             (op.cnz([0, 4], 2), og.cnz([0, 4], 2)), (op.hadamard_transform([2, 0, 3]), og.hadamard_transform([2, 0, 3])),
             (op.u(0.3, 1.1, -0.7, 2), og.u(0.3, 1.1, -0.7, 2)), (op.rz(0.9, 1), og.rz(0.9, 1)), (op.cz(1, 3), og.cz(1, 3)),
             (op.cphase(0.4, 2, 0), og.cphase(0.4, 2, 0)),
             (op.to_controlled(op.hadamard(1), 3), og.to_controlled(og.hadamard(1), 3)),
             (op.to_quantum_gate(lambda x: (x * 7) % 15 if x < 15 else x, [0, 1, 2, 3]),
              og.to_quantum_gate(lambda x: (x * 7) % 15 if x < 15 else x, [0, 1, 2, 3]))]
    for a, (m, t) in pairs:
        assert a.targets() == list(t) and (a.matrix() == m).all()
    assert op.Operation.new(np.eye(4), [0]) is None and op.Operation.new(np.eye(4), [0, 1]) is not None
    assert op.hadamard(0).arity() == 1 and op.cnx([1, 2], 0).arity() == 3


def test_u_gate_matches_openqasm_paper_form():
    """openqasm.rs:145-163: U(pi/2, 0, pi) is H up to the global phase -i... the example program's `h`."""
    m, t = og.u(np.pi / 2, 0.0, np.pi, 0)
    h = og.hadamard(0)[0]
    ratio = m[0, 0] / h[0, 0]
    assert abs(abs(ratio) - 1) < 1e-12 and np.abs(m - ratio * h).max() < 1e-12
