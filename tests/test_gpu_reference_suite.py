"""The reference's own tests (tests/reference_suite.py) on the CUDA engine, fused and unfused, plus the two
algorithm configs of BASELINE.json (Grover target 64; Shor 15 and 21) compared state-for-state and
outcome-for-outcome with the oracle under a shared seeded uniform stream."""
import numpy as np
import pytest

from tests import reference_suite as rs
from tests.util import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fusion", [1, 0])
@pytest.mark.parametrize("test", rs.ALL, ids=lambda t: t.__name__)
def test_reference_suite_on_engine(engine, test, fusion):
    test(rs.EngineBackend(engine, fusion), np.random.default_rng(12345))


def test_period_finder_on_engine(engine):
    """examples/shors.rs:382-421 with 10 trials / 7 successes per (N, a) to bound the run time."""
    rs.t_period_finder_working(rs.EngineBackend(engine), np.random.default_rng(7), trials=10, need=7)


def test_u_gate_on_engine(engine):
    """examples/shors.rs:433-463, every 7th of the 61 440 cases."""
    rs.t_test_u_gate(rs.EngineBackend(engine), None, stride=7)


def test_config_c1_grover_target_64(engine):
    """BASELINE configs[0]: examples/grovers --target 64 (n = 7, 8 iterations): pre-measurement state and all 7
    outcomes identical to the oracle's under the same uniform stream."""
    pre_o, bits_o = rs.grover_full(rs.OracleBackend(), np.random.default_rng(99))
    for fusion in (1, 0):
        pre_g, bits_g = rs.grover_full(rs.EngineBackend(engine, fusion), np.random.default_rng(99))
        assert_close(pre_g, pre_o, "grover pre-measurement state")
        assert bits_g == bits_o


def _shor_slow_both(engine, number, a, seed):
    """find_period_slow (examples/shors.rs:180-234) run gate by gate on both backends."""
    import math
    from oracle import gates as og
    n = math.ceil(math.log2(number + 1))
    regs = [rs.OracleBackend().new([False] * (3 * n)), rs.EngineBackend(engine).new([False] * (3 * n))]
    gates = [og.hadamard_transform(list(range(2 * n))), og.not_(2 * n)]
    targets = list(range(2 * n, 3 * n))
    gates += [og.to_controlled(og.shor_u_gate(targets, number, a, i), i) for i in range(2 * n)]
    gates.append(og.qft(2 * n))
    for g in gates:
        for r in regs:
            r.apply(g)
    assert_close(regs[1].state(), regs[0].state(), f"shor N={number} pre-measurement")
    rng = np.random.default_rng(seed)
    res = [0, 0]
    for i in range(2 * n):
        u = rng.random()
        for j, r in enumerate(regs):
            res[j] |= int(r.measure(i, u)) << i
    assert res[0] == res[1]
    assert_close(regs[1].state(), regs[0].state(), f"shor N={number} collapsed")
    return og.limit_denominator(res[0], 2 ** (2 * n) - 1, number - 1)[1]


def test_config_c2_shor_15(engine):
    """BASELINE configs[1]: Shor N=15, a=7 -> 12 qubits (k=8 Hadamard transform, 8 controlled k=5 U gates, k=8 QFT)."""
    r = _shor_slow_both(engine, 15, 7, seed=1)
    assert 4 % r == 0


def test_config_c2_shor_21(engine):
    """BASELINE configs[1]: Shor N=21, a=2 -> 15 qubits (k=10 matrices of 16 MiB, 10 controlled k=6 U gates)."""
    r = _shor_slow_both(engine, 21, 2, seed=2)
    assert 6 % r == 0


def test_shor_fast_matches_oracle(engine):
    """find_period_fast (examples/shors.rs:287-361): gates interleaved with measurements of qubit 0."""
    for number, a in [(15, 7), (21, 2), (13, 5)]:
        ro = rs.find_period_fast(rs.OracleBackend(), np.random.default_rng(5), number, a)
        rg = rs.find_period_fast(rs.EngineBackend(engine), np.random.default_rng(5), number, a)
        assert ro == rg
