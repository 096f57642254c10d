"""GPU parity of the FUSED path (gate queue -> planner -> tile-pass kernel) against the oracle, on the
synthetic circuits of the BASELINE configs and on adversarial gate mixes, for several tile sizes."""
import numpy as np
import pytest

from oracle.oracle import OracleRegister
from tests import circuits
from tests.util import assert_close, random_state

pytestmark = pytest.mark.gpu

OPT_FUSION, OPT_TILE_BITS, OPT_MAX_QUEUE, OPT_FORCE = 1, 2, 3, 4


def _run(engine, n, ops, seed, tile_bits=12, force=0, batch=False, max_queue=None):
    s = random_state(n, seed)
    reg = engine.Register.new([False] * n)
    reg.set_option(OPT_TILE_BITS, tile_bits)
    if force:
        reg.set_option(OPT_FORCE, force)
    if max_queue:
        reg.set_option(OPT_MAX_QUEUE, max_queue)
    reg.set_state(s)
    ora = OracleRegister(s.copy(), n)
    eops = [engine.Operation(m, t) for m, t in ops]
    if batch:
        reg.apply_batch(eops)
    else:
        for o in eops:
            reg.apply(o)
    for o in ops:
        ora.apply(o)
    return reg, ora


@pytest.mark.parametrize("n,tile_bits", [(3, 12), (6, 8), (9, 8), (10, 12), (12, 9), (14, 11), (16, 12), (16, 13), (18, 10)])
def test_random_layers_fused(engine, n, tile_bits):
    ops = circuits.random_circuit(n, 6, seed=n)
    reg, ora = _run(engine, n, ops, 2000 + n, tile_bits)
    assert_close(reg.state, ora.state, f"n={n} T={tile_bits}")
    assert reg.stat(2) < len(ops), "fusion should need fewer passes than gates"


@pytest.mark.parametrize("n", [2, 5, 8, 12, 15])
def test_qft_fused(engine, n):
    ops = circuits.qft_circuit(n)
    reg, ora = _run(engine, n, ops, 2100 + n, batch=True)
    assert_close(reg.state, ora.state, f"qft n={n}")


@pytest.mark.parametrize("n,tile_bits,seed", [(1, 12, 0), (2, 12, 1), (4, 8, 2), (7, 8, 3), (9, 8, 4), (11, 9, 5), (13, 10, 6),
                                              (14, 12, 7), (16, 11, 8)])
def test_mixed_gates_fused(engine, n, tile_bits, seed):
    ops = circuits.mixed_circuit(n, 120, seed)
    reg, ora = _run(engine, n, ops, 2200 + n, tile_bits)
    assert_close(reg.state, ora.state, f"mixed n={n}")


def test_forced_tile_path_single_gates(engine):
    """Every gate as its own tile pass (force path 2): exercises high tile bits and outer controls."""
    n = 14
    ops = circuits.mixed_circuit(n, 60, 99)
    reg, ora = _run(engine, n, ops, 2300, tile_bits=8, force=2, max_queue=1)
    assert_close(reg.state, ora.state)


def test_fused_equals_unfused_and_measures(engine):
    n = 15
    ops = circuits.random_circuit(n, 8, seed=5)
    reg, ora = _run(engine, n, ops, 2400)
    rng = np.random.default_rng(1)
    for t in [0, 7, 14, 3]:
        u = float(rng.random())
        assert reg.measure(t, u=u) == ora.measure(t, u)
        more = circuits.random_layer(n, rng)
        for m, tg in more:
            reg.apply(engine.Operation(m, tg))
            ora.apply((m, tg))
    assert_close(reg.state, ora.state)


def test_large_fused_vs_oracle(engine):
    """n = 22: two layers of the C3 circuit, fused, full read-back against the oracle."""
    n = 22
    ops = circuits.random_circuit(n, 2, seed=22)
    reg, ora = _run(engine, n, ops, 2500, batch=True)
    assert_close(reg.state, ora.state)
