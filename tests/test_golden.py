"""Golden fixtures (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py from the pinned CPU oracle).
CPU: the oracle reproduces them bit for bit (guards the oracle build against drift).  GPU: the engine matches them
within the north_star tolerance and reproduces every recorded measurement outcome."""
import math
import os

import numpy as np
import pytest

from oracle import gates as og
from oracle.oracle import OracleRegister
from tests import circuits
from tests import reference_suite as rs
from tests.util import assert_close

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
U_SEED = 20261017


def _grover(backend):
    return rs.grover_full(backend, np.random.default_rng(U_SEED))


def _shor15(new_reg, apply, measure):
    number, a = 15, 7
    n = math.ceil(math.log2(number + 1))
    reg = new_reg(3 * n)
    apply(reg, og.hadamard_transform(list(range(2 * n))))
    apply(reg, og.not_(2 * n))
    for i in range(2 * n):
        apply(reg, og.to_controlled(og.shor_u_gate(list(range(2 * n, 3 * n)), number, a, i), i))
    apply(reg, og.qft(2 * n))
    return reg


def test_oracle_reproduces_golden_bit_for_bit():
    pre, bits = _grover(rs.OracleBackend())
    assert (pre == G["grover_t64_state"]).all() and list(map(int, bits)) == list(G["grover_t64_bits"])
    reg = _shor15(lambda n: OracleRegister.new([False] * n), lambda r, g: r.apply(g), None)
    assert (reg.state == G["shor15_a7_state"]).all()
    outs = [reg.measure(i, float(u)) for i, u in enumerate(G["shor15_a7_us"])]
    assert list(map(int, outs)) == list(G["shor15_a7_outcomes"])
    assert int(G["shor15_a7_period"][0]) in (1, 2, 4)
    reg = OracleRegister.new([False] * 10)
    for g in circuits.random_circuit(10, 6, seed=10):
        reg.apply(g)
    assert (reg.state == G["layers_n10_state"]).all()
    reg = OracleRegister.from_int(8, 37)
    for g in circuits.qft_circuit(8):
        reg.apply(g)
    assert (reg.state == G["qft_n8_state"]).all()
    assert abs(np.abs(G["qft_n8_state"]).max() - 1 / 16) < 1e-12        # a basis state's QFT is flat


@pytest.mark.gpu
@pytest.mark.parametrize("fusion", [1, 0])
def test_engine_matches_golden(engine, fusion):
    B = rs.EngineBackend(engine, fusion)
    pre, bits = _grover(B)
    assert_close(pre, G["grover_t64_state"], "grover")
    assert list(map(int, bits)) == list(G["grover_t64_bits"])
    reg = _shor15(lambda n: B.new([False] * n), lambda r, g: r.apply(g), None)
    assert_close(reg.state(), G["shor15_a7_state"], "shor15")
    outs = [reg.measure(i, float(u)) for i, u in enumerate(G["shor15_a7_us"])]
    assert list(map(int, outs)) == list(G["shor15_a7_outcomes"])
    reg = B.new([False] * 10)
    for g in circuits.random_circuit(10, 6, seed=10):
        reg.apply(g)
    assert_close(reg.state(), G["layers_n10_state"], "layers")
    outs = [reg.measure(t, float(u)) for t, u in zip([9, 0, 4, 7], G["layers_n10_us"])]
    assert list(map(int, outs)) == list(G["layers_n10_outcomes"])
    assert_close(reg.state(), G["layers_n10_collapsed"], "layers collapsed")
    reg = B.from_int(8, 37)
    for g in circuits.qft_circuit(8):
        reg.apply(g)
    assert_close(reg.state(), G["qft_n8_state"], "qft")
