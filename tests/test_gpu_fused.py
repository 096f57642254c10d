"""GPU parity of the FUSED path (gate queue -> planner -> tile-pass kernel) against the oracle, on the
synthetic circuits of the BASELINE configs and on adversarial gate mixes, for several tile sizes."""
import os

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


# ---- the register-blocked kernel (2^12 tiles): rounds, thread-phase accumulation, special-cased gates ----------
@pytest.mark.parametrize("n,depth,seed", [(12, 10, 1), (13, 8, 2), (15, 8, 3), (17, 6, 4), (19, 4, 5)])
def test_register_blocked_random_layers(engine, n, depth, seed):
    ops = circuits.random_circuit(n, depth, seed=seed)
    reg, ora = _run(engine, n, ops, 3000 + n, tile_bits=12, batch=True)
    assert_close(reg.state, ora.state, f"v2 layers n={n}")


@pytest.mark.parametrize("n,seed", [(12, 11), (13, 12), (14, 13), (16, 14), (18, 15)])
def test_register_blocked_mixed(engine, n, seed):
    ops = circuits.mixed_circuit(n, 150, seed)
    reg, ora = _run(engine, n, ops, 3100 + n, tile_bits=12)
    assert_close(reg.state, ora.state, f"v2 mixed n={n}")


def test_register_blocked_single_gate_passes(engine):
    """Force the tile kernels even for single gates: every choice of register bits / thread bits / outer bits."""
    n = 15
    ops = circuits.mixed_circuit(n, 90, 21)
    reg, ora = _run(engine, n, ops, 3200, tile_bits=12, force=4, max_queue=1)
    assert_close(reg.state, ora.state)


def test_register_blocked_agrees_with_shared_memory_kernel(engine):
    n = 16
    ops = circuits.random_circuit(n, 6, seed=8) + circuits.qft_circuit(n)
    a, ora = _run(engine, n, ops, 3300, tile_bits=12, batch=True)            # register-blocked
    b, _ = _run(engine, n, ops, 3300, tile_bits=12, force=2, batch=True)     # shared-memory interpreter
    assert_close(a.state, ora.state)
    assert_close(b.state, ora.state)
    assert a.max_abs_diff(b) < 1e-13


@pytest.mark.parametrize("n", [12, 14, 17])
def test_register_blocked_qft_and_diagonals(engine, n):
    rng = np.random.default_rng(n)
    ops = circuits.qft_circuit(n)
    for _ in range(60):          # diagonal-heavy tail: RZ, CZ, controlled phases, 3-qubit diagonals on scrambled bits
        qs = [int(x) for x in rng.permutation(n)[:3]]
        d = np.exp(1j * rng.uniform(0, 6.28, size=8))
        ops.append((np.diag(d), qs))
        ops.append((np.diag(d[:4]), qs[:2]))
        ops.append((np.diag(d[:2]), qs[:1]))
        if rng.random() < 0.3:
            ops.append(((np.sqrt(0.5) * np.array([[1, 1], [1, -1]])).astype(complex), [qs[2]]))
    reg, ora = _run(engine, n, ops, 3400 + n, tile_bits=12, batch=True)
    assert_close(reg.state, ora.state, f"v2 qft+diag n={n}")


@pytest.mark.parametrize("n,kind", [(14, "layers"), (16, "mixed"), (17, "qft"), (20, "layers")])
def test_thorough_planning_forced_on_small_registers(engine, n, kind, monkeypatch):
    """Registers of >= 26 / 28 local qubits (the bench) plan with a look-ahead over the tile bits and search the rounds;
    QB_PLAN_LOOKAHEAD=1 forces the same host logic on at sizes whose state can be compared amplitude by amplitude."""
    monkeypatch.setenv("QB_PLAN_LOOKAHEAD", "1")
    ops = {"layers": circuits.random_circuit(n, 8, seed=n), "qft": circuits.qft_circuit(n), "mixed": circuits.mixed_circuit(n, 150, n)}[kind]
    reg, ora = _run(engine, n, ops, 4000 + n)
    assert_close(reg.state, ora.state, f"thorough planning n={n} {kind}")


@pytest.mark.parametrize("n,kind", [(16, "qft_layers"), (18, "layers"), (17, "mixed")])
def test_plan_cache_and_second_tier_planner_on_repeated_circuits(engine, n, kind, monkeypatch):
    """A gate list that comes back is launched from the plan cache; the second time it is seen the beam-search planner looks for
    a shorter plan (forced on at this size with QB_PLAN_BEAM_MIN_QUBITS).  Four applications on one register against the oracle:
    the state must match after every one, and the later applications never need more sweeps than the first."""
    monkeypatch.setenv("QB_PLAN_BEAM_MIN_QUBITS", "12")
    monkeypatch.setenv("QB_PLAN_LOOKAHEAD", "1")
    ops = {"qft_layers": circuits.qft_circuit(n) + circuits.random_circuit(n, 4, seed=n), "layers": circuits.random_circuit(n, 8, seed=n),
           "mixed": [o for o in circuits.mixed_circuit(n, 200, n) if len(o[1]) <= 2]}[kind]
    s = random_state(n, 4100 + n)
    reg = engine.Register.new([False] * n)
    reg.set_option(6, 2)      # specialised kernels, synchronous
    reg.set_option(7, 12)
    reg.set_state(s)
    ora = OracleRegister(s.copy(), n)
    eops = [engine.Operation(m, t) for m, t in ops]
    sweeps = []
    for rep in range(4):
        p0 = reg.stat(2)
        reg.apply_batch(eops)
        reg.sync()
        sweeps.append(reg.stat(2) - p0)
        for o in ops:
            ora.apply(o)
        assert_close(reg.state, ora.state, f"application {rep + 1}, {kind} n={n}")
    assert sweeps[2] <= sweeps[0] and sweeps[3] == sweeps[2], sweeps


@pytest.mark.parametrize("n,kind,max_queue", [(12, "layers", None), (14, "qft+layers", None), (16, "mixed", None), (15, "layers", 9), (18, "layers", None)])
def test_peephole_rewritten_lists_match_the_oracle(engine, n, kind, max_queue):
    """QB_OPT_PEEPHOLE (on by default only from 28 local qubits): every flushed gate list is rewritten with C^nX = H C^nZ H before
    it is planned (quaru_b200/csrc/qb_gate.cu, peephole_h_cx); the state and the measured outcomes still are the oracle's
    (try_apply_fast / try_measure, /root/reference/src/register.rs:140-191, 297-342) -- also when the queue overflows in the
    middle of the gate stream and held gates meet new ones."""
    ops = {"layers": lambda: circuits.random_circuit(n, 10, seed=40 + n), "mixed": lambda: circuits.mixed_circuit(n, 150, 50 + n),
           "qft+layers": lambda: circuits.qft_circuit(n) + circuits.random_circuit(n, 4, seed=n)}[kind]()
    s = random_state(n, 4100 + n)
    reg = engine.Register.new([False] * n)
    reg.set_option(9, 1)                       # QB_OPT_PEEPHOLE
    if max_queue:
        reg.set_option(OPT_MAX_QUEUE, max_queue)
    reg.set_state(s)
    ora = OracleRegister(s.copy(), n)
    eops = [engine.Operation(m, t) for m, t in ops]
    if max_queue:
        for o in eops:
            reg.apply(o)
    else:
        reg.apply_batch(eops)
    for o in ops:
        ora.apply(o)
    assert_close(reg.state, ora.state, f"peephole on, {kind} n={n}")
    if kind != "mixed":
        assert reg.stat(19) > 0, "the layer circuits have Hadamards next to CNOT targets"      # QB_STAT_PEEPHOLE_REWRITES
    rng = np.random.default_rng(n)
    for q in (0, n - 1, n // 2):
        u = float(rng.random())
        assert reg.try_measure(q, u) == ora.measure(q, u)
    assert_close(reg.state, ora.state, "collapse after peephole-rewritten passes")
