"""GPU parity of the SPECIALISED tile kernels (K2j: qb_jit_gen.cu -> NVRTC -> cuModuleLoadData -> cuLaunchKernel) against the
oracle (try_apply_fast, /root/reference/src/register.rs:140-191 restated in oracle/).  QB_OPT_JIT = 2 makes every pass wait
for its kernel, so what is compared is the generated code, never the interpreter kernel; QB_STAT_JIT_LAUNCHES proves it."""
import numpy as np
import pytest

from oracle.oracle import OracleRegister
from tests import circuits
from tests.util import assert_close, random_state

pytestmark = pytest.mark.gpu

OPT_JIT, OPT_JIT_MIN, OPT_JIT_VARIANT = 6, 7, 8
STAT_PASSES, STAT_JIT_LAUNCHES, STAT_JIT_COMPILES, STAT_JIT_MISSES = 2, 12, 13, 14


def _run(engine, n, ops, seed, variant=2, jit=2, batch=True):
    s = random_state(n, seed)
    reg = engine.Register.new([False] * n)
    reg.set_option(OPT_JIT, jit)
    reg.set_option(OPT_JIT_MIN, 12)
    reg.set_option(OPT_JIT_VARIANT, variant)
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


@pytest.mark.parametrize("n,kind", [(12, "layers"), (14, "mixed"), (15, "qft"), (16, "layers"), (17, "mixed"), (18, "qft"), (20, "layers")])
@pytest.mark.parametrize("variant", [2, 3, 2 | 16, 2 | 16 | 32, 2 | 64])
def test_specialised_kernels_match_the_oracle(engine, n, kind, variant):
    ops = {"layers": circuits.random_circuit(n, 6, seed=n), "qft": circuits.qft_circuit(n), "mixed": circuits.mixed_circuit(n, 150, n)}[kind]
    reg, ora = _run(engine, n, ops, 3000 + n, variant)
    assert_close(reg.state, ora.state, f"jit variant {variant}, {kind} n={n}")
    assert reg.stat(STAT_JIT_LAUNCHES) >= 1, "the specialised kernels must be what ran"
    assert reg.stat(STAT_JIT_MISSES) == 0


def test_measurement_after_specialised_passes(engine):
    n = 16
    ops = circuits.random_circuit(n, 5, seed=4)
    reg, ora = _run(engine, n, ops, 11)
    rng = np.random.default_rng(5)
    for q in (3, 9, 15, 0):
        u = float(rng.random())
        assert reg.try_measure(q, u) == ora.measure(q, u)
    assert_close(reg.state, ora.state, "collapse after jit passes")


def test_cache_hits_on_the_second_application(engine):
    """The same circuit applied twice compiles once: the second application finds every kernel in the cache."""
    n = 14
    ops = circuits.random_circuit(n, 4, seed=9)
    reg, ora = _run(engine, n, ops, 21)
    reg.sync()
    c0, l0 = reg.stat(STAT_JIT_COMPILES), reg.stat(STAT_JIT_LAUNCHES)
    assert l0 >= 1
    reg.apply_batch([engine.Operation(m, t) for m, t in ops])
    for o in ops:
        ora.apply(o)
    assert_close(reg.state, ora.state, "second application")
    assert reg.stat(STAT_JIT_COMPILES) == c0
    assert reg.stat(STAT_JIT_LAUNCHES) == 2 * l0


def test_asynchronous_mode_falls_back_to_the_interpreter_then_hits(engine):
    """QB_OPT_JIT = 1: a pass never waits for the compiler -- the interpreter kernel runs it (a miss) -- and after
    qb_jit_drain() the same program is served by its specialised kernel.  Both must match the oracle."""
    from quaru_b200 import _ffi
    n = 15
    ops = circuits.mixed_circuit(n, 90, 77)
    reg, ora = _run(engine, n, ops, 31, jit=1)
    assert_close(reg.state, ora.state, "first application (interpreter or jit)")
    assert _ffi.lib().qb_jit_drain() == 0
    l0 = reg.stat(STAT_JIT_LAUNCHES)
    reg.apply_batch([engine.Operation(m, t) for m, t in ops])
    for o in ops:
        ora.apply(o)
    assert_close(reg.state, ora.state, "second application (jit)")
    assert reg.stat(STAT_JIT_LAUNCHES) > l0
