"""Edge cases of the boundary on the GPU: empty register, concurrent handles from several threads, queue limits,
large control lists, options and stats."""
import threading

import numpy as np
import pytest

from oracle import gates as og
from oracle.oracle import OracleRegister
from tests import circuits
from tests.util import assert_close, random_state

pytestmark = pytest.mark.gpu


def test_zero_qubit_register(engine):
    """Register::new(&[]) is the 1x1 state [1] (register.rs:43-50 folds over nothing)."""
    reg = engine.Register.new([])
    assert reg.size() == 0
    assert (reg.state == np.array([1.0 + 0j])).all()
    assert reg.norm2() == 1.0
    with pytest.raises(engine.OperationError):
        reg.try_measure(0)
    with pytest.raises(engine.OperationError):
        reg.try_apply(engine.operation.hadamard(0))


def test_handles_are_independent_across_threads(engine):
    """Different handles may be used concurrently (one thread per handle, `&mut self` in the reference)."""
    n, results, errors = 14, {}, []

    def work(seed):
        try:
            ops = circuits.random_circuit(n, 6, seed=seed) + circuits.mixed_circuit(n, 40, seed)
            s = random_state(n, seed)
            reg = engine.Register.new([False] * n)
            reg.set_state(s)
            for m, t in ops:
                reg.apply(engine.Operation(m, t))
            results[seed] = (reg.state, s, ops)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(sd,)) for sd in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for seed, (got, s, ops) in results.items():
        ora = OracleRegister(s.copy(), n)
        for o in ops:
            ora.apply(o)
        assert_close(got, ora.state, f"thread {seed}")


def test_queue_autoflush_and_stats(engine):
    n = 13
    ops = circuits.random_circuit(n, 10, seed=3)
    s = random_state(n, 9)
    reg = engine.Register.new([False] * n)
    reg.set_option(3, 7)                       # QB_OPT_MAX_QUEUE: flush every 7 gates
    reg.set_state(s)
    ora = OracleRegister(s.copy(), n)
    for m, t in ops:
        reg.apply(engine.Operation(m, t))
        ora.apply((m, t))
    assert_close(reg.state, ora.state)
    assert reg.stat(3) == len(ops)             # QB_STAT_GATES
    assert 0 < reg.stat(2) <= len(ops)         # QB_STAT_PASSES
    assert reg.stat(1) >= reg.stat(2)          # QB_STAT_KERNEL_LAUNCHES
    assert reg.stat(4) > 0                     # QB_STAT_BYTES_MOVED
    with pytest.raises(IndexError):
        reg.set_option(2, 99)                  # tile bits out of range
    with pytest.raises(IndexError):
        reg.set_option(77, 1)


def test_reset_and_reuse(engine):
    from quaru_b200 import _ffi
    reg = engine.Register.new([True, False, True, True])
    reg.apply(engine.operation.hadamard(1))
    assert _ffi.lib().qb_reset(reg._h, 5) == 0
    assert (reg.state == OracleRegister.from_int(4, 5).state).all()
    assert _ffi.lib().qb_reset(reg._h, 16) == _ffi.QB_ERR_INVALID_ARGUMENT


@pytest.mark.parametrize("n", [12, 16, 20])
def test_many_controls_through_every_path(engine, n):
    """A Z and an X controlled by every other qubit, fused (tile kernels) and unfused (streaming kernels)."""
    for fusion in (0, 1):
        s = random_state(n, n)
        reg = engine.Register.new([False] * n)
        reg.set_option(1, fusion)
        reg.set_state(s)
        want = s.copy()
        reg.apply_controlled(engine.operation.pauli_z(n - 1), list(range(n - 1)))
        want[-1] = -want[-1]
        reg.apply(engine.operation.hadamard(3))
        reg.apply_controlled(engine.operation.not_(0), list(range(1, n)))
        ora = OracleRegister(want, n)
        ora.apply(og.hadamard(3))
        st = ora.state
        st[[-2, -1]] = st[[-1, -2]]
        assert_close(reg.state, st, f"n={n} fusion={fusion}")


def test_clone_of_a_large_register_is_ordered_after_its_initialisation(engine):
    """qb_clone (#[derive(Clone)], /root/reference/src/register.rs:23) at a size where the new handle's memset takes far longer
    than a launch: the copy must not race with it (the two handles own different streams)."""
    n = 27
    reg = engine.Register.new([False] * n)
    for q in (0, 5, 13, 20, 26):
        reg.apply(engine.operation.hadamard(q))
    for _ in range(3):
        c = reg.clone()
        assert reg.max_abs_diff(c) == 0.0
        assert abs(c.norm2() - 1.0) < 1e-12
        assert abs(c.read(0, 1)[0] - reg.read(0, 1)[0]) == 0.0
        c.close()
