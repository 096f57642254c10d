"""Size-independent properties at the BASELINE's full sizes, where no oracle can follow (the reference itself stops at
30 qubits and a CPU restatement would need days):
  * n = 30 (two 17 GB registers): the fused path -- look-ahead planner, round search, register-blocked tile kernel, all
    of which only switch on for big registers -- against the one-kernel-per-gate path, compared on the device;
  * n = 33 (137.4 GB, the bench configuration): U followed by U^dagger returns |0...0>, the norm stays 1, and after
    measuring qubits of the final basis state the outcomes are the deterministic ones.
Skipped when the device cannot hold the registers."""
import gc

import numpy as np
import pytest

from tests import circuits

pytestmark = pytest.mark.gpu
OPT_FUSION = 1


def _register(engine, n):
    gc.collect()                 # registers of earlier tests that are still waiting for the collector
    try:
        return engine.Register.new([False] * n)
    except Exception as e:       # EngineError (out of memory): not a B200-sized device
        pytest.skip(f"cannot allocate a {n}-qubit register: {e}")


def _bench_like(n, layers):
    return circuits.qft_circuit(n) + circuits.random_circuit(n, layers, seed=0)


def test_fused_equals_unfused_at_30_qubits(engine):
    n = 30
    ops = [engine.Operation(m, t) for m, t in _bench_like(n, 6)]
    fused, plain = _register(engine, n), _register(engine, n)
    plain.set_option(OPT_FUSION, 0)
    fused.apply_batch(ops)
    plain.apply_batch(ops)
    assert fused.stat(2) * 8 < plain.stat(2), "the fused path should need far fewer sweeps"
    assert fused.max_abs_diff(plain) <= 1e-12
    assert abs(fused.norm2() - 1.0) <= 1e-10
    fused.close()
    plain.close()


def test_circuit_then_inverse_is_the_identity_at_33_qubits(engine):
    n = 33
    raw = _bench_like(n, 8)
    reg = _register(engine, n)
    reg.apply_batch([engine.Operation(m, t) for m, t in raw])
    assert abs(reg.norm2() - 1.0) <= 1e-10
    spread = np.abs(reg.read(0, 1 << 10)) ** 2
    assert spread.max() < 1e-6, "after the QFT and the layers the state is spread over all amplitudes"
    reg.apply_batch([engine.Operation(np.asarray(m).conj().T.copy(), t) for m, t in reversed(raw)])
    head = reg.read(0, 4)
    assert abs(head[0] - 1.0) <= 1e-10 and np.abs(head[1:]).max() <= 1e-10
    assert abs(reg.norm2() - 1.0) <= 1e-10          # with amp[0] = 1: everything else is 0
    for q in (0, 17, 32):
        assert reg.measure(q, u=0.999999) is False  # P(1) = 0 on |0...0>
    reg.close()
