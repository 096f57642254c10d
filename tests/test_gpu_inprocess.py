"""In-process multi-GPU registers: `Register.new` (/root/reference/src/register.rs:41-56 keeps its signature) with QB_NUM_GPUS = P
in the environment shards the state over P GPUs of this box, driven by P worker threads of the calling process (qb_group.cu) --
no launcher, no ranks visible to the caller.  Same parity bar as everywhere: amplitudes against the CPU oracle within 1e-12 /
L2 1e-10, measured outcomes equal under a shared `u` stream, error payloads as in the reference.  Needs >= 2 GPUs."""
import numpy as np
import pytest

from oracle import gates as og
from oracle.oracle import OracleRegister
from tests import circuits
from tests.util import assert_close, random_state

pytestmark = pytest.mark.gpu


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _worlds():
    return [w for w in (2, 4, 8)]


@pytest.mark.parametrize("world", _worlds())
def test_stock_constructor_shards_in_process(engine, world, monkeypatch, tmp_path):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    monkeypatch.setenv("QB_NUM_GPUS", str(world))
    g = world.bit_length() - 1
    SWAPS, P2P = 6, 10
    for n, kind in ((3 * g + 1, "layers"), (12, "mixed"), (14, "qft"), (16, "layers"), (20, "layers")):
        ops = {"layers": circuits.random_circuit(n, 4, seed=n), "qft": circuits.qft_circuit(n),
               "mixed": [o for o in circuits.mixed_circuit(n, 100, seed=5 + n) if len(o[1]) <= n - g]}[kind]
        s = random_state(n, 400 + n)
        reg = engine.Register.new([False] * n)
        reg.set_state(s)                                  # every shard takes its slice of the caller's buffer
        ora = OracleRegister(s.copy(), n)
        for m, t in ops:
            reg.apply(engine.Operation(m, t))
            ora.apply((m, t))
        assert_close(reg.state, ora.state, f"in-process world={world} {kind} n={n}")
        rng = np.random.default_rng(n)
        for t in (n - 1, 0, n // 2):
            u = float(rng.random())
            assert reg.try_measure(t, u) == ora.measure(t, u)
        assert_close(reg.state, ora.state, f"after measurement, world={world} n={n}")
        assert abs(reg.norm2() - 1.0) < 1e-10
        if n >= 12 and kind != "qft":
            assert reg.stat(SWAPS) + reg.stat(11) > 0, "a dense gate on a global qubit needs an exchange (or a remote pass)"
        reg.close()
    # error vocabulary through the workers (register.rs:146-153), and raw state files (one per shard)
    n = 12
    reg = engine.Register.new([False] * n)
    with pytest.raises(engine.OperationError) as e:
        reg.try_apply(engine.Operation(og.hadamard(0)[0], [n]))
    assert e.value == engine.OperationError("InvalidTarget", n)
    for m, t in circuits.random_circuit(n, 3, seed=1):
        reg.apply(engine.Operation(m, t))
    want = reg.state
    reg.save_state(tmp_path / "s.qbstate")
    reg.reset(5)
    reg.load_state(tmp_path / "s.qbstate")
    assert np.array_equal(reg.state, want)
    other = engine.Register.new([False] * n)
    other.set_state(want)
    assert reg.max_abs_diff(other) == 0.0 and reg == other
    other.close()
    reg.close()


def test_small_registers_and_one_gpu_stay_single_device(engine, monkeypatch):
    """QB_NUM_GPUS is a request: registers with fewer than 2*log2(P)+1 qubits are built on one GPU (nothing to shard)."""
    monkeypatch.setenv("QB_NUM_GPUS", "8")
    reg = engine.Register.new([False] * 3)
    reg.apply(engine.operation.hadamard(2))
    assert abs(reg.probabilities(2)[0] - 0.5) < 1e-12
    reg.close()
