"""SWAP gates as qubit relabelling, forced on (QB_SWAP_RELABEL=1) at sizes the oracle can follow: big registers (>= 28
local qubits, the bench) have it on by default.  Every entry point that looks at qubits or at raw amplitudes is exercised
after SWAPs have permuted the names: gates, controlled gates, measurement, probabilities, read, write, clone, diff, reset,
save / load.  The host logic itself is covered on CPU by tests/test_relabel.py."""
import numpy as np
import pytest

from oracle import gates as og
from oracle.oracle import OracleRegister
from tests import circuits
from tests.util import assert_close, random_state

pytestmark = pytest.mark.gpu


@pytest.fixture()
def relabel(monkeypatch):
    monkeypatch.setenv("QB_SWAP_RELABEL", "1")      # read when a register is created


def _swap_heavy(n, count, seed):
    rng = np.random.default_rng(seed)
    ops = []
    for o in circuits.mixed_circuit(n, count, seed):
        if len(o[1]) > 3:
            continue
        ops.append(o)
        if rng.random() < 0.35:
            a, b = (int(x) for x in rng.permutation(n)[:2])
            ops.append(og.swap(a, b))
    return ops


@pytest.mark.parametrize("n,kind,fusion", [(12, "qft", 1), (14, "swaps", 1), (16, "bench", 1), (13, "swaps", 0), (20, "qft", 1)])
def test_state_and_outcomes_match_the_oracle(engine, relabel, n, kind, fusion):
    ops = {"qft": circuits.qft_circuit(n), "swaps": _swap_heavy(n, 80, n),
           "bench": circuits.qft_circuit(n) + circuits.random_circuit(n, 4, seed=n) + circuits.qft_circuit(n)}[kind]
    s0 = random_state(n, 300 + n)
    reg = engine.Register.new([False] * n)
    reg.set_option(1, fusion)
    reg.set_state(s0)
    ora = OracleRegister(s0.copy(), n)
    for m, t in ops:
        reg.apply(engine.Operation(m, t))
        ora.apply((m, t))
    for q in (0, n - 1, n // 2):                               # qubit names are the caller's, whatever the engine did
        got, want = reg.probabilities(q), ora.probabilities(q)
        assert abs(got[0] - want[0]) < 1e-12 and abs(got[1] - want[1]) < 1e-12
    assert_close(reg.state, ora.state, f"relabel {kind} n={n} fusion={fusion}")   # raw read: permutation undone first
    rng = np.random.default_rng(5)
    more = circuits.random_circuit(n, 2, seed=3) + [og.swap(0, n - 1), og.swap(1, n // 2)]
    for m, t in more:
        reg.apply(engine.Operation(m, t))
        ora.apply((m, t))
    for q in (n - 1, 1, 0):
        u = float(rng.random())
        assert reg.measure(q, u=u) == ora.measure(q, u)
    assert_close(reg.state, ora.state, f"relabel {kind} n={n}: after measurements")
    reg.close()


def test_swaps_cost_no_sweep(engine, relabel):
    n = 16
    reg = engine.Register.new([False] * n)
    reg.apply(engine.Operation(*og.hadamard(3)))
    reg.flush()
    before = reg.stat(2)
    for a in range(n // 2):
        reg.apply(engine.Operation(*og.swap(a, n - 1 - a)))
    reg.flush()
    assert reg.stat(2) == before, "a SWAP only renames two qubits"
    p = reg.probabilities(n - 1 - 3)                           # the Hadamard's qubit now answers to the name n-1-3
    assert abs(p[0] - 0.5) < 1e-12 and abs(p[1] - 0.5) < 1e-12
    p = reg.probabilities(3)
    assert abs(p[0] - 1.0) < 1e-12 and abs(p[1]) < 1e-12
    st = reg.state
    want = np.zeros(1 << n, dtype=np.complex128)
    want[0] = want[1 << (n - 1 - 3)] = np.sqrt(0.5)
    assert_close(st, want, "state after renaming")
    reg.close()


def test_clone_diff_reset_write_and_files_see_callers_qubits(engine, relabel, tmp_path):
    n = 12
    ops = circuits.qft_circuit(n) + circuits.random_circuit(n, 2, seed=9)      # ends with the names permuted
    ora = OracleRegister.new([False] * n)
    a = engine.Register.new([False] * n)
    for m, t in ops:
        a.apply(engine.Operation(m, t))
        ora.apply((m, t))
    b = a.clone()                                              # same amplitudes, same names
    extra = [og.hadamard(0), og.cnot(0, n - 1), og.swap(2, 5), og.rz(0.3, 5)]
    for m, t in extra:
        a.apply(engine.Operation(m, t))
        b.apply(engine.Operation(m, t))
        ora.apply((m, t))
    assert a.max_abs_diff(b) == 0.0
    assert_close(b.state, ora.state, "clone after more gates")
    path = tmp_path / "r.qbstate"
    a.save_state(path)
    c = engine.Register.new([False] * n)
    c.apply(engine.Operation(*og.swap(0, 1)))                  # names permuted before the load: the file holds the identity
    c.load_state(path)
    assert_close(c.state, ora.state, "loaded image")
    c.apply(engine.Operation(*og.hadamard(0)))
    ora.apply(og.hadamard(0))
    assert_close(c.state, ora.state, "gate after load")
    a.reset(5)
    assert a.probabilities(0) == (0.0, 1.0) and a.probabilities(2) == (0.0, 1.0) and a.probabilities(1) == (1.0, 0.0)   # exact: one amplitude
    s1 = random_state(n, 77)
    b.apply(engine.Operation(*og.swap(3, 9)))
    b.set_state(s1)                                            # raw write: index bit q is the caller's qubit q again
    ora2 = OracleRegister(s1.copy(), n)
    b.apply(engine.Operation(*og.hadamard(3)))
    ora2.apply(og.hadamard(3))
    assert_close(b.state, ora2.state, "write after renaming")
    for r in (a, b, c):
        r.close()
