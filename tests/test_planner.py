"""CPU tests of the host-side planner through the C ABI (qb_plan_debug needs no GPU): the pass assignment must
keep program order between any two gates that share a qubit, every gate must be scheduled exactly once, and
fusion must actually reduce the number of passes on the BASELINE circuits."""
import ctypes

import numpy as np
import pytest

from tests import circuits


def _plan(n, ops, tile_bits=12):
    from quaru_b200 import _ffi
    L = _ffi.lib()
    arr = (_ffi.QbGate * max(len(ops), 1))()
    keep = []
    for i, (m, t) in enumerate(ops):
        mm = np.ascontiguousarray(m, dtype=np.complex128)
        tt = (ctypes.c_uint32 * max(len(t), 1))(*t)
        keep += [mm, tt]
        arr[i].matrix = mm.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        arr[i].targets = tt
        arr[i].k = len(t)
    pass_of = (ctypes.c_int32 * max(len(ops), 1))()
    order = (ctypes.c_int32 * max(len(ops), 1))()
    npass = L.qb_plan_debug(n, arr, len(ops), tile_bits, pass_of, order)
    assert npass >= 0, _ffi.last_error()
    return npass, list(pass_of)[:len(ops)], list(order)[:len(ops)]


def _check_order(ops, pass_of, order):
    """A gate WRITES its dense target qubits and only READS its controls / diagonal qubits.  Program order must
    be kept between two gates whenever one writes a qubit the other touches (reads commute with reads)."""
    from tests.test_dist_cpu import _dense_targets
    last_write, last_touch = {}, {}
    for i, (m, t) in enumerate(ops):
        if pass_of[i] < 0:          # exact identity: dropped
            assert np.array_equal(np.asarray(m), np.eye(len(m)))
            continue
        key = (pass_of[i], order[i])
        writes = set(_dense_targets(m, t))
        for q in t:
            if q in last_write:
                assert last_write[q] < key, f"gate {i} on qubit {q} scheduled before an earlier write"
            if q in writes and q in last_touch:
                assert last_touch[q] < key, f"gate {i} writes qubit {q} before an earlier gate that touches it"
        for q in t:
            last_touch[q] = max(last_touch.get(q, key), key)
            if q in writes:
                last_write[q] = key
    keys = [(p, o) for p, o in zip(pass_of, order) if p >= 0]
    assert len(set(keys)) == len(keys), "two gates share a slot"


@pytest.mark.parametrize("n,tile_bits", [(3, 12), (8, 8), (12, 9), (20, 12), (28, 12), (33, 12), (33, 13), (36, 11)])
def test_plan_preserves_dependencies_random_layers(n, tile_bits):
    ops = circuits.random_circuit(n, 5, seed=n)
    npass, pass_of, order = _plan(n, ops, tile_bits)
    _check_order(ops, pass_of, order)
    if n >= 8:
        assert npass < len(ops) / 2


@pytest.mark.parametrize("n", [5, 12, 24, 33])
def test_plan_qft(n):
    ops = circuits.qft_circuit(n)
    npass, pass_of, order = _plan(n, ops)
    _check_order(ops, pass_of, order)
    assert npass <= 2 * n      # the reference would sweep the state len(ops) ~ n^2/2 times


@pytest.mark.parametrize("seed", range(6))
def test_plan_mixed(seed):
    n = 4 + 3 * seed
    ops = circuits.mixed_circuit(n, 200, seed)
    npass, pass_of, order = _plan(n, ops, 8 + seed % 5)
    _check_order(ops, pass_of, order)


def test_plan_rejects_invalid_targets():
    from quaru_b200 import _ffi
    with pytest.raises(AssertionError):
        _plan(3, [(np.eye(4), [0, 3])])
    assert _ffi.last_error()[1] == 3
    with pytest.raises(AssertionError):
        _plan(3, [(np.eye(4), [1, 1])])
    assert _ffi.last_error()[1] == 1


def test_identity_gates_are_dropped():
    npass, pass_of, order = _plan(4, [(np.eye(2), [1]), (np.eye(4), [0, 2])])
    assert npass == 0 and pass_of == [-1, -1]


@pytest.mark.parametrize("n,depth", [(28, 10), (31, 12), (33, 20)])
def test_lookahead_over_the_tile_bits_needs_fewer_sweeps(n, depth, monkeypatch):
    """Registers of >= 28 qubits choose the high tile bits of each pass with a look-ahead (most frequent upcoming dense
    targets, one-bit exchanges of the greedy set) instead of first come, first served: same ordering guarantees, fewer
    sweeps over the state on the BASELINE's random layers."""
    ops = circuits.random_circuit(n, depth, seed=n)
    monkeypatch.setenv("QB_PLAN_LOOKAHEAD", "0")
    greedy, pass_of, order = _plan(n, ops)
    _check_order(ops, pass_of, order)
    monkeypatch.delenv("QB_PLAN_LOOKAHEAD")
    ahead, pass_of, order = _plan(n, ops)
    _check_order(ops, pass_of, order)
    assert ahead < greedy, (ahead, greedy)
    monkeypatch.setenv("QB_PLAN_LOOKAHEAD", "1")          # forced on for a register below the automatic threshold
    small = circuits.mixed_circuit(16, 300, 4)
    _, pass_of, order = _plan(16, small)
    _check_order(small, pass_of, order)


@pytest.mark.parametrize("n,kind", [(14, "layers"), (15, "qft"), (13, "mixed"), (28, "layers")])
def test_beam_planner_keeps_the_ordering_guarantees(n, kind, monkeypatch):
    """The second-tier planner (beam search over the passes' high tile bits, qb_tile.cu; used when a gate list comes back) obeys
    the same commutation rule as the streaming planner: every pair of non-commuting gates keeps its program order."""
    monkeypatch.setenv("QB_PLAN_DEBUG_BEAM", "4")
    ops = {"layers": circuits.random_circuit(n, 6, seed=n), "qft": circuits.qft_circuit(n), "mixed": circuits.mixed_circuit(n, 200, n)}[kind]
    _, pass_of, order = _plan(n, ops)
    _check_order(ops, pass_of, order)


def test_beam_planner_needs_fewer_sweeps_on_the_baseline_circuit(monkeypatch):
    """BASELINE circuit at 33 qubits (QFT + 20 layers, SWAPs absorbed into the qubit names as the API layer does): the streaming
    planner needs 21 sweeps over the 137 GB state, the beam search 19."""
    import bench
    n = 33
    label, ops = list(range(n)), []
    for m, t in bench.build_circuit(n):
        if m.shape == (4, 4) and np.array_equal(m, bench._SWAP):
            label[t[0]], label[t[1]] = label[t[1]], label[t[0]]
            continue
        ops.append((m, [label[q] for q in t]))
    streaming, pass_of, order = _plan(n, ops)
    monkeypatch.setenv("QB_PLAN_DEBUG_BEAM", "4")
    beam, pass_of, order = _plan(n, ops)
    _check_order(ops, pass_of, order)
    assert beam < streaming and beam <= 19, (beam, streaming)
