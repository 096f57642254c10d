"""SWAP gates as qubit relabelling (registers of >= 28 local qubits; QB_SWAP_RELABEL=1 forces it on): the engine's host logic
(quaru_b200/csrc/qb_gate.cu: is_plain_swap / relabel_gate / canonical_swaps, the same functions the gate queue calls) in a dry
run through qb_relabel_debug, replayed in numpy against the oracle.  No GPU needed."""
import ctypes

import numpy as np
import pytest

from oracle import gates as og
from oracle import oracle_np
from oracle.oracle import OracleRegister
from tests import circuits
from tests.util import assert_close, random_state


def _relabel(n, ops):
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
    cap = 64 * (len(ops) + 2) + 4 * n
    out = (ctypes.c_int32 * cap)()
    cnt = L.qb_relabel_debug(n, arr, len(ops), out, cap)
    assert cnt >= 0, _ffi.last_error()
    flat, recs, i = list(out)[:cnt], [], 0
    while i < len(flat):
        if flat[i] == -3:
            m = flat[i + 1]
            pairs = [(flat[i + 2 + 2 * j], flat[i + 3 + 2 * j]) for j in range(m)]
            assert i + 2 + 2 * m == len(flat)
            return recs, pairs
        if flat[i] < 0:
            recs.append(flat[i])
            i += 1
        else:
            k = flat[i]
            recs.append(flat[i + 1:i + 1 + k])
            i += 1 + k
    raise AssertionError("no final record")


def _swap_heavy(n, count, seed):
    rng = np.random.default_rng(seed)
    ops = []
    for o in circuits.mixed_circuit(n, count, seed):
        ops.append(o)
        if rng.random() < 0.35:
            a, b = (int(x) for x in rng.permutation(n)[:2])
            ops.append(og.swap(a, b))
        if rng.random() < 0.1:
            a, b, c = (int(x) for x in rng.permutation(n)[:3])
            ops.append(og.to_controlled(og.swap(a, b), c))     # a controlled SWAP is a real gate, not a renaming
    return ops


@pytest.mark.parametrize("n,kind", [(6, "qft"), (9, "qft"), (8, "mixed"), (10, "mixed"), (7, "bench")])
def test_relabelled_circuit_replays_to_the_oracle(n, kind):
    if kind == "qft":
        ops = circuits.qft_circuit(n)
    elif kind == "bench":
        ops = circuits.qft_circuit(n) + circuits.random_circuit(n, 4, seed=1) + circuits.qft_circuit(n)
    else:
        ops = _swap_heavy(n, 60, seed=n)
    recs, pairs = _relabel(n, ops)
    assert len(recs) == len(ops)
    s0 = random_state(n, 5 * n)
    ora = OracleRegister(s0.copy(), n)
    st = s0.copy()
    absorbed = 0
    for (m, t), rec in zip(ops, recs):
        ora.apply((m, t))
        if rec == -1:
            absorbed += 1
            assert np.array_equal(np.asarray(m), np.asarray(og.swap(0, 1)[0])) and len(t) == 2
            continue
        if rec == -2:
            continue
        assert len(rec) == len(t) and len(set(rec)) == len(rec)
        st = oracle_np.apply_np(st, n, m, list(rec))
    assert absorbed == sum(1 for m, t in ops if len(t) == 2 and np.array_equal(np.asarray(m), np.asarray(og.swap(0, 1)[0])))
    assert absorbed >= n // 2
    for a, b in pairs:                       # what a raw read does first
        st = oracle_np.apply_np(st, n, og.swap(0, 1)[0], [a, b])
    assert len(pairs) <= n - 1
    assert_close(st, ora.state, f"relabelled replay n={n} {kind}")


def test_measurement_target_follows_the_labels():
    """After SWAP(0, 3) the caller's qubit 0 lives in engine qubit 3: a gate on the caller's qubit 0 is translated to 3."""
    n = 5
    ops = [og.swap(0, 3), og.hadamard(0), og.swap(3, 4), og.cnot(0, 4)]
    recs, pairs = _relabel(n, ops)
    assert recs[0] == -1 and recs[1] == [3] and recs[2] == -1
    # caller's 0 -> engine 3, caller's 3 -> engine 0, then SWAP(3, 4): caller's 3 -> engine 4, caller's 4 -> engine 0
    assert sorted(recs[3]) == sorted([3, 0])
    assert len(pairs) >= 1
