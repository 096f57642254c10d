"""The H / controlled-X / controlled-Z peephole (quaru_b200/csrc/qb_gate.cu: peephole_h_cx, QB_OPT_PEEPHOLE) is pure host logic:
qb_peephole_debug runs it on a gate list and returns the rewritten list in the engine's classified form.  Checked here with the
oracle (try_apply_fast, /root/reference/src/register.rs:140-191): the rewritten list is the same operator as the original one on
random states -- for lists built to hit every rule (H ; CX -> CZ ; H, CX ; H -> H ; CZ, H ; CX ; H -> CZ, H ; CZ ; H -> CX,
H ; H -> nothing, chains of CX on one target, multi-controlled X / Z), for the BASELINE layer circuits, and for random mixes --
and the rules fire where they should."""
import ctypes

import numpy as np
import pytest

from oracle import gates as og
from oracle.oracle import OracleRegister
from tests import circuits
from tests.util import assert_close, random_matrix, random_state


def _rewrite(n, ops):
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
    cap = 64 + sum(8 + 2 * mm.size for mm in keep[0::2]) * 2
    buf = (ctypes.c_double * cap)()
    got = L.qb_peephole_debug(n, arr, len(ops), buf, cap)
    assert got >= 1, _ffi.last_error()
    flat = np.array(buf[:got])
    rewrites, i, out = int(flat[0]), 1, []
    while i < len(flat):
        kind, k, cmask = int(flat[i]), int(flat[i + 1]), int(flat[i + 2])
        targets = [int(x) for x in flat[i + 3:i + 3 + k]]
        m = int(flat[i + 3 + k])
        ent = flat[i + 4 + k:i + 4 + k + 2 * m]
        ent = ent[0::2] + 1j * ent[1::2]
        i += 4 + k + 2 * m
        ctrls = [q for q in range(n) if (cmask >> q) & 1]
        red = np.diag(ent) if kind == 1 else ent.reshape(1 << k, 1 << k)
        dim = 1 << (k + len(ctrls))                      # full operator: identity unless every control bit (the high bits) is 1
        full = np.eye(dim, dtype=np.complex128)
        full[dim - (1 << k):, dim - (1 << k):] = red
        out.append((full, targets + ctrls))
    return rewrites, out


def _same_operator(n, ops, seed):
    rewrites, new = _rewrite(n, ops)
    s = random_state(n, seed)
    a, b = OracleRegister(s.copy(), n), OracleRegister(s.copy(), n)
    for o in ops:
        a.apply(o)
    for o in new:
        b.apply(o)
    assert_close(b.state, a.state, f"peephole changed the operator (n={n}, {len(ops)} -> {len(new)} gates)")
    return rewrites, new


def _count(new, pred):
    return sum(1 for m, t in new if pred(m, t))


def _is_h(m, t):
    return len(t) == 1 and np.allclose(m, og.hadamard(0)[0])


def test_every_rule_fires_and_keeps_the_operator():
    n = 5
    H, CX, CZ = og.hadamard, og.cnot, og.cz
    cases = {
        "H;CX -> CZ;H": ([H(1), CX(0, 1)], 1, 1),
        "CX;H -> H;CZ": ([CX(0, 1), H(1)], 1, 1),
        "H;CX;H -> CZ": ([H(1), CX(0, 1), H(1)], 2, 0),
        "H;CZ;H -> CX": ([H(1), CZ(0, 1), H(1)], 1, 0),
        "H;H -> nothing": ([H(2), og.pauli_z(3), H(2)], 1, 0),
        "chain on one target": ([H(1), CX(0, 1), CX(2, 1), CX(3, 1), H(1)], 4, 0),
        "gates in between that do not touch the target": ([H(1), og.rz(0.3, 0), CX(0, 1), H(3), CX(3, 0), og.rz(1.1, 2), H(1)], 2, 1),
        "toffoli": ([H(2), og.cnx([0, 1], 2), H(2)], 2, 0),
        "ccz": ([H(4), og.cnz([0, 1], 4), H(4)], 1, 0),
        "control side is left alone": ([H(0), CX(0, 1), H(0)], 0, 2),
        "not adjacent on the target": ([H(1), og.rz(0.4, 1), CX(0, 1)], 0, 1),
    }
    for name, (ops, want_rewrites, want_h) in cases.items():
        rewrites, new = _same_operator(n, ops, 7)
        assert rewrites == want_rewrites, (name, rewrites)
        assert _count(new, _is_h) == want_h, (name, [t for _, t in new])


@pytest.mark.parametrize("n,depth,seed", [(6, 12, 1), (9, 10, 2), (12, 8, 3)])
def test_layer_circuits_lose_hadamards_and_cnots(n, depth, seed):
    """The BASELINE layers (H or RZ per qubit, then CNOT / CZ on a random matching): about a quarter of the CNOTs sit between two
    Hadamards of their target, half of them next to one."""
    ops = circuits.random_circuit(n, depth, seed=seed)
    rewrites, new = _same_operator(n, ops, 100 + seed)
    h0, h1 = _count(ops, _is_h), _count(new, _is_h)
    assert rewrites > 0 and h1 < h0, (rewrites, h0, h1)


@pytest.mark.parametrize("n,count,seed", [(5, 120, 11), (8, 200, 12), (10, 150, 13)])
def test_random_mixes(n, count, seed):
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(count):
        kind = int(rng.integers(0, 9))
        qs = [int(x) for x in rng.permutation(n)]
        if kind <= 2: ops.append(og.hadamard(qs[0]))
        elif kind == 3: ops.append(og.cnot(qs[0], qs[1]))
        elif kind == 4: ops.append(og.cz(qs[0], qs[1]))
        elif kind == 5: ops.append(og.cnx(qs[:2], qs[2]))
        elif kind == 6: ops.append(og.cnz(qs[:2], qs[2]))
        elif kind == 7: ops.append(og.rz(float(rng.uniform(0, 6.28)), qs[0]))
        else: ops.append((random_matrix(2, int(rng.integers(1 << 30)), unitary=True), qs[:2]))
    _same_operator(n, ops, 200 + seed)


def test_a_hadamard_that_is_not_normalised_is_left_alone():
    n = 3
    h2 = (og.hadamard(1)[0] * 1.25, [1])
    rewrites, new = _same_operator(n, [h2, og.cnot(0, 1), h2], 5)
    assert rewrites == 0 and len(new) == 3
