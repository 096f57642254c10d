"""CPU-only fuzzers for the engine's HOST logic (no GPU needed; the oracle is the checker).

    python tools/fuzz_host.py rounds  SECONDS [SEED0]   pass planner + round encoder, replayed by tests/rounds_emulator.py
    python tools/fuzz_host.py shard   SECONDS [SEED0]   exchange schedule of the sharded path (qb_shard_plan_debug), all ranks
    python tools/fuzz_host.py shard-reorder SECONDS [SEED0]   the same for the commutation-aware scheduler (the default since the end of round 1)
    python tools/fuzz_host.py qasm    COUNT   [SEED0]   mutated OpenQASM programs: the parser must return, never crash

Round 1: 2700 `rounds`, 26000 `shard` and 160000 `qasm` cases without a finding after the fixes they prompted
(expression recursion guards, expansion cap, EOF inside a gate-body barrier)."""
import ctypes
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from oracle import gates as og  # noqa: E402
from oracle.oracle import OracleRegister  # noqa: E402
from tests.util import random_state  # noqa: E402


def rand_gate(n, rng):
    """One gate of every structural kind the classifier distinguishes, on random qubits of an n-qubit register (n >= 3)."""
    kind = rng.integers(0, 12)
    qs = [int(x) for x in rng.permutation(n)]
    th = float(rng.uniform(0, 6.28))
    if kind == 0:
        return og.hadamard(qs[0])
    if kind == 1:
        return og.rz(th, qs[0])
    if kind == 2:
        return og.cnot(qs[0], qs[1])
    if kind == 3:
        return og.cz(qs[0], qs[1])
    if kind == 4:
        return og.cphase(th, qs[0], qs[1])
    if kind == 5:
        return og.swap(qs[0], qs[1])
    if kind == 6:                                                        # random unitary with 0..3 controls
        g = (np.linalg.qr(rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)))[0], [qs[0]])
        for c in qs[1:1 + int(rng.integers(0, 4))]:
            g = og.to_controlled(g, c)
        return g
    if kind == 7:                                                        # wide diagonal
        k = int(rng.integers(1, 4))
        return (np.diag(np.exp(1j * rng.uniform(0, 6.28, size=1 << k))), qs[:k])
    if kind == 8:                                                        # dense two-qubit unitary
        return (np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0], qs[:2])
    if kind == 9:                                                        # real rotation
        c, s = np.cos(th / 2), np.sin(th / 2)
        return (np.array([[c, -s], [s, c]], dtype=complex), [qs[0]])
    if kind == 10:
        return og.cnx(qs[:2], qs[2])
    g = (np.diag(np.exp(1j * rng.uniform(0, 6.28, size=2))), [qs[0]])    # general one-qubit diagonal, maybe controlled
    for c in qs[1:1 + int(rng.integers(0, 3))]:
        g = og.to_controlled(g, c)
    return g


def fuzz_rounds(budget, seed0):
    from tests import rounds_emulator as em
    t0, it, worst = time.time(), 0, 0.0
    while time.time() - t0 < budget:
        seed = seed0 + it
        it += 1
        rng = np.random.default_rng(seed)
        n = int(rng.integers(12, 17))
        sub = n if rng.random() < 0.5 else int(rng.integers(3, n + 1))    # few qubits: long same-qubit chains
        perm = [int(x) for x in rng.permutation(n)]
        ops = []
        for _ in range(int(rng.integers(5, 300))):
            m, t = rand_gate(sub, rng)
            ops.append((m, [perm[q] for q in t]))
        s0 = random_state(n, seed)
        ora = OracleRegister(s0.copy(), n)
        for o in ops:
            ora.apply(o)
        st = s0.copy()
        for p in em.dump_programs(n, ops):
            if p[0] == "gates":
                tmp = OracleRegister(st, n)
                for tag in p[1]:
                    tmp.apply(ops[tag])
                st = tmp.state
            else:
                st = em.emulate_program(st, n, p[2], p[3])
        err = float(np.max(np.abs(st - ora.state)))
        worst = max(worst, err)
        if err > 1e-11:
            print("FAIL seed", seed, "n", n, "gates", len(ops), "err", err)
    print("rounds: cases", it, "worst error", worst)


def fuzz_shard(budget, seed0, reorder=False):
    from tests.test_dist_cpu import _plan_events, _plan_events_reordered, _replay, _replay_reordered
    t0, it, worst = time.time(), 0, 0.0
    while time.time() - t0 < budget:
        seed = seed0 + it
        it += 1
        rng = np.random.default_rng(seed)
        world = int(rng.choice([2, 4, 8]))
        g = int(np.log2(world))
        n = int(rng.integers(max(2 * g + 1, 4), 13))
        ops = []
        for _ in range(int(rng.integers(1, 120))):
            m, t = rand_gate(n, rng)
            if len(t) <= n - g:
                ops.append((m, t))
        try:
            per_rank = [(_plan_events_reordered if reorder else _plan_events)(n, world, r, ops) for r in range(world)]
            for ev, ph in per_rank[1:]:
                assert ev == per_rank[0][0] and ph == per_rank[0][1], "ranks disagree on the schedule"
            got, want, phys = (_replay_reordered if reorder else _replay)(n, world, ops, per_rank[0][0], seed)
            assert phys == per_rank[0][1], "final placement differs from the replay"
            err = float(np.max(np.abs(got - want)))
            worst = max(worst, err)
            assert err <= 1e-11, f"amplitude error {err}"
        except AssertionError as e:
            print("FAIL seed", seed, "n", n, "world", world, "gates", len(ops), str(e)[:200])
    print("shard-reorder:" if reorder else "shard:", "cases", it, "worst error", worst)


def fuzz_qasm(count, seed0):
    from quaru_b200 import _ffi
    from tests import qasm_programs as qp
    L = _ffi.lib()
    seeds = [v for v in vars(qp).values() if isinstance(v, str) and "OPENQASM" in v]
    toks = ["(", ")", "{", "}", "[", "]", ";", ",", "->", "==", "pi", "gate", "qreg", "creg", "if", "measure", "reset", "barrier",
            "U", "CX", "opaque", "include", '"qelib1.inc"', "-", "+", "*", "/", "^", "sin", "cos", "sqrt", "ln", "exp", "tan",
            "1e308", "0", "99999999999", "q", "c", "//", "\n", "1.5.2", "a", "h"]
    rnd = random.Random(seed0)
    ok = 0
    for _ in range(count):
        s = rnd.choice(seeds)
        for _ in range(rnd.randint(1, 6)):
            m, i = rnd.randint(0, 5), rnd.randint(0, len(s))
            if m == 0:
                s = s[:i] + rnd.choice(toks) + s[i:]
            elif m == 1:
                s = s[:i] + s[min(len(s), i + rnd.randint(1, 12)):]
            elif m == 2:
                j = min(len(s), i + rnd.randint(1, 30))
                s = s[:i] + s[i:j] * 2 + s[j:]
            elif m == 3:
                s = s[:i] + chr(rnd.randint(1, 255)) + s[i + 1:]
            elif m == 4:
                s = s[:i]
            else:
                j = rnd.randint(0, len(s))
                a, b = min(i, j), max(i, j)
                s = s[:a] + s[b:] + s[a:b]
        h = ctypes.c_void_p()
        if L.qb_qasm_parse(s.encode("latin-1", "replace"), ctypes.byref(h)) == 0:
            ok += 1
            L.qb_qasm_free(h)
    print("qasm: cases", count, "accepted", ok, "rejected", count - ok, "(no crash)")


if __name__ == "__main__":
    mode, amount = sys.argv[1], float(sys.argv[2])
    seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    {"rounds": fuzz_rounds, "shard": fuzz_shard, "shard-reorder": lambda a, s: fuzz_shard(a, s, True),
     "qasm": lambda a, s: fuzz_qasm(int(a), s)}[mode](amount, seed0)
