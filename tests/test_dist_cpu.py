"""CPU tests of the multi-GPU (sharded) path's host logic, world_size 2 and 4 over gloo.

The exchange schedule (which local bit positions are traded against which rank bits, and when) is computed
by the same C++ code the GPU path runs (qb_shard_plan_debug = shard_run in dry mode).  Checked here:
  * every rank derives the SAME schedule (otherwise the NCCL send/recv pairs would not match up),
  * replaying the schedule on a full state vector in physical layout (numpy) reproduces the oracle, and every
    dense target is local at the moment its gate runs (controls and diagonal gates may stay global).
"""
import ctypes
import os
import socket

import numpy as np
import pytest

from oracle import oracle_np
from oracle.oracle import OracleRegister
from tests import circuits
from tests.util import assert_close, random_state


def _plan_events(n, world, rank, ops):
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
    cap = 64 * (len(ops) + 1)
    ev = (ctypes.c_int32 * cap)()
    phys = (ctypes.c_uint32 * n)()
    os.environ["QB_SHARD_REORDER"] = "0"          # the first scheduler: exchange records only
    try:
        cnt = L.qb_shard_plan_debug(n, world, rank, arr, len(ops), ev, cap, phys)
    finally:
        del os.environ["QB_SHARD_REORDER"]
    assert cnt >= 0, _ffi.last_error()
    flat = list(ev)[:cnt]
    events, i = [], 0
    while i < len(flat):
        gi, gp = flat[i], flat[i + 1]
        events.append((gi, flat[i + 2:i + 2 + gp], flat[i + 2 + gp:i + 2 + 2 * gp]))
        i += 2 + 2 * gp
    return events, list(phys)


def _dense_targets(m, t):
    """Targets that are neither controls nor covered by a diagonal (numpy restatement of classify_gate)."""
    m = np.asarray(m)
    k = len(t)
    dim = 1 << k
    eye = np.eye(dim)
    ident = [bool((m[r] == eye[r]).all() and (m[:, r] == eye[:, r]).all()) for r in range(dim)]
    ctrl = [j for j in range(k) if all(ident[r] for r in range(dim) if not (r >> j) & 1)]
    keep = [j for j in range(k) if j not in ctrl]
    cb = sum(1 << j for j in ctrl)
    sub = [cb | sum(((x >> i) & 1) << keep[i] for i in range(len(keep))) for x in range(1 << len(keep))]
    red = m[np.ix_(sub, sub)]
    if (red == np.diag(np.diag(red))).all():
        return []
    return [t[j] for j in keep]


def _replay(n, world, ops, events, seed):
    g = int(np.log2(world))
    nl = n - g
    s0 = random_state(n, seed)
    ora = OracleRegister(s0.copy(), n)
    phys = list(range(n))          # logical -> physical
    st = s0.copy()                 # physical layout (identity at start)
    by_gate = {}
    for gi, V, J in events:
        by_gate.setdefault(gi, []).append((V, J))
    for i, (m, t) in enumerate(ops):
        for V, J in by_gate.get(i, []):
            for v, j in zip(V, J):
                assert 0 <= v < nl and 0 <= j < g
                a = st.reshape((2,) * n)
                st = np.ascontiguousarray(np.swapaxes(a, n - 1 - v, n - 1 - (nl + j))).reshape(-1)
                ql, qg = phys.index(v), phys.index(nl + j)
                phys[ql], phys[qg] = nl + j, v
        for q in _dense_targets(m, t):
            assert phys[q] < nl, f"gate {i}: dense target {q} is on a rank bit when the gate runs"
        st = oracle_np.apply_np(st, n, m, [phys[q] for q in t])
        ora.apply((m, t))
    # back to logical order: logical bit q lives at physical bit phys[q]
    a = st.reshape((2,) * n)
    a = np.transpose(a, [n - 1 - phys[n - 1 - ax] for ax in range(n)])
    return np.ascontiguousarray(a).reshape(-1), ora.state, phys


@pytest.mark.parametrize("n,world,kind", [(6, 2, "layers"), (8, 4, "layers"), (9, 8, "layers"), (7, 2, "qft"), (9, 4, "qft"),
                                          (8, 2, "mixed"), (10, 4, "mixed"), (12, 8, "mixed")])
def test_schedule_replay_matches_oracle(n, world, kind):
    if kind == "layers":
        ops = circuits.random_circuit(n, 6, seed=n + world)
    elif kind == "qft":
        ops = circuits.qft_circuit(n)
    else:
        ops = [o for o in circuits.mixed_circuit(n, 80, seed=n) if len(o[1]) <= n - int(np.log2(world))]
    per_rank = [_plan_events(n, world, r, ops) for r in range(world)]
    for ev, ph in per_rank[1:]:
        assert ev == per_rank[0][0] and ph == per_rank[0][1], "ranks disagree on the exchange schedule"
    events, final_phys = per_rank[0]
    got, want, phys = _replay(n, world, ops, events, seed=n * 31 + world)
    assert phys == final_phys
    assert_close(got, want, f"replay n={n} P={world} {kind}")
    if kind == "layers":
        assert len(events) <= 6 + 1, "at most about one exchange per layer is expected"


@pytest.mark.parametrize("world", [4, 8])
def test_controlled_gate_between_rank_bits_keeps_ranks_in_step(world):
    """A CNOT whose control AND target sit on rank bits: ranks where the control bit is 0 have nothing to do,
    but the exchange is a collective, so every rank must schedule it (a hang on real GPUs otherwise)."""
    from oracle import gates as og
    n = 8
    g = int(np.log2(world))
    ops = [og.cnot(n - 1, n - 2), og.hadamard(0), og.to_controlled(og.hadamard(n - 1), n - 2), og.cz(n - 1, n - 2),
           og.cnx([n - 1, 0], n - 2), og.swap(n - 1, n - g)]
    per_rank = [_plan_events(n, world, r, ops) for r in range(world)]
    assert len(per_rank[0][0]) >= 1
    for ev, ph in per_rank[1:]:
        assert ev == per_rank[0][0] and ph == per_rank[0][1], "ranks disagree on the exchange schedule"
    got, want, _ = _replay(n, world, ops, per_rank[0][0], seed=77)
    assert_close(got, want)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ops = circuits.random_circuit(n, 5, seed=3) + circuits.qft_circuit(n)
        events, phys = _plan_events(n, world, rank, ops)                        # the first scheduler
        events2, phys2 = _plan_events_reordered(n, world, rank, ops)            # the default one
        gathered = [None] * world
        dist.all_gather_object(gathered, (events, phys, events2, phys2))
        ok = all(x == gathered[0] for x in gathered)
        # the bench's cross-rank timing reduction (max over ranks), over gloo
        import torch
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and float(t.item()) == float(world)
        q.put((rank, ok, len(events)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_ranks_agree_over_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 10, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert len({cnt for _, _, cnt in res}) == 1


# ---- commutation-aware scheduler (the default) -------------------------------------------------------------------------
def _plan_events_reordered(n, world, rank, ops):
    """Event stream with run records: a negative entry -1 - i means "gate i runs now"; exchange records as before."""
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
    cap = 64 * (len(ops) + 1)
    ev = (ctypes.c_int32 * cap)()
    phys = (ctypes.c_uint32 * n)()
    assert "QB_SHARD_REORDER" not in os.environ and "QB_SHARD_REMOTE" not in os.environ
    cnt = L.qb_shard_plan_debug(n, world, rank, arr, len(ops), ev, cap, phys)
    assert cnt >= 0, _ffi.last_error()
    flat, out, i = list(ev)[:cnt], [], 0
    while i < len(flat):
        if flat[i] < 0:
            out.append(("run", -1 - flat[i]))
            i += 1
        else:
            gp = flat[i + 1]
            out.append(("swap", flat[i + 2:i + 2 + gp], flat[i + 2 + gp:i + 2 + 2 * gp]))
            i += 2 + 2 * gp
    return out, list(phys)


def _replay_reordered(n, world, ops, events, seed):
    g = int(np.log2(world))
    nl = n - g
    s0 = random_state(n, seed)
    ora = OracleRegister(s0.copy(), n)
    for o in ops:
        ora.apply(o)
    phys, st, ran = list(range(n)), s0.copy(), []
    for e in events:
        if e[0] == "swap":
            for v, j in zip(e[1], e[2]):
                assert 0 <= v < nl and 0 <= j < g
                a = st.reshape((2,) * n)
                st = np.ascontiguousarray(np.swapaxes(a, n - 1 - v, n - 1 - (nl + j))).reshape(-1)
                ql, qg = phys.index(v), phys.index(nl + j)
                phys[ql], phys[qg] = nl + j, v
            continue
        m, t = ops[e[1]]
        for q in _dense_targets(m, t):
            assert phys[q] < nl, f"gate {e[1]}: dense target {q} is on a rank bit when the gate runs"
        st = oracle_np.apply_np(st, n, m, [phys[q] for q in t])
        ran.append(e[1])
    identities = [i for i, (m, t) in enumerate(ops) if np.array_equal(np.asarray(m), np.eye(len(m)))]
    assert sorted(ran + identities) == list(range(len(ops))), "every gate runs exactly once"
    a = st.reshape((2,) * n)
    a = np.transpose(a, [n - 1 - phys[n - 1 - ax] for ax in range(n)])
    return np.ascontiguousarray(a).reshape(-1), ora.state, phys


@pytest.mark.parametrize("n,world,kind", [(8, 2, "layers"), (10, 4, "layers"), (12, 8, "layers"), (9, 4, "qft"), (10, 4, "mixed"),
                                          (12, 8, "mixed"), (14, 4, "bench")])
def test_reordering_scheduler_replay_matches_oracle(n, world, kind):
    g = int(np.log2(world))
    if kind == "layers":
        ops = circuits.random_circuit(n, 8, seed=n + world)
    elif kind == "qft":
        ops = circuits.qft_circuit(n)
    elif kind == "bench":
        ops = circuits.qft_circuit(n) + circuits.random_circuit(n, 10, seed=1)
    else:
        ops = [o for o in circuits.mixed_circuit(n, 100, seed=n) if len(o[1]) <= n - g]
    per_rank = [_plan_events_reordered(n, world, r, ops) for r in range(world)]
    for ev, ph in per_rank[1:]:
        assert ev == per_rank[0][0] and ph == per_rank[0][1], "ranks disagree on the schedule"
    events, final_phys = per_rank[0]
    got, want, phys = _replay_reordered(n, world, ops, events, seed=n * 17 + world)
    assert phys == final_phys
    assert_close(got, want, f"reordered replay n={n} P={world} {kind}")
    swaps = sum(1 for e in events if e[0] == "swap")
    base = len(_plan_events(n, world, 0, ops)[0])
    assert swaps <= base, f"deferring must not need more exchanges ({swaps} vs {base})"


# ---- experimental exchange-free scheduler (QB_SHARD_REMOTE=1): rank bits as tile bits of remote tile passes -----------
REMOTE_MARK = -(1 << 30)


def _plan_events_remote(n, world, rank, ops):
    """Event stream of shard_run_remote in dry mode: ("run", i) = gate i through the local path, ("remote", [gates], tile bits) =
    one pass of k2r_tile_pass_remote over the virtual register, ("swap", V, J) = exchange (fallback scheduler only)."""
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
    cap = 64 * (len(ops) + 1)
    ev = (ctypes.c_int32 * cap)()
    phys = (ctypes.c_uint32 * n)()
    os.environ["QB_SHARD_REMOTE"] = "1"
    try:
        cnt = L.qb_shard_plan_debug(n, world, rank, arr, len(ops), ev, cap, phys)
    finally:
        del os.environ["QB_SHARD_REMOTE"]
    assert cnt >= 0, _ffi.last_error()
    flat, out, i = list(ev)[:cnt], [], 0
    while i < len(flat):
        if flat[i] == REMOTE_MARK:
            ng = flat[i + 1]
            out.append(("remote", flat[i + 2:i + 2 + ng], flat[i + 2 + ng:i + 14 + ng]))
            i += 14 + ng
        elif flat[i] < 0:
            out.append(("run", -1 - flat[i]))
            i += 1
        else:
            gp = flat[i + 1]
            out.append(("swap", flat[i + 2:i + 2 + gp], flat[i + 2 + gp:i + 2 + 2 * gp]))
            i += 2 + 2 * gp
    return out, list(phys)


@pytest.mark.parametrize("n,world,kind", [(13, 2, "layers"), (14, 4, "layers"), (15, 8, "layers"), (14, 2, "qft"), (15, 4, "qft"),
                                          (14, 4, "mixed"), (15, 8, "bench")])
def test_remote_scheduler_replay_matches_oracle(n, world, kind):
    """No exchanges: a gate with a dense target on a rank bit runs in a remote pass whose tile contains that rank bit.  All
    ranks must derive the same sequence; replayed in numpy (placement stays the identity) it must reproduce the oracle, with
    every dense target of a remote pass among its tile bits and every dense target of a local run on a local bit."""
    g = int(np.log2(world))
    nl = n - g
    if kind == "layers":
        ops = circuits.random_circuit(n, 6, seed=n + world)
    elif kind == "qft":
        ops = circuits.qft_circuit(n)
    elif kind == "bench":
        ops = circuits.qft_circuit(n) + circuits.random_circuit(n, 6, seed=2)
    else:
        ops = [o for o in circuits.mixed_circuit(n, 120, seed=n) if len(o[1]) <= 2]
    per_rank = [_plan_events_remote(n, world, r, ops) for r in range(world)]
    for ev, ph in per_rank[1:]:
        assert ev == per_rank[0][0] and ph == per_rank[0][1], "ranks disagree on the schedule"
    events, final_phys = per_rank[0]
    assert final_phys == list(range(n)), "remote passes never change the placement"
    assert not any(e[0] == "swap" for e in events), "no exchange is needed for tile-supported gates"
    assert any(e[0] == "remote" for e in events)
    s0 = random_state(n, n * 13 + world)
    ora = OracleRegister(s0.copy(), n)
    for o in ops:
        ora.apply(o)
    st, ran = s0.copy(), []
    for e in events:
        if e[0] == "run":
            m, t = ops[e[1]]
            for q in _dense_targets(m, t):
                assert q < nl, f"gate {e[1]}: dense target {q} is on a rank bit in a local run"
            st = oracle_np.apply_np(st, n, m, list(t))
            ran.append(e[1])
            continue
        tags, tile = e[1], e[2]
        assert tile == sorted(tile) and len(set(tile)) == 12 and any(b >= nl for b in tile)
        for i in tags:
            m, t = ops[i]
            for q in _dense_targets(m, t):
                assert q in tile, f"gate {i}: dense target {q} is not a tile bit of its remote pass"
            st = oracle_np.apply_np(st, n, m, list(t))
            ran.append(i)
    identities = [i for i, (m, t) in enumerate(ops) if np.array_equal(np.asarray(m), np.eye(len(m)))]
    assert sorted(ran + identities) == list(range(len(ops))), "every gate runs exactly once"
    assert_close(st, ora.state, f"remote schedule replay n={n} P={world} {kind}")


def test_remote_scheduler_falls_back_to_exchanges_for_wide_gates():
    """A dense 3-qubit gate cannot go through a tile pass: the whole list goes to the exchange scheduler."""
    from oracle import gates as og
    n, world = 14, 4
    rng = np.random.default_rng(5)
    u8 = np.linalg.qr(rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8)))[0]
    ops = [og.hadamard(13), (u8, [13, 2, 7]), og.cnot(12, 3)]
    events, phys = _plan_events_remote(n, world, 0, ops)
    assert any(e[0] == "swap" for e in events) and not any(e[0] == "remote" for e in events)
