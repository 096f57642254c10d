"""The specialised tile kernels (K2j, quaru_b200/csrc/qb_jit_gen.cu) without a GPU.

For every fused pass of a circuit the library hands out the CUDA source it would compile with NVRTC
(qb_jit_source_debug).  Here that very text is built as host C++ (g++ -DQB_HOST_EMU: one OS thread per CUDA thread, a
pthread barrier for __syncthreads, tests/cpu_kernel/qb_jit_emu.h) and run over a host state vector; the result must match the
CPU oracle (the reference's try_apply_fast, /root/reference/src/register.rs:140-191, restated in oracle/).  A second test
compiles the same text for sm_100a with NVRTC (no device needed) -- what the launch path does on the GPU box.
"""
import ctypes
import hashlib
import os
import subprocess

import numpy as np
import pytest

from oracle import gates as og
from oracle.oracle import OracleRegister
from tests import circuits
from tests import rounds_emulator as em
from tests.util import assert_close, random_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KDIR = os.path.join(ROOT, "tests", "cpu_kernel")
BUILD = os.path.join(KDIR, "_build", "jit")


def jit_source(prog_bytes, min_blocks=2, persistent=0, prefetch=0, remote=0, xchg=None):
    from quaru_b200 import _ffi
    L = _ffi.lib()
    cap = 8 << 20
    buf = ctypes.create_string_buffer(cap)
    vj = None
    if xchg:                                           # ([V...], [J...]): local bits V[i] traded against rank bits J[i]
        vj = bytes(list(xchg[0]) + [0] * (3 - len(xchg[0])) + list(xchg[1]) + [0] * (3 - len(xchg[1])))
    got = L.qb_jit_source_debug(prog_bytes, len(prog_bytes), min_blocks, persistent, prefetch, remote, len(xchg[0]) if xchg else 0, vj, buf, cap)
    assert got > 0, _ffi.last_error()
    return buf.raw[:got].decode()


def build_host(src, remote=False, xchg=False, pipe=False):
    os.makedirs(BUILD, exist_ok=True)
    tag = hashlib.sha1((src + str(remote) + str(xchg) + str(pipe)).encode()).hexdigest()[:16]
    cu, so = os.path.join(BUILD, tag + ".cu"), os.path.join(BUILD, tag + ".so")
    if not os.path.exists(so):
        with open(cu, "w") as f:
            f.write(src)
        cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-ffp-contract=off", "-DQB_HOST_EMU", f'-DQB_JIT_SRC="{cu}"', "-shared", "-fPIC", "-pthread",
               "-I", KDIR, os.path.join(KDIR, "jit_host_runner.cpp"), "-o", so]
        if remote:
            cmd.insert(5, "-DQB_JIT_REMOTE")
        if xchg:
            cmd.insert(5, "-DQB_JIT_XCHG")
        if pipe:
            cmd.insert(5, f"-DQB_JIT_PIPE={pipe}")   # CTA size of the prefetching ring
        subprocess.check_call(cmd)
    lib = ctypes.CDLL(so)
    lib.jit_host_run.restype = ctypes.c_int
    lib.jit_host_run.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint, ctypes.c_uint]
    return lib


def run_circuit(n, ops, seed, grid=0, **gen):
    s0 = random_state(n, seed)
    ora = OracleRegister(s0.copy(), n)
    for o in ops:
        ora.apply(o)
    dev = s0.copy()
    n_prog = 0
    for p in em.dump_programs(n, ops):
        if p[0] == "gates":
            tmp = OracleRegister(dev, n)
            for tag in p[1]:
                tmp.apply(ops[tag])
            dev[:] = tmp.state
            continue
        n_prog += 1
        lib = build_host(jit_source(p[3]["bytes"], **gen), pipe=512 if gen.get("persistent") == 2 else 0)
        assert lib.jit_host_run(dev.ctypes.data, 1 << (n - 12), grid, 0, None, 0, 0) == 0
    return dev, ora.state, n_prog


@pytest.mark.parametrize("n,depth,seed", [(12, 6, 1), (13, 5, 2), (14, 4, 3)])
def test_random_layers(n, depth, seed):
    dev, want, n_prog = run_circuit(n, circuits.random_circuit(n, depth, seed=seed), 50 + seed)
    assert n_prog >= 1
    assert_close(dev, want, f"generated kernel vs oracle, layers n={n}")


@pytest.mark.parametrize("n", [12, 14])
def test_qft(n):
    dev, want, n_prog = run_circuit(n, circuits.qft_circuit(n), 60 + n)
    assert n_prog >= 1
    assert_close(dev, want, f"generated kernel vs oracle, qft n={n}")


@pytest.mark.parametrize("n,seed", [(12, 5), (13, 6), (14, 7)])
def test_mixed_gates(n, seed):
    """Every op kind: dense complex / real / two-qubit gates, controls on register, thread and outer bits, wide diagonals,
    X / CNOT / SWAP (renamings when uncontrolled)."""
    dev, want, n_prog = run_circuit(n, circuits.mixed_circuit(n, 120, seed), 70 + seed)
    assert n_prog >= 1
    assert_close(dev, want, f"generated kernel vs oracle, mixed n={n}")


@pytest.mark.parametrize("n,grid,kind,prefetch", [(14, 3, "layers", 1), (13, 1, "mixed", 0), (14, 2, "qft", 1)])
def test_persistent_variant(n, grid, kind, prefetch):
    ops = {"layers": circuits.random_circuit(n, 4, seed=n), "qft": circuits.qft_circuit(n), "mixed": circuits.mixed_circuit(n, 100, n)}[kind]
    dev, want, n_prog = run_circuit(n, ops, 80 + n, grid=grid, persistent=1, prefetch=prefetch)
    assert n_prog >= 1
    assert_close(dev, want, f"persistent generated kernel vs oracle, {kind} n={n} grid={grid}")


def test_permutation_gates_are_renamings():
    """X, CNOT and SWAP between register bits move no data: the generator only renames the variables a later store reads."""
    n = 12
    ops = [og.hadamard(8), og.hadamard(9), og.not_(9), og.cnot(8, 9), og.rz(0.3, 9), og.hadamard(10), og.swap(9, 10), og.hadamard(9)]
    progs = [p for p in em.dump_programs(n, ops) if p[0] == "prog"]
    assert progs
    src = jit_source(progs[0][3]["bytes"])
    body = src[src.index("QB_KERNEL(qb_jit_pass"):]
    dev, want, _ = run_circuit(n, ops, 5)
    assert_close(dev, want)
    assert " * " not in body.split("// ---- round 0")[0], "no arithmetic before the first round"


def test_nvrtc_compiles_the_generated_source():
    """The text compiles for sm_100a with NVRTC (libnvrtc is loaded with dlopen; no device needed): the launch path's first half."""
    from quaru_b200 import _ffi
    L = _ffi.lib()
    n = 14
    ops = circuits.mixed_circuit(n, 120, 3) + circuits.qft_circuit(n)
    progs = [p for p in em.dump_programs(n, ops) if p[0] == "prog"]
    assert progs
    for variant in ({}, {"persistent": 1, "prefetch": 1}, {"min_blocks": 3}):
        src = jit_source(progs[0][3]["bytes"], **variant)
        ms = ctypes.c_double(0)
        got = L.qb_jit_compile_debug(src.encode(), ctypes.byref(ms), None, 0)
        assert got > 0, _ffi.last_error()


def _remote_fixed_bits(hp, n_local, g, rank):
    """qb_tile2_kernel.cuh remote_fixed_bits: the g fixed top bits of rank `rank`'s tile indices."""
    in_tile = 0
    for b in hp:
        if b >= n_local:
            in_tile |= 1 << (b - n_local)
    fixed = pos = 0
    for J in range(g):
        if (in_tile >> J) & 1:
            fixed |= ((rank >> J) & 1) << pos
            pos += 1
    for J in range(g):
        if not (in_tile >> J) & 1:
            fixed |= ((rank >> J) & 1) << pos
            pos += 1
    return fixed


@pytest.mark.parametrize("world,n,kind", [(2, 14, "layers"), (4, 15, "qft"), (4, 14, "mixed"), (8, 16, "mixed")])
def test_remote_variant_over_separately_allocated_shards(world, n, kind):
    """Sharded registers (QB_SHARD_REMOTE=1): the specialised kernel of a pass whose tile contains RANK bits addresses amplitudes
    through the table of peer-mapped shards.  The register is split into `world` separate host buffers; every rank runs its
    share of the tiles; together they must reproduce the oracle."""
    g = world.bit_length() - 1
    nl = n - g
    ops = {"layers": circuits.random_circuit(n, 5, seed=7 * n), "qft": circuits.qft_circuit(n),
           "mixed": [o for o in circuits.mixed_circuit(n, 120, 3 * n) if len(o[1]) <= 2]}[kind]
    s0 = random_state(n, 40 + n)
    ora = OracleRegister(s0.copy(), n)
    for o in ops:
        ora.apply(o)
    shards = [np.ascontiguousarray(s0[r << nl:(r + 1) << nl]) for r in range(world)]
    bases = (ctypes.c_void_p * world)(*[sh.ctypes.data for sh in shards])
    n_remote = 0
    for p in em.dump_programs(n, ops):
        if p[0] == "gates":
            whole = np.concatenate(shards)
            tmp = OracleRegister(whole, n)
            for tag in p[1]:
                tmp.apply(ops[tag])
            for r in range(world):
                shards[r][:] = tmp.state[r << nl:(r + 1) << nl]
            continue
        hp = p[3]["hp"][:p[3]["h"]]
        n_remote += any(b >= nl for b in hp)
        lib = build_host(jit_source(p[3]["bytes"], remote=1), remote=True)
        for r in range(world):
            tile_or = _remote_fixed_bits(hp, nl, g, r) << (n - 12 - g)
            assert lib.jit_host_run(None, 1 << (nl - 12), 0, tile_or, bases, world, nl) == 0
    assert n_remote >= 1, "at least one pass has a rank bit among its tile bits"
    assert_close(np.concatenate(shards), ora.state, f"remote generated kernel vs oracle, world={world} n={n} {kind}")


@pytest.mark.parametrize("world,n,kind,V,J", [(2, 14, "layers", [12], [0]), (4, 15, "mixed", [9, 12], [1, 0]), (8, 16, "qft", [7, 10, 12], [0, 2, 1]),
                                             (4, 14, "layers", [5, 11], [0, 1]), (8, 15, "layers", [9, 10, 11], [2, 1, 0])])
def test_exchanging_pass_stores_into_the_post_exchange_placement(world, n, kind, V, J):
    """Sharded registers: the last pass before a global<->local exchange carries the exchange in its stores -- every amplitude
    is written where it lives AFTER local bits V[i] have been traded against rank bits J[i], into the destination rank's
    second buffer.  Every rank runs the LOCAL programs of its shard (n - g qubits) with that epilogue; the second buffers,
    read back with the traded bits swapped, must equal what pass-then-exchange gives (the oracle on each shard, permuted)."""
    g = world.bit_length() - 1
    nl = n - g
    ops = {"layers": circuits.random_circuit(nl, 3, seed=3 * n), "qft": circuits.qft_circuit(nl)[:40],
           "mixed": [o for o in circuits.mixed_circuit(nl, 40, 5 * n) if len(o[1]) <= 2]}[kind]
    progs = [p for p in em.dump_programs(nl, ops)]
    assert len(progs) == 1 and progs[0][0] == "prog", "one fused pass per shard"
    A = [random_state(nl, 700 + r) for r in range(world)]
    want_local = []
    for r in range(world):
        ora = OracleRegister(A[r].copy(), nl)
        for o in ops:
            ora.apply(o)
        want_local.append(ora.state)
    B = [np.zeros(1 << nl, dtype=np.complex128) for _ in range(world)]
    bases = (ctypes.c_void_p * world)(*[b.ctypes.data for b in B])
    lib = build_host(jit_source(progs[0][3]["bytes"], xchg=(V, J)), xchg=True)
    for r in range(world):
        assert lib.jit_host_run(A[r].ctypes.data, 1 << (nl - 12), 0, r, bases, world, nl) == 0
    # expected: amplitude (rank r, local l) -> (rank r', local l') with rank bit J[i] <-> local bit V[i]
    for r in range(world):
        l = np.arange(1 << nl, dtype=np.int64)
        r2 = np.full_like(l, r)
        l2 = l.copy()
        for v, j in zip(V, J):
            lb = (l >> v) & 1
            rb = (r >> j) & 1
            r2 = (r2 & ~(1 << j)) | (lb << j)
            l2 = (l2 & ~(1 << v)) | (rb << v)
        for rr in range(world):
            sel = r2 == rr
            assert_close(B[rr][l2[sel]], want_local[r][l[sel]], f"exchanging pass world={world} rank {r} -> {rr}")


@pytest.mark.parametrize("n,grid,kind", [(14, 1, "layers"), (15, 2, "mixed"), (15, 3, "qft"), (13, 1, "mixed"), (16, 5, "layers"), (12, 1, "qft"),
                                         (18, 1, "layers")])
def test_prefetching_ring(n, grid, kind):
    """The prefetching ring (one persistent CTA of 512 threads per SM: two groups of 256 threads on alternate tiles, every thread
    fetches its own amplitudes of the group's next tile with cp.async into a ring of three shared-memory tile buffers; the
    groups tell each other through done[] flags when a buffer may be overwritten; results go straight from registers to the
    state) on CPU threads: tile counts that are smaller than, equal to and not a multiple of the ring and the grid, and 64 tiles
    through one CTA."""
    ops = {"layers": circuits.random_circuit(n, 4, seed=2 * n), "qft": circuits.qft_circuit(n), "mixed": circuits.mixed_circuit(n, 100, 3 * n)}[kind]
    dev, want, n_prog = run_circuit(n, ops, 900 + n, grid=grid, persistent=2)
    assert n_prog >= 1
    assert_close(dev, want, f"prefetching-ring kernel vs oracle, {kind} n={n} grid={grid}")
