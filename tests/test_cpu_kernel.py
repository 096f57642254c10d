"""The register-blocked tile kernel's DEVICE code (quaru_b200/csrc/qb_tile2_kernel.cuh -- the source nvcc compiles for
sm_100a) executed on CPU threads: tests/cpu_kernel/ builds it with g++ -DQB_HOST_EMU (one OS thread per CUDA thread, a
pthread barrier for __syncthreads) and runs the programs the planner + encoder produce (qb_rounds_debug).  The result must
match the CPU oracle, and the NumPy emulator of the program semantics.  This covers the kernel's control flow, indexing,
swizzled transposes and op decoding without a GPU; the inline PTX is replaced by C++ with the same operation order, so
the GPU parity tests remain the final word."""
import ctypes
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
SRC = os.path.join(ROOT, "tests", "cpu_kernel", "k2r_host.cpp")
OUT = os.path.join(ROOT, "tests", "cpu_kernel", "_build", "libk2r_host.so")


@pytest.fixture(scope="module")
def host_kernel():
    deps = [SRC, os.path.join(ROOT, "tests", "cpu_kernel", "qb_host_emu.h"),
            os.path.join(ROOT, "quaru_b200", "csrc", "qb_tile2_kernel.cuh"), os.path.join(ROOT, "quaru_b200", "csrc", "qb_device.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-ffp-contract=off", "-DQB_HOST_EMU", "-shared", "-fPIC", "-pthread",
                               "-x", "c++", "-I", os.path.dirname(SRC), "-I", os.path.join(ROOT, "quaru_b200", "csrc"), SRC, "-o", OUT])
    lib = ctypes.CDLL(OUT)
    lib.k2r_host_run.restype = ctypes.c_int
    lib.k2r_host_run.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int]
    lib.k2r_host_run_remote.restype = ctypes.c_int
    lib.k2r_host_run_remote.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_char_p,
                                        ctypes.c_uint64, ctypes.c_int]
    return lib


def _run(lib, n, ops, seed, small=False):
    s0 = random_state(n, seed)
    ora = OracleRegister(s0.copy(), n)
    for o in ops:
        ora.apply(o)
    dev, emu = s0.copy(), s0.copy()
    n_prog = 0
    for p in em.dump_programs(n, ops):
        if p[0] == "gates":
            for st in (dev, emu):
                tmp = OracleRegister(st, n)
                for tag in p[1]:
                    tmp.apply(ops[tag])
                st[:] = tmp.state
            continue
        n_prog += 1
        assert p[2][:p[3]["lc"]] == list(range(p[3]["lc"]))
        rc = lib.k2r_host_run(dev.ctypes.data, n, p[3]["bytes"], len(p[3]["bytes"]), int(small))
        assert rc == 0, rc
        emu = em.emulate_program(emu, n, p[2], p[3])
    return dev, emu, ora.state, n_prog


@pytest.mark.parametrize("n,depth,seed", [(12, 6, 1), (13, 5, 2), (14, 4, 3)])
def test_random_layers_on_the_host_compiled_kernel(host_kernel, n, depth, seed):
    dev, emu, want, n_prog = _run(host_kernel, n, circuits.random_circuit(n, depth, seed=seed), 50 + seed)
    assert n_prog >= 1
    assert_close(dev, want, f"kernel vs oracle, layers n={n}")
    assert_close(dev, emu, "kernel vs program emulator")


@pytest.mark.parametrize("n", [12, 14])
def test_qft_on_the_host_compiled_kernel(host_kernel, n):
    dev, emu, want, n_prog = _run(host_kernel, n, circuits.qft_circuit(n), 60 + n)
    assert n_prog >= 1
    assert_close(dev, want, f"kernel vs oracle, qft n={n}")


@pytest.mark.parametrize("n,seed", [(12, 5), (13, 6), (14, 7)])
def test_mixed_gates_on_the_host_compiled_kernel(host_kernel, n, seed):
    """Every op kind: dense complex / real / two-qubit gates, controls on register, thread and outer bits, wide diagonals."""
    dev, emu, want, n_prog = _run(host_kernel, n, circuits.mixed_circuit(n, 120, seed), 70 + seed)
    assert n_prog >= 1
    assert_close(dev, want, f"kernel vs oracle, mixed n={n}")
    assert_close(dev, emu, "kernel vs program emulator")


def test_small_program_instantiation(host_kernel):
    """RProgSmall (24 ops / 160 coefficients / 8 rounds) is the instantiation short passes are launched with."""
    n = 13
    ops = [og.hadamard(3), og.cnot(3, 9), og.rz(0.7, 9), og.cz(2, 12), og.hadamard(12), og.cphase(0.4, 0, 11), og.hadamard(11)]
    dev, emu, want, n_prog = _run(host_kernel, n, ops, 9, small=True)
    assert n_prog == 1
    assert_close(dev, want)


@pytest.mark.parametrize("kind,n", [("layers", 14), ("mixed", 13), ("qft", 14)])
def test_programs_of_the_thorough_encoder_on_the_host_compiled_kernel(host_kernel, kind, n, monkeypatch):
    """The round programs big registers get (look-ahead planner, roll-out round selection; forced on here with
    QB_PLAN_LOOKAHEAD=1) through the kernel's own source."""
    monkeypatch.setenv("QB_PLAN_LOOKAHEAD", "1")
    ops = {"layers": circuits.random_circuit(n, 6, seed=3 * n), "qft": circuits.qft_circuit(n), "mixed": circuits.mixed_circuit(n, 120, 2 * n)}[kind]
    dev, emu, want, n_prog = _run(host_kernel, n, ops, 90 + n)
    assert n_prog >= 1
    assert_close(dev, want, f"thorough encoder, kernel vs oracle, {kind} n={n}")
    assert_close(dev, emu, "kernel vs program emulator")


@pytest.mark.parametrize("world,n,kind", [(2, 14, "layers"), (2, 13, "mixed"), (4, 15, "qft"), (4, 14, "mixed"), (8, 15, "layers"), (8, 16, "mixed")])
def test_remote_tile_pass_on_the_host_compiled_kernel(host_kernel, world, n, kind):
    """k2r_tile_pass_remote (QB_SHARD_REMOTE=1): the register is split into `world` shards (separate host buffers reached
    through a pointer table, as the peer-mapped shards are on a GPU box); every rank runs its share of the tiles of the
    n-qubit programs -- whose tiles may contain rank bits -- and together they must reproduce the single-register run bit
    for bit: every tile processed exactly once, whatever mix of rank bits it contains."""
    g = world.bit_length() - 1
    nl = n - g
    ops = {"layers": circuits.random_circuit(n, 5, seed=7 * n), "qft": circuits.qft_circuit(n), "mixed": circuits.mixed_circuit(n, 120, 3 * n)}[kind]
    s0 = random_state(n, 40 + n)
    ora = OracleRegister(s0.copy(), n)
    for o in ops:
        ora.apply(o)
    ref = s0.copy()
    shards = [np.ascontiguousarray(s0[r << nl:(r + 1) << nl]) for r in range(world)]   # separate allocations, like the real shards
    bases = (ctypes.c_void_p * world)(*[sh.ctypes.data for sh in shards])
    n_remote = 0
    for p in em.dump_programs(n, ops):
        if p[0] == "gates":
            whole = np.concatenate(shards)
            for st in (ref, whole):
                tmp = OracleRegister(st, n)
                for tag in p[1]:
                    tmp.apply(ops[tag])
                st[:] = tmp.state
            for r in range(world):
                shards[r][:] = whole[r << nl:(r + 1) << nl]
            continue
        n_remote += any(b >= nl for b in p[2])
        assert host_kernel.k2r_host_run(ref.ctypes.data, n, p[3]["bytes"], len(p[3]["bytes"]), 0) == 0
        for r in range(world):
            assert host_kernel.k2r_host_run_remote(bases, world, r, n, p[3]["bytes"], len(p[3]["bytes"]), 0) == 0
    assert n_remote >= 1, "at least one pass has a rank bit among its tile bits"
    got = np.concatenate(shards)
    assert np.array_equal(got, ref), "same tiles, same arithmetic as the single-register kernel"
    assert_close(got, ora.state, f"remote tile pass vs oracle, world={world} n={n} {kind}")


@pytest.mark.parametrize("io_low,kind,n", [("3", "layers", 14), ("4", "mixed", 13), ("3", "qft", 14)])
def test_relaxed_io_rounds_on_the_host_compiled_kernel(host_kernel, io_low, kind, n, monkeypatch):
    """Experimental QB_TILE_IO_LOW = 4 / 3: tile bits 4 / 3 may be register bits of the HBM-facing rounds (a warp access then
    covers 256- / 128-byte pieces instead of one 512-byte segment), which saves the extra round some passes need only to get
    into a storable layout.  Same kernel, other host-computed offsets and lane mappings."""
    monkeypatch.setenv("QB_TILE_IO_LOW", io_low)
    ops = {"layers": circuits.random_circuit(n, 6, seed=5 * n), "qft": circuits.qft_circuit(n), "mixed": circuits.mixed_circuit(n, 120, 4 * n)}[kind]
    low = [r["R"] for p in em.dump_programs(n, ops) if p[0] == "prog" for r in (p[3]["rounds"][0], p[3]["rounds"][-1])]
    assert any(min(R) < 5 for R in low), "some HBM-facing round uses a register bit below 5"
    dev, emu, want, n_prog = _run(host_kernel, n, ops, 120 + n)
    assert n_prog >= 1
    assert_close(dev, want, f"relaxed io rounds, kernel vs oracle, {kind} n={n}")
    assert_close(dev, emu, "kernel vs program emulator")
