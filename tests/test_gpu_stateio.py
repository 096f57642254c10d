"""Raw state files: save -> load round trips are bit-exact, files are the memory image of the
reference's Complex<f64> vector (register.rs:23-37), and damaged files are rejected."""
import os
import struct

import numpy as np
import pytest

from quaru_b200 import Operation, Register
from quaru_b200.register import EngineError
from tests import circuits
from tests.util import random_state

pytestmark = pytest.mark.gpu


def _prepared(n, seed):
    r = Register.from_int(n, 0)
    r.set_state(random_state(n, seed))
    return r


@pytest.mark.parametrize("n", [1, 5, 12, 20, 23])        # 23 qubits = 128 MiB: two 64 MiB chunks through the pipeline
def test_round_trip_bit_exact(tmp_path, n):
    a = _prepared(n, 40 + n)
    want = a.state
    path = tmp_path / f"s{n}.qbstate"
    a.save_state(path)
    assert os.path.getsize(path) == 64 + 16 * (1 << n)
    b = Register.from_int(n, 1)
    b.load_state(path)
    assert np.array_equal(b.state, want)
    assert a.max_abs_diff(b) == 0.0


def test_file_is_the_reference_memory_image(tmp_path):
    n = 10
    a = _prepared(n, 7)
    path = tmp_path / "img.qbstate"
    a.save_state(path)
    raw = open(path, "rb").read()
    magic, version, dtype, nq, n_local, world, rank, count, norm2, checksum = struct.unpack_from("<8sIIIIIIQdQ", raw, 0)
    assert (magic, version, dtype, nq, n_local, world, rank, count) == (b"QBSTATE1", 1, 0, n, n, 1, 0, 1 << n)
    assert norm2 == a.norm2()
    payload = np.frombuffer(raw, dtype="<f8", offset=64).view(np.complex128)
    assert np.array_equal(payload, a.state)
    words = np.frombuffer(raw, dtype="<u8", offset=64)
    with np.errstate(over="ignore"):
        want = int(np.sum(words * (2 * np.arange(words.size, dtype=np.uint64) + 1), dtype=np.uint64))
    assert checksum == want, "header checksum = sum_j word_j * (2j+1) mod 2^64"


def test_save_flushes_pending_gates_and_load_drops_them(tmp_path):
    n = 12
    ops = [Operation(m, t) for m, t in circuits.random_circuit(n, 4, seed=3)]
    a = Register.from_int(n, 0)
    for op in ops:
        a.apply(op)
    path = tmp_path / "flushed.qbstate"
    a.save_state(path)                              # queued gates must be in the file
    want = a.state
    b = Register.from_int(n, 0)
    b.apply_batch(ops[:5])                          # still queued: replaced by the loaded image
    b.load_state(path)
    assert np.array_equal(b.state, want)
    b.apply_batch(ops[:5])
    a.apply_batch(ops[:5])
    assert a.max_abs_diff(b) == 0.0                 # and the register keeps working afterwards


def test_rejects_missing_wrong_size_truncated_and_corrupt_files(tmp_path):
    n = 12
    a = _prepared(n, 11)
    path = tmp_path / "ok.qbstate"
    a.save_state(path)
    raw = open(path, "rb").read()

    with pytest.raises(EngineError, match="cannot open"):
        a.load_state(tmp_path / "missing.qbstate")
    with pytest.raises(IndexError, match="12-qubit"):
        Register.from_int(n + 1, 0).load_state(path)

    def load(blob, name):
        p = tmp_path / name
        open(p, "wb").write(blob)
        Register.from_int(n, 0).load_state(p)

    with pytest.raises(EngineError, match="header"):
        load(raw[:40], "short.qbstate")
    with pytest.raises(EngineError, match="not a version-1"):
        load(b"NOTSTATE" + raw[8:], "magic.qbstate")
    with pytest.raises(EngineError, match="truncated"):
        load(raw[:-16], "trunc.qbstate")
    with pytest.raises(EngineError, match="trailing"):
        load(raw + b"\0", "long.qbstate")
    flipped = bytearray(raw)
    flipped[64 + 8 * 1001] ^= 0x10                  # one mantissa bit of one amplitude
    with pytest.raises(EngineError, match="checksum"):
        load(bytes(flipped), "flip.qbstate")
    with pytest.raises(EngineError, match="cannot create"):
        a.save_state(tmp_path / "no_such_dir" / "x.qbstate")
    Register.from_int(n, 0).load_state(path)        # the intact file still loads


def test_failed_load_leaves_the_register_as_it_was(tmp_path, monkeypatch):
    """A missing file or a file of another register is detected before the gate queue is dropped and the SWAP relabelling /
    placement is reset (qb_load_state): pending gates still run and the caller's qubit names still match the resident layout."""
    from oracle import gates as og
    from oracle.oracle import OracleRegister
    monkeypatch.setenv("QB_SWAP_RELABEL", "1")
    n = 13
    ops = [og.hadamard(0), og.swap(0, 2), og.not_(1), og.cnot(2, 12), og.hadamard(7)]
    reg = Register.new([False] * n)
    ora = OracleRegister.new([False] * n)
    for g in ops[:3]:
        reg.apply(Operation(g[0], g[1]))
        ora.apply(g)
    reg.sync()
    for g in ops[3:]:
        reg.apply(Operation(g[0], g[1]))            # still queued when the load fails
        ora.apply(g)
    other = tmp_path / "other.qbstate"
    Register.from_int(n + 1, 3).save_state(other)
    with pytest.raises(EngineError):
        reg.load_state(tmp_path / "missing.qbstate")
    with pytest.raises(IndexError):
        reg.load_state(other)
    assert np.abs(reg.state - ora.state).max() < 1e-12
