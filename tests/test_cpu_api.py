"""The engine's C ABI host logic on a machine without a GPU.

tests/cpu_api/ builds the engine's own host translation units (quaru_b200/csrc/qb_api.cu, qb_gate.cu, qb_qasm.cu) with g++
against a stand-in for the CUDA runtime (device memory = host memory) and a host restatement of the kernel-launcher layer
(host_backend.cpp).  Everything above the kernels is the product's code: argument validation and error payloads, the gate
queue, SWAP relabelling and its canonicalisation before raw access, measurement plumbing, clone / reset / diff / state
files, options and stats.  The Python mirror (quaru_b200.Register) is pointed at that library for the duration of a test.
What the kernels compute is NOT covered here (tests/test_cpu_kernel.py, tests/test_rounds_emulator.py and the -m gpu
tests do that)."""
import os
import subprocess

import numpy as np
import pytest

from tests import reference_suite as rs
from tests import test_gpu_relabel as relabel_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests", "cpu_api")
OUT = os.path.join(HERE, "_build", "libquaru_hostapi.so")
CSRC = os.path.join(ROOT, "quaru_b200", "csrc")
SOURCES = [os.path.join(CSRC, f) for f in ("qb_api.cu", "qb_gate.cu", "qb_qasm.cu")] + [os.path.join(HERE, "host_backend.cpp")]


@pytest.fixture(scope="module")
def host_lib():
    deps = SOURCES + [os.path.join(HERE, "fake", "cuda_runtime.h"), os.path.join(ROOT, "include", "quaru_b200.h")] + \
        [os.path.join(CSRC, f) for f in ("qb_common.h", "qb_shard.h", "qb_tile.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-I", os.path.join(HERE, "fake"), "-I", CSRC,
                               "-x", "c++"] + SOURCES + ["-o", OUT])
    return OUT


@pytest.fixture()
def engine(host_lib, monkeypatch):
    """quaru_b200 with its loader pointed at the host build (restored afterwards: other tests use the real library)."""
    import quaru_b200
    from quaru_b200 import _ffi
    monkeypatch.setattr(_ffi, "LIB_PATH", host_lib)
    monkeypatch.setattr(_ffi, "_lib", None)
    return quaru_b200


@pytest.fixture()
def relabel(monkeypatch):
    monkeypatch.setenv("QB_SWAP_RELABEL", "1")


@pytest.mark.parametrize("relabel_env", ["1", "0"])
@pytest.mark.parametrize("fusion", [1, 0])
@pytest.mark.parametrize("test", rs.ALL, ids=lambda t: t.__name__)
def test_reference_suite_through_the_api_layer(engine, monkeypatch, test, fusion, relabel_env):
    """Quaru's own tests (ported in tests/reference_suite.py) through the real C ABI entry points, with SWAP relabelling
    forced on and off."""
    monkeypatch.setenv("QB_SWAP_RELABEL", relabel_env)
    test(rs.EngineBackend(engine, fusion), np.random.default_rng(12345))


@pytest.mark.parametrize("n,kind,fusion", [(8, "qft", 1), (10, "swaps", 1), (9, "bench", 1), (9, "swaps", 0)])
def test_relabelled_state_and_outcomes(engine, relabel, n, kind, fusion):
    relabel_cases.test_state_and_outcomes_match_the_oracle(engine, relabel, n, kind, fusion)


def test_relabelled_swaps_cost_no_sweep(engine, relabel):
    relabel_cases.test_swaps_cost_no_sweep(engine, relabel)


def test_relabelled_clone_diff_reset_write_and_files(engine, relabel, tmp_path):
    relabel_cases.test_clone_diff_reset_write_and_files_see_callers_qubits(engine, relabel, tmp_path)


def test_relabelling_is_off_for_small_registers_by_default(engine, monkeypatch):
    from oracle import gates as og
    monkeypatch.delenv("QB_SWAP_RELABEL", raising=False)
    reg = engine.Register.new([False] * 6)
    reg.apply(engine.Operation(*og.hadamard(1)))
    reg.flush()
    before = reg.stat(2)
    reg.apply(engine.Operation(*og.swap(1, 4)))
    reg.flush()
    assert reg.stat(2) == before + 1, "below 28 local qubits a SWAP is an ordinary gate"
    p = reg.probabilities(4)
    assert abs(p[0] - 0.5) < 1e-12
    reg.close()


def test_error_payloads_and_options(engine):
    from quaru_b200 import _ffi
    from quaru_b200.register import OperationError
    from oracle import gates as og
    reg = engine.Register.new([False] * 3)
    with pytest.raises(OperationError) as e:
        reg.try_apply_fast(engine.Operation(og.hadamard(0)[0], [3]))
    assert e.value == OperationError("InvalidTarget", 3)
    with pytest.raises(OperationError) as e:
        reg.try_apply_fast(engine.Operation(og.cnot(0, 1)[0], [1, 1]))
    assert e.value == OperationError("InvalidTarget", 1)
    with pytest.raises(OperationError):
        reg.try_measure(7, u=0.5)
    with pytest.raises(Exception):
        reg.set_option(_ffi.QB_OPT_FORCE_PATH, 6)
    reg.set_option(_ffi.QB_OPT_FORCE_PATH, 0)
    with pytest.raises(Exception):
        reg.set_option(_ffi.QB_OPT_JIT, 3)
    reg.set_option(_ffi.QB_OPT_JIT, 1)
    assert reg.stat(_ffi.QB_STAT_REMOTE_PASSES) == 0
    reg.close()


def test_bench_diagnostics_row_on_the_host_build(engine, capsys):
    """tools/bench_diagnostics.py (the table bench.py appends to its JSON line): one configuration, in process, on the host
    build -- the happy path of the script itself (circuits, stats, timers, JSON row)."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("bench_diagnostics", os.path.join(ROOT, "tools", "bench_diagnostics.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.one(12, 0, 2)
    out = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("DIAG ")]
    row = json.loads(out[-1][5:])
    assert row["max_abs_diff_vs_interpreter_n24"] == 0.0 and abs(row["norm2_end"] - 1.0) < 1e-9
    for kind in ("layers", "qft", "h4", "h12"):
        assert row[kind]["passes"] >= 1 and row[kind]["gates"] >= 4


@pytest.mark.parametrize("relabel_env,max_queue", [("1", 16), ("0", 16), ("1", 9), ("0", 40)])
def test_overflow_flush_hands_the_last_pass_back(engine, monkeypatch, relabel_env, max_queue):
    """When the queue overflows in the middle of a gate stream the gates of the last planned pass stay queued (in front of
    whatever comes next) instead of being launched half empty.  The host stand-in's run_fused hands back the trailing gates the
    same way, so the queue bookkeeping of the API layer -- order, the `held` count, full flushes at measurements and reads,
    relabelled gates not translated twice -- is what is checked here against the oracle."""
    from quaru_b200 import _ffi
    from oracle import gates as og
    from oracle.oracle import OracleRegister
    from tests import circuits
    from tests.util import assert_close
    monkeypatch.setenv("QB_SWAP_RELABEL", relabel_env)
    n = 9
    ops = relabel_cases._swap_heavy(n, 150, 11) + circuits.qft_circuit(n) + circuits.random_circuit(n, 5, seed=4)
    reg = engine.Register.new([False] * n)
    reg.set_option(_ffi.QB_OPT_MAX_QUEUE, max_queue)
    ora = OracleRegister.new([False] * n)
    rng = np.random.default_rng(8)
    for i, (m, t) in enumerate(ops):
        reg.apply(engine.Operation(m, t))
        ora.apply((m, t))
        if i % 97 == 96:                              # a full flush in the middle of the stream
            q = int(rng.integers(n))
            u = float(rng.random())
            assert reg.measure(q, u=u) == ora.measure(q, u)
        if i % 131 == 130:
            assert_close(reg.state, ora.state, f"after gate {i}")
    assert_close(reg.state, ora.state, "end of the stream")
    reg.close()


@pytest.mark.parametrize("relabel_env", ["0", "1"])
def test_cpp_mirror_through_the_api_layer(host_lib, relabel_env, tmp_path):
    """include/quaru_b200.hpp (the C++ RAII mirror of Register) and its test program, linked against the host build: the whole
    program, not only its --host-only part."""
    exe = str(tmp_path / "test_register_host")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_register.cpp"), "-o", exe, "-L", os.path.dirname(host_lib),
                           "-lquaru_hostapi", f"-Wl,-rpath,{os.path.dirname(host_lib)}"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=dict(os.environ, QB_SWAP_RELABEL=relabel_env))
    assert out.returncode == 0 and "CPP_OK" in out.stdout, out.stdout + out.stderr


def test_failed_load_leaves_queue_and_labels_untouched(engine, relabel, tmp_path):
    """qb_load_state validates the file BEFORE it drops the gate queue and resets the SWAP relabelling / placement: after a
    failed load (missing file) the pending gates are still applied and the caller's qubit names still match the layout."""
    from oracle import gates as og
    from oracle.oracle import OracleRegister
    n = 5
    ops = [og.hadamard(0), og.swap(0, 2), og.not_(1), og.cnot(2, 4)]
    reg = engine.Register.new([False] * n)
    ora = OracleRegister.new([False] * n)
    for g in ops[:3]:
        reg.apply(engine.Operation(g[0], g[1]))
        ora.apply(g)
    reg.sync()                                      # H, SWAP (a relabelling), X executed; the labels are permuted now
    reg.apply(engine.Operation(ops[3][0], ops[3][1]))   # still queued
    ora.apply(ops[3])
    with pytest.raises(Exception):
        reg.load_state(tmp_path / "missing.qbstate")
    assert np.abs(reg.state - ora.state).max() < 1e-12
