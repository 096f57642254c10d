"""The C-ABI library loads on a CPU-only box and exports every symbol include/quaru_b200.h declares; argument
errors that need no device are reported through the documented codes; compute entry points fail loudly (no CPU
fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "quaru_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(engine):
    from quaru_b200 import _ffi
    L = _ffi.lib()
    names = _declared()
    assert len(names) >= 28
    for nm in names:
        assert hasattr(L, nm), f"{nm} declared in the header but not exported"
        assert nm in _ffi.SIGNATURES, f"{nm} has no ctypes signature"
    assert set(_ffi.SIGNATURES) == set(names)
    assert b"sm_100a" in L.qb_version()


def test_no_cpu_fallback_without_device(engine):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from quaru_b200 import EngineError, Register, _ffi
    with pytest.raises(EngineError):
        Register.new([False, False])
    assert "CUDA" in _ffi.last_error()[0]


def test_null_and_range_arguments(engine):
    from quaru_b200 import _ffi
    L = _ffi.lib()
    assert L.qb_create(3, 0, None) == _ffi.QB_ERR_INVALID_ARGUMENT
    h = ctypes.c_void_p()
    assert L.qb_create(3, 8, ctypes.byref(h)) == _ffi.QB_ERR_INVALID_ARGUMENT      # index out of bounds (math.rs:46)
    assert L.qb_create(64, 0, ctypes.byref(h)) == _ffi.QB_ERR_UNSUPPORTED
    assert L.qb_apply(None, None, None, 0) == _ffi.QB_ERR_INVALID_ARGUMENT
    assert L.qb_measure(None, 0, 0.5, None) == _ffi.QB_ERR_INVALID_ARGUMENT
    assert L.qb_num_qubits(None) == 0
    L.qb_destroy(None)                                                              # no-op
    bad = (ctypes.c_double * 8)(1, 0, 0.5, 0, 1, 0, 0, 0)
    assert L.qb_create_product(bad, 2, ctypes.byref(h)) == _ffi.QB_ERR_INVALID_QUBIT
    assert L.qb_last_error_payload() == 0
    assert L.qb_create_sharded(4, 0, 2, 2, None, -1, ctypes.byref(h)) == _ffi.QB_ERR_INVALID_ARGUMENT
    assert L.qb_create_sharded(4, 0, 0, 3, None, -1, ctypes.byref(h)) == _ffi.QB_ERR_INVALID_ARGUMENT
    assert L.qb_save_state(None, b"/tmp/x") == _ffi.QB_ERR_INVALID_ARGUMENT
    assert L.qb_load_state(None, b"/tmp/x") == _ffi.QB_ERR_INVALID_ARGUMENT


def test_garbage_arity_is_an_invalid_target_not_a_crash(engine):
    """An absurd arity must come back through the reference's validation (smallest duplicate, then first out of
    range; register.rs:146-153) without sizing host buffers from it; no exception may cross the C ABI."""
    from quaru_b200 import _ffi
    L = _ffi.lib()
    k = 100000
    targets = (ctypes.c_uint32 * k)(*([7, 3] + [5] * (k - 2)))
    m = (ctypes.c_double * 8)(1, 0, 0, 0, 0, 0, 1, 0)
    gate = (_ffi.QbGate * 1)()
    gate[0].matrix = ctypes.cast(m, ctypes.POINTER(ctypes.c_double))
    gate[0].targets = targets
    gate[0].k = k
    pass_of, order = (ctypes.c_int32 * 1)(), (ctypes.c_int32 * 1)()
    assert L.qb_plan_debug(12, gate, 1, 12, pass_of, order) == _ffi.QB_ERR_INVALID_TARGET
    assert L.qb_last_error_payload() == 5                       # the duplicate
    targets2 = (ctypes.c_uint32 * 64)(*range(64))
    gate[0].targets, gate[0].k = targets2, 64
    assert L.qb_plan_debug(12, gate, 1, 12, pass_of, order) == _ffi.QB_ERR_INVALID_TARGET
    assert L.qb_last_error_payload() == 12                      # first target >= n


def test_oracle_is_not_imported_by_the_product():
    """The product must never route through oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "quaru_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "quaru_oracle" not in src, f
