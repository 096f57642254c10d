"""ctypes loader + Register mirror for the CPU oracle (TEST INFRASTRUCTURE ONLY).

``OracleRegister`` mirrors the reference's ``Register`` API (src/register.rs:39-407) on top of
oracle/quaru_oracle.c so that the ported reference tests read like the originals.  ``measure``
takes the uniform draw ``u`` explicitly (the reference draws it from an unseeded thread_rng,
register.rs:317-318).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; quaru_b200/ never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libquaru_oracle.so")

OK = 0
ERR_INVALID_TARGET = -1
ERR_INVALID_DIMENSIONS = -2
ERR_INVALID_ARITY = -3
ERR_INVALID_QUBIT = -4


class OperationError(Exception):
    """Mirror of register.rs:10-20; ``kind`` in {InvalidTarget, InvalidDimensions, InvalidArity, InvalidQubit}."""

    def __init__(self, kind, payload):
        super().__init__(f"{kind}({payload})")
        self.kind = kind
        self.payload = payload

    def __eq__(self, other):
        return isinstance(other, OperationError) and (self.kind, self.payload) == (other.kind, other.payload)

    __hash__ = None


_KIND = {ERR_INVALID_TARGET: "InvalidTarget", ERR_INVALID_DIMENSIONS: "InvalidDimensions",
         ERR_INVALID_ARITY: "InvalidArity", ERR_INVALID_QUBIT: "InvalidQubit"}


def build(force=False):
    src = os.path.join(_HERE, "quaru_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.orc_try_apply_fast.argtypes = [dp, ctypes.c_uint32, dp, u32p, ctypes.c_uint32, ctypes.c_int, i64p]
        L.orc_try_apply_fast_faithful.argtypes = [dp, ctypes.c_uint32, dp, u32p, ctypes.c_uint32, i64p]
        L.orc_try_apply_all.argtypes = [dp, ctypes.c_uint32, dp, ctypes.c_int]
        L.orc_probabilities.argtypes = [dp, ctypes.c_uint32, ctypes.c_uint32, dp, dp, i64p]
        L.orc_try_measure.argtypes = [dp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_double,
                                      ctypes.POINTER(ctypes.c_int32), i64p]
        L.orc_init_basis.argtypes = [dp, ctypes.c_uint32, ctypes.c_uint64]
        L.orc_init_basis.restype = None
        L.orc_try_from_qubits.argtypes = [dp, ctypes.c_uint32, dp, i64p]
        L.orc_registers_equal.argtypes = [dp, dp, ctypes.c_uint32]
        L.orc_validate_apply.argtypes = [ctypes.c_uint32, u32p, ctypes.c_uint32, i64p]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _raise(rc, payload):
    raise OperationError(_KIND.get(rc, f"Error{rc}"), payload.value)


def apply_raw(state, n, matrix, targets, threads=1, faithful=False):
    """In-place try_apply_fast on a complex128 vector of length 2^n.  Raises OperationError."""
    L = lib()
    m = np.ascontiguousarray(matrix, dtype=np.complex128)
    t = np.ascontiguousarray(targets, dtype=np.uint32)
    pl = ctypes.c_int64(-1)
    tp = t.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
    if faithful:
        rc = L.orc_try_apply_fast_faithful(_dp(state), n, _dp(m), tp, len(t), ctypes.byref(pl))
    else:
        rc = L.orc_try_apply_fast(_dp(state), n, _dp(m), tp, len(t), threads, ctypes.byref(pl))
    if rc != OK:
        _raise(rc, pl)


class OracleRegister:
    """Mirror of quaru::register::Register backed by the C oracle."""

    def __init__(self, state, size):
        self.state = np.ascontiguousarray(state, dtype=np.complex128).reshape(-1)
        self._size = size
        assert self.state.shape[0] == 1 << size

    # register.rs:41
    @classmethod
    def new(cls, input_bits):
        n = len(input_bits)
        idx = sum((1 << q) for q, b in enumerate(input_bits) if b)
        return cls.from_int(n, idx)

    # register.rs:58
    @classmethod
    def from_int(cls, n, k):
        if k >= 1 << n:
            raise IndexError("int_to_state: index out of bounds")      # reference panics (math.rs:46)
        st = np.empty(1 << n, dtype=np.complex128)
        lib().orc_init_basis(_dp(st), n, k)
        return cls(st, n)

    # register.rs:81
    @classmethod
    def try_from_qubits(cls, qubits):
        n = len(qubits)
        for i, q in enumerate(qubits):
            if np.asarray(q).size != 2:                                 # is_qubit len check (:121)
                raise OperationError("InvalidQubit", i)
        q = np.ascontiguousarray(np.array([np.asarray(x).reshape(2) for x in qubits], dtype=np.complex128))
        st = np.empty(1 << n, dtype=np.complex128)
        pl = ctypes.c_int64(-1)
        rc = lib().orc_try_from_qubits(_dp(st), n, _dp(q) if n else None, ctypes.byref(pl))
        if rc != OK:
            _raise(rc, pl)
        return cls(st, n)

    @classmethod
    def from_qubits(cls, qubits):
        try:
            return cls.try_from_qubits(qubits)
        except OperationError as e:
            raise RuntimeError("Incorrect input qubits") from e        # register.rs:75

    def size(self):
        return self._size

    def clone(self):
        return OracleRegister(self.state.copy(), self._size)

    # register.rs:140
    def try_apply_fast(self, op, threads=1, faithful=False):
        matrix, targets = op
        apply_raw(self.state, self._size, matrix, targets, threads, faithful)
        return self

    try_apply = try_apply_fast   # register.rs:193 has the same contract (validated by oracle_np.apply_slow)

    # register.rs:134
    def apply(self, op, **kw):
        try:
            return self.try_apply_fast(op, **kw)
        except OperationError as e:
            raise RuntimeError("Could not apply operation") from e

    # register.rs:267
    def try_apply_all(self, op, literal=None):
        matrix, targets = op
        if len(targets) != 1:
            raise OperationError("InvalidArity", len(targets))
        if literal is None:
            literal = self._size <= 10
        m = np.ascontiguousarray(matrix, dtype=np.complex128)
        rc = lib().orc_try_apply_all(_dp(self.state), self._size, _dp(m), int(bool(literal)))
        assert rc == OK
        return self

    def apply_all(self, op, **kw):
        try:
            return self.try_apply_all(op, **kw)
        except OperationError as e:
            raise RuntimeError("Could not apply operation") from e

    def probabilities(self, target):
        p0, p1, pl = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64(-1)
        rc = lib().orc_probabilities(_dp(self.state), self._size, target, ctypes.byref(p0), ctypes.byref(p1),
                                     ctypes.byref(pl))
        if rc != OK:
            _raise(rc, pl)
        return p0.value, p1.value

    # register.rs:297 (u supplied by the caller)
    def try_measure(self, target, u):
        out, pl = ctypes.c_int32(), ctypes.c_int64(-1)
        rc = lib().orc_try_measure(_dp(self.state), self._size, target, float(u), ctypes.byref(out), ctypes.byref(pl))
        if rc != OK:
            _raise(rc, pl)
        return bool(out.value)

    def measure(self, target, u):
        try:
            return self.try_measure(target, u)
        except OperationError as e:
            raise RuntimeError("Could not measure bit") from e

    # register.rs:345-377 apply_circuit / measure_circuit on (ops, measurement_targets, conditional_ops)
    def apply_circuit(self, ops, measurement_targets, conditional_ops, us):
        for op in ops:
            self.apply(op)
        us = iter(us)
        measures = [(t, self.measure(t, next(us))) for t in measurement_targets]
        for op, conds in conditional_ops:
            if all(c in measures for c in conds):
                self.apply(op)
        return measures

    # register.rs:409
    def __eq__(self, other):
        return bool(lib().orc_registers_equal(_dp(self.state), _dp(other.state), self._size))

    __hash__ = None
