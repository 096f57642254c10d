"""Host-side mirror of quaru::register::Register (src/register.rs:23-413) over the C ABI.

Same method names, argument meaning and error behaviour as the reference; the state is a
device-resident Complex<f64> buffer owned by the CUDA library, so the reference's public `state`
field becomes a download (`state` property) / upload (`set_state`).  `measure` draws its uniform
number from the register's RNG (seedable: the reference uses an unseeded thread_rng,
register.rs:317-318) or takes it explicitly through `u=`.

There is no CPU path here: every method calls libquaru_b200.so.
"""
import ctypes
import os

import numpy as np

from . import _ffi
from .operation import Operation


class OperationError(Exception):
    """enum OperationError (register.rs:10-20)."""

    def __init__(self, kind, payload):
        super().__init__(f"{kind}({payload})")
        self.kind = kind
        self.payload = payload

    def __eq__(self, other):
        return isinstance(other, OperationError) and (self.kind, self.payload) == (other.kind, other.payload)

    __hash__ = None


class EngineError(RuntimeError):
    """CUDA / NCCL / out-of-memory failures: no reference equivalent (the Rust shim panics)."""


_KIND = {_ffi.QB_ERR_INVALID_TARGET: "InvalidTarget", _ffi.QB_ERR_INVALID_DIMENSIONS: "InvalidDimensions",
         _ffi.QB_ERR_INVALID_ARITY: "InvalidArity", _ffi.QB_ERR_INVALID_QUBIT: "InvalidQubit"}


def _check(rc):
    if rc == _ffi.QB_OK:
        return
    msg, payload = _ffi.last_error()
    if rc in _KIND:
        raise OperationError(_KIND[rc], payload)
    if rc == _ffi.QB_ERR_INVALID_ARGUMENT:
        raise IndexError(msg)                      # the reference panics (index out of bounds)
    raise EngineError(f"quaru_b200 error {rc}: {msg}")


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class Register:
    """A quantum register of N qubits whose 2^N amplitudes live in B200 HBM."""

    def __init__(self, handle, seed=None):
        self._h = handle
        self._L = _ffi.lib()
        self._rng = np.random.default_rng(seed)

    # ---- constructors -------------------------------------------------------------------------
    @classmethod
    def new(cls, input_bits, seed=None):
        """Register::new(&[bool]) (register.rs:41): input_bits[q] is qubit q."""
        idx = sum((1 << q) for q, b in enumerate(input_bits) if b)
        return cls.from_int(len(input_bits), idx, seed=seed)

    @classmethod
    def from_int(cls, n, k, seed=None):
        """Register::from_int(n, k) (register.rs:58)."""
        h = ctypes.c_void_p()
        _check(_ffi.lib().qb_create(n, k, ctypes.byref(h)))
        return cls(h, seed)

    @classmethod
    def try_from_qubits(cls, input_bits, seed=None):
        """Register::try_from_qubits (register.rs:81): OperationError('InvalidQubit', i) on bad input."""
        for i, q in enumerate(input_bits):
            if np.asarray(q).size != 2:            # is_qubit length check (register.rs:121)
                raise OperationError("InvalidQubit", i)
        n = len(input_bits)
        q = np.ascontiguousarray(np.array([np.asarray(x).reshape(2) for x in input_bits], dtype=np.complex128)
                                 if n else np.zeros((0, 2), dtype=np.complex128))
        h = ctypes.c_void_p()
        _check(_ffi.lib().qb_create_product(_dp(q) if n else None, n, ctypes.byref(h)))
        return cls(h, seed)

    @classmethod
    def from_qubits(cls, input_bits, seed=None):
        try:
            return cls.try_from_qubits(input_bits, seed=seed)
        except OperationError as e:
            raise RuntimeError("Incorrect input qubits") from e          # register.rs:75

    @classmethod
    def sharded(cls, n, k, rank, world, unique_id, device=-1, seed=None):
        """Rank `rank` of a register sharded over `world` GPUs, driving CUDA device `device`
        (no reference equivalent)."""
        h = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        _check(_ffi.lib().qb_create_sharded(n, k, rank, world, buf, device, ctypes.byref(h)))
        return cls(h, seed)

    def clone(self):
        """#[derive(Clone)] (register.rs:23)."""
        h = ctypes.c_void_p()
        _check(self._L.qb_clone(self._h, ctypes.byref(h)))
        r = Register(h)
        r._rng = np.random.default_rng(self._rng.integers(1 << 63))
        return r

    def close(self):
        if getattr(self, "_h", None):
            self._L.qb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- gates --------------------------------------------------------------------------------
    def try_apply_fast(self, op):
        """Register::try_apply_fast (register.rs:140).  Returns self (chaining, as in the reference)."""
        m = op._matrix
        k = len(op._targets)
        if m.shape != (1 << k, 1 << k):            # cannot be produced through Operation::new
            raise OperationError("InvalidDimensions", m.shape[0])
        t = (ctypes.c_uint32 * max(k, 1))(*op._targets)
        _check(self._L.qb_apply(self._h, _dp(m), t, k))
        return self

    try_apply = try_apply_fast                     # register.rs:193: same contract, same errors

    def apply(self, op):
        """Register::apply (register.rs:134): panics with "Could not apply operation"."""
        try:
            return self.try_apply_fast(op)
        except OperationError as e:
            raise RuntimeError("Could not apply operation") from e

    def apply_controlled(self, op, controls):
        """Extension: `op` applied where every qubit in `controls` is 1, without building the
        2^(k+c) x 2^(k+c) matrix of operation::to_controlled / cnx / cnz."""
        k = len(op._targets)
        t = (ctypes.c_uint32 * max(k, 1))(*op._targets)
        c = (ctypes.c_uint32 * max(len(controls), 1))(*[int(x) for x in controls])
        try:
            _check(self._L.qb_apply_controlled(self._h, _dp(op._matrix), t, k, c, len(controls)))
        except OperationError as e:
            raise RuntimeError("Could not apply operation") from e
        return self

    def try_apply_all(self, operation):
        """Register::try_apply_all (register.rs:267): InvalidArity unless the operation is unary."""
        if operation.arity() != 1:
            raise OperationError("InvalidArity", operation.arity())
        _check(self._L.qb_apply_all(self._h, _dp(operation._matrix)))
        return self

    def apply_all(self, operation):
        try:
            return self.try_apply_all(operation)
        except OperationError as e:
            raise RuntimeError("Could not apply operation") from e

    def apply_batch(self, ops):
        """The gate loop of apply_circuit (register.rs:346-350) as one C-ABI call."""
        n = len(ops)
        arr = (_ffi.QbGate * max(n, 1))()
        keep = []
        for i, op in enumerate(ops):
            t = (ctypes.c_uint32 * max(len(op._targets), 1))(*op._targets)
            keep.append(t)
            arr[i].matrix = _dp(op._matrix)
            arr[i].targets = t
            arr[i].k = len(op._targets)
        _check(self._L.qb_apply_batch(self._h, arr, n))
        return self

    # ---- measurement --------------------------------------------------------------------------
    def seed(self, seed):
        self._rng = np.random.default_rng(seed)

    def probabilities(self, target):
        p0, p1 = ctypes.c_double(), ctypes.c_double()
        _check(self._L.qb_probabilities(self._h, target, ctypes.byref(p0), ctypes.byref(p1)))
        return p0.value, p1.value

    def try_measure(self, target, u=None):
        """Register::try_measure (register.rs:297).  `u` in [0,1) overrides the register's RNG."""
        if u is None:
            u = self._rng.random()
        out = ctypes.c_int32()
        _check(self._L.qb_measure(self._h, target, float(u), ctypes.byref(out)))
        return bool(out.value)

    def measure(self, target, u=None):
        try:
            return self.try_measure(target, u)
        except OperationError as e:
            raise RuntimeError("Could not measure bit") from e           # register.rs:292

    # ---- circuits (register.rs:345-377) -------------------------------------------------------
    def apply_circuit(self, circuit, us=None):
        self.apply_batch([Operation(op._matrix, op._targets) for op in circuit.get_operations()])
        measures = []
        if len(circuit.get_measurement_targets()) > 0:
            measures = self.measure_circuit(circuit, us)
        for op, conds in circuit.get_conditional_operations():
            if all(c in measures for c in conds):
                self.apply(op)
        return measures

    def measure_circuit(self, circuit, us=None):
        it = iter(us) if us is not None else None
        return [(t, self.measure(t, next(it) if it is not None else None)) for t in circuit.get_measurement_targets()]

    # ---- state access (`pub state`, print_*, PartialEq) -----------------------------------------
    def size(self):
        return int(self._L.qb_num_qubits(self._h))

    @property
    def state(self):
        """Download of the reference's `pub state` (register.rs:35) as a (2^n,) complex128 vector."""
        return self.read(0, 1 << self.size())

    def read(self, offset, count):
        out = np.zeros(count, dtype=np.complex128)
        _check(self._L.qb_read_state(self._h, offset, count, _dp(out)))
        return out

    def set_state(self, state, offset=0):
        s = np.ascontiguousarray(state, dtype=np.complex128).reshape(-1)
        _check(self._L.qb_write_state(self._h, offset, s.shape[0], _dp(s)))
        return self

    def reset(self, k=0):
        """Back to the basis state |k> in place (what `*reg = Register::from_int(n, k)` does in the reference's
        drivers, without reallocating 16 * 2^n bytes); pending gates are dropped."""
        _check(self._L.qb_reset(self._h, int(k)))
        return self

    def save_state(self, path):
        """Streams this rank's shard to a raw state file (header + interleaved f64 pairs, see include/quaru_b200.h);
        the device-resident replacement for serialising the reference's `pub state` (register.rs:35)."""
        _check(self._L.qb_save_state(self._h, os.fsencode(path)))
        return self

    def load_state(self, path):
        """Restores a shard written by `save_state` into this register (same n / world / rank); pending gates are
        dropped.  Raises EngineError (QB_ERR_IO) for a missing, truncated or corrupt file."""
        _check(self._L.qb_load_state(self._h, os.fsencode(path)))
        return self

    def norm2(self):
        v = ctypes.c_double()
        _check(self._L.qb_norm2(self._h, ctypes.byref(v)))
        return v.value

    def max_abs_diff(self, other):
        v = ctypes.c_double()
        _check(self._L.qb_max_abs_diff(self._h, other._h, ctypes.byref(v)))
        return v.value

    def l2_diff(self, other):
        v = ctypes.c_double()
        _check(self._L.qb_l2_diff2(self._h, other._h, ctypes.byref(v)))
        return float(np.sqrt(v.value))

    def __eq__(self, other):
        """impl PartialEq (register.rs:409-413): all |a_i - b_i| < 1e-8, evaluated on the device."""
        return isinstance(other, Register) and self.size() == other.size() and self.max_abs_diff(other) < 1e-8

    __hash__ = None

    def _lines(self, fmt, nonzero_only=False, chunk=1 << 20):
        n = self.size()
        for off in range(0, 1 << n, chunk):
            part = self.read(off, min(chunk, (1 << n) - off))
            for j, s in enumerate(part):
                p = s.real * s.real + s.imag * s.imag
                if nonzero_only and not p > 0.0:
                    continue
                yield fmt(off + j, n, s, p)

    def print_probabilities(self):              # register.rs:379
        for line in self._lines(lambda i, n, s, p: f"{i:0{n}b}: {p * 100.0}%"):
            print(line)

    def print_nonzero_probabilities(self):      # register.rs:386
        for line in self._lines(lambda i, n, s, p: f"{i:0{n}b}: {p * 100.0}%", nonzero_only=True):
            print(line)

    def print_state(self):                      # register.rs:396
        for line in self._lines(lambda i, n, s, p: f"{i:0{n}b}: {s.real}{s.imag:+}i"):
            print(line)

    # ---- engine controls (no reference equivalent) ----------------------------------------------
    def set_option(self, option, value):
        _check(self._L.qb_set_option(self._h, option, value))
        return self

    def stat(self, which):
        v = ctypes.c_int64()
        _check(self._L.qb_get_stat(self._h, which, ctypes.byref(v)))
        return v.value

    def flush(self):
        _check(self._L.qb_flush(self._h))

    def sync(self):
        _check(self._L.qb_sync(self._h))

    def timer_start(self):
        _check(self._L.qb_timer_start(self._h))

    def timer_stop(self):
        ms = ctypes.c_float()
        _check(self._L.qb_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value
