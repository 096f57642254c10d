"""Host-side mirror of quaru::openqasm (src/openqasm.rs): `run_openqasm(path)` and `u(theta, phi, lambda, target)`.

Parsing (the un-vendored openqasm-parser crate in the reference) is done natively by the library
(qb_qasm_parse, quaru_b200/csrc/qb_qasm.cu); the interpreter loop below follows openqasm.rs:73-139 line by line,
quirks included: `reset q[i]` re-creates the WHOLE register (:124-130), CX across two registers is a run error
(:101-104), `if (c == v)` compares the little-endian value of the classical register (:175-190, panics above
32 bits).  U parameters arrive as f32 and are widened to f64 (:98).
"""
import ctypes

import numpy as np

from . import _ffi
from .operation import cnot, u  # noqa: F401  (u is re-exported: openqasm::u)
from .register import Register


class OpenQASMParseError(Exception):
    """OpenQASMError::OpenQASMParseError (openqasm.rs:27-30)."""

    def __init__(self, msg, line):
        super().__init__(msg)
        self.line = line


class OpenQASMRunError(Exception):
    """OpenQASMError::OpenQASMRunError(String) (openqasm.rs:32-33)."""


class BasicOp:
    """One entry of get_basic_operations(): kind in {'U','CX','Measure','ResetQ','ResetC'}."""

    __slots__ = ("kind", "params", "a", "b", "condition")

    def __init__(self, kind, params, a, b, condition):
        self.kind, self.params, self.a, self.b, self.condition = kind, params, a, b, condition

    def __repr__(self):
        return f"BasicOp({self.kind}, {self.params}, {self.a}, {self.b}, if={self.condition})"


class OpenQASMProgram:
    """What the reference reads from the parser crate: qregs, cregs (name -> size) and the basic operations."""

    def __init__(self, qregs, cregs, ops):
        self.qregs, self.cregs, self._ops = qregs, cregs, ops

    def get_basic_operations(self):
        return list(self._ops)


_KINDS = {_ffi.QB_QASM_U: "U", _ffi.QB_QASM_CX: "CX", _ffi.QB_QASM_MEASURE: "Measure", _ffi.QB_QASM_RESET_Q: "ResetQ",
          _ffi.QB_QASM_RESET_C: "ResetC"}


def parse_openqasm_source(source):
    L = _ffi.lib()
    h = ctypes.c_void_p()
    rc = L.qb_qasm_parse(source.encode(), ctypes.byref(h))
    if rc != _ffi.QB_OK:
        msg, line = _ffi.last_error()
        raise OpenQASMParseError(msg, line)
    try:
        qn = [L.qb_qasm_qreg_name(h, i).decode() for i in range(L.qb_qasm_num_qregs(h))]
        cn = [L.qb_qasm_creg_name(h, i).decode() for i in range(L.qb_qasm_num_cregs(h))]
        qregs = {nm: int(L.qb_qasm_qreg_size(h, i)) for i, nm in enumerate(qn)}
        cregs = {nm: int(L.qb_qasm_creg_size(h, i)) for i, nm in enumerate(cn)}
        n = L.qb_qasm_num_ops(h)
        raw = L.qb_qasm_ops(h)
        ops = []
        for i in range(n):
            o = raw[i]
            kind = _KINDS[o.kind]
            cond = (cn[o.cond_creg], int(o.cond_val)) if o.cond_creg >= 0 else None
            if kind == "U":
                op = BasicOp(kind, tuple(np.float32(x) for x in o.p), (qn[o.reg_a], o.idx_a), None, cond)
            elif kind == "CX":
                op = BasicOp(kind, (), (qn[o.reg_a], o.idx_a), (qn[o.reg_b], o.idx_b), cond)
            elif kind == "Measure":
                op = BasicOp(kind, (), (qn[o.reg_a], o.idx_a), (cn[o.reg_b], o.idx_b), cond)
            elif kind == "ResetQ":
                op = BasicOp(kind, (), (qn[o.reg_a], o.idx_a), None, cond)
            else:
                op = BasicOp(kind, (), (cn[o.reg_a], o.idx_a), None, cond)
            ops.append(op)
        return OpenQASMProgram(qregs, cregs, ops)
    finally:
        L.qb_qasm_free(h)


def parse_openqasm(path):
    with open(path, "r") as f:
        return parse_openqasm_source(f.read())


class OpenQASMResult:
    """struct OpenQASMResult (openqasm.rs:13-21)."""

    def __init__(self, qregs, cregs, program):
        self.qregs, self.cregs, self.program = qregs, cregs, program


def creg_to_u32(creg):
    """openqasm.rs:175-190: creg[0] is the least significant bit; more than 32 bits is a panic."""
    assert len(creg) <= 32, "Registers larger than 32 bits are not allowed"
    res = 0
    for bit in reversed(creg):
        res = (res << 1) | (1 if bit else 0)
    return res


def run_program(program, new_register, apply_u, apply_cx, measure):
    """The interpreter loop of run_openqasm (openqasm.rs:58-141), parameterised over the register backend so
    that tests can run the identical loop on the CPU oracle."""
    qregs = {name: new_register(size) for name, size in program.qregs.items()}
    cregs = {name: [False] * size for name, size in program.cregs.items()}
    for op in program.get_basic_operations():
        if op.condition is not None:
            creg = cregs[op.condition[0]]
            if creg_to_u32(creg) != op.condition[1]:
                continue
        if op.kind == "U":
            p1, p2, p3 = (float(x) for x in op.params)          # `p1 as f64` (openqasm.rs:98)
            apply_u(qregs[op.a[0]], p1, p2, p3, op.a[1])
        elif op.kind == "CX":
            if op.a[0] != op.b[0]:
                raise OpenQASMRunError("The simulator currently does not support applying CX on qubits in two different registers")
            apply_cx(qregs[op.a[0]], op.a[1], op.b[1])
        elif op.kind == "Measure":
            cregs[op.b[0]][op.b[1]] = measure(qregs[op.a[0]], op.a[1])
        elif op.kind == "ResetQ":
            qregs[op.a[0]] = new_register(program.qregs[op.a[0]])   # resets the whole register (openqasm.rs:129)
        else:
            cregs[op.a[0]] = [False] * len(cregs[op.a[0]])
    return OpenQASMResult(qregs, cregs, program)


def run_openqasm_source(source, seed=None, us=None):
    """Run OpenQASM 2.0 source text.  `us`: optional iterable of uniform draws for the measurements (seeded
    parity runs); otherwise one generator seeded with `seed` feeds every register."""
    program = parse_openqasm_source(source)
    rng = np.random.default_rng(seed)
    it = iter(us) if us is not None else None

    def measure(reg, q):
        return reg.measure(q, u=float(next(it)) if it is not None else float(rng.random()))

    return run_program(program, lambda n: Register.new([False] * n),
                       lambda reg, a, b, c, t: reg.apply(u(a, b, c, t)),
                       lambda reg, c, t: reg.apply(cnot(c, t)), measure)


def run_openqasm(openqasm_file, seed=None, us=None):
    """openqasm::run_openqasm(&Path) (openqasm.rs:52)."""
    with open(openqasm_file, "r") as f:
        return run_openqasm_source(f.read(), seed=seed, us=us)
