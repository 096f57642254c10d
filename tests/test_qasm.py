"""CPU tests of the OpenQASM 2.0 front end (native parser behind qb_qasm_parse) and of the interpreter loop
(quaru_b200/openqasm.py, mirror of src/openqasm.rs:73-139) run over the CPU oracle."""
import math

import numpy as np
import pytest

from oracle import gates as og
from oracle.oracle import OracleRegister
from tests import qasm_programs as P


def _oq():
    from quaru_b200 import openqasm
    return openqasm


def run_on_oracle(source, us):
    """The product's interpreter loop driving OracleRegister (the reference's Register restated)."""
    oq = _oq()
    it = iter(us)
    return oq.run_program(oq.parse_openqasm_source(source), lambda n: OracleRegister.new([False] * n),
                          lambda reg, a, b, c, t: reg.apply(og.u(a, b, c, t)),
                          lambda reg, c, t: reg.apply(og.cnot(c, t)),
                          lambda reg, q: reg.measure(q, float(next(it))))


def test_bell_program_like_the_reference_fixture():
    oq = _oq()
    prog = oq.parse_openqasm_source(P.BELL3)
    assert prog.qregs == {"q": 3} and prog.cregs == {}
    ops = prog.get_basic_operations()
    assert [o.kind for o in ops] == ["U", "CX"]
    assert ops[0].a == ("q", 0) and ops[1].a == ("q", 0) and ops[1].b == ("q", 1)
    th, ph, la = ops[0].params
    assert th == np.float32(math.pi / 2) and ph == np.float32(0.0) and la == np.float32(math.pi)   # f32 parameters
    res = run_on_oracle(P.BELL3, [])
    p = np.abs(res.qregs["q"].state) ** 2
    assert abs(p[0] - 0.5) < 1e-7 and abs(p[3] - 0.5) < 1e-7 and p.sum() - p[0] - p[3] < 1e-12


def test_qelib_include_broadcast_barrier_and_measure():
    oq = _oq()
    prog = oq.parse_openqasm_source(P.QFT4)
    ops = prog.get_basic_operations()
    assert [o.kind for o in ops[-4:]] == ["Measure"] * 4
    assert [(o.a, o.b) for o in ops[-4:]] == [(("q", i), ("c", i)) for i in range(4)]
    assert all(o.kind in ("U", "CX") for o in ops[:-4])
    # cu1 expands to u1, cx, u1, cx, u1 = 3 U + 2 CX
    assert sum(o.kind == "CX" for o in ops) == 2 * 6
    res = run_on_oracle(P.QFT4, [0.1, 0.6, 0.3, 0.9])
    assert res.cregs["c"] in ([False, True, False, True], [True, False, True, False]) or len(res.cregs["c"]) == 4


def test_expressions_and_nested_parameters():
    oq = _oq()
    src = """OPENQASM 2.0;
    gate g(a, b) q { U(a * 2 - b / 4, -(a ^ 2), sqrt(b) + ln(exp(1)) + cos(0) + tan(0)) q; }
    gate outer(x) q { g(x + 1, 16) q; }
    qreg q[1];
    outer(0.5) q[0];
    U(2 ^ 3 ^ 2, -pi, 1e-1 + .5) q[0];
    """
    ops = oq.parse_openqasm_source(src).get_basic_operations()
    a, b = 1.5, 16.0
    want = (a * 2 - b / 4, -(a ** 2), math.sqrt(b) + 1.0 + 1.0 + 0.0)
    assert tuple(float(x) for x in ops[0].params) == tuple(float(np.float32(x)) for x in want)
    assert tuple(float(x) for x in ops[1].params) == tuple(float(np.float32(x)) for x in (2.0 ** 9, -math.pi, 0.6))


def test_conditions_reset_and_interpreter_quirks():
    oq = _oq()
    res = run_on_oracle(P.TELEPORT, [0.2, 0.7, 0.5])
    assert set(res.cregs) == {"c0", "c1", "c2"}
    prog = oq.parse_openqasm_source(P.TELEPORT)
    conds = [o.condition for o in prog.get_basic_operations() if o.condition]
    assert ("c0", 1) in conds and ("c1", 1) in conds
    # reset q[i] re-creates the WHOLE register (openqasm.rs:124-130)
    src = 'OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[2];\ncreg c[2];\nx q[0];\nx q[1];\nreset q[0];\nmeasure q -> c;\n'
    res = run_on_oracle(src, [0.5, 0.5])
    assert res.cregs["c"] == [False, False]
    # CX across two registers is a run error (openqasm.rs:101-104)
    src = 'OPENQASM 2.0;\nqreg a[1];\nqreg b[1];\nCX a[0], b[0];\n'
    with pytest.raises(oq.OpenQASMRunError):
        run_on_oracle(src, [])
    # creg little-endian value (openqasm.rs:175-190)
    assert oq.creg_to_u32([True, False, True]) == 5
    with pytest.raises(AssertionError):
        oq.creg_to_u32([False] * 33)


def test_mixed_program_parses_and_runs():
    res = run_on_oracle(P.MIXED, [0.3, 0.8, 0.45])
    assert abs(np.vdot(res.qregs["a"].state, res.qregs["a"].state).real - 1.0) < 1e-12
    assert len(res.cregs["r"]) == 3 and set(res.qregs) == {"a", "spare"}
    src = 'OPENQASM 2.0;\nqreg q[1];\ncreg c[2];\nU(pi,0,pi) q[0];\nmeasure q[0] -> c[1];\nreset c;\n'
    assert run_on_oracle(src, [0.5]).cregs["c"] == [False, False]          # ResetC (openqasm.rs:131-137)


@pytest.mark.parametrize("src,fragment", [
    ("OPENQASM 3.0;\nqreg q[1];", "2.0"),
    ("qreg q[1];", "OPENQASM"),
    ("OPENQASM 2.0;\nqreg q[2]\nU(0,0,0) q[0];", "expected ';'"),
    ("OPENQASM 2.0;\nqreg q[2];\nfoo q[0];", "not defined"),
    ("OPENQASM 2.0;\nqreg q[2];\nU(0,0,0) r[0];", "undeclared quantum register"),
    ("OPENQASM 2.0;\nqreg q[2];\nU(0,0,0) q[2];", "out of range"),
    ("OPENQASM 2.0;\nqreg q[2];\nU(0,0) q[0];", "3 parameter"),
    ("OPENQASM 2.0;\nqreg q[2];\nCX q[0];", "2 qubit"),
    ("OPENQASM 2.0;\nqreg q[2];\nCX q[1], q[1];", "same qubit"),
    ("OPENQASM 2.0;\nqreg q[2];\nqreg q[3];", "already declared"),
    ("OPENQASM 2.0;\nqreg q[2];\ncreg c[3];\nmeasure q -> c;", "different size"),
    ("OPENQASM 2.0;\nqreg q[2];\nif (c == 1) U(0,0,0) q[0];", "undeclared classical register"),
    ("OPENQASM 2.0;\nopaque magic q;\nqreg q[1];\nmagic q[0];", "opaque"),
    ("OPENQASM 2.0;\ngate g a { h a; }\nqreg q[1];", "not defined"),
    ("OPENQASM 2.0;\ninclude \"other.inc\";", "qelib1.inc"),
    ("OPENQASM 2.0;\nqreg q[1];\nU(0,0,zz) q[0];", "unknown identifier"),
    ("OPENQASM 2.0;\nqreg q[1];\nU(0,0,0) q[0] @", "unexpected character"),
    ("OPENQASM 2.0;\nqreg q[99999999999999999999];", "does not fit 32 bits"),
    ("OPENQASM 2.0;\nqreg q[1];\ncreg c[1];\nif (c == 99999999999) U(0,0,0) q[0];", "does not fit 32 bits"),
    ("OPENQASM 2.0;\ngate g(a,a) x { U(a,0,0) x; }\nqreg q[1];", "used twice"),
    ("OPENQASM 2.0;\ngate g x,x { CX x,x; }\nqreg q[1];", "used twice"),
    ("OPENQASM 2.0;\nqreg q[1];\ngate g a { barrier a", "unterminated barrier"),
    ("OPENQASM 2.0;\nqreg q[1];\ngate a x { a x; }", "not defined"),          # no recursion: a gate sees earlier gates only
])
def test_parse_errors(src, fragment):
    oq = _oq()
    with pytest.raises(oq.OpenQASMParseError) as e:
        oq.parse_openqasm_source(src)
    assert fragment in str(e.value), str(e.value)
    assert e.value.line >= 1


def test_u_matches_the_reference_formula():
    """openqasm::u (openqasm.rs:145-163) on the product side equals the oracle restatement."""
    from quaru_b200 import operation as op
    for th, ph, la in [(0.3, 1.1, -0.7), (math.pi / 2, 0.0, math.pi), (0.0, 0.0, 2.5)]:
        assert (op.u(th, ph, la, 1).matrix() == og.u(th, ph, la, 1)[0]).all()


def test_hostile_programs_are_rejected_not_crashed_on():
    """Recursion and expansion guards of the native parser: inputs that would otherwise overflow the C++ stack or
    exhaust memory come back as parse errors."""
    oq = _oq()
    head = "OPENQASM 2.0;\nqreg q[1];\n"
    for src, fragment in [
        (head + "U(" + "(" * 100000 + "1" + ")" * 100000 + ",0,0) q[0];", "nested too deeply"),
        (head + "U(" + "-" * 200000 + "1,0,0) q[0];", "nested too deeply"),
        (head + "U(" + "+".join(["1"] * 100000) + ",0,0) q[0];", "too long"),
        (head + "gate g0 a { U(0,0,0) a; U(0,0,0) a; }\n" +
         "".join(f"gate g{i} a {{ g{i - 1} a; g{i - 1} a; }}\n" for i in range(1, 40)) + "g39 q[0];", "2^24"),
    ]:
        with pytest.raises(oq.OpenQASMParseError) as e:
            oq.parse_openqasm_source(src)
        assert fragment in str(e.value), str(e.value)
    # generous limits: a 4000-term sum and 150 nested calls are fine
    prog = oq.parse_openqasm_source(head + "U(" + "+".join(["1"] * 4000) + ",0,0) q[0];")
    assert float(prog.get_basic_operations()[0].params[0]) == 4000.0
    oq.parse_openqasm_source(head + "U(" + "sin(" * 150 + "1" + ")" * 150 + ",0,0) q[0];")
