"""OpenQASM 2.0 test programs (written for this repository; the reference's only fixture,
examples/openqasm_example.qasm, is an H + CX Bell pair on q[0], q[1] of a 3-qubit register built from
user-defined u2 / h / cx gates -- BELL3 below exercises the same constructs)."""

BELL3 = """OPENQASM 2.0;
// user-defined gates on top of the two built-ins, as in the reference example
gate u2(phi, lambda) q { U(pi / 2, phi, lambda) q; }
gate h a { u2(0, pi) a; }
gate cx c, t { CX c, t; }
qreg q[3];
h q[0];
cx q[0], q[1];
"""

QFT4 = """OPENQASM 2.0;
include "qelib1.inc";
qreg q[4];
creg c[4];
x q[0];
x q[2];
barrier q;
h q[0];
cu1(pi/2) q[1],q[0];
h q[1];
cu1(pi/4) q[2],q[0];
cu1(pi/2) q[2],q[1];
h q[2];
cu1(pi/8) q[3],q[0];
cu1(pi/4) q[3],q[1];
cu1(pi/2) q[3],q[2];
h q[3];
measure q -> c;
"""

TELEPORT = """OPENQASM 2.0;
include "qelib1.inc";
qreg q[3];
creg c0[1];
creg c1[1];
creg c2[1];
u3(0.3, 0.2, 0.1) q[0];
h q[1];
cx q[1], q[2];
barrier q;
cx q[0], q[1];
h q[0];
measure q[0] -> c0[0];
measure q[1] -> c1[0];
if (c0 == 1) z q[2];
if (c1 == 1) x q[2];
measure q[2] -> c2[0];
"""

MIXED = """OPENQASM 2.0;
include "qelib1.inc";
gate majority a, b, c { cx c, b; cx c, a; ccx a, b, c; }
gate rot(t) a { rx(t / 2) a; ry(-t) a; rz(sin(t) ^ 2 + 0.25e1) a; }
qreg a[6];
qreg spare[2];
creg r[3];
h a;
h spare;
rot(pi / 3) a[1];
cx a[0], a[3];
cx a[1], a[4];
majority a[0], a[4], a[5];
swap a[3], a[5];
crz(0.7) a[2], a[3];
ch a[0], a[1];
measure a[3] -> r[0];
measure a[4] -> r[1];
measure a[5] -> r[2];
if (r == 5) rot(-1.5) a[0];
if (r == 2) x spare[1];
"""
