OPENQASM 2.0;
include "qelib1.inc";
// a Bell pair on q[0], q[1] of a three-qubit register
qreg q[3];
h q[0];
cx q[0], q[1];
