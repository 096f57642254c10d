#!/usr/bin/env python
"""Quantum teleportation and superdense coding on the CUDA engine
(/root/reference/examples/quantum_teleportation.rs:83-140, examples/superdense_coding.rs:28-80).

    python examples/protocols.py teleport -a 0.8 -b 0.0 -c 0.0 -d -0.6 [--circuit]
    python examples/protocols.py superdense -a 1 -b 0
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from quaru_b200 import QuantumCircuit, Register, operation as op


def state_of_qubit(state, n):
    idx = np.arange(state.shape[0])
    beta, alpha = state[((idx >> n) & 1) == 1].sum(), state[((idx >> n) & 1) == 0].sum()
    return np.array([alpha, beta]) / np.sqrt(abs(alpha) ** 2 + abs(beta) ** 2)


def teleport(q0, use_circuit):
    reg = Register.new([False] * 3)
    reg.set_state(np.kron(np.kron([1, 0], [1, 0]), q0))
    if use_circuit:                                              # quantum_teleportation.rs:63-81
        c = QuantumCircuit()
        for g in (op.hadamard(1), op.cnot(1, 2), op.cnot(0, 1), op.hadamard(0)):
            c.add_operation(g)
        c.add_measurement(0)
        c.add_measurement(1)
        c.add_conditional_operation(op.not_(2), [(1, True)])
        c.add_conditional_operation(op.pauli_z(2), [(0, True)])
        reg.apply_circuit(c)
    else:                                                        # quantum_teleportation.rs:83-140
        reg.apply(op.hadamard(1)).apply(op.cnot(1, 2)).apply(op.cnot(0, 1)).apply(op.hadamard(0))
        m0, m1 = reg.measure(0), reg.measure(1)
        if m1:
            reg.apply(op.not_(2))
        if m0:
            reg.apply(op.pauli_z(2))
    return state_of_qubit(reg.state, 2)


def superdense(a, b):                                            # superdense_coding.rs:28-80
    reg = Register.new([False, False])
    reg.apply(op.hadamard(0)).apply(op.cnot(0, 1))
    if b:
        reg.apply(op.not_(0))
    if a:
        reg.apply(op.pauli_z(0))
    reg.apply(op.cnot(0, 1)).apply(op.hadamard(0))
    return reg.measure(0), reg.measure(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("protocol", choices=["teleport", "superdense"])
    ap.add_argument("-a", type=float, default=0.8)
    ap.add_argument("-b", type=float, default=0.0)
    ap.add_argument("-c", type=float, default=0.0)
    ap.add_argument("-d", type=float, default=-0.6)
    ap.add_argument("--circuit", action="store_true")
    a = ap.parse_args()
    if a.protocol == "teleport":
        q0 = np.array([complex(a.a, a.b), complex(a.c, a.d)])
        got = teleport(q0, a.circuit)
        print(f"Expected:\n{q0}\n\nGot:\n{got}\n")
        print("Correct" if np.abs(got - q0).max() < 1e-8 else "WRONG!")
    else:
        x, y = superdense(bool(a.a), bool(a.b))
        print(f"sent ({bool(a.a)}, {bool(a.b)}) received ({x}, {y})")


if __name__ == "__main__":
    main()
