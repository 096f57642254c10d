#!/usr/bin/env python
"""Runs an OpenQASM 2.0 file on the CUDA engine and prints the probabilities of every quantum register
(/root/reference/examples/openqasm_example.rs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quaru_b200 import openqasm

path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "bell.qasm")
registers = openqasm.run_openqasm(path)
for name, qreg in registers.qregs.items():
    print(f"\nQuantum register: {name}")
    qreg.print_probabilities()
    print()
for name, creg in registers.cregs.items():
    print(f"Classical register {name}: {creg}")
