#!/usr/bin/env python
"""Grover search on the CUDA engine -- the workload of /root/reference/examples/grovers.rs with its flags
(--target, --list, --statistics; grovers.rs:16-28).  The reference builds the oracle / diffuser as full
2^n x 2^n diagonal matrices (grovers.rs:181-218), which caps it at ~12 qubits; here the same diagonals are handed
over for n <= 12 (bit-identical inputs to the reference) and, with --decomposed (or n > 12), as multi-controlled Z
gates (cnz, operation.rs:520, through qb_apply_controlled: no dense matrix), so that registers far beyond the reference's reach work.

    python examples/grovers.py --target 64                  # n = 7, list = 0..127 (BASELINE config C1)
    python examples/grovers.py --target 123456 --qubits 24 --decomposed
"""
import argparse
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from quaru_b200 import Operation, Register, operation as op


def oracle_operation(regsize, winner, lst):                      # grovers.rs:181
    d = np.ones(1 << regsize)
    for i, item in enumerate(lst):
        if item == winner and i < (1 << regsize):
            d[i] = -1.0
    return Operation(np.diag(d), list(range(regsize)))


def diffusion_operation(regsize):                                # grovers.rs:207
    d = -np.ones(1 << regsize)
    d[0] = 1.0
    return Operation(np.diag(d), list(range(regsize)))


def phase_flip_on(reg, n, index):
    """-1 on basis state `index`: X on the zero bits, cnz over all qubits, X back (decomposed oracle)."""
    zeros = [q for q in range(n) if not (index >> q) & 1]
    for q in zeros:
        reg.apply(op.not_(q))
    reg.apply_controlled(op.pauli_z(0), list(range(1, n)))     # cnz without the 2^n x 2^n matrix
    for q in zeros:
        reg.apply(op.not_(q))


def grovers_algorithm(winner, n, lst=None, decomposed=False, seed=None):      # grovers.rs:121-168
    reg = Register.new([False] * n, seed=seed)
    start = time.time()
    reg.apply_all(op.hadamard(0))
    iters = int(math.floor(math.pi / 4.0 * math.sqrt((1 << n) / 1)))          # iterations_floor, m = 1
    if not decomposed:
        oracle, diffuser = oracle_operation(n, winner, lst), diffusion_operation(n)
    for _ in range(iters):
        if decomposed:
            phase_flip_on(reg, n, winner)
        else:
            reg.apply(oracle)
        reg.apply_all(op.hadamard(0))                                           # diffuser(), grovers.rs:172-176
        if decomposed:
            phase_flip_on(reg, n, 0)          # -(1 - 2|0><0|): same as diag(1,-1,..,-1) up to a global phase
        else:
            reg.apply(diffuser)
        reg.apply_all(op.hadamard(0))
    measured = [reg.measure(i) for i in reversed(range(n))]
    return measured, iters, time.time() - start, reg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--target", type=int, required=True)
    ap.add_argument("--list", type=int, nargs="*", default=None, help="the list to search (default 0..2^n-1)")
    ap.add_argument("--qubits", type=int, default=None)
    ap.add_argument("--decomposed", action="store_true")
    ap.add_argument("--statistics", type=int, default=1, help="number of runs")
    ap.add_argument("--seed", type=int, default=None)
    a = ap.parse_args()
    n = a.qubits or max(2, math.ceil(math.log2(a.target + 1)))
    lst = a.list if a.list is not None else list(range(1 << n)) if n <= 20 else None
    decomposed = a.decomposed or n > 12
    if not decomposed and a.target not in lst:
        print("target is not in the list")            # grovers.rs: returns None when m == 0
        return 1
    hits = 0
    for r in range(a.statistics):
        bits, iters, dt, reg = grovers_algorithm(a.target if decomposed else a.target, n, lst, decomposed,
                                                 None if a.seed is None else a.seed + r)
        result = sum(int(b) << i for i, b in zip(reversed(range(n)), bits))
        found = (lst.index(a.target) if (not decomposed and lst) else a.target)
        hits += result == found
        print(f"Grover's Algorithm executed with the following results:\nTarget: {a.target}\nResult: {result} ({bits})\n"
              f"Iterations of oracle + diffuser: {iters}\nTime elapsed: {dt * 1e3:.0f} ms "
              f"({reg.stat(3)} gates, {reg.stat(2)} passes, {reg.stat(1)} kernel launches)")
    if a.statistics > 1:
        print(f"{hits}/{a.statistics} runs found the target")
    return 0


if __name__ == "__main__":
    sys.exit(main())
