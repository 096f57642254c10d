#!/usr/bin/env python
"""Shor's algorithm on the CUDA engine -- the workload of /root/reference/examples/shors.rs with its flags
(--number, --n-times, --fast, --use-circuit; shors.rs:13-30): period finding with 3n qubits (slow, shors.rs:180-234),
n+1 qubits with a recycled control qubit (fast, :287-361), or through a QuantumCircuit (:236-285).

    python examples/shors.py --number 15            # BASELINE config C2
    python examples/shors.py --number 21 --fast
"""
import argparse
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from quaru_b200 import Operation, QuantumCircuit, Register, operation as op


def modpow(base, exponent, modulus):                             # math.rs:70
    result = 1
    while exponent > 0:
        if exponent % 2 == 1:
            result = (result * base) % modulus
        exponent //= 2
        base = base * base % modulus
    return result


def limit_denominator(m, n, l):                                  # math.rs:55
    a, b, p, q, r, s, v = n, m % n, 1, 0, m // n, 1, 1
    while 0 < b and q + a // b * s <= l:
        a, b, p, q, r, s, v = b, a % b, r, s, p + a // b * r, q + a // b * s, -v
    t, u = p + (l - q) // s * r, q + (l - q) // s * s
    return (r, s) if 2 * b * u <= n else (t, u)


def u_gate(targets, modulus, a, i):                              # shors.rs:79
    a_pow_mod = modpow(a, 1 << i, modulus)
    return op.to_quantum_gate(lambda x: (x * a_pow_mod) % modulus, targets)


def qft(n):                                                      # shors.rs:363
    m = 1 << n
    w = np.exp(1j * (2.0 * math.pi / m))
    idx = np.arange(m)
    return Operation(w ** np.outer(idx, idx) * (1.0 / math.sqrt(m)), list(range(n)))


def _gates_slow(number, a, n):
    gates = [op.hadamard_transform(list(range(2 * n))), op.not_(2 * n)]
    targets = list(range(2 * n, 3 * n))
    gates += [op.to_controlled(u_gate(targets, number, a, i), i) for i in range(2 * n)]
    gates.append(qft(2 * n))
    return gates


def find_period_slow(number, a, rng):                            # shors.rs:180
    n = math.ceil(math.log2(number + 1))
    reg = Register.new([False] * (3 * n))
    for g in _gates_slow(number, a, n):
        reg.apply(g)
    res = 0
    for i in range(2 * n):
        res |= int(reg.measure(i, u=float(rng.random()))) << i
    return limit_denominator(res, 2 ** (2 * n) - 1, number - 1)[1]


def find_period_circuit(number, a, rng):                         # shors.rs:236
    n = math.ceil(math.log2(number + 1))
    circ = QuantumCircuit()
    for g in _gates_slow(number, a, n):
        circ.add_operation(g)
    for i in range(2 * n):
        circ.add_measurement(i)
    reg = Register.new([False] * (3 * n))
    measures = reg.apply_circuit(circ, us=[float(rng.random()) for _ in range(2 * n)])
    res = 0
    for i, m in measures:
        res |= int(m) << i
    return limit_denominator(res, 2 ** (2 * n) - 1, number - 1)[1]


def find_period_fast(number, a, rng):                            # shors.rs:287
    n = math.ceil(math.log2(number + 1))
    reg = Register.new([False] * (n + 1))
    reg.apply(op.not_(1))
    res, sm = 0, 0.0
    targets = list(range(1, n + 1))
    for i in range(2 * n):
        reg.apply(op.hadamard(0))
        reg.apply(op.to_controlled(u_gate(targets, number, a, 2 * n - 1 - i), 0))
        reg.apply(Operation(np.array([[1, 0], [0, np.exp(1j * (-2.0 * math.pi * sm))]]), [0]))
        reg.apply(op.hadamard(0))
        m = reg.measure(0, u=float(rng.random()))
        sm /= 2.0
        if m:
            reg.apply(op.not_(0))
            res |= 1 << i
            sm += 0.25
    return limit_denominator(res, 2 ** (2 * n) - 1, number - 1)[1]


def shors(number, fast, use_circuit, rng):                       # shors.rs:94-166
    if number % 2 == 0:
        return 2
    for k in range(2, int(math.log2(number)) + 1):
        c = int(number ** (1.0 / k) + 1e-9)
        if c ** k == number:
            return c
    while True:
        a = int(rng.integers(2, number))
        k = math.gcd(a, number)
        if k != 1:
            return k
        r = find_period_circuit(number, a, rng) if use_circuit else (find_period_fast if fast else find_period_slow)(number, a, rng)
        if r % 2 == 0:
            k = modpow(a, r // 2, number)
            if k != number - 1:
                for f in (math.gcd(k - 1, number), math.gcd(k + 1, number)):
                    if f != 1 and f != number:
                        return f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--number", type=int, default=15)
    ap.add_argument("--n-times", type=int, default=1)
    ap.add_argument("--fast", action="store_true")
    ap.add_argument("--use-circuit", action="store_true")
    ap.add_argument("--seed", type=int, default=None)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    times = []
    for _ in range(a.n_times):
        t0 = time.time()
        d1 = shors(a.number, a.fast, a.use_circuit, rng)
        dt = (time.time() - t0) * 1e3
        times.append(dt)
        print(f"The factors of {a.number} are {d1} and {a.number // d1}")
        print(f"Time elapsed: {dt:.0f} ms")
        assert a.number % d1 == 0
    print(f"Total time: {sum(times):.0f} ms after {a.n_times} runs\nAverage time: {sum(times) / len(times):.0f} ms\n"
          f"Best time: {min(times):.0f} ms")


if __name__ == "__main__":
    main()
