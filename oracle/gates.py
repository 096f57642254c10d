"""Oracle-side restatement of Quaru's gate constructors (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/operation.rs:340-536, src/openqasm.rs:145-163 and the
matrix builders of examples/{grovers,shors}.rs.  A gate is a pair
``(matrix: complex128 ndarray of shape (2^k, 2^k), targets: list[int])`` with
matrix index bit j <-> targets[j] (register.rs:168-173).

The product has its own constructors (quaru_b200/operation.py and
include/quaru_b200.hpp); tests compare the two.  Nothing in quaru_b200/ imports
this module.
"""
import math

import numpy as np

C = np.complex128


def _real(m):
    return np.array(m, dtype=np.float64).astype(C)


def kron(a, b):
    """ndarray::linalg::kron(a, b): out[i*rb + k, j*cb + l] = a[i,j] * b[k,l]."""
    return np.kron(a, b)


# --- operation.rs:340-536 -------------------------------------------------------------------

def identity(target):                                   # :340
    return _real([[1.0, 0.0], [0.0, 1.0]]), [target]


def hadamard(target):                                   # :350  FRAC_1_SQRT_2 * [[1,1],[1,-1]]
    return (math.sqrt(0.5) * np.array([[1.0, 1.0], [1.0, -1.0]])).astype(C), [target]


def hadamard_transform(targets):                        # :358  kron(H_t, acc) for t in targets[1:]
    m = hadamard(targets[0])[0]
    for t in targets[1:]:
        m = kron(hadamard(t)[0], m)
    return m, list(targets)


def cnot(control, target):                              # :372  targets = [target, control]
    return _real([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]]), [target, control]


def to_quantum_gate(f, targets):                        # :386  column c = e_{f(c)}
    n = 1 << len(targets)
    m = np.zeros((n, n), dtype=C)
    for c in range(n):
        m[f(c), c] = 1.0                                # int_to_state panics if f(c) >= n
    return m, list(targets)


def to_controlled(op, control):                         # :406  I (+) M, control appended as MSB
    mat, targets = op
    old = mat.shape[0]
    m = np.zeros((2 * old, 2 * old), dtype=C)
    m[:old, :old] = np.eye(old)
    m[old:, old:] = mat
    return m, list(targets) + [control]


def swap(t1, t2):                                       # :427
    return _real([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]), [t1, t2]


def phase(target):                                      # :443  diag(1, i)
    return np.array([[1, 0], [0, 1j]], dtype=C), [target]


def not_(target):                                       # :458
    return _real([[0, 1], [1, 0]]), [target]


def pauli_y(target):                                    # :468
    return np.array([[0, -1j], [1j, 0]], dtype=C), [target]


def pauli_z(target):                                    # :481
    return _real([[1, 0], [0, -1]]), [target]


def cnx(controls, target):                              # :491  targets = [target, controls...]
    targets = [target] + list(controls)
    n = 1 << len(targets)
    m = np.zeros((n, n))
    for i in range(n - 2):
        m[i, i] = 1.0
    m[n - 1, n - 2] = 1.0
    m[n - 2, n - 1] = 1.0
    return m.astype(C), targets


def cnz(controls, target):                              # :520
    targets = [target] + list(controls)
    n = 1 << len(targets)
    m = np.eye(n)
    m[n - 1, n - 1] = -1.0
    return m.astype(C), targets


# --- openqasm.rs:145-163 --------------------------------------------------------------------

def u(theta, phi, lam, target):
    i = 1j
    return np.array([
        [np.exp(-i * (phi + lam) / 2.0) * np.cos(theta / 2.0), -np.exp(-i * (phi - lam) / 2.0) * np.sin(theta / 2.0)],
        [np.exp(i * (phi - lam) / 2.0) * np.sin(theta / 2.0), np.exp(i * (phi + lam) / 2.0) * np.cos(theta / 2.0)],
    ], dtype=C), [target]


# --- gates the BASELINE configs build through Operation::new ------------

def rz(theta, target):
    """diag(e^{-i theta/2}, e^{i theta/2}) == u(0, 0, theta, target)."""
    return np.array([[np.exp(-0.5j * theta), 0], [0, np.exp(0.5j * theta)]], dtype=C), [target]


def cz(control, target):
    return cnz([control], target)


def cphase(angle, control, target):
    """to_controlled(diag(1, e^{i angle}) on target, control) (QFT cascade)."""
    return to_controlled((np.array([[1, 0], [0, np.exp(1j * angle)]], dtype=C), [target]), control)


# --- examples/grovers.rs:181-218 ------------------------------------------------------------

def grover_oracle(regsize, winner, lst):                # :181
    n = 1 << regsize
    d = np.ones(n)
    for i, item in enumerate(lst):
        if item == winner and i < n:
            d[i] = -1.0
    return np.diag(d).astype(C), list(range(regsize))


def grover_diffusion(regsize):                          # :207
    n = 1 << regsize
    d = -np.ones(n)
    d[0] = 1.0
    return np.diag(d).astype(C), list(range(regsize))


# --- examples/shors.rs:79-92, 363-373; src/math.rs:55-82 --------------------------------------

def modpow(base, exponent, modulus):                    # math.rs:70 (u32 arithmetic, wraps are not hit for n<=5)
    result = 1
    while exponent > 0:
        if exponent % 2 == 1:
            result = (result * base) % modulus
        exponent //= 2
        base = base * base % modulus
    return result


def limit_denominator(m, n, l):                         # math.rs:55
    a, b, p, q, r, s, v = n, m % n, 1, 0, m // n, 1, 1
    while 0 < b and q + a // b * s <= l:
        a, b, p, q, r, s, v = b, a % b, r, s, p + a // b * r, q + a // b * s, -v
    t, u_ = p + (l - q) // s * r, q + (l - q) // s * s
    return (r, s) if 2 * b * u_ <= n else (t, u_)


def shor_u_gate(targets, modulus, a, i):                # shors.rs:79
    a_pow_mod = modpow(a, 1 << i, modulus)
    return to_quantum_gate(lambda x: (x * a_pow_mod) % modulus, targets)


def qft(n):                                             # shors.rs:363  w^(i*j)/sqrt(m), w = e^{2 pi i/m}
    m = 1 << n
    w = np.exp(1j * (2.0 * math.pi / m))
    idx = np.arange(m)
    return (w ** np.outer(idx, idx) * (1.0 / math.sqrt(m))).astype(C), list(range(n))
