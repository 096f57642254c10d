"""Seeded synthetic circuits of the BASELINE configs, as lists of oracle gates
``(matrix, targets)``.  Used identically by the oracle and (through quaru_b200.Operation) the engine."""
import numpy as np

from oracle import gates as og


def random_layer(n, rng):
    """One layer of config C3: per qubit H (p=1/2) or RZ(theta~U[0,2pi)); then a uniformly random perfect
    matching, each pair CNOT or CZ (p=1/2) with random orientation."""
    ops = []
    for q in range(n):
        if rng.random() < 0.5:
            ops.append(og.hadamard(q))
        else:
            ops.append(og.rz(float(rng.uniform(0.0, 2.0 * np.pi)), q))
    perm = [int(x) for x in rng.permutation(n)]
    for a, b in zip(perm[0::2], perm[1::2]):
        if rng.random() < 0.5:
            ops.append(og.cnot(a, b))
        else:
            ops.append(og.cz(a, b))
    return ops


def random_circuit(n, depth, seed):
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(depth):
        ops += random_layer(n, rng)
    return ops


def qft_circuit(n):
    """Textbook QFT (config C4): H + controlled-phase cascade + final swaps."""
    ops = []
    for t in range(n - 1, -1, -1):
        ops.append(og.hadamard(t))
        for c in range(t - 1, -1, -1):
            ops.append(og.cphase(2.0 * np.pi / (1 << (t - c + 1)), c, t))
    for a in range(n // 2):
        ops.append(og.swap(a, n - 1 - a))
    return ops


def mixed_circuit(n, count, seed):
    """Adversarial mix for the planner: dense 1q/2q with scrambled targets, controlled, multi-controlled,
    diagonal and (for small n) wide gates."""
    rng = np.random.default_rng(seed)
    ops = []
    for i in range(count):
        kind = int(rng.integers(9))
        qs = [int(x) for x in rng.permutation(n)]
        if kind == 0 or n == 1:
            m = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
            ops.append((np.linalg.qr(m)[0], [qs[0]]))
        elif kind == 1:
            m = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
            ops.append((np.linalg.qr(m)[0], qs[:2]))
        elif kind == 2:
            ops.append(og.cnot(qs[0], qs[1]))
        elif kind == 3:
            ops.append(og.cz(qs[0], qs[1]))
        elif kind == 4:
            ops.append(og.swap(qs[0], qs[1]))
        elif kind == 5:
            ops.append(og.rz(float(rng.uniform(0, 6.28)), qs[0]))
        elif kind == 6:
            ops.append(og.cphase(float(rng.uniform(0, 6.28)), qs[0], qs[1]))
        elif kind == 7 and n >= 3:
            nc = int(rng.integers(2, min(n, 5)))
            ops.append(og.cnx(qs[1:1 + nc], qs[0]) if rng.random() < 0.5 else og.cnz(qs[1:1 + nc], qs[0]))
        elif kind == 8 and n >= 3:
            k = int(rng.integers(3, min(n, 5) + 1))
            if rng.random() < 0.5:
                ops.append((np.diag(np.exp(1j * rng.uniform(0, 6.28, size=1 << k))), qs[:k]))
            else:
                m = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
                ops.append((np.linalg.qr(m)[0], qs[:k]))
        else:
            ops.append(og.hadamard(qs[0]))
    return ops
