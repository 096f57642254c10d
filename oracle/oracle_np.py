"""numpy twin of the oracle (TEST INFRASTRUCTURE ONLY) -- an independent second restatement.

``apply_np`` computes the same map as register.rs:140-191 through a tensor contraction instead of
the gather / mat-vec / scatter loop, and ``apply_slow_np`` restates the reference's *other* public
path, ``try_apply`` (register.rs:193-250: qubit permutation + kron(I, M) . state), literally.  Tests
require C oracle == apply_np == apply_slow_np before the oracle is trusted.
"""
import numpy as np


def apply_np(state, n, matrix, targets):
    """new[p_r] = sum_c M[r][c] old[p_c], matrix bit j <-> targets[j]; returns a new vector."""
    k = len(targets)
    psi = np.asarray(state, dtype=np.complex128).reshape((2,) * n)          # axis a <-> qubit n-1-a
    m = np.asarray(matrix, dtype=np.complex128).reshape((2,) * (2 * k))     # row bits MSB..LSB, col bits MSB..LSB
    # matrix axis for bit j (of row index) is k-1-j; state axis of qubit q is n-1-q
    col_axes = [k + (k - 1 - j) for j in range(k)]
    st_axes = [n - 1 - targets[j] for j in range(k)]
    out = np.tensordot(m, psi, axes=(col_axes, st_axes))                    # axes: row bits (MSB..LSB), then rest
    # move row-bit axis (k-1-j) back to state axis n-1-targets[j]
    out = np.moveaxis(out, [k - 1 - j for j in range(k)], st_axes)
    return np.ascontiguousarray(out).reshape(-1)


def apply_slow_np(state, n, matrix, targets):
    """register.rs:209-247 literally.  O(4^n) memory: tiny n only."""
    N = 1 << n
    state = np.asarray(state, dtype=np.complex128).reshape(-1)
    perm = list(targets) + [i for i in range(n) if i not in targets]       # :211-216
    idx = np.arange(N)
    j = np.zeros(N, dtype=np.int64)
    for kk, v in enumerate(perm[:n]):                                       # :226-229
        j |= ((idx >> kk) & 1) << v
    permuted = state[j]                                                     # :231
    big = np.kron(np.eye(1 << (n - len(targets))), np.asarray(matrix, dtype=np.complex128))  # :235
    permuted = big @ permuted                                               # :238
    j2 = np.zeros(N, dtype=np.int64)
    for kk, v in enumerate(perm[:n]):                                       # :243-245
        j2 |= ((idx >> v) & 1) << kk
    return permuted[j2]                                                     # :246


def probabilities_np(state, n, target):
    p = np.abs(np.asarray(state).reshape(-1)) ** 2
    bit = (np.arange(1 << n) >> target) & 1
    return float(p[bit == 0].sum()), float(p[bit == 1].sum())


def measure_np(state, n, target, u):
    """register.rs:297-342 with the draw supplied; returns (outcome, new_state)."""
    p0, p1 = probabilities_np(state, n, target)
    res = u > p0
    total = p1 if res else p0
    bit = (np.arange(1 << n) >> target) & 1
    out = np.where(bit == int(res), np.asarray(state).reshape(-1) / np.sqrt(total), 0.0)
    return bool(res), out
