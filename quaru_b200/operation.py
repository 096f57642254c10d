"""Host-side mirror of quaru::operation (src/operation.rs:34-70, 325-536) and openqasm::u
(src/openqasm.rs:145-163): the `Operation` payload that crosses the C ABI and the reference's gate
constructors, with the reference's names, argument order and target conventions.

These are tiny host matrices (the reference builds them with ndarray on the host too); the device
work happens in Register.apply.  The same constructors exist for C++ callers in
include/quaru_b200.hpp and for the Rust shim the reference's own src/operation.rs is kept as is.
"""
import math

import numpy as np

C = np.complex128


class Operation:
    """`struct Operation {matrix, targets}` (operation.rs:49-52) + trait QuantumOperation (:34-38)."""

    __slots__ = ("_matrix", "_targets")

    def __init__(self, matrix, targets):
        self._matrix = np.ascontiguousarray(matrix, dtype=C)
        self._targets = [int(t) for t in targets]

    @staticmethod
    def new(matrix, targets):
        """Operation::new (operation.rs:60-69): None unless the matrix is 2^k x 2^k for k = len(targets)."""
        m = np.asarray(matrix)
        side = 2 ** len(targets)
        if m.ndim != 2 or m.shape[0] != side or m.shape[1] != side:
            return None
        return Operation(m, targets)

    def matrix(self):
        return self._matrix.copy()

    def targets(self):
        return list(self._targets)

    def arity(self):
        return len(self._targets)

    def __eq__(self, other):   # #[derive(PartialEq)] operation.rs:48
        return (isinstance(other, Operation) and self._targets == other._targets
                and self._matrix.shape == other._matrix.shape and bool((self._matrix == other._matrix).all()))

    __hash__ = None

    def __repr__(self):
        return f"Operation(targets={self._targets}, matrix={self._matrix.shape[0]}x{self._matrix.shape[1]})"


def _real(rows):
    return np.array(rows, dtype=np.float64).astype(C)


def identity(target):                       # operation.rs:340
    return Operation(_real([[1, 0], [0, 1]]), [target])


def hadamard(target):                       # operation.rs:350
    return Operation((math.sqrt(0.5) * np.array([[1.0, 1.0], [1.0, -1.0]])).astype(C), [target])


def hadamard_transform(targets):            # operation.rs:358
    m = hadamard(targets[0])._matrix
    for t in targets[1:]:
        m = np.kron(hadamard(t)._matrix, m)
    return Operation(m, list(targets))


def cnot(control, target):                  # operation.rs:372 -- targets = [target, control]
    return Operation(_real([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]]), [target, control])


def to_quantum_gate(f, targets):            # operation.rs:386 -- column c is the basis vector f(c)
    n = 1 << len(targets)
    m = np.zeros((n, n), dtype=C)
    for c in range(n):
        v = f(c)
        if not 0 <= v < n:
            raise IndexError("to_quantum_gate: f(c) out of range")      # reference: index panic (math.rs:46)
        m[v, c] = 1.0
    return Operation(m, list(targets))


def to_controlled(op, control):             # operation.rs:406 -- I (+) M, control becomes the MSB
    old = op._matrix.shape[0]
    m = np.zeros((2 * old, 2 * old), dtype=C)
    m[:old, :old] = np.eye(old)
    m[old:, old:] = op._matrix
    return Operation(m, op._targets + [control])


def swap(target_1, target_2):               # operation.rs:427
    return Operation(_real([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]), [target_1, target_2])


def phase(target):                          # operation.rs:443
    return Operation(np.array([[1, 0], [0, 1j]], dtype=C), [target])


def not_(target):                           # operation.rs:458 (`not` is a Python keyword)
    return Operation(_real([[0, 1], [1, 0]]), [target])


def pauli_y(target):                        # operation.rs:468
    return Operation(np.array([[0, -1j], [1j, 0]], dtype=C), [target])


def pauli_z(target):                        # operation.rs:481
    return Operation(_real([[1, 0], [0, -1]]), [target])


def cnx(controls, target):                  # operation.rs:491 -- targets = [target, controls...]
    targets = [target] + list(controls)
    n = 1 << len(targets)
    m = np.zeros((n, n))
    for i in range(n - 2):
        m[i, i] = 1.0
    m[n - 1, n - 2] = 1.0
    m[n - 2, n - 1] = 1.0
    return Operation(m.astype(C), targets)


def cnz(controls, target):                  # operation.rs:520
    targets = [target] + list(controls)
    n = 1 << len(targets)
    m = np.eye(n)
    m[n - 1, n - 1] = -1.0
    return Operation(m.astype(C), targets)


def u(theta, phi, lam, target):             # openqasm.rs:145
    i = 1j
    return Operation.new(np.array([
        [np.exp(-i * (phi + lam) / 2.0) * np.cos(theta / 2.0), -np.exp(-i * (phi - lam) / 2.0) * np.sin(theta / 2.0)],
        [np.exp(i * (phi - lam) / 2.0) * np.sin(theta / 2.0), np.exp(i * (phi + lam) / 2.0) * np.cos(theta / 2.0)],
    ], dtype=C), [target])


# The synthetic BASELINE circuits use RZ / CZ / controlled phase; the reference has no constructor
# for them and builds them through Operation::new / cnz / to_controlled.

def rz(theta, target):
    return Operation(np.array([[np.exp(-0.5j * theta), 0], [0, np.exp(0.5j * theta)]], dtype=C), [target])


def cz(control, target):
    return cnz([control], target)


def cphase(angle, control, target):
    return to_controlled(Operation(np.array([[1, 0], [0, np.exp(1j * angle)]], dtype=C), [target]), control)
