"""Host-side mirror of quaru::operation::QuantumCircuit (src/operation.rs:79-322).

Builder API (`add_operation`, `add_measurement`, `add_conditional_operation`, getters) with the reference's
panics, plus the four reducers (`reduce_*`, operation.rs:114-269).  The reducers are host-side algebra on the
small gate matrices; they are optional here -- the library already fuses queued gates into multi-gate HBM passes --
and are re-stated from their *documented intent* with a correct program-order product, because the reference
implementations have defects: `combine_ops` applies the later gate first in one branch and
mis-sizes two overlapping 2-qubit gates, `reduce_circuit_cancel_gates` cancels ANY two adjacent equal real
matrices (also non-involutory ones).  What the reference's own tests pin (operation.rs:600-651: H CNOT CNOT H -> 0
ops, H X on one target -> 1 op, H H -> 0, CNOT CNOT -> 0) holds here too.
"""
import numpy as np

from .operation import Operation


def _embed(matrix, targets, union):
    """The matrix of `targets`-gate as an operator on the ordered qubit list `union` (matrix bit j <-> union[j])."""
    k = len(union)
    pos = [union.index(t) for t in targets]
    dim = 1 << k
    out = np.zeros((dim, dim), dtype=np.complex128)
    rest = [j for j in range(k) if j not in pos]
    sub = 1 << len(targets)
    for r_rest in range(1 << len(rest)):
        base = sum(((r_rest >> i) & 1) << rest[i] for i in range(len(rest)))
        idx = [base | sum(((x >> j) & 1) << pos[j] for j in range(len(targets))) for x in range(sub)]
        out[np.ix_(idx, idx)] = matrix
    return out


def _product(first, second):
    """`second` after `first` (program order) as one Operation on the union of their targets."""
    union = list(first._targets) + [t for t in second._targets if t not in first._targets]
    m = _embed(second._matrix, second._targets, union) @ _embed(first._matrix, first._targets, union)
    return Operation(m, union)


class QuantumCircuit:
    def __init__(self):                              # QuantumCircuit::new (operation.rs:87)
        self.operations = []
        self._measurement_targets = []
        self._conditional_operations = []

    def reset_register(self):                        # operation.rs:95
        self.operations.clear()
        self._measurement_targets.clear()
        self._conditional_operations.clear()
        return self

    def get_operations(self):                        # operation.rs:101
        return self.operations

    def get_measurement_targets(self):               # operation.rs:104
        return self._measurement_targets

    def get_conditional_operations(self):            # operation.rs:107
        return self._conditional_operations

    def clear_operations(self):                      # operation.rs:271
        self.operations.clear()

    def add_operation(self, operation):              # operation.rs:276
        if any(t in self._measurement_targets for t in operation.targets()):
            raise RuntimeError("Cannot add operation after measurement")
        if len(self._conditional_operations) > 0:
            raise RuntimeError("Cannot add operation after conditional operation (yet...)")
        self.operations.append(operation)

    def add_measurement(self, target):               # operation.rs:292
        if target in self._measurement_targets:
            raise RuntimeError("Cannot add measurement after measurement")
        self._measurement_targets.append(target)

    def add_conditional_operation(self, operation, condition):   # operation.rs:302
        if any(t in self._measurement_targets for t in operation.targets()):
            raise RuntimeError("Cannot add operation after measurement")
        if not all(c[0] in self._measurement_targets for c in condition):
            raise RuntimeError("Condition must be a measurement")
        self._conditional_operations.append((operation, list(condition)))

    # ---- reducers (operation.rs:114-269), correct program-order semantics -----------------------------------
    def reduce_circuit_cancel_gates(self):
        """operation.rs:189: drop adjacent pairs of identical gates that cancel.  The reference drops any two
        adjacent equal REAL matrices; here the pair is dropped only if the matrix really squares to the identity."""
        changed = True
        while changed and len(self.operations) > 1:
            changed = False
            out, i = [], 0
            ops = self.operations
            while i < len(ops):
                if i + 1 < len(ops) and ops[i] == ops[i + 1]:
                    m = ops[i]._matrix
                    if np.abs(m @ m - np.eye(m.shape[0])).max() < 1e-12:     # H.H = 1 only up to rounding
                        i += 2
                        changed = True
                        continue
                out.append(ops[i])
                i += 1
            self.operations = out
        return self

    def reduce_circuit_gates_with_same_targets(self):
        """operation.rs:218: multiply runs of consecutive gates with identical target lists (later . earlier)."""
        out = []
        for op in self.operations:
            if out and out[-1]._targets == op._targets:
                out[-1] = Operation(op._matrix @ out[-1]._matrix, op._targets)
            else:
                out.append(Operation(op._matrix, op._targets))
        self.operations = out
        return self

    def reduce_non_overlapping_gates(self, max_size=5):
        """operation.rs:243: put consecutive gates on disjoint qubits into one time slot (kron).  The reference grows
        the combined gate without bound (2^k x 2^k matrices); `max_size` caps the arity of a slot."""
        out = []
        for op in self.operations:
            if out and not set(out[-1]._targets) & set(op._targets) and len(out[-1]._targets) + len(op._targets) <= max_size:
                prev = out[-1]
                out[-1] = Operation(np.kron(op._matrix, prev._matrix), prev._targets + op._targets)
            else:
                out.append(Operation(op._matrix, op._targets))
        self.operations = out
        return self

    def reduce_gates_with_one_off_size(self, max_size):
        """operation.rs:114: first merge same-target runs, then combine each adjacent OVERLAPPING pair whose gates
        both act on at most `max_size` qubits into one gate on the union of their targets."""
        self.reduce_circuit_gates_with_same_targets()
        out, i = [], 0
        ops = self.operations
        while i < len(ops):
            if i + 1 < len(ops):
                a, b = ops[i], ops[i + 1]
                union = len(set(a._targets) | set(b._targets))
                if len(a._targets) <= max_size and len(b._targets) <= max_size and set(a._targets) & set(b._targets) \
                        and union <= max(max_size, 2):
                    out.append(_product(a, b))
                    i += 2
                    continue
            out.append(ops[i])
            i += 1
        self.operations = out
        return self
