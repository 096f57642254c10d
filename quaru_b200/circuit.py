"""Host-side mirror of quaru::operation::QuantumCircuit's builder API (src/operation.rs:79-109, 271-321).

The reference's reducers (operation.rs:114-269) fuse gates by multiplying dense matrices on the host;
in this engine fusion happens inside the library (gates queued by Register.apply / apply_batch are
planned into multi-gate HBM passes), so a circuit is just an ordered gate list plus measurement
targets and conditional operations, with the reference's panics kept.
"""


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
