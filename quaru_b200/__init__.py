"""quaru_b200 -- B200-native state-vector engine behind Quaru's Register / Operation API.

Host-side mirror of the reference's interface (names as in src/register.rs, src/operation.rs);
all state-vector work runs in hand-written sm_100a CUDA kernels reached through the C ABI in
include/quaru_b200.h (quaru_b200/lib/libquaru_b200.so).  There is no CPU fallback.
"""
from . import openqasm, operation
from .circuit import QuantumCircuit
from .operation import Operation
from .register import EngineError, OperationError, Register

__all__ = ["Register", "Operation", "OperationError", "EngineError", "QuantumCircuit", "operation", "openqasm"]
