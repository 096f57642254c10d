"""ctypes binding of include/quaru_b200.h.

The CUDA library is the product: if it is missing this module raises -- there is no CPU fallback
and nothing here imports oracle/.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QUARU_B200_LIB") or os.path.join(_HERE, "lib", "libquaru_b200.so")   # env: kernel experiments

QB_OK = 0
QB_ERR_INVALID_TARGET = -1
QB_ERR_INVALID_DIMENSIONS = -2
QB_ERR_INVALID_ARITY = -3
QB_ERR_INVALID_QUBIT = -4
QB_ERR_INVALID_ARGUMENT = -5
QB_ERR_CUDA, QB_ERR_NCCL, QB_ERR_OOM, QB_ERR_NO_DEVICE, QB_ERR_UNSUPPORTED = 1, 2, 3, 4, 5

QB_OPT_FUSION, QB_OPT_TILE_BITS, QB_OPT_MAX_QUEUE, QB_OPT_FORCE_PATH, QB_OPT_PROFILE = 1, 2, 3, 4, 5
QB_OPT_JIT, QB_OPT_JIT_MIN_QUBITS, QB_OPT_JIT_VARIANT = 6, 7, 8
(QB_STAT_JIT_LAUNCHES, QB_STAT_JIT_COMPILES, QB_STAT_JIT_MISSES, QB_STAT_JIT_COMPILE_MS, QB_STAT_JIT_PENDING,
 QB_STAT_JIT_FAILED) = 12, 13, 14, 15, 16, 17
QB_STAT_FUSED_EXCHANGES = 18
QB_OPT_PEEPHOLE = 9
QB_STAT_PEEPHOLE_REWRITES = 19
(QB_STAT_KERNEL_LAUNCHES, QB_STAT_PASSES, QB_STAT_GATES, QB_STAT_BYTES_MOVED, QB_STAT_SWAP_BYTES,
 QB_STAT_SWAPS, QB_STAT_TILE_LAUNCHES, QB_STAT_TILE_NS, QB_STAT_ROUNDS, QB_STAT_P2P_EXCHANGES,
 QB_STAT_REMOTE_PASSES) = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11


QB_ERR_QASM = -6
QB_QASM_U, QB_QASM_CX, QB_QASM_MEASURE, QB_QASM_RESET_Q, QB_QASM_RESET_C = 0, 1, 2, 3, 4


class QbQasmOp(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("p", ctypes.c_float * 3), ("reg_a", ctypes.c_int32), ("idx_a", ctypes.c_int32),
                ("reg_b", ctypes.c_int32), ("idx_b", ctypes.c_int32), ("cond_creg", ctypes.c_int32),
                ("cond_val", ctypes.c_uint32)]


class QbGate(ctypes.Structure):
    _fields_ = [("matrix", ctypes.POINTER(ctypes.c_double)),
                ("targets", ctypes.POINTER(ctypes.c_uint32)),
                ("k", ctypes.c_uint32)]


_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_reg = ctypes.c_void_p
_regp = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); every symbol include/quaru_b200.h declares
SIGNATURES = {
    "qb_last_error": (ctypes.c_char_p, []),
    "qb_last_error_payload": (ctypes.c_int64, []),
    "qb_version": (ctypes.c_char_p, []),
    "qb_create": (ctypes.c_int, [ctypes.c_uint32, ctypes.c_uint64, _regp]),
    "qb_create_product": (ctypes.c_int, [_dp, ctypes.c_uint32, _regp]),
    "qb_create_sharded": (ctypes.c_int, [ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int32, ctypes.c_int32,
                                         ctypes.c_void_p, ctypes.c_int32, _regp]),
    "qb_nccl_unique_id": (ctypes.c_int, [ctypes.c_void_p]),
    "qb_clone": (ctypes.c_int, [_reg, _regp]),
    "qb_destroy": (None, [_reg]),
    "qb_reset": (ctypes.c_int, [_reg, ctypes.c_uint64]),
    "qb_num_qubits": (ctypes.c_uint32, [_reg]),
    "qb_apply": (ctypes.c_int, [_reg, _dp, _u32p, ctypes.c_uint32]),
    "qb_apply_controlled": (ctypes.c_int, [_reg, _dp, _u32p, ctypes.c_uint32, _u32p, ctypes.c_uint32]),
    "qb_apply_all": (ctypes.c_int, [_reg, _dp]),
    "qb_apply_batch": (ctypes.c_int, [_reg, ctypes.POINTER(QbGate), ctypes.c_uint64]),
    "qb_probabilities": (ctypes.c_int, [_reg, ctypes.c_uint32, _dp, _dp]),
    "qb_measure": (ctypes.c_int, [_reg, ctypes.c_uint32, ctypes.c_double, _i32p]),
    "qb_read_state": (ctypes.c_int, [_reg, ctypes.c_uint64, ctypes.c_uint64, _dp]),
    "qb_write_state": (ctypes.c_int, [_reg, ctypes.c_uint64, ctypes.c_uint64, _dp]),
    "qb_save_state": (ctypes.c_int, [_reg, ctypes.c_char_p]),
    "qb_load_state": (ctypes.c_int, [_reg, ctypes.c_char_p]),
    "qb_norm2": (ctypes.c_int, [_reg, _dp]),
    "qb_max_abs_diff": (ctypes.c_int, [_reg, _reg, _dp]),
    "qb_l2_diff2": (ctypes.c_int, [_reg, _reg, _dp]),
    "qb_flush": (ctypes.c_int, [_reg]),
    "qb_sync": (ctypes.c_int, [_reg]),
    "qb_timer_start": (ctypes.c_int, [_reg]),
    "qb_timer_stop": (ctypes.c_int, [_reg, ctypes.POINTER(ctypes.c_float)]),
    "qb_set_option": (ctypes.c_int, [_reg, ctypes.c_int32, ctypes.c_int64]),
    "qb_get_stat": (ctypes.c_int, [_reg, ctypes.c_int32, _i64p]),
    "qb_qasm_parse": (ctypes.c_int, [ctypes.c_char_p, _regp]),
    "qb_qasm_free": (None, [_reg]),
    "qb_qasm_num_qregs": (ctypes.c_uint32, [_reg]),
    "qb_qasm_num_cregs": (ctypes.c_uint32, [_reg]),
    "qb_qasm_qreg_name": (ctypes.c_char_p, [_reg, ctypes.c_uint32]),
    "qb_qasm_creg_name": (ctypes.c_char_p, [_reg, ctypes.c_uint32]),
    "qb_qasm_qreg_size": (ctypes.c_uint32, [_reg, ctypes.c_uint32]),
    "qb_qasm_creg_size": (ctypes.c_uint32, [_reg, ctypes.c_uint32]),
    "qb_qasm_num_ops": (ctypes.c_uint64, [_reg]),
    "qb_qasm_ops": (ctypes.POINTER(QbQasmOp), [_reg]),
    "qb_rounds_debug": (ctypes.c_int64, [ctypes.c_uint32, ctypes.POINTER(QbGate), ctypes.c_uint64, ctypes.c_void_p,
                                         ctypes.c_uint64]),
    "qb_shard_plan_debug": (ctypes.c_int64, [ctypes.c_uint32, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(QbGate),
                                             ctypes.c_uint64, _i32p, ctypes.c_uint64, _u32p]),
    "qb_relabel_debug": (ctypes.c_int64, [ctypes.c_uint32, ctypes.POINTER(QbGate), ctypes.c_uint64, _i32p, ctypes.c_uint64]),
    "qb_peephole_debug": (ctypes.c_int64, [ctypes.c_uint32, ctypes.POINTER(QbGate), ctypes.c_uint64, ctypes.POINTER(ctypes.c_double), ctypes.c_uint64]),
    "qb_plan_debug": (ctypes.c_int64, [ctypes.c_uint32, ctypes.POINTER(QbGate), ctypes.c_uint64, ctypes.c_uint32,
                                       _i32p, _i32p]),
    "qb_jit_source_debug": (ctypes.c_int64, [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                             ctypes.c_int32, ctypes.c_int32, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint64]),
    "qb_jit_compile_debug": (ctypes.c_int64, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double), ctypes.c_void_p, ctypes.c_uint64]),
    "qb_jit_drain": (ctypes.c_int, []),
}


def lib():
    """Load the engine.  Raises if the CUDA library has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C quaru_b200/csrc).  quaru_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    L = lib()
    return L.qb_last_error().decode(), L.qb_last_error_payload()
