/*
 * quaru_b200.h -- C ABI of the B200-native state-vector engine behind Quaru's
 * Register / Operation API (the drop-in boundary).
 *
 * The reference (DATX11-22/quaru) has no FFI of its own: its API *is* the Rust
 * `impl Register` (src/register.rs:39-407).  Each entry point below names the
 * reference item it replaces; a Rust `register.rs` shim (rust/src/register.rs)
 * binds exactly these symbols and keeps the reference's method signatures.
 *
 * Conventions
 *   - return 0 = QB_OK; negative = mirror of `OperationError` (register.rs:10-20);
 *     positive = device/runtime failures that have no reference equivalent (the
 *     Rust shim panics on them: there is no CPU fallback).
 *   - complex numbers are interleaved (re, im) doubles = num::Complex<f64> (#[repr(C)]).
 *   - amplitude index bit q  <->  qubit q (register.rs:47-50).
 *   - gate matrices are row-major 2^k x 2^k; matrix index bit j <-> targets[j]
 *     (register.rs:168-173); targets are in arbitrary order, not sorted.
 *   - every host pointer is borrowed for the duration of the call only.
 *   - a handle may be used by one thread at a time (`&mut self` in the reference);
 *     different handles may be used concurrently.
 *   - qb_apply / qb_apply_all / qb_apply_batch enqueue work (gates are buffered
 *     and fused into multi-gate passes); every call that returns data
 *     (measure, probabilities, read, norm, diff) or qb_flush/qb_sync first
 *     drains the queue.  QB_OPT_FUSION=0 makes every gate its own HBM pass.
 *   - no exception or panic crosses this boundary.
 */
#ifndef QUARU_B200_H
#define QUARU_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qb_reg qb_reg;

/* ---- status codes ------------------------------------------------------------------------ */
#define QB_OK 0
#define QB_ERR_INVALID_TARGET (-1)     /* OperationError::InvalidTarget(usize)        register.rs:13 */
#define QB_ERR_INVALID_DIMENSIONS (-2) /* OperationError::InvalidDimensions(r, c)     register.rs:15 */
#define QB_ERR_INVALID_ARITY (-3)      /* OperationError::InvalidArity(usize)         register.rs:17 */
#define QB_ERR_INVALID_QUBIT (-4)      /* OperationError::InvalidQubit(usize)         register.rs:19 */
#define QB_ERR_INVALID_ARGUMENT (-5)   /* NULL pointer / out-of-range index: a Rust panic (index OOB) in the reference */
#define QB_ERR_QASM (-6)               /* OpenQASMError::OpenQASMParseError (openqasm.rs:27-30); payload = line */
#define QB_ERR_CUDA 1
#define QB_ERR_NCCL 2
#define QB_ERR_OOM 3
#define QB_ERR_NO_DEVICE 4
#define QB_ERR_UNSUPPORTED 5
#define QB_ERR_IO 6                    /* state file missing / unreadable / corrupt; payload = errno or 0 */

/* Message / payload (the usize carried by the OperationError variant) of the last failing call
 * on the calling thread. */
const char* qb_last_error(void);
int64_t qb_last_error_payload(void);

/* ---- construction / destruction ---------------------------------------------------------- */

/* Register::new(&[bool]) register.rs:41-56 and Register::from_int(n, k) register.rs:58-63.
 * The shim folds the bools into basis_index = sum b_q 2^q.  basis_index >= 2^n is the
 * reference's index-out-of-bounds panic -> QB_ERR_INVALID_ARGUMENT.  Uses the current CUDA
 * device (cudaGetDevice). */
int qb_create(uint32_t n, uint64_t basis_index, qb_reg** out);

/* Register::try_from_qubits(&[Array2<c64>]) register.rs:81-102: `qubits` is n x 2 interleaved
 * c64, qubit i = (qubits[4i..4i+1], qubits[4i+2..4i+3]).  Validates |sum|a|^2 - 1| < 1e-6
 * (register.rs:107-115) -> QB_ERR_INVALID_QUBIT with payload i. */
int qb_create_product(const double* qubits, uint32_t n, qb_reg** out);

/* Sharded register: rank `rank` of `world` (power of two) holds amplitudes whose top
 * log2(world) index bits equal `rank`.  `nccl_unique_id` is the 128-byte ncclUniqueId produced
 * by qb_nccl_unique_id() on rank 0 and broadcast by the caller (torch.distributed in bench.py).
 * All subsequent calls on the handle are collective: every rank makes the same calls with the
 * same arguments.  `device` is the CUDA ordinal this rank drives (LOCAL_RANK under torchrun; -1 = the
 * library's current device -- note the library links its own CUDA runtime, so torch.cuda.set_device does
 * not change it).  No reference equivalent (the reference is single-process). */
int qb_create_sharded(uint32_t n, uint64_t basis_index, int32_t rank, int32_t world,
                      const void* nccl_unique_id, int32_t device, qb_reg** out);
int qb_nccl_unique_id(void* out128);

/* #[derive(Clone)] register.rs:23 ; Drop.  qb_destroy(NULL) is a no-op. */
int qb_clone(qb_reg* src, qb_reg** out);
void qb_destroy(qb_reg* reg);

/* `*qreg = Register::new(&vec![false; n])` (openqasm.rs:129) without reallocating. */
int qb_reset(qb_reg* reg, uint64_t basis_index);

/* Register::size() register.rs:404-406 */
uint32_t qb_num_qubits(const qb_reg* reg);

/* ---- gate application -------------------------------------------------------------------- */

/* Register::apply / try_apply_fast / try_apply (register.rs:134-250).
 * Validation in the reference's order: smallest duplicate target -> QB_ERR_INVALID_TARGET
 * (payload = duplicate), first target >= n in `targets` order -> QB_ERR_INVALID_TARGET
 * (payload = that target).  No unitarity is assumed. */
int qb_apply(qb_reg* reg, const double* matrix, const uint32_t* targets, uint32_t k);

/* Extension (no reference equivalent): the gate `matrix` on `targets`, applied only where every qubit in
 * `controls` is 1 -- what operation::to_controlled / cnx / cnz (operation.rs:406, 491, 520) express by building a
 * 2^(k+c) x 2^(k+c) matrix, without the matrix (a 30-control Z would need 2^62 entries).  Validation as qb_apply over
 * targets followed by controls. */
int qb_apply_controlled(qb_reg* reg, const double* matrix, const uint32_t* targets, uint32_t k,
                        const uint32_t* controls, uint32_t n_controls);

/* Register::apply_all / try_apply_all (register.rs:259-281): the 2x2 gate on every qubit.
 * The arity check (InvalidArity) belongs to the shim, which owns the Operation. */
int qb_apply_all(qb_reg* reg, const double* m2x2);

/* One gate of a batch: the (matrix, targets) payload of `Operation` (operation.rs:49-52). */
typedef struct qb_gate {
    const double* matrix;    /* row-major 2^k x 2^k interleaved c64 */
    const uint32_t* targets; /* k entries */
    uint32_t k;
} qb_gate;

/* The gate loop of Register::apply_circuit (register.rs:346-350): validates all gates first
 * (nothing is applied if any is invalid; payload = offending qubit), then enqueues them. */
int qb_apply_batch(qb_reg* reg, const qb_gate* gates, uint64_t count);

/* ---- measurement ------------------------------------------------------------------------- */

/* First loop of try_measure (register.rs:303-315): p0/p1 = sum |a_i|^2 by bit `target`.
 * Deterministic fixed-shape tree reduction (independent of the number of GPUs). */
int qb_probabilities(qb_reg* reg, uint32_t target, double* p0, double* p1);

/* Register::measure / try_measure (register.rs:291-342).  `u` is the uniform [0,1) draw the
 * reference takes from thread_rng (register.rs:317-318); the shim supplies it.
 * outcome = (u > p0); rejected half zeroed; kept half divided by sqrt(p_outcome). */
int qb_measure(qb_reg* reg, uint32_t target, double u, int32_t* outcome);

/* ---- state access (`pub state`, print_*, PartialEq; register.rs:35, 379-413) -------------- */

/* Copy amplitudes [offset, offset+count) to / from host (2*count doubles).  On a sharded
 * register the range is in GLOBAL index space; only the part owned by this rank is
 * transferred (read: the rest of `out` is left untouched). */
int qb_read_state(qb_reg* reg, uint64_t offset, uint64_t count, double* out);
int qb_write_state(qb_reg* reg, uint64_t offset, uint64_t count, const double* in);

/* Raw state files (checkpoint / restore).  The reference's `pub state` is a host array any driver can serialise
 * (register.rs:23-37); for a device-resident state the engine streams it.  One file = one rank's shard: a 64-byte
 * header ("QBSTATE1", version, n, n_local, world, rank, count, sum|a|^2, checksum) followed by 2^n_local interleaved
 * little-endian (re, im) f64 pairs.  Sharded registers: every rank calls with ITS OWN path (collective, because the
 * qubit placement is first restored to the identity).  qb_load_state needs a handle with the same n / world / rank
 * and verifies the uploaded payload against the header's checksum (sum_j word_j*(2j+1) mod 2^64 over the raw 64-bit
 * words, computed on the device); on QB_ERR_IO the
 * register contents are unspecified and must be reset or reloaded. */
int qb_save_state(qb_reg* reg, const char* path);
int qb_load_state(qb_reg* reg, const char* path);

/* sum |a_i|^2 ; max_i |a_i - b_i| (PartialEq uses < 1e-8, register.rs:411) ; sum |a_i-b_i|^2 */
int qb_norm2(qb_reg* reg, double* out);
int qb_max_abs_diff(qb_reg* a, qb_reg* b, double* out);
int qb_l2_diff2(qb_reg* a, qb_reg* b, double* out);

/* ---- queue / stream control -------------------------------------------------------------- */
int qb_flush(qb_reg* reg); /* plan + launch everything queued; asynchronous */
int qb_sync(qb_reg* reg);  /* flush + wait for the handle's stream */

/* CUDA-event stopwatch on the handle's own stream (torch.cuda.Event cannot see it).
 * start: flush, record; stop: flush, record, synchronize, return elapsed ms. */
int qb_timer_start(qb_reg* reg);
int qb_timer_stop(qb_reg* reg, float* ms);

#define QB_OPT_FUSION 1        /* 1 (default): buffer gates and fuse into tile passes; 0: one pass per gate */
#define QB_OPT_TILE_BITS 2     /* log2 amplitudes per shared-memory tile (default 12) */
#define QB_OPT_MAX_QUEUE 3     /* auto-flush after this many buffered gates (default 4096) */
#define QB_OPT_FORCE_PATH 4    /* tests only: 0 auto, 1 generic kernel (in place), 2 shared-memory tile kernel, 4 tile kernels even for single gates,
                                  5 generic kernel out of place (the k > 13 path: needs a second state buffer) */
#define QB_OPT_PROFILE 5       /* 1: bracket every tile-pass launch with CUDA events (QB_STAT_TILE_NS) */
#define QB_OPT_JIT 6           /* specialised tile kernels (one NVRTC-compiled kernel per distinct pass program, cached by program hash):
                                  0 off; 1 (default, env QB_JIT) asynchronous -- a pass runs through the interpreter kernel until its
                                  kernel is ready; 2 synchronous -- wait for the compiler, a compile failure is an error (tests) */
#define QB_OPT_JIT_MIN_QUBITS 7 /* local qubits from which passes are specialised (default 26, env QB_JIT_MIN_QUBITS) */
#define QB_OPT_JIT_VARIANT 8   /* kernel shape: bits 0-3 = CTAs per SM the register budget is set for (2 -> 128 registers, 3 -> 80),
                                  bit 4 = persistent CTAs walking the tiles, bit 5 = with an L2 prefetch of the next tile,
                                  bit 6 = the prefetching ring: one persistent CTA of 512 threads per SM, two groups on alternate tiles,
                                  every thread fetches its amplitudes of the group's next tile with cp.async into a ring of three
                                  shared-memory tile buffers (measured slower than the default shape) */
#define QB_OPT_PEEPHOLE 9      /* H / controlled-X / controlled-Z rewriting of every flushed gate list (C^nX = H C^nZ H: H ; CX -> CZ ; H,
                                  CX ; H -> H ; CZ, H ; CX ; H -> CZ, H ; CZ ; H -> CX, H ; H -> nothing): 0 off, 1 on.  Default: on for
                                  registers of >= 28 local qubits (env QB_PEEPHOLE=1 / 0 forces it on / off); only with fusion on */
int qb_set_option(qb_reg* reg, int32_t option, int64_t value);

#define QB_STAT_KERNEL_LAUNCHES 1 /* kernels launched by this handle since creation */
#define QB_STAT_PASSES 2          /* state sweeps (HBM read+write passes) */
#define QB_STAT_GATES 3           /* gates accepted */
#define QB_STAT_BYTES_MOVED 4     /* physical HBM bytes the launched passes read+wrote (model) */
#define QB_STAT_SWAP_BYTES 5      /* bytes this rank sent over NVLink in global<->local swaps */
#define QB_STAT_SWAPS 6           /* number of global<->local exchanges */
#define QB_STAT_TILE_LAUNCHES 7   /* tile-pass kernel launches (profiled ones when QB_OPT_PROFILE=1) */
#define QB_STAT_TILE_NS 8         /* summed device time of the profiled tile-pass launches, ns */
#define QB_STAT_ROUNDS 9          /* register rounds executed by the register-blocked tile kernel */
#define QB_STAT_P2P_EXCHANGES 10  /* exchanges done by the peer-memory kernel (the rest went through NCCL send/recv) */
#define QB_STAT_REMOTE_PASSES 11  /* tile passes that reached peers' shards directly (QB_SHARD_REMOTE=1; their NVLink bytes count in QB_STAT_SWAP_BYTES) */
#define QB_STAT_JIT_LAUNCHES 12   /* tile passes run by a specialised kernel */
#define QB_STAT_JIT_COMPILES 13   /* distinct pass programs this handle handed to the compiler */
#define QB_STAT_JIT_MISSES 14     /* passes the interpreter kernel ran because the specialised kernel was not ready (or failed) */
#define QB_STAT_JIT_COMPILE_MS 15 /* process-wide: summed NVRTC time, ms */
#define QB_STAT_JIT_PENDING 16    /* process-wide: compiles queued or running */
#define QB_STAT_JIT_FAILED 17     /* process-wide: compiles that failed */
#define QB_STAT_FUSED_EXCHANGES 18 /* sharded registers: exchanges carried by the stores of a tile pass (counted in QB_STAT_SWAPS too) */
#define QB_STAT_PEEPHOLE_REWRITES 19 /* rewrites of the H / CX / CZ peephole so far */
int qb_get_stat(qb_reg* reg, int32_t stat, int64_t* value);

/* Blocks until no specialised kernel is being compiled (process-wide).  A driver that is about to time a circuit it has
 * already applied once calls this so that every pass finds its kernel in the cache. */
int qb_jit_drain(void);

/* ---- planner introspection (host only, no GPU needed) ------------------------------------ */
/* Plans `count` gates for an n-qubit register with the given tile parameters exactly as
 * qb_flush would, and writes for each gate the index of the pass it was assigned to
 * (pass_of_gate[count]) and its position inside the pass (order_in_pass[count]).
 * Returns the number of passes, or a negative error. */
int64_t qb_plan_debug(uint32_t n, const qb_gate* gates, uint64_t count, uint32_t tile_bits,
                      int32_t* pass_of_gate, int32_t* order_in_pass);

/* ---- OpenQASM 2.0 front end ------------------------------------------------------------------
 * Replaces the un-vendored crate the reference parses with (openqasm-parser =^1.0.0, Cargo.toml:21;
 * call sites src/openqasm.rs:53-54, 73, 91-137): source text -> register declarations + the flat list of
 * basic operations (gate definitions inlined, register-wide statements broadcast), i.e. what
 * `OpenQASMProgram::{qregs, cregs, get_basic_operations()}` gives the reference's interpreter loop.
 * The loop itself (openqasm.rs:73-139) stays on the host side: the Rust `openqasm.rs` is kept as is and
 * quaru_b200/openqasm.py mirrors it.  U parameters are f32 (BasicOp::U(f32, f32, f32, _)). */
typedef struct qb_qasm qb_qasm;
#define QB_QASM_U 0        /* BasicOp::U(theta, phi, lambda, qubit a)       */
#define QB_QASM_CX 1       /* BasicOp::CX(control a, target b)              */
#define QB_QASM_MEASURE 2  /* BasicOp::Measure(qubit a, classical bit b)    */
#define QB_QASM_RESET_Q 3  /* BasicOp::ResetQ(qubit a)                      */
#define QB_QASM_RESET_C 4  /* BasicOp::ResetC(classical bit a)              */
typedef struct qb_qasm_op {
    int32_t kind;
    float p[3];            /* U: theta, phi, lambda */
    int32_t reg_a, idx_a;  /* register index (declaration order) and element */
    int32_t reg_b, idx_b;
    int32_t cond_creg;     /* `if (creg == cond_val)`: classical register index, or -1 */
    uint32_t cond_val;
} qb_qasm_op;
int qb_qasm_parse(const char* source, qb_qasm** out);  /* QB_ERR_QASM + qb_last_error() on syntax / semantic errors */
void qb_qasm_free(qb_qasm* prog);
uint32_t qb_qasm_num_qregs(const qb_qasm* prog);
uint32_t qb_qasm_num_cregs(const qb_qasm* prog);
const char* qb_qasm_qreg_name(const qb_qasm* prog, uint32_t i);
const char* qb_qasm_creg_name(const qb_qasm* prog, uint32_t i);
uint32_t qb_qasm_qreg_size(const qb_qasm* prog, uint32_t i);
uint32_t qb_qasm_creg_size(const qb_qasm* prog, uint32_t i);
uint64_t qb_qasm_num_ops(const qb_qasm* prog);
const qb_qasm_op* qb_qasm_ops(const qb_qasm* prog);

/* Serialises the register-blocked round programs the fused path would launch for `gates` on an n-qubit register
 * (format: quaru_b200/csrc/qb_tile.cu rounds_debug).  tests/test_rounds_emulator.py replays them with a NumPy
 * emulator of the kernel's semantics against the CPU oracle, so the host-side encoder is covered without a GPU.
 * Returns the number of bytes written or a negative error. */
int64_t qb_rounds_debug(uint32_t n, const qb_gate* gates, uint64_t count, void* out, uint64_t capacity);

/* Dry run of the multi-GPU placement logic: the global<->local exchanges rank `rank` of `world` would
 * perform for this gate list, as a flat int32 list of records [gate index before which the exchange
 * happens, g', V[0..g') local bit positions, J[0..g') rank-bit indices], and the final logical->physical
 * placement.  Returns the number of int32 written or a negative error.  Every rank must produce the same
 * list (tests/test_dist_cpu.py checks this over gloo). */
int64_t qb_shard_plan_debug(uint32_t n, int32_t world, int32_t rank, const qb_gate* gates, uint64_t count,
                            int32_t* events, uint64_t capacity, uint32_t* final_phys_of);

/* The CUDA source of the specialised kernel for one serialised round program (a `kind 0` record of qb_rounds_debug), and
 * an NVRTC compile of such a source for sm_100a; neither needs a GPU.  tests/test_jit_cpu.py builds the source as host C++
 * (-DQB_HOST_EMU) and replays it against the oracle. */
int64_t qb_jit_source_debug(const void* program, uint64_t len, int32_t min_blocks, int32_t persistent, int32_t prefetch,
                            int32_t remote, int32_t xchg_bits, const uint8_t* xchg_vj /* V[3] then J[3] */, char* out,
                            uint64_t capacity);
int64_t qb_jit_compile_debug(const char* source, double* ms, void* cubin_out, uint64_t capacity);

const char* qb_version(void);

/* Dry run of the SWAP relabelling (registers of >= 28 local qubits, or QB_SWAP_RELABEL=1): a SWAP gate
 * (src/operation.rs:427) only exchanges the names of two qubits inside the engine; later gates and measurements are
 * translated, raw state access first undoes the permutation with real SWAP passes.  Output, one record per gate:
 * [-2] exact identity (dropped), [-1] SWAP absorbed, or [k, engine qubit of targets[0], ..., of targets[k-1]]; then
 * [-3, m, a_0, b_0, ..., a_{m-1}, b_{m-1}]: the engine-qubit pairs that are swapped, in order, before a raw read.
 * Returns the number of int32 written or a negative error (tests/test_relabel.py replays it against the oracle). */
int64_t qb_relabel_debug(uint32_t n, const qb_gate* gates, uint64_t count, int32_t* out, uint64_t capacity);

/* Dry run of the H / CX / CZ peephole (QB_OPT_PEEPHOLE) on a gate list: out[0] = number of rewrites, then one record per gate
 * of the rewritten list: [kind (0 dense, 1 diagonal), k, control mask, targets[0..k), m, re/im of the m matrix (dense: 2^k x 2^k
 * row-major; diagonal: 2^k) entries] -- the engine's classified form (controls extracted).  Returns the number of doubles
 * written or a negative error (tests/test_peephole.py applies both lists with the oracle). */
int64_t qb_peephole_debug(uint32_t n, const qb_gate* gates, uint64_t count, double* out, uint64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* QUARU_B200_H */
