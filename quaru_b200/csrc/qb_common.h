// qb_common.h -- internal types shared by the engine's translation units.
//
// The engine replaces the state engine of the reference (src/register.rs): the state vector
// lives in HBM as 2^n double2 (= num::Complex<f64>) and every Register method becomes one or
// more sm_100a kernels.  Nothing here is visible through the C ABI (include/quaru_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/quaru_b200.h"

namespace qb {

typedef double2 c64;  // {x = re, y = im}; layout-identical to num::Complex<f64>

// ---- thread-local error record (qb_last_error / qb_last_error_payload) -----------------------
int set_error(int code, int64_t payload, const char* fmt, ...);
int fail_cuda(cudaError_t e, const char* what, const char* file, int line);

#define QB_CUDA(expr)                                                               \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) return qb::fail_cuda(_e, #expr, __FILE__, __LINE__); \
    } while (0)

// No C++ exception may cross the C ABI: every extern "C" entry point that can allocate is a function-try-block
// ending in QB_CATCH_ALL (host allocation failure -> QB_ERR_OOM, anything else -> QB_ERR_UNSUPPORTED).
int translate_exception() noexcept;
#define QB_CATCH_ALL \
    catch (...) { return qb::translate_exception(); }

#define QB_TRY(expr)                  \
    do {                              \
        int _rc = (expr);             \
        if (_rc != QB_OK) return _rc; \
    } while (0)

// ---- gate model --------------------------------------------------------------------------------
// Structure the planner detects in the dense matrix an `Operation` carries (operation.rs:49-52).
enum GateKind : uint8_t {
    GK_DENSE = 0,  // general 2^k x 2^k block on `targets`, applied where all ctrl_mask bits are 1
    GK_DIAG = 1,   // diagonal on `targets` (RZ, phase, CZ, cnz, Grover oracle/diffuser, QFT phases)
};

struct Gate {
    uint32_t k = 0;                 // arity after control extraction (may be 0 for a pure masked phase)
    std::vector<uint32_t> targets;  // PHYSICAL bit positions; matrix bit j <-> targets[j]
    std::vector<c64> matrix;        // GK_DENSE: 2^k x 2^k row-major; GK_DIAG: 2^k diagonal entries
    uint64_t ctrl_mask = 0;         // bits that must all be 1 for the gate to act (extracted controls)
    GateKind kind = GK_DENSE;
    int64_t tag = -1;               // caller's index of the gate (planner introspection)
    uint64_t target_mask() const {
        uint64_t m = 0;
        for (uint32_t t : targets) m |= 1ull << t;
        return m;
    }
    uint64_t qubit_mask() const { return target_mask() | ctrl_mask; }
};

// Analyse a raw (matrix, targets) payload (already validated): pull out control qubits (bits on
// which the matrix is the identity whenever the bit is 0), detect diagonal structure.
// Returns false if the gate is exactly the identity (nothing to do).
bool classify_gate(const double* matrix, const uint32_t* targets, uint32_t k, Gate* out);

// ---- SWAP gates as qubit relabelling (qb_gate.cu; pure host logic, checked on CPU through qb_relabel_debug) -------------
// A SWAP of two qubits moves no amplitude if the engine simply exchanges the two qubits' names: `label[q]` is the engine
// qubit that currently holds the caller's qubit q.  Later gates and measurements go through the map; raw state access first
// restores the identity with real SWAP passes (canonical_swaps).  On the BASELINE circuit (QFT = 16 final SWAPs, each a dense
// two-qubit gate whose two targets must both be tile bits) this removes 2 of 25 sweeps and the 16 heaviest ops of a step.
bool is_plain_swap(const Gate& g);                                     // uncontrolled dense 2-qubit gate equal to SWAP, entry by entry
void relabel_gate(const std::vector<uint32_t>& label, Gate* g);        // caller's qubits -> engine qubits (targets and controls)
void canonical_swaps(std::vector<uint32_t>* label, std::vector<uint32_t>* pairs);  // engine qubit pairs (a0, b0, a1, b1, ...) to SWAP,
                                                                                   // in order, to make the map the identity again
void make_swap_gate(uint32_t a, uint32_t b, Gate* out);

// ---- H / controlled-X / controlled-Z peephole over a gate list (qb_gate.cu; pure host logic, checked on CPU through
// qb_peephole_debug).  C^nX(t) = H(t) C^nZ H(t), so on the qubit t of a (multi-)controlled X
//     H ; CX      ->  CZ ; H          CX ; H   ->  H ; CZ          H ; CX ; H  ->  CZ          H ; CZ ; H  ->  CX      H ; H  ->  (nothing)
// where ";" is program order and "adjacent" means adjacent among the gates that touch t.  A Hadamard costs a tile pass 32
// FP64 instructions per thread, a CX whose control is a thread bit 64 selects, a CZ next to nothing (a sign in a per-thread
// phase table), so every rule removes work; the unitary is unchanged up to the rounding of 2 * (1/sqrt 2)^2.
// Returns the number of rewrites.
size_t peephole_h_cx(std::vector<Gate>* gates);

// ---- the handle ----------------------------------------------------------------------------------
struct Shard;  // multi-GPU state (qb_shard.cu)
struct XchgArgs;  // an exchange carried by the stores of a tile pass (qb_jit.h)
struct Group;  // P shards driven from one process (qb_group.cu)

struct Reg {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t n = 0;        // logical qubits
    uint32_t n_local = 0;  // index bits held by this rank (== n on one GPU)
    c64* state = nullptr;  // 2^n_local amplitudes
    c64* scratch = nullptr;  // lazily allocated second buffer (out-of-place generic path only)
    double2* partials = nullptr;  // reduction workspace
    size_t partials_len = 0;
    c64* dmat = nullptr;  // device arena for large matrices / diagonals / offset tables
    size_t dmat_bytes = 0;
    void* hpin = nullptr;  // pinned host staging for pass programs
    size_t hpin_bytes = 0;
    void* dprog = nullptr;  // device copy of the pass program
    cudaEvent_t prog_free = nullptr;  // recorded after the last kernel that reads dprog/hpin
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<Gate> queue;  // gates accepted but not yet launched
    size_t held = 0;          // how many of them an overflow flush handed back (see run_fused)
    // options
    int fusion = 1;
    uint32_t tile_bits = 12;
    uint64_t max_queue = 4096;
    int force_path = 0;
    int profile = 0;
    // specialised (JIT-compiled) tile kernels, qb_jit.cu: 0 off, 1 asynchronous (the interpreter kernel runs a pass until its
    // specialised kernel is ready), 2 synchronous (wait for the compiler; failures are errors)
    int jit = 1;
    uint32_t jit_min_qubits = 26;   // local qubits from which a pass is worth a compile (its sweeps take >= ~1 ms)
    int jit_min_blocks = 2, jit_persistent = 0, jit_prefetch = 0, jit_pipelined = 0;
    // stats
    int64_t st_launches = 0, st_passes = 0, st_gates = 0, st_bytes = 0, st_swap_bytes = 0, st_swaps = 0;
    int64_t st_tile_launches = 0, st_tile_ns = 0, st_rounds = 0, st_p2p = 0, st_remote_passes = 0;
    int64_t st_jit_launches = 0, st_jit_compiles = 0, st_jit_misses = 0;
    std::vector<cudaEvent_t> tile_events;  // start/stop pairs of profiled tile launches
    size_t tile_events_used = 0;
    // sharding
    int rank = 0, world = 1;
    Shard* shard = nullptr;
    Group* group = nullptr;  // set on every shard of an in-process (QB_NUM_GPUS) register
    // The shard scheduler asks the LAST pass of the next run_fused to carry a global<->local exchange in its stores
    // (xchg_request); run_fused raises xchg_last_pass around that launch; the launcher reports xchg_done.
    const XchgArgs* xchg_request = nullptr;
    bool xchg_last_pass = false, xchg_done = false;
    int64_t st_fused_exchanges = 0;
    std::vector<uint32_t> phys_of;  // logical qubit -> physical bit position
    // SWAP relabelling (above the placement): caller's qubit q is engine ("logical") qubit label[q]
    bool relabel = false;
    std::vector<uint32_t> label;
    // H / CX / CZ peephole over every flushed gate list (qb_gate.cu); default: registers whose sweeps are expensive
    bool peephole = false;
    int64_t st_peephole = 0;
};

// ---- kernels' host launchers (qb_kernels.cu) -------------------------------------------------------
int k_init_basis(Reg* r, uint64_t local_index, bool owned);
int k_init_product(Reg* r, const double* qubits_host);
int k_apply_gate_unfused(Reg* r, const Gate& g);  // one gate = at most one sweep over the shard
int k_probabilities(Reg* r, int target_phys /* -1: everything into p0 */, double* p0, double* p1);
int k_collapse(Reg* r, uint32_t target_phys, int outcome, double total_prob);
int k_scale_or_zero_all(Reg* r, bool keep, double total_prob);  // measured qubit is a rank bit
int k_diff(Reg* a, Reg* b, double* max_abs, double* l2sq);
int ensure_dmat(Reg* r, size_t bytes);

// ---- raw state files (qb_stateio.cu); the caller has flushed the queue and restored the identity placement ------
int state_save(Reg* r, const char* path);
int state_load(Reg* r, const char* path);
int state_load_check(Reg* r, const char* path);  // header validation only; the register is not touched

// ---- fused tile passes (qb_tile.cu) ----------------------------------------------------------------
int drain_tile_events(Reg* r);  // fold finished profiling events into st_tile_ns
int run_fused(Reg* r, std::vector<Gate>& gates, std::vector<Gate>* hold = nullptr);  // plan + launch; consumes `gates`.  With `hold`
                                                   // (the queue overflowed, more gates are coming) the gates of the LAST planned pass are
                                                   // handed back instead of launched: that pass is usually half empty, and the next flush
                                                   // can fill it (3 steps of the bench at 33 qubits: 69 -> 68 sweeps, 6 steps: 135 -> 133)
int run_local(Reg* r, std::vector<Gate>& gates);  // fused or unfused according to r->fusion (qb_api.cu)
int64_t plan_debug(uint32_t n, std::vector<Gate>& gates, uint32_t tile_bits, int32_t* pass_of_gate,
                   int32_t* order_in_pass);

// ---- in-process multi-GPU registers (qb_group.cu) ---------------------------------------------------
Group* group_create(int world);             // spawns one worker thread per rank (device = rank)
void group_destroy(Group* g);               // joins the workers
Group* group_of(Reg* r);                    // the register's group when called from OUTSIDE its workers, else nullptr
int group_world(const Group* g);
Reg* group_member(Group* g, int rank);
void group_set_member(Group* g, int rank, Reg* r);
void group_barrier(Group* g);               // workers only
int group_requested_world();                // QB_NUM_GPUS (2, 4, 8), else 1
}  // namespace qb
#include <functional>
namespace qb {
int group_run(Group* g, const std::function<int(Reg*, int)>& fn);   // fn(shard, rank) on every worker; first failure in rank order

// ---- specialised tile kernels (qb_jit.cu): process-wide compiler state ------------------------------
int jit_drain();           // wait until no compile is pending
void jit_global_stats(int64_t* compiled, int64_t* failed, int64_t* pending, int64_t* compile_ms);

}  // namespace qb
