// qb_api.cu -- the extern "C" boundary (include/quaru_b200.h): handle lifetime, argument
// validation in the reference's order, the gate queue, measurement and state access.
// Each entry point cites the reference item it stands in for (src/register.rs).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>

#include "qb_common.h"
#include "qb_shard.h"
#include "qb_tile.h"

namespace qb {
const char* last_error_str();
int64_t last_error_payload();

// register.rs:146-153 (get_duplicate :418-429 first, then the range check in `targets` order)
static int validate_targets(uint32_t n, const uint32_t* targets, uint32_t k) {
    if (k > 0 && !targets) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "targets is NULL");
    // n <= 40, so any 41 targets already contain a duplicate or an out-of-range qubit: looking at no more than that
    // keeps the sort below from sizing a buffer with a garbage arity
    k = std::min(k, 41u);
    if (k >= 2) {
        std::vector<uint32_t> s(targets, targets + k);
        std::sort(s.begin(), s.end());
        for (uint32_t i = 1; i < k; ++i)
            if (s[i] == s[i - 1]) return set_error(QB_ERR_INVALID_TARGET, s[i], "InvalidTarget(%u): duplicate target", s[i]);
    }
    for (uint32_t j = 0; j < k; ++j)
        if (targets[j] >= n)
            return set_error(QB_ERR_INVALID_TARGET, targets[j], "InvalidTarget(%u): register has %u qubits", targets[j], n);
    return QB_OK;
}

static int alloc_common(Reg* r) {
    QB_CUDA(cudaGetDevice(&r->device));
    QB_CUDA(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    QB_CUDA(cudaEventCreate(&r->ev0));
    QB_CUDA(cudaEventCreate(&r->ev1));
    QB_CUDA(cudaEventCreateWithFlags(&r->prog_free, cudaEventDisableTiming));
    const uint64_t n_amps = 1ull << r->n_local;
    cudaError_t e = cudaMalloc((void**)&r->state, n_amps * sizeof(c64));
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(e == cudaErrorMemoryAllocation ? QB_ERR_OOM : QB_ERR_CUDA, (int64_t)e,
                         "cannot allocate %.3f GB for a %u-qubit shard: %s", n_amps * 16.0 / 1e9, r->n_local,
                         cudaGetErrorString(e));
    }
    r->phys_of.resize(r->n);
    for (uint32_t q = 0; q < r->n; ++q) r->phys_of[q] = q;
    r->label.resize(r->n);
    for (uint32_t q = 0; q < r->n; ++q) r->label[q] = q;
    // SWAP gates as relabelling: for registers whose sweeps are expensive (QB_SWAP_RELABEL=1 / 0 forces it on / off)
    const char* rl = getenv("QB_SWAP_RELABEL");
    r->relabel = rl ? rl[0] == '1' : r->n_local >= 28;
    const char* ph = getenv("QB_PEEPHOLE");   // H ; CX -> CZ ; H and friends: same policy, QB_PEEPHOLE=1 / 0 forces it on / off
    r->peephole = ph ? ph[0] == '1' : r->n_local >= 28;
    if (const char* j = getenv("QB_JIT")) r->jit = (j[0] >= '0' && j[0] <= '2') ? j[0] - '0' : 1;
    if (const char* j = getenv("QB_JIT_MIN_QUBITS")) r->jit_min_qubits = (uint32_t)atoi(j);
    if (const char* j = getenv("QB_JIT_VARIANT")) {
        const int v = atoi(j);
        r->jit_min_blocks = (v & 15) ? (v & 15) : 2; r->jit_persistent = (v >> 4) & 1; r->jit_prefetch = (v >> 5) & 1;
        r->jit_pipelined = (v >> 6) & 1;
    }
    return QB_OK;
}

static void free_reg(Reg* r) {
    if (!r) return;
    cudaSetDevice(r->device);
    if (r->stream) cudaStreamSynchronize(r->stream);
    shard_destroy(r);
    if (r->state) cudaFree(r->state);
    if (r->scratch) cudaFree(r->scratch);
    if (r->partials) cudaFree(r->partials);
    if (r->dmat) cudaFree(r->dmat);
    if (r->dprog) cudaFree(r->dprog);
    if (r->hpin) cudaFreeHost(r->hpin);
    for (cudaEvent_t e : r->tile_events) cudaEventDestroy(e);
    if (r->ev0) cudaEventDestroy(r->ev0);
    if (r->ev1) cudaEventDestroy(r->ev1);
    if (r->prog_free) cudaEventDestroy(r->prog_free);
    if (r->stream) cudaStreamDestroy(r->stream);
    delete r;
}

static int create_impl(uint32_t n, uint64_t basis_index, int rank, int world, const void* uid, int device, qb_reg** out,
                       Group* group = nullptr) {
    if (!out) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "out is NULL");
    *out = nullptr;
    if (n > 40) return set_error(QB_ERR_UNSUPPORTED, n, "%u qubits: at most 40 are supported", n);
    if (n < 64 && basis_index >= (1ull << n))
        return set_error(QB_ERR_INVALID_ARGUMENT, (int64_t)basis_index,
                         "basis index %llu out of range for %u qubits (reference: index out of bounds panic, math.rs:46)",
                         (unsigned long long)basis_index, n);
    if (world < 1 || (world & (world - 1)) || rank < 0 || rank >= world)
        return set_error(QB_ERR_INVALID_ARGUMENT, world, "world must be a power of two and 0 <= rank < world");
    uint32_t g = 0;
    while ((1 << g) < world) ++g;
    if (g > n) return set_error(QB_ERR_INVALID_ARGUMENT, world, "%d ranks need at least %u qubits", world, g);
    if (device >= 0) QB_CUDA(cudaSetDevice(device));
    Reg* r = new (std::nothrow) Reg();
    if (!r) return set_error(QB_ERR_OOM, -1, "host allocation failed");
    r->n = n;
    r->n_local = n - g;
    r->rank = rank;
    r->world = world;
    r->group = group;
    int rc = alloc_common(r);
    if (group) {
        // in-process register: every rank publishes its shard, and all of them learn whether every allocation succeeded
        // BEFORE anybody enters a collective (a rank that failed would leave the others waiting in ncclCommInitRank)
        group_set_member(group, rank, r);
        group_barrier(group);
        for (int p = 0; p < world && rc == QB_OK; ++p)
            if (!group_member(group, p) || !group_member(group, p)->state) rc = set_error(QB_ERR_OOM, p, "shard %d of the in-process register could not be allocated", p);
        group_barrier(group);
    }
    if (rc == QB_OK && world > 1) rc = shard_init(r, uid);
    if (rc == QB_OK) {
        const uint64_t owner = basis_index >> r->n_local;
        rc = k_init_basis(r, basis_index & ((1ull << r->n_local) - 1), owner == (uint64_t)rank);
    }
    if (rc != QB_OK) {
        if (group) group_set_member(group, rank, nullptr);
        free_reg(r);
        return rc;
    }
    *out = reinterpret_cast<qb_reg*>(r);
    return QB_OK;
}

// Register::new with QB_NUM_GPUS = P in the environment: the same sharded register as qb_create_sharded, its P shards driven
// by P worker threads of this process (qb_group.cu).  The handle is rank 0's shard.
static int create_group(uint32_t n, uint64_t basis_index, int world, qb_reg** out) {
    if (!out) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    QB_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < world) return set_error(QB_ERR_NO_DEVICE, ndev, "QB_NUM_GPUS=%d but only %d CUDA device(s) are visible", world, ndev);
    unsigned char uid[128];
    QB_TRY(shard_unique_id(uid));
    Group* G = group_create(world);
    const int rc = group_run(G, [&](Reg*, int rank) -> int {
        qb_reg* m = nullptr;
        return create_impl(n, basis_index, rank, world, uid, rank, &m, G);
    });
    if (rc != QB_OK) {
        // ranks that did succeed are released by their own workers (shard_destroy is collective over the survivors only when
        // all of them survived; after a failed creation the communicator is abandoned, not torn down collectively)
        for (int p = 0; p < world; ++p)
            if (Reg* m = group_member(G, p)) { m->shard = nullptr; }
        group_run(G, [&](Reg* m, int) -> int { if (m) free_reg(m); return QB_OK; });
        group_destroy(G);
        return rc;
    }
    *out = reinterpret_cast<qb_reg*>(group_member(G, 0));
    return QB_OK;
}

// Launch everything queued.  Fusion on: plan tile passes (qb_tile.cu); off: one kernel per gate.  `more_coming`: the queue
// overflowed in the middle of a gate stream -- the last planned pass, usually half empty, stays queued for the next flush.
static int flush(Reg* r, bool more_coming = false) {
    if (r->queue.empty()) return QB_OK;
    QB_CUDA(cudaSetDevice(r->device));
    std::vector<Gate> gates;
    gates.swap(r->queue);
    r->held = 0;
    if (r->peephole && r->fusion) {
        r->st_peephole += (int64_t)peephole_h_cx(&gates);
        if (gates.empty()) return QB_OK;   // everything cancelled
    }
    if (r->world > 1) return shard_run(r, gates);
    if (r->fusion && more_coming) {
        std::vector<Gate> hold;
        QB_TRY(run_fused(r, gates, &hold));
        r->queue.insert(r->queue.begin(), hold.begin(), hold.end());   // already in engine qubits: not through enqueue_gate again
        r->held = hold.size();
        return QB_OK;
    }
    if (r->fusion) return run_fused(r, gates);
    for (const Gate& g : gates) QB_TRY(k_apply_gate_unfused(r, g));
    return QB_OK;
}

int run_local(Reg* r, std::vector<Gate>& gates) {
    if (r->fusion) return run_fused(r, gates);
    for (const Gate& g : gates) QB_TRY(k_apply_gate_unfused(r, g));
    return QB_OK;
}

// Common tail of every gate entry point: a plain SWAP only renames two qubits (when relabelling is on), everything else is
// translated to engine qubits and queued.
static int enqueue_gate(Reg* r, Gate&& g) {
    if (r->relabel) {
        if (is_plain_swap(g)) {
            std::swap(r->label[g.targets[0]], r->label[g.targets[1]]);
            return QB_OK;
        }
        relabel_gate(r->label, &g);
    }
    r->queue.push_back(std::move(g));
    if (!r->fusion) return flush(r);
    if (r->queue.size() >= r->max_queue + r->held) return flush(r, true);   // `held` gates came back from the last overflow
                                                                             // flush: only new ones count
    return QB_OK;
}

static int enqueue(Reg* r, const double* matrix, const uint32_t* targets, uint32_t k) {
    Gate g;
    r->st_gates++;
    if (!classify_gate(matrix, targets, k, &g)) return QB_OK;  // exact identity
    return enqueue_gate(r, std::move(g));
}

static void reset_labels(Reg* r) {
    for (uint32_t q = 0; q < r->n; ++q) r->label[q] = q;
}

// Before raw state access: pending gates launched, the caller's qubit q back in index bit q (real SWAP passes undo the
// relabelling, then the sharded placement is restored).
static int canonical(Reg* r) {
    if (r->relabel) {
        // the swaps are derived on a COPY of the label map; the map itself becomes the identity only once the SWAP passes have
        // been launched -- if the flush fails the labels still describe the resident layout (the gates stay dropped)
        std::vector<uint32_t> pairs, label = r->label;
        canonical_swaps(&label, &pairs);
        for (size_t i = 0; i + 1 < pairs.size(); i += 2) {
            Gate g;
            make_swap_gate(pairs[i], pairs[i + 1], &g);
            r->queue.push_back(std::move(g));
        }
        QB_TRY(flush(r));
        r->label = label;
    }
    QB_TRY(flush(r));
    if (r->world > 1) QB_TRY(shard_canonicalize(r));
    return QB_OK;
}

}  // namespace qb

using namespace qb;

#define REG(p) reinterpret_cast<qb::Reg*>(p)
// every entry point validates the handle and makes its device current (the library's CUDA runtime keeps its
// own per-thread current device)
#define CHECK_REG(p)                                                                           \
    if (!(p)) return qb::set_error(QB_ERR_INVALID_ARGUMENT, -1, "register handle is NULL");    \
    QB_CUDA(cudaSetDevice(REG(p)->device))

// In-process multi-GPU registers: a call made from outside the register's worker threads is executed by every worker on its
// own shard (`m`), collectively; the workers' own (nested) calls take the ordinary path below.
#define GROUP_CALL(reg, expr)                                                                        \
    if (qb::Group* G_ = qb::group_of(REG(reg)))                                                      \
        return qb::group_run(G_, [&](qb::Reg* m_, int rank_) -> int {                                \
            qb_reg* m = reinterpret_cast<qb_reg*>(m_);                                               \
            (void)rank_;                                                                             \
            return (expr);                                                                           \
        })

extern "C" {

const char* qb_last_error(void) { return qb::last_error_str(); }
int64_t qb_last_error_payload(void) { return qb::last_error_payload(); }
int64_t qb_shard_plan_debug(uint32_t n, int32_t world, int32_t rank, const qb_gate* gates, uint64_t count, int32_t* events,
                            uint64_t capacity, uint32_t* final_phys_of) try {
    if ((count && !gates) || (capacity && !events)) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "NULL argument");
    if (world < 1 || (world & (world - 1)) || rank < 0 || rank >= world)
        return set_error(QB_ERR_INVALID_ARGUMENT, world, "world must be a power of two and 0 <= rank < world");
    std::vector<Gate> gs;
    for (uint64_t i = 0; i < count; ++i) {
        int rc = validate_targets(n, gates[i].targets, gates[i].k);
        if (rc != QB_OK) return rc;
        Gate g;
        if (!classify_gate(gates[i].matrix, gates[i].targets, gates[i].k, &g)) continue;
        g.tag = (int64_t)i;
        gs.push_back(std::move(g));
    }
    return shard_plan_debug(n, world, rank, gs, events, capacity, final_phys_of);
} QB_CATCH_ALL

int64_t qb_rounds_debug(uint32_t n, const qb_gate* gates, uint64_t count, void* out, uint64_t capacity) try {
    if ((count && !gates) || (capacity && !out)) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "NULL argument");
    std::vector<Gate> gs;
    for (uint64_t i = 0; i < count; ++i) {
        int rc = validate_targets(n, gates[i].targets, gates[i].k);
        if (rc != QB_OK) return rc;
        Gate g;
        if (!classify_gate(gates[i].matrix, gates[i].targets, gates[i].k, &g)) continue;
        g.tag = (int64_t)i;
        gs.push_back(std::move(g));
    }
    return rounds_debug(n, gs, (unsigned char*)out, capacity);
} QB_CATCH_ALL

const char* qb_version(void) { return "quaru_b200 0.1 (sm_100a)"; }

int qb_create(uint32_t n, uint64_t basis_index, qb_reg** out) try {
    const int world = group_requested_world();   // QB_NUM_GPUS: Register::new keeps its signature (register.rs:41)
    uint32_t g = 0;
    while ((1 << g) < world) ++g;
    if (world > 1 && n >= 2 * g + 1 && n <= 40 && (n >= 64 || basis_index < (1ull << n))) return create_group(n, basis_index, world, out);
    return create_impl(n, basis_index, 0, 1, nullptr, -1, out);
} QB_CATCH_ALL

int qb_create_sharded(uint32_t n, uint64_t basis_index, int32_t rank, int32_t world, const void* uid, int32_t device,
                      qb_reg** out) try {
    if (world > 1 && !uid) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "nccl_unique_id is NULL");
    return create_impl(n, basis_index, rank, world, uid, device, out);
} QB_CATCH_ALL

int qb_nccl_unique_id(void* out128) try { return shard_unique_id(out128); } QB_CATCH_ALL

// Register::try_from_qubits (register.rs:81-102)
int qb_create_product(const double* qubits, uint32_t n, qb_reg** out) try {
    if (!out) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "out is NULL");
    *out = nullptr;
    if (n > 0 && !qubits) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "qubits is NULL");
    for (uint32_t i = 0; i < n; ++i) {  // is_one (register.rs:107-115)
        double total_prob = -1.0;
        for (int j = 0; j < 2; ++j) {
            const double re = qubits[4 * i + 2 * j], im = qubits[4 * i + 2 * j + 1];
            total_prob += re * re + im * im;
        }
        if (!(fabs(total_prob) < 0.000001)) return set_error(QB_ERR_INVALID_QUBIT, i, "InvalidQubit(%u)", i);
    }
    int rc = create_impl(n, 0, 0, 1, nullptr, -1, out);
    if (rc != QB_OK) return rc;
    rc = k_init_product(REG(*out), qubits);
    if (rc != QB_OK) { free_reg(REG(*out)); *out = nullptr; }
    return rc;
} QB_CATCH_ALL

int qb_clone(qb_reg* src_, qb_reg** out) try {
    CHECK_REG(src_);
    if (!out) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "out is NULL");
    Reg* src = REG(src_);
    *out = nullptr;
    if (src->world > 1) return set_error(QB_ERR_UNSUPPORTED, -1, "clone of a sharded register is not supported");
    QB_TRY(flush(src));
    int rc = create_impl(src->n, 0, 0, 1, nullptr, src->device, out);
    if (rc != QB_OK) return rc;
    Reg* dst = REG(*out);
    dst->fusion = src->fusion; dst->tile_bits = src->tile_bits; dst->max_queue = src->max_queue; dst->force_path = src->force_path;
    dst->jit = src->jit; dst->jit_min_qubits = src->jit_min_qubits; dst->jit_min_blocks = src->jit_min_blocks;
    dst->jit_persistent = src->jit_persistent; dst->jit_prefetch = src->jit_prefetch; dst->jit_pipelined = src->jit_pipelined;
    dst->relabel = src->relabel; dst->label = src->label;   // same amplitudes, same qubit names
    dst->peephole = src->peephole;
    // create_impl queued dst's initialisation (memset + the single 1) on dst->stream, a different non-blocking stream: the copy
    // is ordered after both streams' work by issuing it on dst->stream once src->stream has drained
    cudaError_t e = cudaStreamSynchronize(src->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst->state, src->state, sizeof(c64) << src->n_local, cudaMemcpyDeviceToDevice, dst->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(dst->stream);
    if (e != cudaSuccess) { free_reg(dst); *out = nullptr; return fail_cuda(e, "clone copy", __FILE__, __LINE__); }
    return QB_OK;
} QB_CATCH_ALL

void qb_destroy(qb_reg* reg) {
    if (Group* G = group_of(REG(reg))) {
        group_run(G, [&](Reg* m, int) -> int { free_reg(m); return QB_OK; });   // shard_destroy is collective
        group_destroy(G);
        return;
    }
    free_reg(REG(reg));
}

int qb_reset(qb_reg* reg, uint64_t basis_index) try {
    if (reg) GROUP_CALL(reg, qb_reset(m, basis_index));
    CHECK_REG(reg);
    Reg* r = REG(reg);
    if (basis_index >= (1ull << r->n)) return set_error(QB_ERR_INVALID_ARGUMENT, (int64_t)basis_index, "basis index out of range");
    r->queue.clear();
    r->held = 0;
    shard_reset_placement(r);
    reset_labels(r);
    const uint64_t owner = basis_index >> r->n_local;
    return k_init_basis(r, basis_index & ((1ull << r->n_local) - 1), owner == (uint64_t)r->rank);
} QB_CATCH_ALL

uint32_t qb_num_qubits(const qb_reg* reg) { return reg ? reinterpret_cast<const qb::Reg*>(reg)->n : 0; }

// Register::apply -> try_apply_fast (register.rs:134-191)
int qb_apply(qb_reg* reg, const double* matrix, const uint32_t* targets, uint32_t k) try {
    if (reg) GROUP_CALL(reg, qb_apply(m, matrix, targets, k));
    CHECK_REG(reg);
    Reg* r = REG(reg);
    if (!matrix) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "matrix is NULL");
    QB_TRY(validate_targets(r->n, targets, k));
    return enqueue(r, matrix, targets, k);
} QB_CATCH_ALL

int qb_apply_controlled(qb_reg* reg, const double* matrix, const uint32_t* targets, uint32_t k, const uint32_t* controls,
                        uint32_t n_controls) try {
    if (reg) GROUP_CALL(reg, qb_apply_controlled(m, matrix, targets, k, controls, n_controls));
    CHECK_REG(reg);
    Reg* r = REG(reg);
    if (!matrix) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "matrix is NULL");
    if (n_controls && !controls) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "controls is NULL");
    if (k && !targets) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "targets is NULL");
    if (k > 40 || n_controls > 40) return set_error(QB_ERR_INVALID_TARGET, -1, "InvalidTarget: more targets or controls than qubits");
    std::vector<uint32_t> all(targets, targets + k);
    all.insert(all.end(), controls, controls + n_controls);
    QB_TRY(validate_targets(r->n, all.data(), (uint32_t)all.size()));
    Gate g;
    r->st_gates++;
    if (!classify_gate(matrix, targets, k, &g)) return QB_OK;
    for (uint32_t i = 0; i < n_controls; ++i) g.ctrl_mask |= 1ull << controls[i];
    return enqueue_gate(r, std::move(g));
} QB_CATCH_ALL

// Register::try_apply_all (register.rs:267-281) == the 2x2 gate on every qubit
int qb_apply_all(qb_reg* reg, const double* m2x2) try {
    if (reg) GROUP_CALL(reg, qb_apply_all(m, m2x2));
    CHECK_REG(reg);
    Reg* r = REG(reg);
    if (!m2x2) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "matrix is NULL");
    for (uint32_t q = 0; q < r->n; ++q) QB_TRY(enqueue(r, m2x2, &q, 1));
    return QB_OK;
} QB_CATCH_ALL

// gate loop of Register::apply_circuit (register.rs:346-350)
int qb_apply_batch(qb_reg* reg, const qb_gate* gates, uint64_t count) try {
    if (reg) GROUP_CALL(reg, qb_apply_batch(m, gates, count));
    CHECK_REG(reg);
    Reg* r = REG(reg);
    if (count && !gates) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "gates is NULL");
    for (uint64_t i = 0; i < count; ++i) {
        if (!gates[i].matrix) return set_error(QB_ERR_INVALID_ARGUMENT, (int64_t)i, "gate %llu: matrix is NULL", (unsigned long long)i);
        QB_TRY(validate_targets(r->n, gates[i].targets, gates[i].k));
    }
    for (uint64_t i = 0; i < count; ++i) QB_TRY(enqueue(r, gates[i].matrix, gates[i].targets, gates[i].k));
    return QB_OK;
} QB_CATCH_ALL

// first loop of try_measure (register.rs:303-315)
int qb_probabilities(qb_reg* reg, uint32_t target, double* p0, double* p1) try {
    if (reg && p0 && p1)
        if (Group* G_ = group_of(REG(reg)))
            return group_run(G_, [&](Reg* m_, int rank_) -> int {
                double a, b;
                return qb_probabilities(reinterpret_cast<qb_reg*>(m_), target, rank_ ? &a : p0, rank_ ? &b : p1);
            });
    CHECK_REG(reg);
    Reg* r = REG(reg);
    if (!p0 || !p1) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "output pointer is NULL");
    if (target >= r->n) return set_error(QB_ERR_INVALID_TARGET, target, "InvalidTarget(%u): register has %u qubits", target, r->n);
    QB_TRY(flush(r));
    const uint32_t t = r->relabel ? r->label[target] : target;   // the engine qubit that holds the caller's qubit
    if (r->world > 1) return shard_probabilities(r, t, p0, p1);
    return k_probabilities(r, (int)t, p0, p1);
} QB_CATCH_ALL

// Register::try_measure (register.rs:297-342)
int qb_measure(qb_reg* reg, uint32_t target, double u, int32_t* outcome) try {
    if (reg && outcome)
        if (Group* G_ = group_of(REG(reg)))
            return group_run(G_, [&](Reg* m_, int rank_) -> int {
                int32_t o;   // every rank derives the same outcome from the same probabilities and the same u
                return qb_measure(reinterpret_cast<qb_reg*>(m_), target, u, rank_ ? &o : outcome);
            });
    CHECK_REG(reg);
    Reg* r = REG(reg);
    if (!outcome) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "outcome is NULL");
    double p0, p1;
    QB_TRY(qb_probabilities(reg, target, &p0, &p1));
    const int res = u > p0;                    // :321 (strict)
    const double total = res ? p1 : p0;        // :323
    *outcome = res;
    const uint32_t t = r->relabel ? r->label[target] : target;   // valid: qb_probabilities checked the range
    if (r->world > 1) return shard_collapse(r, t, res, total);
    return k_collapse(r, t, res, total);
} QB_CATCH_ALL

int qb_read_state(qb_reg* reg, uint64_t offset, uint64_t count, double* out) try {
    if (reg) GROUP_CALL(reg, qb_read_state(m, offset, count, out));
    CHECK_REG(reg);
    Reg* r = REG(reg);
    const uint64_t N = 1ull << r->n;
    if (offset > N || count > N - offset) return set_error(QB_ERR_INVALID_ARGUMENT, (int64_t)offset, "range exceeds the state");
    if (count && !out) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "out is NULL");
    QB_TRY(canonical(r));
    const uint64_t lo = (uint64_t)r->rank << r->n_local, hi = lo + (1ull << r->n_local);
    const uint64_t a = std::max(offset, lo), b = std::min(offset + count, hi);
    if (a < b)
        QB_CUDA(cudaMemcpyAsync(out + 2 * (a - offset), r->state + (a - lo), (b - a) * sizeof(c64), cudaMemcpyDeviceToHost, r->stream));
    QB_CUDA(cudaStreamSynchronize(r->stream));
    return QB_OK;
} QB_CATCH_ALL

int qb_write_state(qb_reg* reg, uint64_t offset, uint64_t count, const double* in) try {
    if (reg) GROUP_CALL(reg, qb_write_state(m, offset, count, in));
    CHECK_REG(reg);
    Reg* r = REG(reg);
    const uint64_t N = 1ull << r->n;
    if (offset > N || count > N - offset) return set_error(QB_ERR_INVALID_ARGUMENT, (int64_t)offset, "range exceeds the state");
    if (count && !in) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "in is NULL");
    QB_TRY(canonical(r));
    const uint64_t lo = (uint64_t)r->rank << r->n_local, hi = lo + (1ull << r->n_local);
    const uint64_t a = std::max(offset, lo), b = std::min(offset + count, hi);
    if (a < b)
        QB_CUDA(cudaMemcpyAsync(r->state + (a - lo), in + 2 * (a - offset), (b - a) * sizeof(c64), cudaMemcpyHostToDevice, r->stream));
    QB_CUDA(cudaStreamSynchronize(r->stream));
    return QB_OK;
} QB_CATCH_ALL

int qb_save_state(qb_reg* reg, const char* path) try {
    if (reg && path)
        if (Group* G_ = group_of(REG(reg)))   // one file per shard: <path>.r<rank>
            return group_run(G_, [&](Reg* m_, int rank_) -> int {
                const std::string p = std::string(path) + ".r" + std::to_string(rank_);
                return qb_save_state(reinterpret_cast<qb_reg*>(m_), p.c_str());
            });
    CHECK_REG(reg);
    Reg* r = REG(reg);
    if (!path) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "path is NULL");
    QB_TRY(canonical(r));
    return state_save(r, path);
} QB_CATCH_ALL

int qb_load_state(qb_reg* reg, const char* path) try {
    if (reg && path)
        if (Group* G_ = group_of(REG(reg)))   // one file per shard: <path>.r<rank>
            return group_run(G_, [&](Reg* m_, int rank_) -> int {
                const std::string p = std::string(path) + ".r" + std::to_string(rank_);
                return qb_load_state(reinterpret_cast<qb_reg*>(m_), p.c_str());
            });
    CHECK_REG(reg);
    Reg* r = REG(reg);
    if (!path) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "path is NULL");
    QB_TRY(state_load_check(r, path));   // missing file / wrong register: nothing below happens, the handle is untouched
    r->queue.clear();  // the loaded image replaces whatever the pending gates would have produced
    r->held = 0;
    shard_reset_placement(r);  // files hold the identity placement
    reset_labels(r);
    return state_load(r, path);
} QB_CATCH_ALL

int qb_norm2(qb_reg* reg, double* out) try {
    if (reg && out)
        if (Group* G_ = group_of(REG(reg)))
            return group_run(G_, [&](Reg* m_, int rank_) -> int {
                double o;
                return qb_norm2(reinterpret_cast<qb_reg*>(m_), rank_ ? &o : out);
            });
    CHECK_REG(reg);
    Reg* r = REG(reg);
    if (!out) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "out is NULL");
    QB_TRY(flush(r));
    double p0, p1;
    if (r->world > 1) QB_TRY(shard_probabilities(r, UINT32_MAX, &p0, &p1));
    else QB_TRY(k_probabilities(r, -1, &p0, &p1));
    *out = p0;
    return QB_OK;
} QB_CATCH_ALL

static int diff_impl(qb_reg* a_, qb_reg* b_, double* mx, double* l2) {
    if (a_ && b_)
        if (Group* Ga = group_of(REG(a_))) {
            Group* Gb = REG(b_)->group;
            if (!Gb || group_world(Gb) != group_world(Ga)) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "registers differ in size, sharding or device");
            return group_run(Ga, [&](Reg* m_, int rank_) -> int {
                double x, y;
                return diff_impl(reinterpret_cast<qb_reg*>(m_), reinterpret_cast<qb_reg*>(group_member(Gb, rank_)), rank_ ? &x : mx, rank_ ? &y : l2);
            });
        }
    CHECK_REG(a_);
    CHECK_REG(b_);
    Reg *a = REG(a_), *b = REG(b_);
    if (a->n != b->n || a->world != b->world || a->device != b->device)
        return set_error(QB_ERR_INVALID_ARGUMENT, -1, "registers differ in size, sharding or device");
    QB_TRY(canonical(a));
    QB_TRY(canonical(b));
    double m, l;
    QB_TRY(k_diff(a, b, &m, &l));
    if (a->world > 1) QB_TRY(shard_reduce_diff(a, &m, &l));
    *mx = m;
    *l2 = l;
    return QB_OK;
}

int qb_max_abs_diff(qb_reg* a, qb_reg* b, double* out) try {
    if (!out) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "out is NULL");
    double l2;
    return diff_impl(a, b, out, &l2);
} QB_CATCH_ALL

int qb_l2_diff2(qb_reg* a, qb_reg* b, double* out) try {
    if (!out) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "out is NULL");
    double mx;
    return diff_impl(a, b, &mx, out);
} QB_CATCH_ALL

int qb_flush(qb_reg* reg) try {
    if (reg) GROUP_CALL(reg, qb_flush(m));
    CHECK_REG(reg);
    return flush(REG(reg));
} QB_CATCH_ALL

int qb_sync(qb_reg* reg) try {
    if (reg) GROUP_CALL(reg, qb_sync(m));
    CHECK_REG(reg);
    Reg* r = REG(reg);
    QB_TRY(flush(r));
    QB_CUDA(cudaStreamSynchronize(r->stream));
    return QB_OK;
} QB_CATCH_ALL

int qb_timer_start(qb_reg* reg) try {
    if (reg) GROUP_CALL(reg, qb_timer_start(m));
    CHECK_REG(reg);
    Reg* r = REG(reg);
    QB_TRY(flush(r));
    QB_CUDA(cudaEventRecord(r->ev0, r->stream));
    return QB_OK;
} QB_CATCH_ALL

int qb_timer_stop(qb_reg* reg, float* ms) try {
    if (reg && ms)
        if (Group* G_ = group_of(REG(reg))) {
            float per_rank[8] = {0};
            QB_TRY(group_run(G_, [&](Reg* m_, int rank_) -> int { return qb_timer_stop(reinterpret_cast<qb_reg*>(m_), &per_rank[rank_]); }));
            *ms = 0.f;
            for (int p = 0; p < group_world(G_); ++p) *ms = std::max(*ms, per_rank[p]);   // max over the shards
            return QB_OK;
        }
    CHECK_REG(reg);
    Reg* r = REG(reg);
    if (!ms) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "ms is NULL");
    QB_TRY(flush(r));
    QB_CUDA(cudaEventRecord(r->ev1, r->stream));
    QB_CUDA(cudaEventSynchronize(r->ev1));
    QB_CUDA(cudaEventElapsedTime(ms, r->ev0, r->ev1));
    return QB_OK;
} QB_CATCH_ALL

int qb_set_option(qb_reg* reg, int32_t option, int64_t value) try {
    if (reg) GROUP_CALL(reg, qb_set_option(m, option, value));
    CHECK_REG(reg);
    Reg* r = REG(reg);
    QB_TRY(flush(r));
    switch (option) {
        case QB_OPT_FUSION: r->fusion = value != 0; return QB_OK;
        case QB_OPT_TILE_BITS:
            if (value < 8 || value > 13) return set_error(QB_ERR_INVALID_ARGUMENT, value, "tile bits must be in [8, 13]");
            r->tile_bits = (uint32_t)value;
            return QB_OK;
        case QB_OPT_MAX_QUEUE:
            if (value < 1) return set_error(QB_ERR_INVALID_ARGUMENT, value, "max queue must be >= 1");
            r->max_queue = (uint64_t)value;
            return QB_OK;
        case QB_OPT_PROFILE: r->profile = value != 0; return QB_OK;
        case QB_OPT_PEEPHOLE: r->peephole = value != 0; return QB_OK;
        case QB_OPT_JIT:
            if (value < 0 || value > 2) return set_error(QB_ERR_INVALID_ARGUMENT, value, "jit must be 0, 1 or 2");
            r->jit = (int)value;
            return QB_OK;
        case QB_OPT_JIT_MIN_QUBITS: r->jit_min_qubits = (uint32_t)std::max<int64_t>(value, 12); return QB_OK;
        case QB_OPT_JIT_VARIANT:
            if ((value & 15) < 1 || (value & 15) > 4 || value > 127) return set_error(QB_ERR_INVALID_ARGUMENT, value, "bad jit variant");
            r->jit_min_blocks = (int)(value & 15); r->jit_persistent = (int)((value >> 4) & 1); r->jit_prefetch = (int)((value >> 5) & 1);
            r->jit_pipelined = (int)((value >> 6) & 1);
            return QB_OK;
        case QB_OPT_FORCE_PATH:
            if (value < 0 || value > 5) return set_error(QB_ERR_INVALID_ARGUMENT, value, "force path must be in [0, 5]");
            r->force_path = (int)value;
            return QB_OK;
    }
    return set_error(QB_ERR_INVALID_ARGUMENT, option, "unknown option %d", option);
} QB_CATCH_ALL

int qb_get_stat(qb_reg* reg, int32_t stat, int64_t* value) try {
    if (reg && value)
        if (Group* G_ = group_of(REG(reg)))
            return group_run(G_, [&](Reg* m_, int rank_) -> int {
                int64_t v;   // rank 0's counters (every rank launches the same passes and exchanges)
                return qb_get_stat(reinterpret_cast<qb_reg*>(m_), stat, rank_ ? &v : value);
            });
    CHECK_REG(reg);
    Reg* r = REG(reg);
    if (!value) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "value is NULL");
    switch (stat) {
        case QB_STAT_KERNEL_LAUNCHES: *value = r->st_launches; return QB_OK;
        case QB_STAT_PASSES: *value = r->st_passes; return QB_OK;
        case QB_STAT_GATES: *value = r->st_gates; return QB_OK;
        case QB_STAT_BYTES_MOVED: *value = r->st_bytes; return QB_OK;
        case QB_STAT_SWAP_BYTES: *value = r->st_swap_bytes; return QB_OK;
        case QB_STAT_SWAPS: *value = r->st_swaps; return QB_OK;
        case QB_STAT_TILE_LAUNCHES: QB_TRY(drain_tile_events(r)); *value = r->st_tile_launches; return QB_OK;
        case QB_STAT_ROUNDS: *value = r->st_rounds; return QB_OK;
        case QB_STAT_P2P_EXCHANGES: *value = r->st_p2p; return QB_OK;
        case QB_STAT_REMOTE_PASSES: *value = r->st_remote_passes; return QB_OK;
        case QB_STAT_TILE_NS: QB_TRY(drain_tile_events(r)); *value = r->st_tile_ns; return QB_OK;
        case QB_STAT_JIT_LAUNCHES: *value = r->st_jit_launches; return QB_OK;
        case QB_STAT_JIT_COMPILES: *value = r->st_jit_compiles; return QB_OK;
        case QB_STAT_JIT_MISSES: *value = r->st_jit_misses; return QB_OK;
        case QB_STAT_FUSED_EXCHANGES: *value = r->st_fused_exchanges; return QB_OK;
        case QB_STAT_PEEPHOLE_REWRITES: *value = r->st_peephole; return QB_OK;
        case QB_STAT_JIT_COMPILE_MS:
        case QB_STAT_JIT_PENDING:
        case QB_STAT_JIT_FAILED: {
            int64_t c, f, p, ms;
            jit_global_stats(&c, &f, &p, &ms);
            *value = stat == QB_STAT_JIT_COMPILE_MS ? ms : stat == QB_STAT_JIT_PENDING ? p : f;
            return QB_OK;
        }
    }
    return set_error(QB_ERR_INVALID_ARGUMENT, stat, "unknown stat %d", stat);
} QB_CATCH_ALL

int64_t qb_relabel_debug(uint32_t n, const qb_gate* gates, uint64_t count, int32_t* out, uint64_t capacity) try {
    if ((count && !gates) || !out) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "NULL argument");
    if (n == 0 || n > 40) return set_error(QB_ERR_INVALID_ARGUMENT, n, "1..40 qubits");
    std::vector<uint32_t> label(n);
    for (uint32_t q = 0; q < n; ++q) label[q] = q;
    std::vector<int32_t> rec;
    for (uint64_t i = 0; i < count; ++i) {
        QB_TRY(validate_targets(n, gates[i].targets, gates[i].k));
        Gate g;
        if (!classify_gate(gates[i].matrix, gates[i].targets, gates[i].k, &g)) { rec.push_back(-2); continue; }
        if (is_plain_swap(g)) {   // the same two steps as enqueue_gate
            std::swap(label[g.targets[0]], label[g.targets[1]]);
            rec.push_back(-1);
            continue;
        }
        relabel_gate(label, &g);
        // cross-check: classifying the gate on the translated qubits must give the same gate
        std::vector<uint32_t> tt(gates[i].k);
        for (uint32_t j = 0; j < gates[i].k; ++j) tt[j] = label[gates[i].targets[j]];
        Gate c;
        classify_gate(gates[i].matrix, tt.data(), gates[i].k, &c);
        if (c.targets != g.targets || c.ctrl_mask != g.ctrl_mask || c.kind != g.kind || c.k != g.k)
            return set_error(QB_ERR_UNSUPPORTED, (int64_t)i, "internal: relabelled gate %llu differs from the gate on the translated qubits", (unsigned long long)i);
        rec.push_back((int32_t)gates[i].k);
        for (uint32_t t : tt) rec.push_back((int32_t)t);
    }
    std::vector<uint32_t> pairs;
    canonical_swaps(&label, &pairs);
    for (uint32_t q = 0; q < n; ++q)
        if (label[q] != q) return set_error(QB_ERR_UNSUPPORTED, q, "internal: canonical_swaps left qubit %u out of place", q);
    rec.push_back(-3);
    rec.push_back((int32_t)(pairs.size() / 2));
    for (uint32_t p : pairs) rec.push_back((int32_t)p);
    if (rec.size() > capacity) return set_error(QB_ERR_INVALID_ARGUMENT, (int64_t)rec.size(), "buffer too small");
    for (size_t i = 0; i < rec.size(); ++i) out[i] = rec[i];
    return (int64_t)rec.size();
} QB_CATCH_ALL

int64_t qb_peephole_debug(uint32_t n, const qb_gate* gates, uint64_t count, double* out, uint64_t capacity) try {
    if ((count && !gates) || !out) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "NULL argument");
    if (n == 0 || n > 40) return set_error(QB_ERR_INVALID_ARGUMENT, n, "1..40 qubits");
    std::vector<Gate> gs;
    for (uint64_t i = 0; i < count; ++i) {
        QB_TRY(validate_targets(n, gates[i].targets, gates[i].k));
        Gate g;
        if (!classify_gate(gates[i].matrix, gates[i].targets, gates[i].k, &g)) continue;
        g.tag = (int64_t)i;
        gs.push_back(std::move(g));
    }
    const size_t rewrites = peephole_h_cx(&gs);
    std::vector<double> rec;
    rec.push_back((double)rewrites);
    for (const Gate& g : gs) {
        rec.push_back(g.kind == GK_DIAG ? 1.0 : 0.0);
        rec.push_back((double)g.k);
        rec.push_back((double)g.ctrl_mask);   // n <= 40: exact in a double
        for (uint32_t t : g.targets) rec.push_back((double)t);
        rec.push_back((double)g.matrix.size());
        for (const c64& e : g.matrix) { rec.push_back(e.x); rec.push_back(e.y); }
    }
    if (rec.size() > capacity) return set_error(QB_ERR_INVALID_ARGUMENT, (int64_t)rec.size(), "buffer too small");
    for (size_t i = 0; i < rec.size(); ++i) out[i] = rec[i];
    return (int64_t)rec.size();
} QB_CATCH_ALL

int64_t qb_plan_debug(uint32_t n, const qb_gate* gates, uint64_t count, uint32_t tile_bits, int32_t* pass_of_gate,
                      int32_t* order_in_pass) try {
    if (count && (!gates || !pass_of_gate || !order_in_pass)) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "NULL argument");
    std::vector<Gate> gs;
    std::vector<int64_t> src;
    for (uint64_t i = 0; i < count; ++i) {
        int rc = validate_targets(n, gates[i].targets, gates[i].k);
        if (rc != QB_OK) return rc;
        pass_of_gate[i] = -1;
        order_in_pass[i] = -1;
        Gate g;
        if (!classify_gate(gates[i].matrix, gates[i].targets, gates[i].k, &g)) continue;
        g.tag = (int64_t)i;
        gs.push_back(std::move(g));
    }
    return plan_debug(n, gs, tile_bits, pass_of_gate, order_in_pass);
} QB_CATCH_ALL

}  // extern "C"
