// qb_gate.cu -- host-side analysis of the (matrix, targets) payload of an `Operation`
// (reference: src/operation.rs:49-52; constructors :340-536).  The reference always applies the
// dense 2^k x 2^k matrix (src/register.rs:160-188).  Here the structure the constructors put into
// those matrices is recovered so that less of the state has to move:
//   * control qubits  (cnot :372, to_controlled :406, cnx :491): the gate only touches amplitudes
//     whose control bits are all 1;
//   * diagonal gates  (phase :443, pauli_z :481, cnz :520, RZ/CZ/controlled phases, Grover's
//     oracle/diffuser examples/grovers.rs:181-218): no pairing, pure element-wise multiply;
// Both are detected with exact comparisons against 0.0 / 1.0, so the result is the same linear map
// as the dense product, entry for entry.
#include <stdarg.h>
#include <stdio.h>

#include <exception>
#include <new>

#include "qb_common.h"

namespace qb {

static thread_local char g_err[512] = "";
static thread_local int64_t g_payload = -1;

int set_error(int code, int64_t payload, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    g_payload = payload;
    return code;
}

int fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
    int code = (e == cudaErrorMemoryAllocation) ? QB_ERR_OOM
               : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? QB_ERR_NO_DEVICE
                                                                              : QB_ERR_CUDA;
    return set_error(code, (int64_t)e, "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what,
                     file, line);
}

int translate_exception() noexcept {
    try {
        throw;
    } catch (const std::bad_alloc&) {
        return set_error(QB_ERR_OOM, -1, "host allocation failed");
    } catch (const std::exception& e) {
        return set_error(QB_ERR_UNSUPPORTED, -1, "internal error: %s", e.what());
    } catch (...) {
        return set_error(QB_ERR_UNSUPPORTED, -1, "internal error (unknown exception)");
    }
}

const char* last_error_str() { return g_err; }
int64_t last_error_payload() { return g_payload; }

static inline bool is_one(const double* e) { return e[0] == 1.0 && e[1] == 0.0; }
static inline bool is_zero(const double* e) { return e[0] == 0.0 && e[1] == 0.0; }

bool classify_gate(const double* M, const uint32_t* targets, uint32_t k, Gate* out) {
    const uint64_t dim = 1ull << k;
    // --- control bits: matrix bit j such that every row and column with bit j == 0 is the identity's.
    // Row/column r is "identity-like" iff M[r][r] == 1 and all other entries of row r and column r are 0.
    std::vector<uint8_t> ident(dim, 1);
    for (uint64_t r = 0; r < dim; ++r) {
        const double* row = M + 2 * r * dim;
        if (!is_one(row + 2 * r)) { ident[r] = 0; continue; }
        for (uint64_t c = 0; c < dim && ident[r]; ++c) {
            if (c == r) continue;
            if (!is_zero(row + 2 * c) || !is_zero(M + 2 * (c * dim + r))) ident[r] = 0;
        }
    }
    uint64_t ctrl_bits = 0;  // in matrix-bit space
    for (uint32_t j = 0; j < k; ++j) {
        bool ctrl = true;
        for (uint64_t r = 0; r < dim && ctrl; ++r)
            if (!((r >> j) & 1) && !ident[r]) ctrl = false;
        if (ctrl) ctrl_bits |= 1ull << j;
    }
    // --- reduced gate on the non-control bits, inside the all-controls-set block
    std::vector<uint32_t> keep;  // matrix bits kept
    for (uint32_t j = 0; j < k; ++j)
        if (!((ctrl_bits >> j) & 1)) keep.push_back(j);
    const uint32_t kk = (uint32_t)keep.size();
    const uint64_t rdim = 1ull << kk;
    auto expand = [&](uint64_t x) {
        uint64_t v = ctrl_bits;
        for (uint32_t j = 0; j < kk; ++j)
            if ((x >> j) & 1) v |= 1ull << keep[j];
        return v;
    };
    std::vector<uint64_t> ex(rdim);
    for (uint64_t x = 0; x < rdim; ++x) ex[x] = expand(x);

    bool diag = true, identity = true;
    for (uint64_t r = 0; r < rdim && diag; ++r)
        for (uint64_t c = 0; c < rdim; ++c) {
            const double* e = M + 2 * (ex[r] * dim + ex[c]);
            if (r != c && !is_zero(e)) { diag = false; identity = false; break; }
            if (r == c && !is_one(e)) identity = false;
        }
    if (diag && identity) return false;

    out->k = kk;
    out->targets.resize(kk);
    for (uint32_t j = 0; j < kk; ++j) out->targets[j] = targets[keep[j]];
    out->ctrl_mask = 0;
    for (uint32_t j = 0; j < k; ++j)
        if ((ctrl_bits >> j) & 1) out->ctrl_mask |= 1ull << targets[j];
    if (diag) {
        out->kind = GK_DIAG;
        out->matrix.resize(rdim);
        for (uint64_t r = 0; r < rdim; ++r) {
            const double* e = M + 2 * (ex[r] * dim + ex[r]);
            out->matrix[r] = make_double2(e[0], e[1]);
        }
    } else {
        out->kind = GK_DENSE;
        out->matrix.resize(rdim * rdim);
        for (uint64_t r = 0; r < rdim; ++r)
            for (uint64_t c = 0; c < rdim; ++c) {
                const double* e = M + 2 * (ex[r] * dim + ex[c]);
                out->matrix[r * rdim + c] = make_double2(e[0], e[1]);
            }
    }
    return true;
}


// ---- SWAP relabelling ---------------------------------------------------------------------------------------------------
static const double kSwap[32] = {1, 0, 0, 0, 0, 0, 0, 0,  0, 0, 0, 0, 1, 0, 0, 0,  0, 0, 1, 0, 0, 0, 0, 0,  0, 0, 0, 0, 0, 0, 1, 0};

bool is_plain_swap(const Gate& g) {
    if (g.kind != GK_DENSE || g.k != 2 || g.ctrl_mask != 0 || g.targets.size() != 2 || g.matrix.size() != 16) return false;
    for (int i = 0; i < 16; ++i)
        if (g.matrix[i].x != kSwap[2 * i] || g.matrix[i].y != kSwap[2 * i + 1]) return false;
    return true;
}

void relabel_gate(const std::vector<uint32_t>& label, Gate* g) {
    for (uint32_t& t : g->targets) t = label[t];
    uint64_t c = 0;
    for (uint64_t m = g->ctrl_mask; m; m &= m - 1) c |= 1ull << label[__builtin_ctzll(m)];
    g->ctrl_mask = c;
}

void canonical_swaps(std::vector<uint32_t>* label, std::vector<uint32_t>* pairs) {
    std::vector<uint32_t>& L = *label;
    const uint32_t n = (uint32_t)L.size();
    std::vector<uint32_t> at(n);   // engine qubit -> caller's qubit held there
    for (uint32_t q = 0; q < n; ++q) at[L[q]] = q;
    for (uint32_t q = 0; q < n; ++q) {
        if (L[q] == q) continue;
        const uint32_t p = L[q];       // the caller's qubit q lives in engine qubit p: exchange the contents of p and q
        const uint32_t other = at[q];  // the caller's qubit that lives in engine qubit q moves to p
        pairs->push_back(q);
        pairs->push_back(p);
        L[q] = q; at[q] = q;
        L[other] = p; at[p] = other;
    }
}

void make_swap_gate(uint32_t a, uint32_t b, Gate* out) {
    const uint32_t t[2] = {a, b};
    classify_gate(kSwap, t, 2, out);
}

}  // namespace qb
