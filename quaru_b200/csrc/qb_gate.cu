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

#include <math.h>

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

// ---- H / CX / CZ peephole --------------------------------------------------------------------------------------------------
namespace {
bool ph_is_h(const Gate& g, double* s) {
    if (g.kind != GK_DENSE || g.k != 1 || g.ctrl_mask != 0 || g.matrix.size() != 4) return false;
    for (const c64& e : g.matrix) if (e.y != 0.0) return false;
    const double v = g.matrix[0].x;
    if (!(g.matrix[1].x == v && g.matrix[2].x == v && g.matrix[3].x == -v)) return false;
    if (fabs(2.0 * v * v - 1.0) > 1e-15) return false;   // the normalised Hadamard (operation.rs:350), to the last bit or two
    *s = v;
    return true;
}
bool ph_is_cx(const Gate& g) {   // X on one target under >= 1 controls
    if (g.kind != GK_DENSE || g.k != 1 || g.ctrl_mask == 0 || g.matrix.size() != 4) return false;
    return g.matrix[0].x == 0.0 && g.matrix[0].y == 0.0 && g.matrix[1].x == 1.0 && g.matrix[1].y == 0.0 && g.matrix[2].x == 1.0 &&
           g.matrix[2].y == 0.0 && g.matrix[3].x == 0.0 && g.matrix[3].y == 0.0;
}
bool ph_is_cz(const Gate& g) {   // the phase -1 where >= 2 qubits are all 1 (what classify_gate makes of CZ / cnz)
    return g.kind == GK_DIAG && g.k == 0 && g.matrix.size() == 1 && g.matrix[0].x == -1.0 && g.matrix[0].y == 0.0 &&
           __builtin_popcountll(g.ctrl_mask) >= 2;
}
inline bool ph_touches(const Gate& g, uint32_t t) { return (g.qubit_mask() >> t) & 1ull; }
long ph_last_touching(const std::vector<Gate>& v, long before, uint32_t t) {
    for (long i = before - 1; i >= 0; --i)
        if (ph_touches(v[(size_t)i], t)) return i;
    return -1;
}
Gate ph_make_cz(uint64_t qubits, int64_t tag) {
    Gate z;
    z.k = 0;
    z.kind = GK_DIAG;
    z.matrix.assign(1, make_double2(-1.0, 0.0));
    z.ctrl_mask = qubits;
    z.tag = tag;
    return z;
}
Gate ph_make_cx(uint64_t ctrl, uint32_t t, int64_t tag) {
    Gate x;
    x.k = 1;
    x.kind = GK_DENSE;
    x.targets.assign(1, t);
    x.matrix = {make_double2(0, 0), make_double2(1, 0), make_double2(1, 0), make_double2(0, 0)};
    x.ctrl_mask = ctrl;
    x.tag = tag;
    return x;
}
}  // namespace

size_t peephole_h_cx(std::vector<Gate>* gates) {
    std::vector<Gate> out;
    out.reserve(gates->size() + 8);
    size_t rewrites = 0;
    // the Hadamard at out[i] (on t) meets what precedes it on t: H -> both vanish; CX on t -> the H moves in front of it, the
    // CX becomes a CZ, and the H goes on looking
    auto settle_h = [&](long i, uint32_t t) {
        for (;;) {
            double s_i = 0, s_p = 0;
            ph_is_h(out[(size_t)i], &s_i);
            const long p = ph_last_touching(out, i, t);
            if (p < 0) return;
            const Gate& prev = out[(size_t)p];
            if (ph_is_h(prev, &s_p) && fabs(2.0 * s_i * s_p - 1.0) <= 1e-15) {
                out.erase(out.begin() + i);
                out.erase(out.begin() + p);
                ++rewrites;
                return;
            }
            if (ph_is_cx(prev) && prev.targets[0] == t) {          // CX ; H  ->  H ; CZ   (gates between them do not touch t)
                Gate h = out[(size_t)i];
                out.erase(out.begin() + i);
                const Gate z = ph_make_cz(prev.ctrl_mask | (1ull << t), prev.tag);
                out[(size_t)p] = z;
                out.insert(out.begin() + p, h);
                ++rewrites;
                i = p;
                continue;
            }
            if (ph_is_cz(prev) && ((prev.ctrl_mask >> t) & 1ull)) {   // H ; CZ ; H  ->  CX
                const long q = ph_last_touching(out, p, t);
                double s_q = 0;
                if (q >= 0 && ph_is_h(out[(size_t)q], &s_q) && fabs(2.0 * s_i * s_q - 1.0) <= 1e-15) {
                    const Gate x = ph_make_cx(prev.ctrl_mask & ~(1ull << t), t, prev.tag);
                    out[(size_t)p] = x;
                    out.erase(out.begin() + i);
                    out.erase(out.begin() + q);
                    ++rewrites;
                }
            }
            return;
        }
    };
    for (Gate& g : *gates) {
        double s = 0;
        if (ph_is_cx(g)) {
            const uint32_t t = g.targets[0];
            const long p = ph_last_touching(out, (long)out.size(), t);
            double s_p = 0;
            if (p >= 0 && ph_is_h(out[(size_t)p], &s_p)) {           // H ; CX  ->  CZ ; H
                Gate h = out[(size_t)p];
                out.erase(out.begin() + p);
                out.push_back(ph_make_cz(g.ctrl_mask | (1ull << t), g.tag));
                out.push_back(std::move(h));
                ++rewrites;
                continue;
            }
            out.push_back(std::move(g));
            continue;
        }
        if (ph_is_h(g, &s)) {
            const uint32_t t = g.targets[0];
            out.push_back(std::move(g));
            settle_h((long)out.size() - 1, t);
            continue;
        }
        out.push_back(std::move(g));
    }
    gates->swap(out);
    return rewrites;
}

}  // namespace qb
