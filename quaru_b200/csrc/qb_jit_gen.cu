// qb_jit_gen.cu -- K2j source generator: one round program (qb_tile2_kernel.cuh) -> straight-line CUDA C++ for exactly
// that pass.  It is a partial evaluation of the interpreter body (qb_tile2_body.inc) with the program as a constant:
//   * the 16 amplitudes of a thread are 32 scalar variables; which variable holds logical register j is tracked HERE, so
//     X / CNOT / SWAP between register bits cost nothing (the next store simply reads the other variable);
//   * every coefficient is a literal (hex float: exact), zero terms are dropped, +-1 factors are additions; the only data
//     tables left are the per-thread phase tables indexed by thread / tile-index bits (read through the L1 path);
//   * controls on thread bits are predicates around straight-line code, controls on tile-index bits are uniform branches;
//   * tile base, thread -> tile-local mapping, swizzled shared-memory offsets and HBM offsets are shifts and masks with
//     literal operands.
// No op decode, no switch, no loop-carried register file: what the interpreter spends ~48 integer instructions per op on is
// gone, and ptxas schedules across op boundaries.  The same text compiles as host C++ (-DQB_HOST_EMU, tests/cpu_kernel/) so
// the generator is checked against the oracle without a GPU.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <utility>
#include <vector>

#include "qb_jit.h"

namespace qb {

namespace {

const char* kDevicePrelude = R"QB(
typedef unsigned int u32;
typedef unsigned long long u64;
#define QB_LD(X, Y, P) asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(X), "=d"(Y) : "l"(P))
#define QB_ST(P, X, Y) asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(P), "d"(X), "d"(Y) : "memory")
#define QB_PREFETCH(P) asm volatile("prefetch.global.L2 [%0];" ::"l"(P))
#define QB_SYNC() __syncthreads()
#define QB_TABLE_DECL __device__ __align__(16) const double QB_TAB[]
#define QB_TAB_AT(i) __ldg(reinterpret_cast<const double2*>(QB_TAB) + (i))
#define QB_SMEM(t) extern __shared__ __align__(16) double2 t[]
#define QB_TID threadIdx.x
#define QB_BID blockIdx.x
#define QB_GRID gridDim.x
struct QbRemote { double2* base[8]; u32 n_local; u32 pad; };
#define QB_KERNEL(name, minb) extern "C" __global__ void __launch_bounds__(256, minb) name(double2* __restrict__ st, u32 tile_or, u32 n_tiles)
#define QB_KERNEL_REMOTE(name, minb) extern "C" __global__ void __launch_bounds__(256, minb) name(const __grid_constant__ QbRemote ra, u32 tile_or, u32 n_tiles)
#define QB_KERNEL_RING(name) extern "C" __global__ void __launch_bounds__(512, 1) name(double2* __restrict__ st, u32 tile_or, u32 n_tiles)
#define QB_SMEM_RAW(t) extern __shared__ __align__(128) unsigned char t[]
static __device__ __forceinline__ u32 qb_saddr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
#define QB_GROUP_SYNC(G) asm volatile("bar.sync %0, 256;" ::"r"(1u + (G)) : "memory")
#define QB_BLOCK_SYNC() __syncthreads()
#define QB_CP_ASYNC16(D, S) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(qb_saddr(D)), "l"(S) : "memory")
#define QB_CP_ASYNC_COMMIT() asm volatile("cp.async.commit_group;" ::: "memory")
#define QB_CP_ASYNC_WAIT_ALL() asm volatile("cp.async.wait_group 0;" ::: "memory")
#define QB_SPIN() __nanosleep(20)
#define QB_FLAG_LD(P) (*(volatile u32*)(P))
#define QB_FLAG_ST(P, V) do { __threadfence_block(); *(volatile u32*)(P) = (V); } while (0)
#define QB_KERNEL_XCHG(name, minb) extern "C" __global__ void __launch_bounds__(256, minb) name(const double2* __restrict__ st, const __grid_constant__ QbRemote out, u32 rank, u32 tile_or, u32 n_tiles)
)QB";

struct Gen {
    std::string s;
    const ProgView& p;
    JitOpts o;
    int nm[16];
    bool tph_one = true;
    bool uses_table = false;
    uint32_t gpos[RT];   // global bit position of tile-local bit b

    Gen(const ProgView& pv, const JitOpts& jo) : p(pv), o(jo) {
        for (int j = 0; j < 16; ++j) nm[j] = j;
        for (uint32_t b = 0; b < RT; ++b) gpos[b] = b < p.lc ? b : p.hp[b - p.lc];
    }
    void f(const char* fmt, ...) __attribute__((format(printf, 2, 3))) {
        char buf[1024];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        s += buf;
    }
    static std::string lit(double v) {
        char buf[64];
        if (v == 0.0) return "0.0";
        if (v == 1.0) return "1.0";
        if (v == -1.0) return "(-1.0)";
        snprintf(buf, sizeof(buf), v < 0 ? "(%a)" : "%a", v);
        return buf;
    }
    std::string X(int j) const { return "x" + std::to_string(nm[j]); }
    std::string Y(int j) const { return "y" + std::to_string(nm[j]); }

    // sum of coef * var, dropping zero terms; "0.0" if nothing is left
    static std::string sum(const std::vector<std::pair<double, std::string>>& terms) {
        std::string e;
        for (const auto& t : terms) {
            if (t.first == 0.0) continue;
            if (t.first == 1.0) e += (e.empty() ? "" : " + ") + t.second;
            else if (t.first == -1.0) e += (e.empty() ? "-" : " - ") + t.second;
            else if (e.empty()) e += lit(t.first) + " * " + t.second;
            else if (t.first < 0) e += " - " + lit(-t.first) + " * " + t.second;
            else e += " + " + lit(t.first) + " * " + t.second;
        }
        return e.empty() ? "0.0" : e;
    }
    // real / imaginary part of sum_k c_k * (x_k + i y_k)
    std::string re_of(const std::vector<std::pair<c64, int>>& t) const {
        std::vector<std::pair<double, std::string>> v;
        for (const auto& e : t) { v.push_back({e.first.x, X(e.second)}); v.push_back({-e.first.y, Y(e.second)}); }
        return sum(v);
    }
    std::string im_of(const std::vector<std::pair<c64, int>>& t) const {
        std::vector<std::pair<double, std::string>> v;
        for (const auto& e : t) { v.push_back({e.first.x, Y(e.second)}); v.push_back({e.first.y, X(e.second)}); }
        return sum(v);
    }

    // ---- conditions -----------------------------------------------------------------------------------------------------
    // "" = unconditional; else a C expression, true where the op acts
    static std::string cond_of(uint32_t tmask, uint64_t omask) {
        char buf[160];
        std::string c;
        if (omask) { snprintf(buf, sizeof(buf), "(tix & 0x%llxull) == 0x%llxull", (unsigned long long)omask, (unsigned long long)omask); c = buf; }
        if (tmask) {
            snprintf(buf, sizeof(buf), "(tl & 0x%xu) == 0x%xu", tmask, tmask);
            c = c.empty() ? std::string(buf) : "(" + c + ") && (" + buf + ")";
        }
        return c;
    }
    void open(const std::string& c) { if (!c.empty()) f("    if (%s) {\n", c.c_str()); else s += "    {\n"; }
    void close() { s += "    }\n"; }

    // ---- register-file updates ----------------------------------------------------------------------------------------
    // logical registers `regs` <- M * regs (M row-major, dimension regs.size()), under `c`
    void dense(const std::vector<int>& regs, const c64* M, const std::string& c) {
        const size_t d = regs.size();
        // a permutation matrix (X, CNOT, SWAP ... with unit entries): rename when unconditional
        bool perm = true;
        std::vector<int> src(d, -1);
        for (size_t r = 0; r < d && perm; ++r) {
            int ones = 0;
            for (size_t k = 0; k < d; ++k) {
                const c64 e = M[r * d + k];
                if (e.x == 1.0 && e.y == 0.0) { ++ones; src[r] = (int)k; }
                else if (e.x != 0.0 || e.y != 0.0) perm = false;
            }
            if (ones != 1) perm = false;
        }
        if (perm) {
            bool ident = true;
            for (size_t r = 0; r < d; ++r) ident = ident && src[r] == (int)r;
            if (ident) return;
            if (c.empty()) {
                int old[16];
                for (size_t r = 0; r < d; ++r) old[r] = nm[regs[r]];
                for (size_t r = 0; r < d; ++r) nm[regs[r]] = old[src[r]];
                return;
            }
        }
        open(c);
        for (size_t r = 0; r < d; ++r) {
            std::vector<std::pair<c64, int>> t;
            for (size_t k = 0; k < d; ++k) t.push_back({M[r * d + k], regs[k]});
            f("      const double nx%zu = %s, ny%zu = %s;\n", r, re_of(t).c_str(), r, im_of(t).c_str());
        }
        for (size_t r = 0; r < d; ++r) f("      %s = nx%zu; %s = ny%zu;\n", X(regs[r]).c_str(), r, Y(regs[r]).c_str(), r);
        close();
    }
    void hadamard(int a) {   // (a0, a1) <- (a0 + a1, a0 - a1); the scalar is folded into the pass factor by the encoder
        s += "    {\n";
        for (int j0 = 0; j0 < 16; ++j0) {
            if ((j0 >> a) & 1) continue;
            const int j1 = j0 | (1 << a);
            f("      { const double tx = %s + %s, ty = %s + %s; %s = %s - %s; %s = %s - %s; %s = tx; %s = ty; }\n", X(j0).c_str(), X(j1).c_str(),
              Y(j0).c_str(), Y(j1).c_str(), X(j1).c_str(), X(j0).c_str(), X(j1).c_str(), Y(j1).c_str(), Y(j0).c_str(), Y(j1).c_str(),
              X(j0).c_str(), Y(j0).c_str());
        }
        close();
    }
    // register j <- (fr + i fi) * register j, fr / fi C expressions (literals or variables)
    void cmul_reg_expr(int j, const std::string& fr, const std::string& fi) {
        f("      { const double tx = %s * %s - %s * %s; %s = %s * %s + %s * %s; %s = tx; }\n", fr.c_str(), X(j).c_str(), fi.c_str(), Y(j).c_str(),
          Y(j).c_str(), fr.c_str(), Y(j).c_str(), fi.c_str(), X(j).c_str(), X(j).c_str());
    }
    void cmul_reg_lit(int j, c64 v) {
        if (v.x == 1.0 && v.y == 0.0) return;
        std::vector<std::pair<c64, int>> t{{v, j}};
        f("      { const double tx = %s; %s = %s; %s = tx; }\n", re_of(t).c_str(), Y(j).c_str(), im_of(t).c_str(), X(j).c_str());
    }
    // ---- the per-thread phase ------------------------------------------------------------------------------------------
    void tph_mul_lit(c64 v, const std::string& c) {
        if (v.x == 1.0 && v.y == 0.0) return;
        if (tph_one) {
            if (c.empty()) f("    px = %s; py = %s;\n", lit(v.x).c_str(), lit(v.y).c_str());
            else f("    { const bool c_ = %s; px = c_ ? %s : 1.0; py = c_ ? %s : 0.0; }\n", c.c_str(), lit(v.x).c_str(), lit(v.y).c_str());
            tph_one = false;
            return;
        }
        open(c);
        std::vector<std::pair<double, std::string>> re{{v.x, "px"}, {-v.y, "py"}}, im{{v.x, "py"}, {v.y, "px"}};
        f("      const double tx = %s; py = %s; px = tx;\n", sum(re).c_str(), sum(im).c_str());
        close();
    }
    void tph_mul_table(const std::string& index_expr, const std::string& c) {
        uses_table = true;
        if (tph_one) {
            if (c.empty()) f("    { const double2 f_ = QB_TAB_AT(%s); px = f_.x; py = f_.y; }\n", index_expr.c_str());
            else f("    { px = 1.0; py = 0.0; if (%s) { const double2 f_ = QB_TAB_AT(%s); px = f_.x; py = f_.y; } }\n", c.c_str(), index_expr.c_str());
            tph_one = false;
            return;
        }
        open(c);
        f("      const double2 f_ = QB_TAB_AT(%s); const double tx = f_.x * px - f_.y * py; py = f_.x * py + f_.y * px; px = tx;\n", index_expr.c_str());
        close();
    }
    // index bit from a selector byte (0x80 | tile-index bit, else tile-local thread bit in the low 6 bits)
    static std::string sel_bit(uint32_t sc, int shift) {
        char buf[96];
        if (shift) snprintf(buf, sizeof(buf), "(((%s >> %u) & 1u) << %d)", (sc & 0x80) ? "tix" : "tl", sc & 0x3f, shift);
        else snprintf(buf, sizeof(buf), "((%s >> %u) & 1u)", (sc & 0x80) ? "tix" : "tl", sc & 0x3f);
        return buf;
    }

    // ---- addressing ---------------------------------------------------------------------------------------------------
    // tl = tile-local index bits of this thread in round rd (register bits 0)
    void emit_tl(const RoundDesc& rd) {
        std::string e;
        for (int i = 0; i < 8;) {
            int len = 1;
            while (i + len < 8 && rd.tb[i + len] == rd.tb[i] + len) ++len;
            char buf[96];
            if (rd.tb[i] == i) snprintf(buf, sizeof(buf), "(tid & 0x%xu)", ((1u << len) - 1u) << i);
            else if (rd.tb[i] > i) snprintf(buf, sizeof(buf), "((tid & 0x%xu) << %d)", ((1u << len) - 1u) << i, rd.tb[i] - i);
            else snprintf(buf, sizeof(buf), "((tid & 0x%xu) >> %d)", ((1u << len) - 1u) << i, i - rd.tb[i]);
            e += (e.empty() ? "" : " | ") + std::string(buf);
            i += len;
        }
        f("    tl = %s;\n", e.c_str());
    }
    // offset (in amplitudes) of the tile-local thread part from the tile base, as an expression of tl
    std::string goff_expr() const {
        char buf[128];
        std::string e;
        snprintf(buf, sizeof(buf), "(u64)(tl & 0x%xu)", (1u << p.lc) - 1u);
        e = buf;
        for (uint32_t i = 0; i < p.h;) {
            uint32_t len = 1;
            while (i + len < p.h && p.hp[i + len] == p.hp[i] + len) ++len;
            snprintf(buf, sizeof(buf), " | ((u64)(tl & 0x%xu) << %u)", ((1u << len) - 1u) << (p.lc + i), p.hp[i] - (p.lc + i));
            e += buf;
            i += len;
        }
        return e;
    }
    // tile base: the bits of the tile index `tix` deposited into the non-tile positions
    std::string base_expr(const char* tix) const {
        uint64_t tile_mask = 0;
        for (uint32_t b = 0; b < RT; ++b) tile_mask |= 1ull << gpos[b];
        std::vector<uint32_t> pos;   // position of tile-index bit i
        for (uint32_t b = 0; b < 64 && pos.size() < 40; ++b)
            if (!((tile_mask >> b) & 1)) pos.push_back(b);
        std::string e;
        char buf[160];
        const uint32_t top = 31;   // tile indices have fewer than 32 bits
        for (uint32_t i = 0; i <= top;) {
            uint32_t len = 1;
            while (i + len <= top && pos[i + len] == pos[i] + len) ++len;
            if (i + len > top) snprintf(buf, sizeof(buf), "((u64)(%s >> %u) << %u)", tix, i, pos[i]);
            else snprintf(buf, sizeof(buf), "((u64)((%s >> %u) & 0x%xu) << %u)", tix, i, (1u << len) - 1u, pos[i]);
            e += (e.empty() ? "" : " | ") + std::string(buf);
            i += len;
        }
        return e;
    }
    std::string amp_ptr(const char* off) const {
        if (!o.remote) return std::string("(st + (") + off + "))";
        return std::string("(ra.base[(") + off + ") >> ra.n_local] + ((" + off + ") & lmask))";
    }

    // ---- one op ---------------------------------------------------------------------------------------------------------
    void emit_op(const ROp& op) {
        const c64* M = p.mats + op.mat;
        const int code = op.code;
        const std::string c = cond_of(op.tmask, op.omask);
        auto pairs_of = [](int a, uint32_t rmask, auto&& fn) {
            for (int j0 = 0; j0 < 16; ++j0) {
                if ((j0 >> a) & 1) continue;
                if ((j0 & rmask) != rmask) continue;
                fn(j0, j0 | (1 << a));
            }
        };
        if (code < RC_D2) {
            const int a = code & 3, kind = code & ~3;
            const uint32_t rmask = (kind >= RC_D1C_P) ? op.rmask : 0u;
            if (kind == RC_D1C || kind == RC_D1R || kind == RC_D1C_P || kind == RC_D1R_P) {
                c64 m[4] = {M[0], M[1], M[2], M[3]};
                if (kind == RC_D1R || kind == RC_D1R_P)
                    for (c64& e : m) e.y = 0.0;   // the interpreter reads only the real parts
                pairs_of(a, rmask, [&](int j0, int j1) { dense({j0, j1}, m, c); });
            } else {   // RC_DG1R / RC_DG1R_P: diag(d0, d1) on register bit a
                open(c);
                for (int j = 0; j < 16; ++j) {
                    if ((j & rmask) != rmask) continue;
                    cmul_reg_lit(j, ((j >> a) & 1) ? M[1] : M[0]);
                }
                close();
            }
            return;
        }
        if (code >= RC_H && code < RC_H + 4) { hadamard(code - RC_H); return; }
        if (code >= RC_PH_FLUSH && code < RC_PH_FLUSH + 4) {
            const int b = code - RC_PH_FLUSH;
            if (!tph_one) {
                s += "    {\n";
                for (int j = 0; j < 16; ++j)
                    if ((j >> b) & 1) cmul_reg_expr(j, "px", "py");
                close();
            }
            tph_one = true;
            return;
        }
        switch (code) {
            case RC_D2: {
                const std::string c2 = cond_of(op.tmask, op.omask);
                for (int j = 0; j < 16; ++j) {
                    if (((j >> op.a) & 1) || ((j >> op.b) & 1)) continue;
                    if ((j & op.rmask) != op.rmask) continue;
                    dense({j, j | (1 << op.a), j | (1 << op.b), j | (1 << op.a) | (1 << op.b)}, M, c2);
                }
            } break;
            case RC_PH_T:
            case RC_PH_ACC: tph_mul_lit(M[0], c); break;
            case RC_PH_NEG:
                open(c);
                for (int j = 0; j < 16; ++j)
                    if ((j & op.rmask) == op.rmask) f("      %s = -%s; %s = -%s;\n", X(j).c_str(), X(j).c_str(), Y(j).c_str(), Y(j).c_str());
                close();
                break;
            case RC_PH_R:
                open(c);
                for (int j = 0; j < 16; ++j)
                    if ((j & op.rmask) == op.rmask) cmul_reg_lit(j, M[0]);
                close();
                break;
            case RC_TAB:
            case RC_TAB_ACC: {
                const uint32_t sel[4] = {op.src[0], op.src[1], op.src[2], op.b};
                const std::string oc = cond_of(0, op.omask);
                if (op.k == 0) { tph_mul_lit(M[0], oc); break; }
                std::string idx = std::to_string(op.mat) + "u";
                for (int i = 0; i < (int)op.k && i < 4; ++i) idx += " + " + sel_bit(sel[i], i);
                tph_mul_table(idx, oc);
            } break;
            case RC_LAYER: {
                const uint32_t present = op.k & 15u, hmask = op.k >> 4;
                const uint32_t tm[4] = {op.tmask, op.rmask, (uint32_t)op.src[0] | ((uint32_t)op.src[1] << 8), (uint32_t)op.src[2] | ((uint32_t)op.b << 8)};
                for (int a = 0; a < 4; ++a) {
                    if (!((present >> a) & 1u)) continue;
                    if ((hmask >> a) & 1u) { hadamard(a); continue; }
                    const c64 c0 = M[2 * a], c1 = M[2 * a + 1];
                    const c64 m[4] = {make_double2(c0.x, 0.0), make_double2(c0.y, 0.0), make_double2(c1.x, 0.0), make_double2(c1.y, 0.0)};
                    const std::string lc_ = cond_of(tm[a], 0);
                    pairs_of(a, 0, [&](int j0, int j1) { dense({j0, j1}, m, lc_); });
                }
            } break;
            case RC_DG1T: {
                const uint32_t s0 = op.src[0];
                tph_mul_table(std::to_string(op.mat) + "u + " + sel_bit(s0, 0), c);
            } break;
            default: {   // RC_DG
                uint32_t regmask[3] = {0, 0, 0};
                std::string dyn;
                bool any_reg = false;
                for (int i = 0; i < (int)op.k && i < 3; ++i) {
                    const uint32_t sc = op.src[i];
                    if (sc & 0xc0) dyn += (dyn.empty() ? "" : " | ") + sel_bit(sc & 0x80 ? sc : (sc & 0x3f), i);
                    else { regmask[i] = 1u << (sc & 3u); any_reg = true; }
                }
                if (!any_reg && op.rmask == 0) {
                    if (dyn.empty()) tph_mul_lit(M[0], c);
                    else tph_mul_table(std::to_string(op.mat) + "u + (" + dyn + ")", c);
                    break;
                }
                open(c);
                if (!dyn.empty()) { uses_table = true; f("      const u32 dx = %s;\n", dyn.c_str()); }
                for (int j = 0; j < 16; ++j) {
                    if ((j & op.rmask) != op.rmask) continue;
                    const uint32_t xj = ((j & regmask[0]) ? 1u : 0u) | ((j & regmask[1]) ? 2u : 0u) | ((j & regmask[2]) ? 4u : 0u);
                    if (dyn.empty()) cmul_reg_lit(j, M[xj]);
                    else {
                        f("      { const double2 f_ = QB_TAB_AT(%uu + (dx | %uu));\n  ", op.mat, xj);
                        cmul_reg_expr(j, "f_.x", "f_.y");
                        s += "      }\n";
                    }
                }
                close();
            } break;
        }
    }

    // tile-local offset of register j in a round whose register bits are R[0..3]
    static uint32_t reg_local(const RoundDesc& rd, int j) {
        uint32_t l = 0;
        for (int i = 0; i < 4; ++i)
            if ((j >> i) & 1) l |= 1u << rd.R[i];
        return l;
    }

    // ---- the prefetching ring (pipelined = 1) ------------------------------------------------------------------------------
    // One persistent CTA of 512 threads per SM, two groups of 256 threads on alternate tiles, three 64 KB buffers: tile k lives
    // in buffer k % 3.  Every thread fetches ITS OWN 16 amplitudes of its group's next tile (k + 2) with 16-byte cp.async
    // (LDGSTS: HBM -> shared memory without passing through registers) while it computes tile k, and reads them back with
    // cp.async.wait_group + 16 LDS at the top of tile k + 2: a per-thread prefetch queue in shared memory, so no cross-thread
    // completion tracking is needed -- no mbarrier, no bulk copies.  Buffer (k + 2) % 3 is the buffer of the OTHER group's tile
    // k - 1; that group publishes "my shared-memory use of tile t is over" (after the last transpose read) in done[group], and
    // the prefetch is issued at the first round boundary at which done[other] >= k holds (at the latest before the stores,
    // spinning).  Results go from registers straight to HBM as in the default shape.
    void emit_prefetch_lambda() {
        s += "    auto prefetch = [&](u32 kk) {\n";
        s += "      const u32 ntix = (QB_BID + kk * QB_GRID) | tile_or;\n";
        f("      const u64 ngt = (%s) + goff0;\n", base_expr("ntix").c_str());
        s += "      double2* const nt = bufs + (kk % 3u) * 4096u + tl0;\n";
        for (int j = 0; j < 16; ++j)
            f("      QB_CP_ASYNC16(nt + 0x%xu, st + (ngt + 0x%llxull));\n", reg_local(p.rounds[0], j), (unsigned long long)p.gro_first[j]);
        s += "      QB_CP_ASYNC_COMMIT();\n    };\n";
    }
    std::string run_ring() {
        s += "QB_KERNEL_RING(qb_jit_pass)\n{\n";
        s += "    QB_SMEM_RAW(smem_raw);\n";
        s += "    double2* const bufs = reinterpret_cast<double2*>(smem_raw);\n";
        s += "    u32* const done = reinterpret_cast<u32*>(smem_raw + 3u * 65536u);\n";
        s += "    const u32 tid_all = QB_TID, grp = tid_all >> 8, tid = tid_all & 255u;\n";
        s += "    if (tid_all < 2u) QB_FLAG_ST(&done[tid_all], 0u);\n";
        s += "    QB_BLOCK_SYNC();\n";
        s += "    const u32 n_mine = QB_BID < n_tiles ? (n_tiles - QB_BID + QB_GRID - 1u) / QB_GRID : 0u;\n";
        s += "    u32 tl;\n";
        emit_tl(p.rounds[0]);
        s += "    const u32 tl0 = tl;\n";
        f("    const u64 goff0 = %s;\n", goff_expr().c_str());
        emit_prefetch_lambda();
        s += "    if (grp < n_mine) prefetch(grp);\n";
        s += "    for (u32 k = grp; k < n_mine; k += 2u) {\n";
        s += "    double2* const tile = bufs + (k % 3u) * 4096u;\n";
        s += "    const u32 tix = (QB_BID + k * QB_GRID) | tile_or;\n";
        f("    const u64 base = %s;\n", base_expr("tix").c_str());
        s += "    bool pf = k + 2u >= n_mine;   // true: tile k + 2 has been requested (or does not exist)\n";
        s += "    double";
        for (int j = 0; j < 16; ++j) f(" x%d, y%d%s", j, j, j == 15 ? ";\n" : ",");
        s += "    double px = 1.0, py = 0.0;\n    tl = tl0;\n";
        for (int j = 0; j < 16; ++j) nm[j] = j;
        s += "    QB_CP_ASYNC_WAIT_ALL();\n";   // the only copies in flight are this thread's amplitudes of tile k
        for (int j = 0; j < 16; ++j)
            f("    { const double2 t_ = tile[tl0 | 0x%xu]; x%d = t_.x; y%d = t_.y; }\n", reg_local(p.rounds[0], j), j, j);
        const char* publish = "    if (tid == 0u) QB_FLAG_ST(&done[grp], k + 1u);\n";
        const char* try_pf = "    if (!pf && QB_FLAG_LD(&done[grp ^ 1u]) >= k) { prefetch(k + 2u); pf = true; }\n";
        if (p.n_rounds > 1) s += "    QB_GROUP_SYNC(grp);\n";   // everybody has its amplitudes: the buffer is the transposes' scratch from here on
        else { s += "    QB_GROUP_SYNC(grp);\n"; s += publish; }
        s += try_pf;
        for (uint32_t r = 0; r < p.n_rounds; ++r) {
            const RoundDesc& rd = p.rounds[r];
            f("    // ---- round %u: register bits %u %u %u %u ----\n", r, rd.R[0], rd.R[1], rd.R[2], rd.R[3]);
            if (r > 0) {
                const RoundDesc& prev = p.rounds[r - 1];
                s += "    {\n      const u32 s_t = qb_swz(tl);\n";
                for (int j = 0; j < 16; ++j) f("      tile[s_t ^ 0x%xu] = make_double2(%s, %s);\n", prev.sro[j], X(j).c_str(), Y(j).c_str());
                s += "    }\n    QB_GROUP_SYNC(grp);\n";
                emit_tl(rd);
                s += "    {\n      const u32 s_n = qb_swz(tl);\n";
                for (int j = 0; j < 16; ++j) {
                    nm[j] = j;
                    f("      { const double2 t_ = tile[s_n ^ 0x%xu]; x%d = t_.x; y%d = t_.y; }\n", rd.sro[j], j, j);
                }
                s += "    }\n    QB_GROUP_SYNC(grp);\n";   // the buffer is written again by the next transpose, or handed to the other group
                if (r + 1 == p.n_rounds) s += publish;
                s += try_pf;
            }
            s += "    px = 1.0; py = 0.0;\n";
            tph_one = true;
            for (uint32_t oi = rd.op_begin; oi < rd.op_end; ++oi) emit_op(p.ops[oi]);
            if (rd.has_tph && !tph_one) {
                s += "    {\n";
                for (int j = 0; j < 16; ++j) cmul_reg_expr(j, "px", "py");
                close();
            }
        }
        // the other group may still be transposing in the buffer tile k + 2 goes to: wait for it here at the latest
        s += "    if (!pf) { while (QB_FLAG_LD(&done[grp ^ 1u]) < k) QB_SPIN(); prefetch(k + 2u); }\n";
        f("    {\n      const u64 gt = base + (%s);\n", goff_expr().c_str());
        for (int j = 0; j < 16; ++j)
            f("      QB_ST(st + (gt + 0x%llxull), %s, %s);\n", (unsigned long long)p.gro_last[j], X(j).c_str(), Y(j).c_str());
        s += "    }\n    }\n}\n";
        return finish();
    }

    std::string run() {
        if (o.pipelined && !o.remote && o.xchg_bits == 0) return run_ring();
        // ---- kernel body ----
        f("%s(qb_jit_pass, %d)\n{\n", o.xchg_bits > 0 ? "QB_KERNEL_XCHG" : o.remote ? "QB_KERNEL_REMOTE" : "QB_KERNEL", o.min_blocks);
        s += "    QB_SMEM(tile);\n    const u32 tid = QB_TID;\n";
        if (o.remote) s += "    const u64 lmask = (1ull << ra.n_local) - 1ull;\n";
        if (o.persistent) s += "    for (u32 t0 = QB_BID; t0 < n_tiles; t0 += QB_GRID) {\n    const u32 tix = t0 | tile_or;\n";
        else s += "    const u32 tix = QB_BID | tile_or;\n    (void)n_tiles;\n";
        f("    const u64 base = %s;\n", base_expr("tix").c_str());
        s += "    double";
        for (int j = 0; j < 16; ++j) f(" x%d, y%d%s", j, j, j == 15 ? ";\n" : ",");
        s += "    double px = 1.0, py = 0.0;\n    u32 tl;\n";
        for (int j = 0; j < 16; ++j) nm[j] = j;
        // round 0: HBM -> registers
        emit_tl(p.rounds[0]);
        f("    {\n      const u64 gt = base + (%s);\n", goff_expr().c_str());
        for (int j = 0; j < 16; ++j) {
            char off[64];
            snprintf(off, sizeof(off), "gt + 0x%llxull", (unsigned long long)p.gro_first[j]);
            f("      QB_LD(x%d, y%d, %s);\n", j, j, amp_ptr(off).c_str());
        }
        s += "    }\n";
        if (o.persistent && o.prefetch && p.rounds[0].tb[0] == 0 && p.rounds[0].tb[1] == 1 && p.rounds[0].tb[2] == 2 && !o.remote) {
            // one thread per 128-byte line of the next tile (lanes 0..2 = tile bits 0..2 = 8 amplitudes)
            s += "    if (t0 + QB_GRID < n_tiles && (tid & 7u) == 0u) {\n      const u32 ntix = (t0 + QB_GRID) | tile_or;\n";
            f("      const u64 ngt = (%s) + (%s);\n", base_expr("ntix").c_str(), goff_expr().c_str());
            for (int j = 0; j < 16; ++j) f("      QB_PREFETCH(st + (ngt + 0x%llxull));\n", (unsigned long long)p.gro_first[j]);
            s += "    }\n";
        }
        for (uint32_t r = 0; r < p.n_rounds; ++r) {
            const RoundDesc& rd = p.rounds[r];
            f("    // ---- round %u: register bits %u %u %u %u ----\n", r, rd.R[0], rd.R[1], rd.R[2], rd.R[3]);
            if (r > 0) {
                const RoundDesc& prev = p.rounds[r - 1];
                s += "    {\n      const u32 s_t = qb_swz(tl);\n";
                for (int j = 0; j < 16; ++j) f("      tile[s_t ^ 0x%xu] = make_double2(%s, %s);\n", prev.sro[j], X(j).c_str(), Y(j).c_str());
                s += "    }\n    QB_SYNC();\n";
                emit_tl(rd);
                s += "    {\n      const u32 s_n = qb_swz(tl);\n";
                for (int j = 0; j < 16; ++j) {
                    nm[j] = j;
                    f("      { const double2 t_ = tile[s_n ^ 0x%xu]; x%d = t_.x; y%d = t_.y; }\n", rd.sro[j], j, j);
                }
                s += "    }\n";
                // the buffer is written again by the next transpose of this tile, or (persistent CTAs) of the next tile
                bool more = o.persistent != 0;
                for (uint32_t r2 = r + 1; r2 < p.n_rounds; ++r2) more = true;
                if (more) s += "    QB_SYNC();\n";
            }
            s += "    px = 1.0; py = 0.0;\n";
            tph_one = true;
            for (uint32_t oi = rd.op_begin; oi < rd.op_end; ++oi) emit_op(p.ops[oi]);
            if (rd.has_tph && !tph_one) {
                s += "    {\n";
                for (int j = 0; j < 16; ++j) cmul_reg_expr(j, "px", "py");
                close();
            }
        }
        // last round: registers -> HBM
        f("    {\n      const u64 gt = base + (%s);\n", goff_expr().c_str());
        if (o.xchg_bits > 0) {
            // The exchanging epilogue (sharded registers): the amplitude at local index l of this rank is stored where it lives
            // AFTER the global<->local exchange of local bits V[i] with rank bits J[i] -- rank bit J[i] of the destination is bit
            // V[i] of l, bit V[i] of the destination index is this rank's bit J[i] -- into the destination rank's OTHER state
            // buffer (out.base[]), which nobody reads during this pass.  The transfer rides on the pass's stores (posted writes
            // over NVLink), tile by tile, while other CTAs compute.
            uint32_t jmask = 0;
            for (int i = 0; i < o.xchg_bits; ++i) jmask |= 1u << o.xchg_J[i];
            f("      const u32 rk0 = rank & ~0x%xu;\n", jmask);
            for (int i = 0; i < o.xchg_bits; ++i) f("      const u64 rb%d = (u64)((rank >> %u) & 1u) << %u;\n", i, o.xchg_J[i], o.xchg_V[i]);
            for (int j = 0; j < 16; ++j) {
                f("      { const u64 l_ = gt + 0x%llxull; u32 dr_ = rk0; u64 dl_ = l_;\n", (unsigned long long)p.gro_last[j]);
                for (int i = 0; i < o.xchg_bits; ++i)
                    f("        dr_ |= ((u32)(l_ >> %u) & 1u) << %u; dl_ = (dl_ & ~(1ull << %u)) | rb%d;\n", o.xchg_V[i], o.xchg_J[i], o.xchg_V[i], i);
                f("        QB_ST(out.base[dr_] + dl_, %s, %s); }\n", X(j).c_str(), Y(j).c_str());
            }
        } else {
            for (int j = 0; j < 16; ++j) {
                char off[64];
                snprintf(off, sizeof(off), "gt + 0x%llxull", (unsigned long long)p.gro_last[j]);
                f("      QB_ST(%s, %s, %s);\n", amp_ptr(off).c_str(), X(j).c_str(), Y(j).c_str());
            }
        }
        s += "    }\n";
        if (o.persistent) s += "    }\n";
        s += "}\n";
        return finish();
    }

    // ---- prelude + tables in front of the kernel text ----
    std::string finish() {
        std::string head = "// generated by quaru_b200 (qb_jit_gen.cu): one fused tile pass, specialised\n#ifdef QB_HOST_EMU\n#include \"qb_jit_emu.h\"\n#else\n";
        head += kDevicePrelude;
        head += "#endif\nstatic __device__ __forceinline__ u32 qb_swz(u32 l) { return l ^ (((l >> 3) ^ (l >> 6) ^ (l >> 9)) & 7u); }\n";
        if (uses_table) {
            head += "QB_TABLE_DECL = {\n";
            char buf[96];
            for (uint32_t i = 0; i < p.n_mats; ++i) {
                snprintf(buf, sizeof(buf), "  %a, %a,\n", p.mats[i].x, p.mats[i].y);
                head += buf;
            }
            head += "};\n";
        }
        return head + s;
    }
};

}  // namespace

std::string jit_generate(const ProgView& p, const JitOpts& o) {
    Gen g(p, o);
    return g.run();
}

}  // namespace qb
