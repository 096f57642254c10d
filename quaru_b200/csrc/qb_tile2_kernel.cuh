// qb_tile2_kernel.cuh -- device side of K2r, the register-blocked fused tile pass: the program format shared with
// the host-side encoder (qb_tile2.cu) and the kernel itself.  See qb_tile2.cu for the design notes.
#pragma once
#include <stdint.h>

#include "qb_device.cuh"

namespace qb {

typedef double2 c64;  // as in qb_common.h

// flat op codes: (kind, register bit) combinations are separate switch cases
enum RCode : uint8_t {
    RC_D1C = 0, RC_D1R = 4, RC_DG1R = 8,          // + register bit, no control among the register bits
    RC_D1C_P = 16, RC_D1R_P = 20, RC_DG1R_P = 24,  // + register bit, predicated on rmask
    RC_D2 = 32, RC_PH_T = 33, RC_PH_NEG = 34, RC_PH_R = 35, RC_DG = 36, RC_DG1T = 37,
    RC_H = 40,                                     // + register bit: scalar multiple of the Hadamard matrix, uncontrolled
    RC_PH_ACC = 44,                                // phase whose only register-bit control is one bit: accumulate per thread ...
    RC_PH_FLUSH = 48,                              // + register bit: ... and apply the product to the 8 registers with that bit set
    RC_TAB = 52, RC_TAB_ACC = 53,                  // merged thread-uniform phases: table over <= 4 thread / outer bits -> tph
    RC_LAYER = 54                                  // up to four real 2x2 gates, one per register bit, decoded once
};

struct ROp {
    uint8_t code;      // RCode
    uint8_t a, b;      // register-bit indices of the dense targets (matrix bit 0, matrix bit 1)
    uint8_t flags;
    uint16_t rmask;    // controls that are register bits (mask over the 4 register bits)
    uint16_t tmask;    // controls that are thread bits (mask over the 12 tile-local bits)
    uint32_t mat;      // offset into mats
    uint8_t src[3];    // RK_DG: source of index bit i: 0x00|reg bit, 0x40|tile-local bit (thread), 0x80|global bit (outer)
    uint8_t k;         // RK_DG: number of index bits (1..3)
    uint64_t omask;    // controls outside the tile, in TILE-INDEX space (bit i = i-th non-tile bit = bit i of blockIdx.x)
};
static_assert(sizeof(ROp) == 24, "ROp layout");

struct RoundDesc {
    uint8_t R[4];      // ascending tile-local bits held in registers
    uint8_t tb[8];     // thread-id bit i <-> tile-local bit tb[i]
    uint16_t op_begin, op_end;
    uint8_t has_tph, pad[3];
    uint16_t sro[16];  // swizzled shared-memory offset of register j (host-computed: uniform, ~270 instructions per
                       // thread and transpose when derived on the device)
};

template <int NOPS, int NMAT, int NROUNDS>
struct RProgram {
    uint32_t lc, h, n_rounds, n_mats;   // n_mats: coefficients in use
    uint8_t hp[8];
    uint64_t gro_first[16], gro_last[16];  // offset from the tile base of register j in the first / last round
    RoundDesc rounds[NROUNDS];
    ROp ops[NOPS];
    c64 mats[NMAT];
};
typedef RProgram<24, 160, 8> RProgSmall;      // ~3.4 KB
typedef RProgram<192, 1200, 48> RProgLarge;   // ~25 KB

constexpr int RT = 12;           // tile bits
constexpr int RTHREADS = 256;

__device__ __forceinline__ uint32_t swz(uint32_t l) { return l ^ (((l >> 3) ^ (l >> 6) ^ (l >> 9)) & 7u); }

// ---- in-place primitives ----------------------------------------------------------------------------
// ncu showed that compiler-scheduled updates of the loop-carried register file end every switch case with
// up to 64 register MOVs (results computed out of place, then copied back): 45 % of all issued instructions.
// These PTX blocks fix the update ORDER so that every result is written over a value that is already dead,
// which lets ptxas keep each amplitude in one physical register across the whole op loop.
#ifdef QB_HOST_EMU   // the same operations in the same order, for the CPU harness (tests/cpu_kernel)
__device__ __forceinline__ void ip_cmul(c64& a, const c64 f) {
    const double t = -(f.y * a.y), u = f.y * a.x;
    a.x = fma(f.x, a.x, t);
    a.y = fma(f.x, a.y, u);
}
__device__ __forceinline__ void ip_pair_real(c64& a0, c64& a1, double p, double q, double r, double s) {
    const double t = r * a0.x, u = r * a0.y;
    a0.x = fma(q, a1.x, p * a0.x);
    a0.y = fma(q, a1.y, p * a0.y);
    a1.x = fma(s, a1.x, t);
    a1.y = fma(s, a1.y, u);
}
__device__ __forceinline__ void ip_pair_cplx(c64& a0, c64& a1, const c64 m00, const c64 m01, const c64 m10, const c64 m11) {
    double p = m10.x * a0.x, q = m10.x * a0.y, t = -(m00.y * a0.y), u = m00.y * a0.x;
    q = fma(m10.y, a0.x, q);
    p = fma(-m10.y, a0.y, p);
    a0.x = fma(m00.x, a0.x, t);
    a0.y = fma(m00.x, a0.y, u);
    a0.x = fma(m01.x, a1.x, a0.x);
    a0.y = fma(m01.x, a1.y, a0.y);
    a0.x = fma(-m01.y, a1.y, a0.x);
    a0.y = fma(m01.y, a1.x, a0.y);
    t = m11.y * a1.y;
    u = m11.y * a1.x;
    a1.x = fma(m11.x, a1.x, p) - t;
    a1.y = fma(m11.x, a1.y, q) + u;
}
__device__ __forceinline__ void ip_pair_h(c64& a0, c64& a1) {
    a0.x = a0.x + a1.x;
    a0.y = a0.y + a1.y;
    a1.x = fma(a1.x, -2.0, a0.x);
    a1.y = fma(a1.y, -2.0, a0.y);
}
#else
__device__ __forceinline__ void ip_cmul(c64& a, const c64 f) {   // a <- f * a
    asm volatile("{\n\t.reg .f64 t, u;\n\t"
                 "mul.f64 t, %3, %1;\n\t"            // t = f.y * a.y
                 "mul.f64 u, %3, %0;\n\t"            // u = f.y * a.x
                 "neg.f64 t, t;\n\t"
                 "fma.rn.f64 %0, %2, %0, t;\n\t"     // a.x = f.x * a.x - f.y * a.y
                 "fma.rn.f64 %1, %2, %1, u;\n\t"     // a.y = f.x * a.y + f.y * a.x
                 "}" : "+d"(a.x), "+d"(a.y) : "d"(f.x), "d"(f.y));
}
__device__ __forceinline__ void ip_pair_real(c64& a0, c64& a1, double p, double q, double r, double s) {
    // (a0, a1) <- [[p, q], [r, s]] (a0, a1), real matrix
    asm volatile("{\n\t.reg .f64 t, u;\n\t"
                 "mul.f64 t, %6, %0;\n\t"            // t = r * a0.x
                 "mul.f64 u, %6, %1;\n\t"            // u = r * a0.y
                 "mul.f64 %0, %4, %0;\n\t"           // a0.x = p * a0.x
                 "mul.f64 %1, %4, %1;\n\t"
                 "fma.rn.f64 %0, %5, %2, %0;\n\t"    // a0.x += q * a1.x
                 "fma.rn.f64 %1, %5, %3, %1;\n\t"
                 "fma.rn.f64 %2, %7, %2, t;\n\t"     // a1.x = s * a1.x + r * a0.x(old)
                 "fma.rn.f64 %3, %7, %3, u;\n\t"
                 "}" : "+d"(a0.x), "+d"(a0.y), "+d"(a1.x), "+d"(a1.y) : "d"(p), "d"(q), "d"(r), "d"(s));
}
__device__ __forceinline__ void ip_pair_cplx(c64& a0, c64& a1, const c64 m00, const c64 m01, const c64 m10, const c64 m11) {
    asm volatile("{\n\t.reg .f64 p, q, t, u;\n\t"
                 "mul.f64 p, %8, %0;\n\t"            // p = m10x*a0x
                 "mul.f64 q, %8, %1;\n\t"            // q = m10x*a0y
                 "mul.f64 t, %5, %1;\n\t"            // t = m00y*a0y
                 "mul.f64 u, %5, %0;\n\t"            // u = m00y*a0x
                 "fma.rn.f64 q, %9, %0, q;\n\t"      // q += m10y*a0x
                 "neg.f64 t, t;\n\t"
                 "fma.rn.f64 p, %13, %1, p;\n\t"     // p -= m10y*a0y
                 "fma.rn.f64 %0, %4, %0, t;\n\t"     // a0x = m00x*a0x - m00y*a0y
                 "fma.rn.f64 %1, %4, %1, u;\n\t"     // a0y = m00x*a0y + m00y*a0x
                 "fma.rn.f64 %0, %6, %2, %0;\n\t"    // a0x += m01x*a1x
                 "fma.rn.f64 %1, %6, %3, %1;\n\t"    // a0y += m01x*a1y
                 "fma.rn.f64 %0, %12, %3, %0;\n\t"   // a0x -= m01y*a1y
                 "fma.rn.f64 %1, %7, %2, %1;\n\t"    // a0y += m01y*a1x
                 "mul.f64 t, %11, %3;\n\t"           // t = m11y*a1y
                 "mul.f64 u, %11, %2;\n\t"           // u = m11y*a1x
                 "fma.rn.f64 %2, %10, %2, p;\n\t"    // a1x = m11x*a1x + p
                 "fma.rn.f64 %3, %10, %3, q;\n\t"    // a1y = m11x*a1y + q
                 "sub.f64 %2, %2, t;\n\t"
                 "add.f64 %3, %3, u;\n\t"
                 "}" : "+d"(a0.x), "+d"(a0.y), "+d"(a1.x), "+d"(a1.y)
                     : "d"(m00.x), "d"(m00.y), "d"(m01.x), "d"(m01.y), "d"(m10.x), "d"(m10.y), "d"(m11.x), "d"(m11.y),
                       "d"(-m01.y), "d"(-m10.y));
}

__device__ __forceinline__ void ip_pair_h(c64& a0, c64& a1) {
    // (a0, a1) <- (a0 + a1, a0 - a1): two FP64 instructions per component; the common scalar (1/sqrt 2) of the
    // Hadamard matrix is deferred into the per-thread phase of the round
    asm volatile("{\n\t"
                 "add.f64 %0, %0, %2;\n\t"           // a0.x = a0.x + a1.x
                 "add.f64 %1, %1, %3;\n\t"
                 "fma.rn.f64 %2, %2, 0dC000000000000000, %0;\n\t"   // a1.x = (a0.x + a1.x) - 2 a1.x = a0.x - a1.x up to the rounding of the sum
                 "fma.rn.f64 %3, %3, 0dC000000000000000, %1;\n\t"
                 "}" : "+d"(a0.x), "+d"(a0.y), "+d"(a1.x), "+d"(a1.y));
}
#endif

template <int A>
__device__ __forceinline__ void r_h(c64 (&v)[16]) {
#pragma unroll
    for (int j0 = 0; j0 < 16; ++j0) {
        if ((j0 >> A) & 1) continue;
        ip_pair_h(v[j0], v[j0 | (1 << A)]);
    }
}

template <int A>
__device__ __forceinline__ void r_d1r_vals(c64 (&v)[16], double p, double q, double r, double s) {
#pragma unroll
    for (int j0 = 0; j0 < 16; ++j0) {
        if ((j0 >> A) & 1) continue;
        ip_pair_real(v[j0], v[j0 | (1 << A)], p, q, r, s);
    }
}

// one slot of a layer op: a real 2x2 gate on register bit A with its own thread-bit control mask, or a Hadamard
template <int A>
__device__ __forceinline__ void r_layer_slot(c64 (&v)[16], uint32_t present, uint32_t hmask, const c64* __restrict__ C, uint32_t tm,
                                             uint32_t tl) {
    if (!((present >> A) & 1u)) return;                    // uniform
    if ((hmask >> A) & 1u) { r_h<A>(v); return; }          // uniform
    const bool act = (tl & tm) == tm;                      // per thread, never branched on
    const c64 c0 = C[2 * A], c1 = C[2 * A + 1];
    r_d1r_vals<A>(v, act ? c0.x : 1.0, act ? c0.y : 0.0, act ? c1.x : 0.0, act ? c1.y : 1.0);
}

// phase on the registers whose index contains the (compile-time) mask RM
template <int RM, bool NEG>
__device__ __forceinline__ void r_ph(c64 (&v)[16], const c64 d, const double sgn) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if ((j & RM) != RM) continue;
        if (NEG) { v[j].x *= sgn; v[j].y *= sgn; } else ip_cmul(v[j], d);
    }
}
template <bool NEG>
__device__ __forceinline__ void r_ph_dispatch(c64 (&v)[16], uint32_t rmask, const c64 d, const double sgn) {
    switch (rmask) {
        case 1: r_ph<1, NEG>(v, d, sgn); break;   case 2: r_ph<2, NEG>(v, d, sgn); break;   case 3: r_ph<3, NEG>(v, d, sgn); break;
        case 4: r_ph<4, NEG>(v, d, sgn); break;   case 5: r_ph<5, NEG>(v, d, sgn); break;   case 6: r_ph<6, NEG>(v, d, sgn); break;
        case 7: r_ph<7, NEG>(v, d, sgn); break;   case 8: r_ph<8, NEG>(v, d, sgn); break;   case 9: r_ph<9, NEG>(v, d, sgn); break;
        case 10: r_ph<10, NEG>(v, d, sgn); break; case 11: r_ph<11, NEG>(v, d, sgn); break; case 12: r_ph<12, NEG>(v, d, sgn); break;
        case 13: r_ph<13, NEG>(v, d, sgn); break; case 14: r_ph<14, NEG>(v, d, sgn); break; default: r_ph<15, NEG>(v, d, sgn); break;
    }
}

// ---- register-resident gate kernels (compile-time register indices) ---------------------------------
// PRED = false is the common case (no control among the register bits): straight-line code, no predicates.
template <int A, bool REAL, bool PRED>
__device__ __forceinline__ void r_d1(c64 (&v)[16], const c64* __restrict__ m, uint32_t rmask, bool act) {
    // a thread whose thread-bit controls are not satisfied applies the identity instead of branching around
    // the update: a divergent branch here makes the compiler copy the whole register file every iteration
    (void)act;  // the caller already pointed `m` at the identity (mats[0..3]) for inactive threads
    const c64 m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma unroll
    for (int j0 = 0; j0 < 16; ++j0) {
        if ((j0 >> A) & 1) continue;
        const int j1 = j0 | (1 << A);
        if (PRED && (j0 & rmask) != rmask) continue;
        if (REAL) ip_pair_real(v[j0], v[j1], m00.x, m01.x, m10.x, m11.x);
        else ip_pair_cplx(v[j0], v[j1], m00, m01, m10, m11);
    }
}

// one-qubit diagonal (RZ, phase...) on register bit A: the factor of register j is known at compile time
template <int A, bool PRED>
__device__ __forceinline__ void r_dg1(c64 (&v)[16], c64 d0, c64 d1, uint32_t rmask, bool act) {
    (void)act;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (PRED && (j & rmask) != rmask) continue;
        ip_cmul(v[j], ((j >> A) & 1) ? d1 : d0);
    }
}

template <int A, int B>  // A = register bit of matrix bit 0, B = of matrix bit 1, A != B
__device__ __forceinline__ void r_d2(c64 (&v)[16], const c64* __restrict__ M, uint32_t rmask) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (((j >> A) & 1) || ((j >> B) & 1)) continue;
        if (rmask != 0 && (j & rmask) != rmask) continue;
        const int i1 = j | (1 << A), i2 = j | (1 << B), i3 = i1 | (1 << B);
        const c64 a0 = v[j], a1 = v[i1], a2 = v[i2], a3 = v[i3];
        c64 r0 = cmul2(M[0], a0, M[1], a1);   r0 = cfma(M[2], a2, r0);   r0 = cfma(M[3], a3, r0);
        c64 r1 = cmul2(M[4], a0, M[5], a1);   r1 = cfma(M[6], a2, r1);   r1 = cfma(M[7], a3, r1);
        c64 r2 = cmul2(M[8], a0, M[9], a1);   r2 = cfma(M[10], a2, r2);  r2 = cfma(M[11], a3, r2);
        c64 r3 = cmul2(M[12], a0, M[13], a1); r3 = cfma(M[14], a2, r3);  r3 = cfma(M[15], a3, r3);
        v[j] = r0; v[i1] = r1; v[i2] = r2; v[i3] = r3;
    }
}

__device__ __forceinline__ void r_d2_dispatch(c64 (&v)[16], const c64* M, uint32_t a, uint32_t b, uint32_t rmask) {
    switch (a * 4 + b) {
        case 0 * 4 + 1: r_d2<0, 1>(v, M, rmask); break;
        case 0 * 4 + 2: r_d2<0, 2>(v, M, rmask); break;
        case 0 * 4 + 3: r_d2<0, 3>(v, M, rmask); break;
        case 1 * 4 + 0: r_d2<1, 0>(v, M, rmask); break;
        case 1 * 4 + 2: r_d2<1, 2>(v, M, rmask); break;
        case 1 * 4 + 3: r_d2<1, 3>(v, M, rmask); break;
        case 2 * 4 + 0: r_d2<2, 0>(v, M, rmask); break;
        case 2 * 4 + 1: r_d2<2, 1>(v, M, rmask); break;
        case 2 * 4 + 3: r_d2<2, 3>(v, M, rmask); break;
        case 3 * 4 + 0: r_d2<3, 0>(v, M, rmask); break;
        case 3 * 4 + 1: r_d2<3, 1>(v, M, rmask); break;
        default: r_d2<3, 2>(v, M, rmask); break;
    }
}

#ifdef QB_HOST_EMU
#define QB_DYNAMIC_SMEM(name) unsigned char* name = qb_emu_shared()
#else
#define QB_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// The interpreter kernel: one CTA per tile, the pass program in the kernel-parameter (constant) bank.  It runs a pass until
// (or unless) the pass's specialised kernel is available (qb_jit.cu), and every pass of registers too small to be worth a compile.
template <class Prog>
__global__ void __launch_bounds__(RTHREADS, 2) k2r_tile_pass(c64* __restrict__ st, const __grid_constant__ Prog prog) {
    QB_DYNAMIC_SMEM(smem_raw);
    c64* tile = reinterpret_cast<c64*>(smem_raw);
#define QB_TILE_IX blockIdx.x
#define QB_AMP(off) (st + (off))
#define QB_MATS prog.mats
#include "qb_tile2_body.inc"
#undef QB_TILE_IX
#undef QB_AMP
#undef QB_MATS
}

// ---- sharded registers: a tile pass whose tile may contain RANK bits (experimental, QB_SHARD_REMOTE=1) --------------------
// The P = 2^g shards of one register are treated as ONE virtual register of n = n_local + g index bits (amplitude index
// `off`: rank = off >> n_local, offset in that rank's shard = the low n_local bits); every rank maps all shards (CUDA IPC,
// qb_shard.cu).  The pass program is the one the planner and the encoder produce for an n-qubit register: rank bits can be
// tile bits like any other high bit, so a gate with a dense target on a global qubit runs WITHOUT the global<->local
// exchange that otherwise precedes it -- the tile's amplitudes that live on peers are loaded from and stored to peer memory
// over NVLink by the same kernel that applies the gates, tile by tile, while the other resident CTA of the SM computes.
// Work split: the virtual register has 2^(n - 12) tiles; each rank takes the 2^(n_local - 12) tiles whose top g tile-index
// bits equal `fixed` (remote_fixed_bits): a rank bit outside the tile keeps this rank's value (all of the tile's
// amplitudes with that bit are local), a rank bit inside the tile is replaced by the highest local non-tile bits, so that
// the ranks' tile sets are disjoint and cover everything.  Controls and diagonal gates on rank bits outside the tile are
// ordinary "outer" bits of the tile index.  No rank-dependent decision is taken anywhere but in `fixed`.
struct RemoteArgs {
    c64* base[8];        // shard of rank i (base[own rank] = own shard)
    uint32_t n_local;    // index bits per shard
    uint32_t tile_or;    // fixed << (n - 12 - g): OR-ed into blockIdx.x to give the tile index in the virtual register
};

// Host and device: the g fixed top bits of this rank's tile indices, from the high tile bits `hp` (ascending positions).
// Bit order, LSB first: the rank bits that ARE tile bits (their substitutes are the highest local non-tile bits, which sit
// right below the non-tile rank bits in tile-index space), then the rank bits that are not.
inline uint32_t remote_fixed_bits(const uint8_t* hp, uint32_t h, uint32_t n_local, uint32_t g, uint32_t rank) {
    uint32_t in_tile = 0;
    for (uint32_t j = 0; j < h; ++j)
        if (hp[j] >= n_local) in_tile |= 1u << (hp[j] - n_local);
    uint32_t fixed = 0, pos = 0;
    for (uint32_t J = 0; J < g; ++J)
        if ((in_tile >> J) & 1u) fixed |= ((rank >> J) & 1u) << pos++;
    for (uint32_t J = 0; J < g; ++J)
        if (!((in_tile >> J) & 1u)) fixed |= ((rank >> J) & 1u) << pos++;
    return fixed;
}

__device__ __forceinline__ c64* remote_amp(const RemoteArgs& ra, uint64_t lmask, uint64_t off) {
    return ra.base[off >> ra.n_local] + (off & lmask);
}

template <class Prog>
__global__ void __launch_bounds__(RTHREADS, 2) k2r_tile_pass_remote(const __grid_constant__ RemoteArgs ra, const __grid_constant__ Prog prog) {
    QB_DYNAMIC_SMEM(smem_raw);
    c64* tile = reinterpret_cast<c64*>(smem_raw);
    const uint32_t tile_ix = blockIdx.x | ra.tile_or;
    const uint64_t lmask = (1ull << ra.n_local) - 1ull;
#define QB_TILE_IX tile_ix
#define QB_AMP(off) remote_amp(ra, lmask, (off))
#define QB_MATS prog.mats
#include "qb_tile2_body.inc"
#undef QB_TILE_IX
#undef QB_AMP
#undef QB_MATS
}

}  // namespace qb
