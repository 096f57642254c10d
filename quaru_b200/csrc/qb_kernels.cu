// qb_kernels.cu -- single-gate ("unfused") kernels, measurement, init and comparison kernels.
//
// Every kernel here is HBM-bound streaming work over the 2^n_local double2 amplitudes of this
// rank: 256-bit (LDG.E.256 / STG.E.256, sm_100+) accesses, two amplitudes per access, grid sized
// as a multiple of the SM count with a grid-stride loop, and two independent work items in flight
// per thread.  Reference semantics being replaced are cited per kernel (src/register.rs).
#include <algorithm>

#include "qb_common.h"
#include "qb_device.cuh"

namespace qb {

static int num_sms() {  // of the current device (every entry point makes the handle's device current)
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cached = 0;
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        if (cached <= 0) cached = 148;
        cached_dev = dev;
    }
    return cached;
}

// grid for a grid-stride kernel: enough CTAs to cover `items` with `per_thread` items each, capped
// at `ctas_per_sm` resident CTAs per SM (a multiple of the SM count when capped).
static unsigned grid_for(uint64_t items, int threads, int per_thread, int ctas_per_sm) {
    uint64_t need = (items + (uint64_t)threads * per_thread - 1) / ((uint64_t)threads * per_thread);
    uint64_t cap = (uint64_t)num_sms() * ctas_per_sm;
    if (need < 1) need = 1;
    return (unsigned)std::min<uint64_t>(need, cap);
}

// ================================================================================================
// Index enumeration: the indices a controlled gate touches are those with every control bit 1
// and (for the pair kernels) the target bit 0.  Work item w is expanded by inserting a bit at each
// listed position (ascending), which is the reference's `indecies` filter (register.rs:156-157)
// turned inside out: no 2^n scan, O(#inserted bits) per item.
// ================================================================================================
struct InsertList {
    uint32_t cnt;
    uint32_t pos[40];
};

static InsertList make_inserts(uint64_t mask) {
    InsertList l;
    l.cnt = 0;
    for (uint32_t b = 0; b < 64 && l.cnt < 40; ++b)
        if ((mask >> b) & 1) l.pos[l.cnt++] = b;
    return l;
}

__device__ __forceinline__ uint64_t expand_index(uint64_t w, const InsertList& ins) {
#pragma unroll 1
    for (uint32_t j = 0; j < ins.cnt; ++j) w = insert_zero_bit(w, ins.pos[j]);
    return w;
}

struct Mat2 {
    c64 m00, m01, m10, m11;
};

// ================================================================================================
// K1: one dense single-qubit gate (optionally controlled), in place.
// Replaces try_apply_fast with k = 1 (register.rs:140-191): new[p0] = m00 a0 + m01 a1,
// new[p1] = m10 a0 + m11 a1 for every pair that differs in bit `t`.
//   MODE 0: bit 0 is neither target nor control -> a work item is two adjacent pairs:
//           256-bit load of (a[i], a[i+1]) from the bit_t=0 half and from the bit_t=1 half.
//   MODE 1: t == 0 -> the pair itself is one 256-bit vector.
//   MODE 2: a control sits on bit 0 (t >= 1) -> 128-bit accesses.
// ================================================================================================
template <int MODE>
__global__ void __launch_bounds__(256) k1_apply_1q(c64* __restrict__ st, uint64_t items, uint32_t t,
                                                   uint64_t ctrl_mask, InsertList ins, Mat2 m) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t tbit = 1ull << t;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < items; w += 2 * stride) {
        const uint64_t w2 = w + stride;
        const bool two = w2 < items;
        if (MODE == 0) {
            const uint64_t iA = expand_index(w << 1, ins) | ctrl_mask;
            const uint64_t jA = expand_index((two ? w2 : w) << 1, ins) | ctrl_mask;
            c64x2 A = ld256(st + iA), B = ld256(st + (iA | tbit));
            c64x2 C = ld256(st + jA), D = ld256(st + (jA | tbit));
            c64x2 oA, oB, oC, oD;
            oA.a = cmul2(m.m00, A.a, m.m01, B.a); oB.a = cmul2(m.m10, A.a, m.m11, B.a);
            oA.b = cmul2(m.m00, A.b, m.m01, B.b); oB.b = cmul2(m.m10, A.b, m.m11, B.b);
            oC.a = cmul2(m.m00, C.a, m.m01, D.a); oD.a = cmul2(m.m10, C.a, m.m11, D.a);
            oC.b = cmul2(m.m00, C.b, m.m01, D.b); oD.b = cmul2(m.m10, C.b, m.m11, D.b);
            st256(st + iA, oA); st256(st + (iA | tbit), oB);
            if (two) { st256(st + jA, oC); st256(st + (jA | tbit), oD); }
        } else if (MODE == 1) {
            const uint64_t iA = expand_index(w, ins) | ctrl_mask;  // ins contains position 0 -> even
            const uint64_t jA = expand_index(two ? w2 : w, ins) | ctrl_mask;
            c64x2 A = ld256(st + iA), C = ld256(st + jA);
            c64x2 oA, oC;
            oA.a = cmul2(m.m00, A.a, m.m01, A.b); oA.b = cmul2(m.m10, A.a, m.m11, A.b);
            oC.a = cmul2(m.m00, C.a, m.m01, C.b); oC.b = cmul2(m.m10, C.a, m.m11, C.b);
            st256(st + iA, oA);
            if (two) st256(st + jA, oC);
        } else {
            const uint64_t iA = expand_index(w, ins) | ctrl_mask;
            const uint64_t jA = expand_index(two ? w2 : w, ins) | ctrl_mask;
            c64 a0 = ld128(st + iA), a1 = ld128(st + (iA | tbit));
            c64 c0 = ld128(st + jA), c1 = ld128(st + (jA | tbit));
            st128(st + iA, cmul2(m.m00, a0, m.m01, a1)); st128(st + (iA | tbit), cmul2(m.m10, a0, m.m11, a1));
            if (two) { st128(st + jA, cmul2(m.m00, c0, m.m01, c1)); st128(st + (jA | tbit), cmul2(m.m10, c0, m.m11, c1)); }
        }
    }
}

// ================================================================================================
// K4: diagonal gates (no reference constructor is special-cased in Quaru -- it multiplies the
// dense matrix; here amp[i] *= D[bits of i at `targets`], only where all control bits are 1).
//   VEC: bit 0 is not a control -> 256-bit accesses of (a[i], a[i+1]).
// D comes by value for k <= 2 (RZ, phases, CZ-like) and from global memory otherwise.
// ================================================================================================
struct DiagSmall {
    c64 d[4];
};
// k <= 2: select without dynamic indexing of a by-value struct (that would force a local-memory copy)
__device__ __forceinline__ c64 pick_small(const DiagSmall& ds, uint32_t i) {
    const c64 lo = (i & 1) ? ds.d[1] : ds.d[0];
    const c64 hi = (i & 1) ? ds.d[3] : ds.d[2];
    return (i & 2) ? hi : lo;
}
struct TargetList {
    uint32_t k;
    uint32_t t[40];
};

__device__ __forceinline__ uint32_t gather_bits(uint64_t i, const TargetList& tl) {
    uint32_t r = 0;
#pragma unroll 1
    for (uint32_t j = 0; j < tl.k; ++j) r |= (uint32_t)((i >> tl.t[j]) & 1) << j;
    return r;
}

template <bool VEC, bool SMALL>
__global__ void __launch_bounds__(256) k4_apply_diag(c64* __restrict__ st, uint64_t items, uint64_t ctrl_mask,
                                                     InsertList ins, TargetList tl, DiagSmall ds,
                                                     const c64* __restrict__ dg) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < items; w += 2 * stride) {
        const uint64_t w2 = w + stride;
        const bool two = w2 < items;
        if (VEC) {
            const uint64_t i = expand_index(w << 1, ins) | ctrl_mask;
            const uint64_t j = expand_index((two ? w2 : w) << 1, ins) | ctrl_mask;
            c64x2 A = ld256(st + i), C = ld256(st + j);
            const uint32_t ia = gather_bits(i, tl), ib = gather_bits(i | 1, tl);
            const uint32_t ja = gather_bits(j, tl), jb = gather_bits(j | 1, tl);
            c64 fa, fb, ga, gb;
            if (SMALL) { fa = pick_small(ds, ia); fb = pick_small(ds, ib); ga = pick_small(ds, ja); gb = pick_small(ds, jb); }
            else { fa = __ldg(dg + ia); fb = __ldg(dg + ib); ga = __ldg(dg + ja); gb = __ldg(dg + jb); }
            A.a = cmul(fa, A.a); A.b = cmul(fb, A.b);
            C.a = cmul(ga, C.a); C.b = cmul(gb, C.b);
            st256(st + i, A);
            if (two) st256(st + j, C);
        } else {
            const uint64_t i = expand_index(w, ins) | ctrl_mask;
            const uint64_t j = expand_index(two ? w2 : w, ins) | ctrl_mask;
            c64 a = ld128(st + i), c = ld128(st + j);
            const uint32_t ia = gather_bits(i, tl), ja = gather_bits(j, tl);
            c64 fa = SMALL ? pick_small(ds, ia) : __ldg(dg + ia);
            c64 ga = SMALL ? pick_small(ds, ja) : __ldg(dg + ja);
            st128(st + i, cmul(fa, a));
            if (two) st128(st + j, cmul(ga, c));
        }
    }
}

// ================================================================================================
// K3: general k-qubit dense gate, any k <= n_local, matrix and offset table in global memory,
// out of place (in -> out), LANES lanes cooperate on one output amplitude.
// Replaces try_apply_fast for the wide gates of the Grover / Shor configs (k = 5..10 on <= 15
// qubits; examples/shors.rs:187-214) and is the correctness fallback for every other shape.
//   out[o] = sum_c M[row(o)][c] * in[(o & ~tmask) | offs[c]]   if all control bits of o are 1
//   out[o] = in[o]                                              otherwise
// ================================================================================================
template <int LANES>
__global__ void __launch_bounds__(256) k3_apply_general(const c64* __restrict__ in, c64* __restrict__ out,
                                                        uint64_t n_amps, const c64* __restrict__ M,
                                                        const uint64_t* __restrict__ offs, uint32_t k,
                                                        uint64_t tmask, uint64_t ctrl_mask, TargetList tl) {
    const uint64_t gthread = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per_iter = ((uint64_t)gridDim.x * blockDim.x) / LANES;
    const uint64_t iters = (n_amps + per_iter - 1) / per_iter;  // same trip count for every thread (shuffles below)
    const uint32_t sub = threadIdx.x % LANES;
    const uint64_t dim = 1ull << k;
    for (uint64_t it = 0; it < iters; ++it) {
        const uint64_t o = it * per_iter + gthread / LANES;
        const bool valid = o < n_amps;
        const bool act = valid && (o & ctrl_mask) == ctrl_mask;
        c64 acc = make_double2(0.0, 0.0);
        if (act) {
            const uint64_t base = o & ~tmask;
            const c64* Mr = M + (uint64_t)gather_bits(o, tl) * dim;
            for (uint64_t c = sub; c < dim; c += LANES) acc = cfma(Mr[c], in[base | offs[c]], acc);
        }
#pragma unroll
        for (int s = LANES / 2; s >= 1; s >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, s);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, s);
        }
        if (valid && sub == 0) out[o] = act ? acc : in[o];
    }
}

// ================================================================================================
// K3i: the same gate IN PLACE, for k <= 13 -- no second state buffer, so a dense 3+-qubit gate (hadamard_transform on three
// targets, Shor's to_quantum_gate, an unfused dense 2-qubit gate) works on a register that fills the GPU.
// A chunk = 2^gl amplitude groups (a group = the 2^k amplitudes that differ only in the target bits; only groups whose control
// bits are all 1 are enumerated).  A CTA stages a chunk in shared memory (group index fastest across threads, so loads follow
// the non-target low bits), then LANES lanes cooperate on each output exactly as K3 does (same per-lane strided partial sums,
// same xor tree: bit-identical results) and write it straight back: every input of the chunk is already in shared memory.
// ================================================================================================
template <int LANES>
__global__ void __launch_bounds__(256) k3_apply_general_inplace(c64* __restrict__ st, uint64_t n_chunks, uint32_t gl,
                                                                const c64* __restrict__ M, const uint64_t* __restrict__ offs,
                                                                uint32_t k, uint64_t ctrl_mask, InsertList ins) {
    extern __shared__ __align__(16) c64 k3_sh[];   // [group][c]
    constexpr uint32_t SLICES = 256 / LANES;
    const uint32_t dim = 1u << k, gpc = 1u << gl, ch = dim << gl;
    const uint32_t sub = threadIdx.x % LANES, slice = threadIdx.x / LANES;
    const uint32_t iters = (ch + SLICES - 1) / SLICES;   // same trip count for every thread (shuffles below)
    for (uint64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        for (uint32_t e = threadIdx.x; e < ch; e += 256) {
            const uint32_t g = e & (gpc - 1u), c = e >> gl;
            const uint64_t base = expand_index((chunk << gl) | g, ins) | ctrl_mask;
            k3_sh[(g << k) | c] = st[base | offs[c]];
        }
        __syncthreads();
        for (uint32_t it = 0; it < iters; ++it) {
            const uint32_t e = it * SLICES + slice;
            const bool valid = e < ch;
            const uint32_t g = e & (gpc - 1u), row = e >> gl;
            c64 acc = make_double2(0.0, 0.0);
            if (valid) {
                const c64* Mr = M + (uint64_t)row * dim;
                const c64* in = k3_sh + ((uint64_t)g << k);
                for (uint32_t c = sub; c < dim; c += LANES) acc = cfma(Mr[c], in[c], acc);
            }
#pragma unroll
            for (int sft = LANES / 2; sft >= 1; sft >>= 1) {
                acc.x += __shfl_xor_sync(0xffffffffu, acc.x, sft);
                acc.y += __shfl_xor_sync(0xffffffffu, acc.y, sft);
            }
            if (valid && sub == 0) st[(expand_index((chunk << gl) | g, ins) | ctrl_mask) | offs[row]] = acc;
        }
        __syncthreads();   // the next chunk's staging overwrites the shared-memory copy
    }
}

// ================================================================================================
// K6: measurement probabilities -- first loop of try_measure (register.rs:303-315).
// The sum is an exact balanced ADJACENT-PAIR binary tree over the amplitude index (leaves are
// |a|^2 = rn(rn(re*re) + rn(im*im)) put into bucket bit_t(i); missing leaves are +0.0), so the
// result is bit-reproducible, independent of grid shape and of how many GPUs share the state
// (each rank owns an aligned subtree).  tests/ reproduces it bit-for-bit in numpy.
//   level 0: one CTA per chunk of 2^14 amplitudes: 2-amp leaf per lane -> xor-shuffle tree over the
//            32 lanes (xor shuffles form the adjacent-pair tree) -> 256 warp results per CTA ->
//            adjacent-pair tree in shared memory.
//   level>0: one CTA per 1024 adjacent partials.
// ================================================================================================
constexpr int RED_THREADS = 256;
constexpr int RED_ROUNDS = 32;
constexpr uint64_t RED_CHUNK = 2ull * 32 * (RED_THREADS / 32) * RED_ROUNDS;  // 16384 amplitudes

__device__ __forceinline__ double norm_sqr_rn(c64 a) { return __dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y)); }

__device__ __forceinline__ double2 pair_tree_smem(double2* s, int count) {
    // s[0..count) holds the inputs (count is a power of two <= blockDim.x); returns the root in thread 0.
    for (int len = count >> 1; len >= 1; len >>= 1) {
        double2 v = make_double2(0.0, 0.0);
        if ((int)threadIdx.x < len) {
            double2 a = s[2 * threadIdx.x], b = s[2 * threadIdx.x + 1];
            v = make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y));
        }
        __syncthreads();
        if ((int)threadIdx.x < len) s[threadIdx.x] = v;
        __syncthreads();
    }
    return s[0];
}

__global__ void __launch_bounds__(RED_THREADS) k6_prob_level0(const c64* __restrict__ st, uint64_t n_amps, int t,
                                                               double2* __restrict__ partials) {
    __shared__ double2 s[RED_THREADS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t chunk_leaf0 = (uint64_t)blockIdx.x * (RED_CHUNK / 2);
#pragma unroll 4
    for (int j = 0; j < RED_ROUNDS; ++j) {
        const int b = j * (RED_THREADS / 32) + warp;
        const uint64_t i = (chunk_leaf0 + (uint64_t)b * 32 + lane) * 2;
        double2 v = make_double2(0.0, 0.0);
        if (i + 1 < n_amps) {
            c64x2 A = ld256(st + i);
            const double q0 = norm_sqr_rn(A.a), q1 = norm_sqr_rn(A.b);
            if (t == 0) v = make_double2(q0, q1);
            else if (t > 0 && ((i >> t) & 1)) v = make_double2(0.0, __dadd_rn(q0, q1));
            else v = make_double2(__dadd_rn(q0, q1), 0.0);
        } else if (i < n_amps) {  // n_amps == 1
            v = make_double2(norm_sqr_rn(st[i]), 0.0);
        }
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) {
            const double ox = __shfl_xor_sync(0xffffffffu, v.x, m), oy = __shfl_xor_sync(0xffffffffu, v.y, m);
            v.x = __dadd_rn(v.x, ox);
            v.y = __dadd_rn(v.y, oy);
        }
        if (lane == 0) s[b] = v;
    }
    __syncthreads();
    double2 root = pair_tree_smem(s, RED_THREADS);
    if (threadIdx.x == 0) partials[blockIdx.x] = root;
}

__global__ void __launch_bounds__(1024) k6_tree_level(const double2* __restrict__ in, uint64_t count,
                                                      double2* __restrict__ out) {
    __shared__ double2 s[1024];
    const uint64_t i = (uint64_t)blockIdx.x * 1024 + threadIdx.x;
    s[threadIdx.x] = i < count ? in[i] : make_double2(0.0, 0.0);
    __syncthreads();
    double2 root = pair_tree_smem(s, 1024);
    if (threadIdx.x == 0) out[blockIdx.x] = root;
}

// ================================================================================================
// K7: collapse + renormalise -- second loop of try_measure (register.rs:324-339): amplitudes whose
// target bit != outcome become 0 (stored without being read), the others are DIVIDED by
// sqrt(total_prob) (division as in the reference, not a reciprocal multiply).
// ================================================================================================
__global__ void __launch_bounds__(256) k7_collapse(c64* __restrict__ st, uint64_t items, uint32_t t, int res, double s) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const c64 zero = make_double2(0.0, 0.0);
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < items; w += stride) {
        const uint64_t i = w << 1;
        c64x2 o;
        if (t == 0) {
            c64x2 A = ld256(st + i);
            o.a = res ? zero : make_double2(A.a.x / s, A.a.y / s);
            o.b = res ? make_double2(A.b.x / s, A.b.y / s) : zero;
        } else if ((int)((i >> t) & 1) != res) {
            o.a = zero; o.b = zero;
        } else {
            c64x2 A = ld256(st + i);
            o.a = make_double2(A.a.x / s, A.a.y / s);
            o.b = make_double2(A.b.x / s, A.b.y / s);
        }
        st256(st + i, o);
    }
}

__global__ void k7_collapse_tiny(c64* st, uint64_t n_amps, uint32_t t, int res, double s) {
    for (uint64_t i = threadIdx.x; i < n_amps; i += blockDim.x) {
        if ((int)((i >> t) & 1) != res) st[i] = make_double2(0.0, 0.0);
        else st[i] = make_double2(st[i].x / s, st[i].y / s);
    }
}

__global__ void __launch_bounds__(256) k7_scale_all(c64* __restrict__ st, uint64_t n_amps, int keep, double s) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
        if (keep) { c64 a = st[i]; st[i] = make_double2(a.x / s, a.y / s); }
        else st[i] = make_double2(0.0, 0.0);
    }
}

// ================================================================================================
// K8: initial states.  Register::new / from_int (register.rs:41-63) = zeros + one 1.0;
// try_from_qubits (register.rs:91-96) folds kron(q_i, acc), i.e.
// amp[x] = q_{n-1}[b_{n-1}] * ( ... (q_1[b_1] * (q_0[b_0] * 1)) ), multiplied in that order with
// num-complex's unfused formula so the result is bit-identical to the reference's.
// ================================================================================================
__global__ void k8_set_one(c64* st, uint64_t index) { st[index] = make_double2(1.0, 0.0); }

struct ProductQubits {
    uint32_t n;
    c64 q[40][2];
};

__device__ __forceinline__ c64 cmul_unfused(c64 b, c64 a) {
    return make_double2(__dsub_rn(__dmul_rn(b.x, a.x), __dmul_rn(b.y, a.y)),
                        __dadd_rn(__dmul_rn(b.x, a.y), __dmul_rn(b.y, a.x)));
}

__global__ void __launch_bounds__(256) k8_init_product(c64* __restrict__ st, uint64_t n_amps, uint64_t index_base,
                                                       ProductQubits pq) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
        const uint64_t g = index_base + i;
        c64 acc = make_double2(1.0, 0.0);
        for (uint32_t q = 0; q < pq.n; ++q) acc = cmul_unfused(pq.q[q][(g >> q) & 1], acc);
        st[i] = acc;
    }
}

// ================================================================================================
// K10: comparison reductions for parity checks (PartialEq, register.rs:409-413): max |a-b| and
// sum |a-b|^2.  Not order-sensitive (max) / test-only (l2), so a plain two-level reduction.
// ================================================================================================
__global__ void __launch_bounds__(256) k10_diff(const c64* __restrict__ a, const c64* __restrict__ b, uint64_t n_amps,
                                                 double2* __restrict__ partials) {
    __shared__ double2 s[256];
    double mx = 0.0, l2 = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
        const double dx = a[i].x - b[i].x, dy = a[i].y - b[i].y;
        const double d2 = dx * dx + dy * dy;
        // NaN-propagating max: a NaN difference must not be hidden by fmax
        mx = (d2 != d2 || mx != mx) ? __longlong_as_double(0x7ff8000000000000ll) : fmax(mx, d2);
        l2 += d2;
    }
    s[threadIdx.x] = make_double2(mx, l2);
    __syncthreads();
    for (int len = 128; len >= 1; len >>= 1) {
        if ((int)threadIdx.x < len) {
            double2 x = s[threadIdx.x], y = s[threadIdx.x + len];
            const double m2 = (x.x != x.x || y.x != y.x) ? __longlong_as_double(0x7ff8000000000000ll) : fmax(x.x, y.x);
            s[threadIdx.x] = make_double2(m2, x.y + y.y);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = s[0];
}

// ================================================================================================
// host launchers
// ================================================================================================
int ensure_dmat(Reg* r, size_t bytes) {
    if (r->dmat_bytes >= bytes) return QB_OK;
    // stream-ordered: earlier kernels may still be reading the old arena
    if (r->dmat) QB_CUDA(cudaFreeAsync(r->dmat, r->stream));
    r->dmat = nullptr;
    r->dmat_bytes = 0;
    size_t want = std::max<size_t>(bytes, 1 << 20);
    QB_CUDA(cudaMallocAsync((void**)&r->dmat, want, r->stream));
    r->dmat_bytes = want;
    return QB_OK;
}

static int ensure_partials(Reg* r, size_t len) {
    if (r->partials_len >= len) return QB_OK;
    if (r->partials) QB_CUDA(cudaFreeAsync(r->partials, r->stream));
    r->partials = nullptr;
    r->partials_len = 0;
    QB_CUDA(cudaMallocAsync((void**)&r->partials, len * sizeof(double2), r->stream));
    r->partials_len = len;
    return QB_OK;
}

int k_init_basis(Reg* r, uint64_t local_index, bool owned) {
    const uint64_t n_amps = 1ull << r->n_local;
    QB_CUDA(cudaMemsetAsync(r->state, 0, n_amps * sizeof(c64), r->stream));
    if (owned) {
        k8_set_one<<<1, 1, 0, r->stream>>>(r->state, local_index);
        QB_CUDA(cudaGetLastError());
        r->st_launches++;
    }
    return QB_OK;
}

int k_init_product(Reg* r, const double* q) {
    ProductQubits pq;
    pq.n = r->n;
    for (uint32_t i = 0; i < r->n; ++i) {
        pq.q[i][0] = make_double2(q[4 * i + 0], q[4 * i + 1]);
        pq.q[i][1] = make_double2(q[4 * i + 2], q[4 * i + 3]);
    }
    const uint64_t n_amps = 1ull << r->n_local;
    k8_init_product<<<grid_for(n_amps, 256, 4, 8), 256, 0, r->stream>>>(r->state, n_amps,
                                                                        (uint64_t)r->rank << r->n_local, pq);
    QB_CUDA(cudaGetLastError());
    r->st_launches++;
    return QB_OK;
}

constexpr uint32_t K3I_MAX_K = 13;   // 2^13 amplitudes = 128 KB of shared memory per group

static int launch_general(Reg* r, const Gate& g) {
    const uint64_t n_amps = 1ull << r->n_local;
    const bool in_place = g.k <= K3I_MAX_K && r->force_path != 5;
    if (!in_place && !r->scratch) {
        // out-of-place (k > 13, or QB_OPT_FORCE_PATH = 5): needs a second state-sized buffer
        cudaError_t e = cudaMalloc((void**)&r->scratch, n_amps * sizeof(c64));
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_error(QB_ERR_OOM, (int64_t)g.k,
                             "a dense %u-qubit gate on %u qubits needs a second state buffer (%.1f GB) that does not fit",
                             g.k, r->n_local, n_amps * 16.0 / 1e9);
        }
    }
    const uint64_t dim = 1ull << g.k;
    const size_t mat_bytes = dim * dim * sizeof(c64), off_bytes = dim * sizeof(uint64_t);
    QB_TRY(ensure_dmat(r, mat_bytes + off_bytes));
    std::vector<uint64_t> offs(dim);
    for (uint64_t c = 0; c < dim; ++c) {
        uint64_t o = 0;
        for (uint32_t j = 0; j < g.k; ++j)
            if ((c >> j) & 1) o |= 1ull << g.targets[j];
        offs[c] = o;
    }
    QB_CUDA(cudaMemcpyAsync(r->dmat, g.matrix.data(), mat_bytes, cudaMemcpyHostToDevice, r->stream));
    uint64_t* doffs = (uint64_t*)((char*)r->dmat + mat_bytes);
    QB_CUDA(cudaMemcpyAsync(doffs, offs.data(), off_bytes, cudaMemcpyHostToDevice, r->stream));
    QB_CUDA(cudaStreamSynchronize(r->stream));  // `offs` is a pageable local: make sure it has been consumed
    const uint64_t tmask = g.target_mask();
    if (in_place) {
        const uint32_t nc = (uint32_t)__builtin_popcountll(g.ctrl_mask);
        const uint32_t free_bits = r->n_local - g.k - nc;   // bits of the group index
        const uint32_t gl = std::min<uint32_t>(free_bits, g.k < 11 ? 11 - g.k : 0);   // chunks of 2048 amplitudes (32 KB)
        const uint64_t n_chunks = 1ull << (free_bits - gl);
        const size_t smem = sizeof(c64) << (g.k + gl);
        const InsertList ins = make_inserts(tmask | g.ctrl_mask);
        const unsigned grid = (unsigned)std::min<uint64_t>(n_chunks, (uint64_t)num_sms() * 4);
#define QB_GENI(L)                                                                                                              \
    do {                                                                                                                        \
        if (smem > 48 * 1024)                                                                                                   \
            QB_CUDA(cudaFuncSetAttribute(k3_apply_general_inplace<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k3_apply_general_inplace<L><<<grid, 256, smem, r->stream>>>(r->state, n_chunks, gl, r->dmat, doffs, g.k, g.ctrl_mask, ins); \
    } while (0)
        if (g.k >= 5) QB_GENI(32);
        else if (g.k == 4) QB_GENI(16);
        else if (g.k == 3) QB_GENI(8);
        else if (g.k == 2) QB_GENI(4);
        else if (g.k == 1) QB_GENI(2);
        else QB_GENI(1);
#undef QB_GENI
        QB_CUDA(cudaGetLastError());
        r->st_launches++;
        r->st_passes++;
        r->st_bytes += (int64_t)((n_amps >> nc) * 32);
        return QB_OK;
    }
    TargetList tl;
    tl.k = g.k;
    for (uint32_t j = 0; j < g.k; ++j) tl.t[j] = g.targets[j];
#define QB_GEN(L)                                                                                              \
    k3_apply_general<L><<<grid_for(n_amps * L, 256, 1, 8), 256, 0, r->stream>>>(r->state, r->scratch, n_amps, \
                                                                                r->dmat, doffs, g.k, tmask,  \
                                                                                g.ctrl_mask, tl)
    if (g.k >= 5) QB_GEN(32);
    else if (g.k == 4) QB_GEN(16);
    else if (g.k == 3) QB_GEN(8);
    else if (g.k == 2) QB_GEN(4);
    else if (g.k == 1) QB_GEN(2);
    else QB_GEN(1);
#undef QB_GEN
    QB_CUDA(cudaGetLastError());
    if (r->world > 1) {   // peers have mapped r->state (CUDA IPC): keep the pointer, copy the result back
        QB_CUDA(cudaMemcpyAsync(r->state, r->scratch, n_amps * sizeof(c64), cudaMemcpyDeviceToDevice, r->stream));
    } else {
        std::swap(r->state, r->scratch);
    }
    r->st_launches++;
    r->st_passes++;
    r->st_bytes += (int64_t)(n_amps * 32);
    return QB_OK;
}

static int launch_diag(Reg* r, const Gate& g) {
    const uint64_t n_amps = 1ull << r->n_local;
    const uint32_t nc = (uint32_t)__builtin_popcountll(g.ctrl_mask);
    const bool vec = !(g.ctrl_mask & 1) && r->n_local >= 1 + nc;
    InsertList ins = make_inserts(g.ctrl_mask);
    TargetList tl;
    tl.k = g.k;
    for (uint32_t j = 0; j < g.k; ++j) tl.t[j] = g.targets[j];
    DiagSmall ds;
    const bool small = g.k <= 2;
    const c64* dg = nullptr;
    if (small) {
        for (uint32_t j = 0; j < 4; ++j) ds.d[j] = j < (1u << g.k) ? g.matrix[j] : make_double2(0, 0);
    } else {
        QB_TRY(ensure_dmat(r, g.matrix.size() * sizeof(c64)));
        QB_CUDA(cudaMemcpyAsync(r->dmat, g.matrix.data(), g.matrix.size() * sizeof(c64), cudaMemcpyHostToDevice,
                                r->stream));
        dg = r->dmat;
    }
    const uint64_t items = vec ? (n_amps >> (nc + 1)) : (n_amps >> nc);
    const unsigned grid = grid_for(items, 256, 2, 8);
    if (vec && small) k4_apply_diag<true, true><<<grid, 256, 0, r->stream>>>(r->state, items, g.ctrl_mask, ins, tl, ds, dg);
    else if (vec) k4_apply_diag<true, false><<<grid, 256, 0, r->stream>>>(r->state, items, g.ctrl_mask, ins, tl, ds, dg);
    else if (small) k4_apply_diag<false, true><<<grid, 256, 0, r->stream>>>(r->state, items, g.ctrl_mask, ins, tl, ds, dg);
    else k4_apply_diag<false, false><<<grid, 256, 0, r->stream>>>(r->state, items, g.ctrl_mask, ins, tl, ds, dg);
    QB_CUDA(cudaGetLastError());
    r->st_launches++;
    r->st_passes++;
    // physical bytes: touched amplitudes, rounded up to whole 32-byte sectors when bit 0 is a control
    r->st_bytes += (int64_t)((n_amps >> (nc - ((g.ctrl_mask & 1) ? 1 : 0))) * 32);
    return QB_OK;
}

static int launch_1q(Reg* r, const Gate& g) {
    const uint64_t n_amps = 1ull << r->n_local;
    const uint32_t t = g.targets[0];
    const uint32_t nc = (uint32_t)__builtin_popcountll(g.ctrl_mask);
    Mat2 m{g.matrix[0], g.matrix[1], g.matrix[2], g.matrix[3]};
    InsertList ins = make_inserts(g.ctrl_mask | (1ull << t));
    const uint64_t pairs = n_amps >> (nc + 1);
    if (t == 0) {
        k1_apply_1q<1><<<grid_for(pairs, 256, 2, 8), 256, 0, r->stream>>>(r->state, pairs, t, g.ctrl_mask, ins, m);
    } else if (g.ctrl_mask & 1) {
        k1_apply_1q<2><<<grid_for(pairs, 256, 2, 8), 256, 0, r->stream>>>(r->state, pairs, t, g.ctrl_mask, ins, m);
    } else {
        const uint64_t items = pairs >> 1;  // >= 1: bit 0 is free, so n_local >= nc + 2
        k1_apply_1q<0><<<grid_for(items, 256, 2, 8), 256, 0, r->stream>>>(r->state, items, t, g.ctrl_mask, ins, m);
    }
    QB_CUDA(cudaGetLastError());
    r->st_launches++;
    r->st_passes++;
    r->st_bytes += (int64_t)((n_amps >> (nc - ((g.ctrl_mask & 1) ? 1 : 0))) * 32);
    return QB_OK;
}

int k_apply_gate_unfused(Reg* r, const Gate& g) {
    if (r->force_path == 1 || r->force_path == 5) {
        if (g.kind == GK_DIAG) {  // the generic kernel takes dense matrices: expand the diagonal
            Gate d = g;
            d.kind = GK_DENSE;
            const uint64_t dim = 1ull << g.k;
            d.matrix.assign(dim * dim, make_double2(0, 0));
            for (uint64_t i = 0; i < dim; ++i) d.matrix[i * dim + i] = g.matrix[i];
            return launch_general(r, d);
        }
        return launch_general(r, g);
    }
    if (g.kind == GK_DIAG) return launch_diag(r, g);
    if (g.k == 1) return launch_1q(r, g);
    return launch_general(r, g);
}

int k_probabilities(Reg* r, int t, double* p0, double* p1) {
    const uint64_t n_amps = 1ull << r->n_local;
    uint64_t count = (n_amps + RED_CHUNK - 1) / RED_CHUNK;
    size_t need = count;
    for (uint64_t c = count; c > 1; c = (c + 1023) / 1024) need += (c + 1023) / 1024;
    QB_TRY(ensure_partials(r, need + 1));
    k6_prob_level0<<<(unsigned)count, RED_THREADS, 0, r->stream>>>(r->state, n_amps, t, r->partials);
    QB_CUDA(cudaGetLastError());
    r->st_launches++;
    double2* cur = r->partials;
    while (count > 1) {
        const uint64_t nxt = (count + 1023) / 1024;
        k6_tree_level<<<(unsigned)nxt, 1024, 0, r->stream>>>(cur, count, cur + count);
        QB_CUDA(cudaGetLastError());
        r->st_launches++;
        cur += count;
        count = nxt;
    }
    double2 h;
    QB_CUDA(cudaMemcpyAsync(&h, cur, sizeof(double2), cudaMemcpyDeviceToHost, r->stream));
    QB_CUDA(cudaStreamSynchronize(r->stream));
    *p0 = h.x;
    *p1 = h.y;
    r->st_passes++;
    r->st_bytes += (int64_t)(n_amps * 16);
    return QB_OK;
}

int k_collapse(Reg* r, uint32_t t, int outcome, double total_prob) {
    const uint64_t n_amps = 1ull << r->n_local;
    const double s = sqrt(total_prob);
    if (n_amps < 2) {
        k7_collapse_tiny<<<1, 32, 0, r->stream>>>(r->state, n_amps, t, outcome, s);
    } else {
        const uint64_t items = n_amps >> 1;
        k7_collapse<<<grid_for(items, 256, 2, 8), 256, 0, r->stream>>>(r->state, items, t, outcome, s);
    }
    QB_CUDA(cudaGetLastError());
    r->st_launches++;
    r->st_passes++;
    r->st_bytes += (int64_t)(t == 0 ? n_amps * 32 : n_amps * 24);
    return QB_OK;
}

int k_scale_or_zero_all(Reg* r, bool keep, double total_prob) {
    const uint64_t n_amps = 1ull << r->n_local;
    k7_scale_all<<<grid_for(n_amps, 256, 4, 8), 256, 0, r->stream>>>(r->state, n_amps, keep ? 1 : 0, sqrt(total_prob));
    QB_CUDA(cudaGetLastError());
    r->st_launches++;
    r->st_passes++;
    r->st_bytes += (int64_t)(keep ? n_amps * 32 : n_amps * 16);
    return QB_OK;
}

int k_diff(Reg* a, Reg* b, double* max_abs, double* l2sq) {
    const uint64_t n_amps = 1ull << a->n_local;
    const unsigned grid = grid_for(n_amps, 256, 8, 8);
    QB_TRY(ensure_partials(a, grid + 1));
    // b's pending work must be visible to a's stream
    QB_CUDA(cudaStreamSynchronize(b->stream));
    k10_diff<<<grid, 256, 0, a->stream>>>(a->state, b->state, n_amps, a->partials);
    QB_CUDA(cudaGetLastError());
    a->st_launches++;
    std::vector<double2> h(grid);
    QB_CUDA(cudaMemcpyAsync(h.data(), a->partials, grid * sizeof(double2), cudaMemcpyDeviceToHost, a->stream));
    QB_CUDA(cudaStreamSynchronize(a->stream));
    double mx = 0.0, l2 = 0.0;
    for (unsigned i = 0; i < grid; ++i) {
        if (h[i].x != h[i].x || mx != mx) mx = NAN; else mx = std::max(mx, h[i].x);
        l2 += h[i].y;
    }
    *max_abs = sqrt(mx);
    *l2sq = l2;
    return QB_OK;
}

}  // namespace qb
