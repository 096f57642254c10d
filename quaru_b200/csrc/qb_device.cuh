// qb_device.cuh -- device helpers: 256-bit global accesses (sm_100+), complex arithmetic, bit tricks.
#pragma once
#ifdef QB_HOST_EMU
#include "qb_host_emu.h"  // tests/cpu_kernel: runs device code on CPU threads (test infrastructure, never in the product build)
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace qb {

struct __align__(32) c64x2 {
    double2 a, b;
};

// Streaming accesses: every amplitude is read once and written once per pass and the state is far
// larger than L2, so loads do not allocate in L1 and stores are plain write-back.
#ifdef QB_HOST_EMU
__device__ __forceinline__ void prefetch_l2(const void*) {}
__device__ __forceinline__ c64x2 ld256(const double2* p) { c64x2 v; v.a = p[0]; v.b = p[1]; return v; }
__device__ __forceinline__ void st256(double2* p, const c64x2& v) { p[0] = v.a; p[1] = v.b; }
__device__ __forceinline__ double2 ld128(const double2* p) { return *p; }
__device__ __forceinline__ void st128(double2* p, const double2& v) { *p = v; }
#else
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ c64x2 ld256(const double2* p) {
    c64x2 v;
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v.a.x), "=d"(v.a.y), "=d"(v.b.x), "=d"(v.b.y)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st256(double2* p, const c64x2& v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.a.x), "d"(v.a.y), "d"(v.b.x), "d"(v.b.y)
                 : "memory");
}
__device__ __forceinline__ double2 ld128(const double2* p) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st128(double2* p, const double2& v) {
    asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
#endif

__device__ __forceinline__ uint64_t insert_zero_bit(uint64_t x, uint32_t p) {
    const uint64_t lo = x & ((1ull << p) - 1ull);
    return ((x >> p) << (p + 1)) | lo;
}

// (a.x + i a.y)(b.x + i b.y)
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// acc + a*b
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 acc) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
    return acc;
}
// m0*a0 + m1*a1
__device__ __forceinline__ double2 cmul2(double2 m0, double2 a0, double2 m1, double2 a1) {
    return cfma(m1, a1, cmul(m0, a0));
}

}  // namespace qb
