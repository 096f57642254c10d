// qb_host_emu.h -- just enough of the CUDA execution model to run the engine's DEVICE code on CPU threads.
//
// Test infrastructure (tests/cpu_kernel): the kernels' headers are compiled with g++ and -DQB_HOST_EMU, one OS thread
// per CUDA thread of a CTA, a pthread barrier for __syncthreads(), CTAs one after the other.  It checks the kernels'
// control flow, indexing and program decoding against the oracle on a machine without a GPU; the inline-PTX
// primitives are replaced by C++ with the same operation order (qb_tile2_kernel.cuh).  It says nothing about
// performance, memory-model subtleties or the PTX itself -- the -m gpu tests remain the parity tests proper.
#pragma once
#include <math.h>
#include <pthread.h>
#include <stdint.h>

struct double2 {
    double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct ulonglong2 {
    unsigned long long x, y;
};
struct qb_emu_dim3 {
    unsigned x, y, z;
};

#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) alignas(n)

extern thread_local qb_emu_dim3 threadIdx, blockIdx;
extern qb_emu_dim3 blockDim, gridDim;
extern pthread_barrier_t qb_emu_barrier;
unsigned char* qb_emu_shared();

static inline void __syncthreads() { pthread_barrier_wait(&qb_emu_barrier); }
