// qb_jit_emu.h -- host stand-ins for the few CUDA constructs the generated tile kernels (quaru_b200/csrc/qb_jit_gen.cu) use,
// so that the GENERATED SOURCE ITSELF can be compiled with g++ -DQB_HOST_EMU and replayed against the oracle without a GPU:
// one OS thread per CUDA thread of a CTA, a pthread barrier for __syncthreads(), CTAs one after the other
// (jit_host_runner.cpp).  Test infrastructure only.
#pragma once
#include <pthread.h>
#include <stdint.h>

typedef unsigned int u32;
typedef unsigned long long u64;
struct alignas(16) double2 {
    double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct qb_emu_dim3 {
    unsigned x, y, z;
};
extern thread_local qb_emu_dim3 qb_emu_thread, qb_emu_block;
extern qb_emu_dim3 qb_emu_grid;
extern pthread_barrier_t qb_emu_barrier;
extern double2* qb_emu_shared;
struct QbRemote {
    double2* base[8];
    u32 n_local;
    u32 pad;
};

#define __device__
#define __forceinline__ inline
#define QB_LD(X, Y, P) do { const double2* p_ = (P); (X) = p_->x; (Y) = p_->y; } while (0)
#define QB_ST(P, X, Y) do { double2* p_ = (P); p_->x = (X); p_->y = (Y); } while (0)
#define QB_PREFETCH(P) ((void)(P))
#define QB_SYNC() pthread_barrier_wait(&qb_emu_barrier)
#define QB_TABLE_DECL static const double QB_TAB[]
#define QB_TAB_AT(i) (make_double2(QB_TAB[2 * (i)], QB_TAB[2 * (i) + 1]))
#define QB_SMEM(t) double2* t = qb_emu_shared
#define QB_TID qb_emu_thread.x
#define QB_BID qb_emu_block.x
#define QB_GRID qb_emu_grid.x
#define QB_KERNEL(name, minb) extern "C" void name(double2* st, u32 tile_or, u32 n_tiles)
#define QB_KERNEL_REMOTE(name, minb) extern "C" void name(const QbRemote ra, u32 tile_or, u32 n_tiles)
#define QB_KERNEL_XCHG(name, minb) extern "C" void name(const double2* st, const QbRemote out, u32 rank, u32 tile_or, u32 n_tiles)

// ---- the prefetching ring (two groups of 256 threads, per-thread cp.async into a ring of three tile buffers, done[] flags
// between the groups): named barriers are pthread barriers, cp.async is an immediate copy, the flags are atomics.  A prefetch
// that is issued before the other group has left the buffer shows up as wrong amplitudes.
#include <string.h>
#include <sched.h>
extern pthread_barrier_t qb_emu_group_barrier[2], qb_emu_block_barrier;
extern unsigned char* qb_emu_shared_raw;
#define QB_KERNEL_RING(name) extern "C" void name(double2* st, u32 tile_or, u32 n_tiles)
#define QB_SMEM_RAW(t) unsigned char* t = qb_emu_shared_raw
#define QB_GROUP_SYNC(G) pthread_barrier_wait(&qb_emu_group_barrier[(G)])
#define QB_BLOCK_SYNC() pthread_barrier_wait(&qb_emu_block_barrier)
#define QB_CP_ASYNC16(D, S) memcpy((void*)(D), (const void*)(S), 16)
#define QB_CP_ASYNC_COMMIT() ((void)0)
#define QB_CP_ASYNC_WAIT_ALL() ((void)0)
#define QB_FLAG_LD(P) __atomic_load_n((P), __ATOMIC_ACQUIRE)
#define QB_FLAG_ST(P, V) __atomic_store_n((P), (V), __ATOMIC_RELEASE)
#define QB_SPIN() sched_yield()
