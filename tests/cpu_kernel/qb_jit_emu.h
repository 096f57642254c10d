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

// ---- the pipelined shape (producer warp + bulk-async tile ring + two consumer groups): mbarriers, bulk copies and named
// barriers emulated with mutexes, memcpy and pthread barriers.  A protocol error (a count that never reaches zero, a parity
// mix-up) shows up here as a test that hangs or as wrong amplitudes.
#include <string.h>
#include <atomic>
#include <mutex>
#include <sched.h>
struct QB_MBAR_T {
    std::mutex mu;
    int count = 0, pending = 0;
    long long tx = 0;
    unsigned phase = 0;
    char pad[24];
};
static inline void qb_emu_mbar_check(QB_MBAR_T* b) {   // b->mu held
    if (b->pending == 0 && b->tx == 0) { b->phase++; b->pending = b->count; }
}
static inline void qb_emu_mbar_init(QB_MBAR_T* b, int c) { new (b) QB_MBAR_T(); b->count = b->pending = c; }
#include <stdio.h>
#include <stdlib.h>
static inline void qb_emu_mbar_fail(const char* what) {   // what the hardware answers with a trap ("Warp Illegal Instruction")
    fprintf(stderr, "qb_jit_emu: mbarrier protocol error: %s\n", what);
    abort();
}
static inline void qb_emu_mbar_arrive(QB_MBAR_T* b, long long expect) {
    std::lock_guard<std::mutex> lk(b->mu);
    if (b->pending <= 0) qb_emu_mbar_fail("arrival on a phase whose arrivals are complete (pending count underflow)");
    b->tx += expect;
    if (b->tx > 1048575 || b->tx < -1048575) qb_emu_mbar_fail("tx-count out of range");
    b->pending--;
    qb_emu_mbar_check(b);
}
static inline void qb_emu_mbar_complete(QB_MBAR_T* b, long long bytes) {
    std::lock_guard<std::mutex> lk(b->mu);
    b->tx -= bytes;
    if (b->tx > 1048575 || b->tx < -1048575) qb_emu_mbar_fail("tx-count out of range");
    qb_emu_mbar_check(b);
}
static inline void qb_emu_mbar_wait(QB_MBAR_T* b, unsigned parity) {   // returns once the phase of that parity has completed
    for (;;) {
        { std::lock_guard<std::mutex> lk(b->mu); if ((b->phase & 1u) != parity) return; }
        sched_yield();
    }
}
extern pthread_barrier_t qb_emu_group_barrier[2], qb_emu_block_barrier, qb_emu_warp_barrier[17];
extern unsigned char* qb_emu_shared_raw;
#define QB_KERNEL_PIPE(name) extern "C" void name(double2* st, u32 tile_or, u32 n_tiles)
#define QB_KERNEL_RING(name) extern "C" void name(double2* st, u32 tile_or, u32 n_tiles)
#define QB_SMEM_RAW(t) unsigned char* t = qb_emu_shared_raw
#define QB_MBAR_INIT(B, C) qb_emu_mbar_init((B), (C))
#define QB_MBAR_INIT_FENCE() ((void)0)
#define QB_MBAR_EXPECT_TX(B, N) qb_emu_mbar_arrive((B), (N))
#define QB_MBAR_ARRIVE(B) qb_emu_mbar_arrive((B), 0)
#define QB_MBAR_WAIT(B, P) qb_emu_mbar_wait((B), (P))
#define QB_BULK_G2S(D, S, N, B) do { memcpy((D), (S), (N)); qb_emu_mbar_complete((B), (N)); } while (0)
#define QB_BULK_S2G(D, S, N) memcpy((D), (S), (N))
#define QB_BULK_COMMIT() ((void)0)
#define QB_BULK_WAIT_READ() ((void)0)
#define QB_FENCE_ASYNC() std::atomic_thread_fence(std::memory_order_seq_cst)
#define QB_GROUP_SYNC(G) pthread_barrier_wait(&qb_emu_group_barrier[(G)])
#define QB_SYNCWARP() pthread_barrier_wait(&qb_emu_warp_barrier[QB_TID >> 5])   /* lanes are OS threads here: a real barrier */
#define QB_BLOCK_SYNC() pthread_barrier_wait(&qb_emu_block_barrier)
// the prefetching ring (pipelined = 3): cp.async is an immediate copy here; the done[] flags are atomics
#define QB_CP_ASYNC16(D, S) memcpy((void*)(D), (const void*)(S), 16)
#define QB_CP_ASYNC_COMMIT() ((void)0)
#define QB_CP_ASYNC_WAIT_ALL() ((void)0)
#define QB_FLAG_LD(P) __atomic_load_n((P), __ATOMIC_ACQUIRE)
#define QB_FLAG_ST(P, V) __atomic_store_n((P), (V), __ATOMIC_RELEASE)
#define QB_SPIN() sched_yield()
