// jit_host_runner.cpp -- runs ONE generated tile kernel (the text qb_jit_source_debug returns, included below through
// -DQB_JIT_SRC="...") on CPU threads over a host state vector.  Built per program by tests/test_jit_cpu.py with
//   g++ -std=c++17 -O1 -ffp-contract=off -DQB_HOST_EMU -DQB_JIT_SRC='"<file>"' -shared -fPIC -pthread -I tests/cpu_kernel
// Test infrastructure; see qb_jit_emu.h.
#include <thread>
#include <vector>

#include "qb_jit_emu.h"

thread_local qb_emu_dim3 qb_emu_thread, qb_emu_block;
qb_emu_dim3 qb_emu_grid;
pthread_barrier_t qb_emu_barrier;
alignas(128) static unsigned char g_shared_raw[3 * 65536 + 1024];
double2* qb_emu_shared = reinterpret_cast<double2*>(g_shared_raw);
unsigned char* qb_emu_shared_raw = g_shared_raw;
pthread_barrier_t qb_emu_group_barrier[2], qb_emu_block_barrier;

#include QB_JIT_SRC

// grid = 0: one CTA per tile; else `grid` persistent CTAs.  remote: `bases` = world shard pointers, the tiles of rank `rank`.
// exchanging pass (-DQB_JIT_XCHG): `state` = this rank's shard (read), `bases` = every rank's SECOND buffer (written),
// `tile_or` carries the rank.
extern "C" int jit_host_run(double* state, unsigned n_tiles, unsigned grid, unsigned tile_or, void* const* bases, unsigned world,
                            unsigned n_local) {
#ifdef QB_JIT_PIPE
    // prefetching ring: `grid` persistent CTAs of QB_JIT_PIPE = 512 threads (two groups of 256), one after the other
    (void)bases; (void)world; (void)n_local;
    const unsigned ctas_p = grid ? grid : 1;
    qb_emu_grid = {ctas_p, 1, 1};
    for (unsigned b = 0; b < ctas_p; ++b) {
        pthread_barrier_init(&qb_emu_block_barrier, nullptr, QB_JIT_PIPE);
        pthread_barrier_init(&qb_emu_group_barrier[0], nullptr, 256);
        pthread_barrier_init(&qb_emu_group_barrier[1], nullptr, 256);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < QB_JIT_PIPE; ++t)
            th.emplace_back([=]() {
                qb_emu_thread = {t, 0, 0};
                qb_emu_block = {b, 0, 0};
                qb_jit_pass(reinterpret_cast<double2*>(state), tile_or, n_tiles);
            });
        for (std::thread& x : th) x.join();
        pthread_barrier_destroy(&qb_emu_block_barrier);
        pthread_barrier_destroy(&qb_emu_group_barrier[0]);
        pthread_barrier_destroy(&qb_emu_group_barrier[1]);
    }
    return 0;
#else
    const unsigned ctas = grid ? grid : n_tiles;
    qb_emu_grid = {ctas, 1, 1};
    pthread_barrier_init(&qb_emu_barrier, nullptr, 256);
    std::vector<std::thread> threads;
    for (unsigned t = 0; t < 256; ++t)
        threads.emplace_back([=]() {
            for (unsigned b = 0; b < ctas; ++b) {
                qb_emu_thread = {t, 0, 0};
                qb_emu_block = {b, 0, 0};
#if defined(QB_JIT_XCHG)
                QbRemote out;
                for (unsigned p = 0; p < 8; ++p) out.base[p] = p < world ? reinterpret_cast<double2*>(bases[p]) : nullptr;
                out.n_local = n_local;
                out.pad = 0;
                qb_jit_pass(reinterpret_cast<const double2*>(state), out, tile_or, 0u, n_tiles);
#elif defined(QB_JIT_REMOTE)
                QbRemote ra;
                for (unsigned p = 0; p < 8; ++p) ra.base[p] = p < world ? reinterpret_cast<double2*>(bases[p]) : nullptr;
                ra.n_local = n_local;
                ra.pad = 0;
                qb_jit_pass(ra, tile_or, n_tiles);
#else
                (void)bases; (void)world; (void)n_local;
                qb_jit_pass(reinterpret_cast<double2*>(state), tile_or, n_tiles);
#endif
                pthread_barrier_wait(&qb_emu_barrier);   // the next CTA reuses the shared-memory array
            }
        });
    for (std::thread& th : threads) th.join();
    pthread_barrier_destroy(&qb_emu_barrier);
    return 0;
#endif
}
