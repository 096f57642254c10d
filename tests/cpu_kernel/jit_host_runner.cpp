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
alignas(32) static double2 g_shared[4096];
double2* qb_emu_shared = g_shared;

#include QB_JIT_SRC

// grid = 0: one CTA per tile; else `grid` persistent CTAs.  remote: `bases` = world shard pointers, the tiles of rank `rank`.
// exchanging pass (-DQB_JIT_XCHG): `state` = this rank's shard (read), `bases` = every rank's SECOND buffer (written),
// `tile_or` carries the rank.
extern "C" int jit_host_run(double* state, unsigned n_tiles, unsigned grid, unsigned tile_or, void* const* bases, unsigned world,
                            unsigned n_local) {
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
}
