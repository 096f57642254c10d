// k2r_host.cpp -- runs the register-blocked tile kernel (quaru_b200/csrc/qb_tile2_kernel.cuh, the code that is
// compiled for sm_100a) on CPU threads, on programs serialised by qb_rounds_debug.  Test infrastructure; see
// qb_host_emu.h.  Built by tests/test_cpu_kernel.py with
//   g++ -std=c++17 -O1 -DQB_HOST_EMU -shared -fPIC -pthread -I tests/cpu_kernel -I quaru_b200/csrc
#include <string.h>

#include <thread>
#include <vector>

#include "qb_tile2_kernel.cuh"

thread_local qb_emu_dim3 threadIdx, blockIdx;
qb_emu_dim3 blockDim, gridDim;
pthread_barrier_t qb_emu_barrier;
alignas(32) static unsigned char g_shared[227 * 1024];
unsigned char* qb_emu_shared() { return g_shared; }

using namespace qb;

// Rebuilds the kernel-parameter struct from the serialised form (layout: qb_tile2.cu, encode_rounds_debug).
template <class Prog>
static bool deserialise(const unsigned char* bytes, uint64_t len, Prog* prog) {
    memset(prog, 0, sizeof(*prog));
    if (len < 28) return false;
    uint32_t hdr[5];
    memcpy(hdr, bytes, 20);
    const uint32_t n_rounds = hdr[2], n_ops = hdr[3], n_mats = hdr[4];
    if (n_rounds > sizeof(prog->rounds) / sizeof(RoundDesc) || n_ops > sizeof(prog->ops) / sizeof(ROp) ||
        n_mats > sizeof(prog->mats) / sizeof(c64))
        return false;
    const uint64_t need = 28 + (uint64_t)n_rounds * sizeof(RoundDesc) + (uint64_t)n_ops * sizeof(ROp) + (uint64_t)n_mats * sizeof(c64) + 256;
    if (len != need) return false;
    prog->lc = hdr[0];
    prog->h = hdr[1];
    prog->n_rounds = n_rounds;
    prog->n_mats = n_mats;
    const unsigned char* r = bytes + 20;
    memcpy(prog->hp, r, 8); r += 8;
    memcpy(prog->rounds, r, n_rounds * sizeof(RoundDesc)); r += n_rounds * sizeof(RoundDesc);
    memcpy(prog->ops, r, n_ops * sizeof(ROp)); r += n_ops * sizeof(ROp);
    memcpy(prog->mats, r, n_mats * sizeof(c64)); r += n_mats * sizeof(c64);
    memcpy(prog->gro_first, r, 128); r += 128;
    memcpy(prog->gro_last, r, 128);
    return true;
}

template <class Prog>
static void run_grid(double* state, uint32_t n_tiles, const Prog& prog) {
    blockDim = {RTHREADS, 1, 1};
    gridDim = {n_tiles, 1, 1};
    pthread_barrier_init(&qb_emu_barrier, nullptr, RTHREADS);
    std::vector<std::thread> threads;
    for (unsigned t = 0; t < RTHREADS; ++t)
        threads.emplace_back([=, &prog]() {
            for (unsigned b = 0; b < n_tiles; ++b) {   // CTAs one after the other; all threads of a CTA in lock step at barriers
                threadIdx = {t, 0, 0};
                blockIdx = {b, 0, 0};
                k2r_tile_pass<Prog>(reinterpret_cast<c64*>(state), prog);
                pthread_barrier_wait(&qb_emu_barrier);   // the next CTA reuses the shared-memory array
            }
        });
    for (std::thread& th : threads) th.join();
    pthread_barrier_destroy(&qb_emu_barrier);
}

extern "C" int k2r_host_run(double* state, uint32_t n, const unsigned char* program, uint64_t len, int use_small) {
    if (n < RT) return -1;
    const uint32_t n_tiles = 1u << (n - RT);
    if (use_small) {
        static RProgSmall small;
        if (!deserialise(program, len, &small)) return -2;
        run_grid(state, n_tiles, small);
    } else {
        static RProgLarge large;
        if (!deserialise(program, len, &large)) return -2;
        run_grid(state, n_tiles, large);
    }
    return 0;
}

// the remote tile pass (sharded registers): `bases[p]` = shard of rank p, a host buffer of 2^(n_virtual - g) amplitudes;
// runs the tiles that rank `rank` of `world` = 2^g would process, with the product's own work split (remote_fixed_bits)
template <class Prog>
static void run_grid_remote(double* const* bases, uint32_t world, uint32_t rank, uint32_t n_virtual, const Prog& prog) {
    uint32_t g = 0;
    while ((1u << g) < world) ++g;
    RemoteArgs ra;
    memset(&ra, 0, sizeof(ra));
    for (uint32_t p = 0; p < world; ++p) ra.base[p] = reinterpret_cast<c64*>(bases[p]);
    ra.n_local = n_virtual - g;
    ra.tile_or = remote_fixed_bits(prog.hp, prog.h, ra.n_local, g, rank) << (n_virtual - RT - g);
    const uint32_t n_tiles = 1u << (ra.n_local - RT);
    blockDim = {RTHREADS, 1, 1};
    gridDim = {n_tiles, 1, 1};
    pthread_barrier_init(&qb_emu_barrier, nullptr, RTHREADS);
    std::vector<std::thread> threads;
    for (unsigned t = 0; t < RTHREADS; ++t)
        threads.emplace_back([=, &prog]() {
            for (unsigned b = 0; b < n_tiles; ++b) {
                threadIdx = {t, 0, 0};
                blockIdx = {b, 0, 0};
                k2r_tile_pass_remote<Prog>(ra, prog);
                pthread_barrier_wait(&qb_emu_barrier);
            }
        });
    for (std::thread& th : threads) th.join();
    pthread_barrier_destroy(&qb_emu_barrier);
}

extern "C" int k2r_host_run_remote(double* const* bases, uint32_t world, uint32_t rank, uint32_t n_virtual, const unsigned char* program,
                                   uint64_t len, int use_small) {
    if (world < 2 || world > 8 || (world & (world - 1)) || rank >= world) return -1;
    uint32_t g = 0;
    while ((1u << g) < world) ++g;
    if (n_virtual < RT + g) return -1;
    if (use_small) {
        static RProgSmall small;
        if (!deserialise(program, len, &small)) return -2;
        run_grid_remote(bases, world, rank, n_virtual, small);
    } else {
        static RProgLarge large;
        if (!deserialise(program, len, &large)) return -2;
        run_grid_remote(bases, world, rank, n_virtual, large);
    }
    return 0;
}
