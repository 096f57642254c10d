// qb_jit.h -- K2j: per-pass specialised tile kernels.  The round program of a fused pass (qb_tile2_kernel.cuh: rounds, ops,
// coefficient tables) is translated into straight-line CUDA source -- register indices, coefficients, masks and shared-memory
// offsets are literals, X / CNOT / SWAP on register bits are compile-time renamings -- compiled for sm_100a with NVRTC on a
// background thread pool and cached by program hash.  Until a program's kernel is ready the interpreter kernel
// (k2r_tile_pass) runs the pass, so nothing ever waits for the compiler unless asked to (QB_OPT_JIT = 2).
#pragma once
#include <string>

#include "qb_common.h"
#include "qb_tile2_kernel.cuh"

namespace qb {

struct ProgView {   // RProgram<...> without its template parameters
    uint32_t lc, h, n_rounds, n_mats, n_ops;
    const uint8_t* hp;
    const uint64_t *gro_first, *gro_last;
    const RoundDesc* rounds;
    const ROp* ops;
    const c64* mats;
};
template <class Prog>
inline ProgView make_view(const Prog& p) {
    return ProgView{p.lc, p.h, p.n_rounds, p.n_mats, p.rounds[p.n_rounds - 1].op_end, p.hp, p.gro_first, p.gro_last, p.rounds, p.ops, p.mats};
}

struct JitOpts {
    int min_blocks = 2;   // __launch_bounds__(256, min_blocks): 2 -> 128 registers, 3 -> 80
    int persistent = 0;   // CTAs walk the tiles (grid = resident CTAs) instead of one CTA per tile
    int prefetch = 0;     // persistent only: prefetch.global.L2 of the next tile after the current tile's loads were issued
    int pipelined = 0;    // the prefetching ring: one persistent CTA of 512 threads per SM, two groups on alternate tiles, every thread
                          // fetches its amplitudes of the group's next tile with cp.async into a ring of three tile buffers
    int remote = 0;       // sharded registers: amplitudes addressed through the table of peer-mapped shards
    // sharded registers, the exchanging epilogue: the pass stores every amplitude where it lives after the exchange of local
    // bits xchg_V[i] with rank bits xchg_J[i], into the destination rank's other state buffer (0 = an ordinary pass)
    int xchg_bits = 0;
    uint8_t xchg_V[3] = {0, 0, 0}, xchg_J[3] = {0, 0, 0};
};

// what the launch of an exchanging pass needs beside the program
struct XchgArgs {
    int gp = 0;
    uint8_t V[3] = {0, 0, 0}, J[3] = {0, 0, 0};   // ascending V
    c64* out[8] = {nullptr};                      // every rank's other buffer (out[own rank] = own other buffer)
    uint32_t rank = 0;
};

std::string jit_generate(const ProgView& p, const JitOpts& o);

// Launch path (qb_tile2.cu).  *launched = true if a specialised kernel ran the pass; false: the caller runs the interpreter
// kernel (compile pending, JIT off, or unavailable).  With r->jit == 2 the call waits for the compiler and a failure is an error.
int jit_try_launch(Reg* r, const ProgView& p, const RemoteArgs* remote, bool* launched, const XchgArgs* xchg = nullptr);

}  // namespace qb
