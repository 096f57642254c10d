// qb_tile.h -- shared between the two tile-pass kernels (qb_tile.cu: shared-memory interpreter, any tile size;
// qb_tile2.cu: register-blocked kernel for 2^12-amplitude tiles).
#pragma once
#include "qb_common.h"

namespace qb {
struct PlannedPass {
    std::vector<uint32_t> tile_pos;  // ascending global positions of the tile bits
    uint32_t lc = 0;
    std::vector<size_t> gate_idx;    // indices into the gate list, in execution order
    bool unfused = false;            // a single gate the tile kernel cannot take
};
bool tile_supported(const Gate& g);  // can go through a tile pass at all (dense k <= 2, diagonal k <= 3)
void plan_register_passes(const std::vector<Gate>& gates, uint32_t n, std::vector<PlannedPass>* out);
int tile_profile_begin(Reg* r);  // QB_OPT_PROFILE: CUDA-event bracket around one tile launch
int tile_profile_end(Reg* r);
int run_pass_v2(Reg* r, const std::vector<Gate>& gates, const std::vector<size_t>& gate_idx,
                const std::vector<uint32_t>& tile_pos, uint32_t lc, bool* handled);
int run_pass_remote(Reg* r, uint32_t n_virtual, const std::vector<Gate>& gates, const std::vector<size_t>& gate_idx,
                    const std::vector<uint32_t>& tile_pos, uint32_t lc, c64* const* bases, bool* handled);
// true if the register-blocked encoder takes the pass as it is (planners: a pass it refuses is cut in two sweeps at launch)
bool pass_encodes_v2(uint32_t n_local, const std::vector<Gate>& gates, const std::vector<size_t>& gate_idx,
                     const std::vector<uint32_t>& tile_pos, uint32_t lc);
int64_t encode_rounds_debug(uint32_t n_local, const std::vector<Gate>& gates, const std::vector<size_t>& gate_idx,
                            const std::vector<uint32_t>& tile_pos, uint32_t lc, unsigned char* out, uint64_t cap);
int64_t rounds_debug(uint32_t n, std::vector<Gate>& gates, unsigned char* out, uint64_t cap);
}  // namespace qb
