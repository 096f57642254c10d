// qb_tile.h -- shared between the two tile-pass kernels (qb_tile.cu: shared-memory interpreter, any tile size;
// qb_tile2.cu: register-blocked kernel for 2^12-amplitude tiles).
#pragma once
#include "qb_common.h"

namespace qb {
int tile_profile_begin(Reg* r);  // QB_OPT_PROFILE: CUDA-event bracket around one tile launch
int tile_profile_end(Reg* r);
int run_pass_v2(Reg* r, const std::vector<Gate>& gates, const std::vector<size_t>& gate_idx,
                const std::vector<uint32_t>& tile_pos, uint32_t lc, bool* handled);
int run_pass_remote(Reg* r, uint32_t n_virtual, const std::vector<Gate>& gates, const std::vector<size_t>& gate_idx,
                    const std::vector<uint32_t>& tile_pos, uint32_t lc, c64* const* bases, bool* handled);
int64_t encode_rounds_debug(uint32_t n_local, const std::vector<Gate>& gates, const std::vector<size_t>& gate_idx,
                            const std::vector<uint32_t>& tile_pos, uint32_t lc, unsigned char* out, uint64_t cap);
int64_t rounds_debug(uint32_t n, std::vector<Gate>& gates, unsigned char* out, uint64_t cap);
}  // namespace qb
