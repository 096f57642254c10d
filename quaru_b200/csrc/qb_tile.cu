// qb_tile.cu -- fused tile passes (placeholder: one kernel per gate until the tile kernel lands).
#include "qb_common.h"
namespace qb {
int run_fused(Reg* r, std::vector<Gate>& gates) {
    for (const Gate& g : gates) QB_TRY(k_apply_gate_unfused(r, g));
    return QB_OK;
}
int64_t plan_debug(uint32_t, std::vector<Gate>& gates, uint32_t, int32_t* pass_of_gate, int32_t* order_in_pass) {
    int64_t p = 0;
    for (const Gate& g : gates) { pass_of_gate[g.tag] = (int32_t)p++; order_in_pass[g.tag] = 0; }
    return p;
}
}  // namespace qb
