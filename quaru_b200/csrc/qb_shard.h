// qb_shard.h -- multi-GPU sharding of one register over the ranks of one NVLink/NVSwitch box.
// No reference equivalent (the reference is single-process); see DESIGN.md "Multi-GPU".
#pragma once
#include "qb_common.h"

namespace qb {
int shard_unique_id(void* out128);
int shard_init(Reg* r, const void* unique_id);
void shard_destroy(Reg* r);
int shard_run(Reg* r, std::vector<Gate>& gates);  // gates carry LOGICAL targets; maps, swaps, runs locally
int shard_probabilities(Reg* r, uint32_t logical_target /* UINT32_MAX: norm */, double* p0, double* p1);
int shard_collapse(Reg* r, uint32_t logical_target, int outcome, double total_prob);
int shard_canonicalize(Reg* r);  // restore the identity qubit placement (needed before raw state access)
void shard_reset_placement(Reg* r);  // identity placement, no data movement (the state is about to be overwritten)
int shard_reduce_diff(Reg* r, double* max_abs, double* l2sq);
int64_t shard_plan_debug(uint32_t n, int world, int rank, std::vector<Gate>& gates, int32_t* events, uint64_t capacity,
                         uint32_t* final_phys_of);
}  // namespace qb
