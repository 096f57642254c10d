// qb_shard.cu -- placeholder until the NCCL exchange lands (single-GPU build).
#include "qb_shard.h"
namespace qb {
int shard_unique_id(void*) { return set_error(QB_ERR_UNSUPPORTED, -1, "sharding not built yet"); }
int shard_init(Reg*, const void*) { return set_error(QB_ERR_UNSUPPORTED, -1, "sharding not built yet"); }
void shard_destroy(Reg*) {}
int shard_run(Reg*, std::vector<Gate>&) { return set_error(QB_ERR_UNSUPPORTED, -1, "sharding not built yet"); }
int shard_probabilities(Reg*, uint32_t, double*, double*) { return set_error(QB_ERR_UNSUPPORTED, -1, "sharding not built yet"); }
int shard_collapse(Reg*, uint32_t, int, double) { return set_error(QB_ERR_UNSUPPORTED, -1, "sharding not built yet"); }
int shard_canonicalize(Reg*) { return QB_OK; }
int shard_reduce_diff(Reg*, double*, double*) { return QB_OK; }
}  // namespace qb
