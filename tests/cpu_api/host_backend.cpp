// host_backend.cpp -- the engine's kernel-launcher layer (qb_common.h: k_*, run_fused, state files, sharding) on HOST memory,
// for tests/test_cpu_api.py.  Test infrastructure: a plain restatement of what the CUDA kernels compute (gather / matrix-vector /
// scatter per amplitude group with control masks, diagonal tables, probability sums, collapse), so that the REAL API translation
// units (qb_api.cu, qb_gate.cu) can be driven through the C ABI without a GPU.  Single rank only; sharded entry points fail.
#include <math.h>
#include <stdio.h>

#include <vector>

#include "qb_common.h"
#include "qb_shard.h"
#include "qb_tile.h"

namespace qb {

int k_init_basis(Reg* r, uint64_t local_index, bool owned) {
    const uint64_t n = 1ull << r->n_local;
    for (uint64_t i = 0; i < n; ++i) r->state[i] = make_double2(0.0, 0.0);
    if (owned) r->state[local_index] = make_double2(1.0, 0.0);
    r->st_launches++;
    return QB_OK;
}

int k_init_product(Reg* r, const double* q) {   // register.rs:81-102: amp[i] = q_{n-1}[bit] * (... * (q_0[bit] * 1))
    const uint64_t n = 1ull << r->n_local;
    for (uint64_t i = 0; i < n; ++i) {
        double re = 1.0, im = 0.0;
        for (uint32_t b = 0; b < r->n; ++b) {
            const double* f = q + 4 * b + 2 * ((i >> b) & 1);
            const double nr = f[0] * re - f[1] * im, ni = f[0] * im + f[1] * re;
            re = nr; im = ni;
        }
        r->state[i] = make_double2(re, im);
    }
    return QB_OK;
}

int k_apply_gate_unfused(Reg* r, const Gate& g) {
    const uint64_t n = 1ull << r->n_local, dim = 1ull << g.k;
    const uint64_t tmask = g.target_mask();
    std::vector<c64> in(dim), out(dim);
    for (uint64_t base = 0; base < n; ++base) {
        if (base & tmask) continue;
        if ((base & g.ctrl_mask) != g.ctrl_mask) continue;
        auto idx = [&](uint64_t c) {
            uint64_t o = base;
            for (uint32_t j = 0; j < g.k; ++j) if ((c >> j) & 1) o |= 1ull << g.targets[j];
            return o;
        };
        if (g.kind == GK_DIAG) {
            for (uint64_t c = 0; c < dim; ++c) {
                c64& a = r->state[idx(c)];
                const c64 d = g.matrix[c];
                a = make_double2(d.x * a.x - d.y * a.y, d.x * a.y + d.y * a.x);
            }
            continue;
        }
        for (uint64_t c = 0; c < dim; ++c) in[c] = r->state[idx(c)];
        for (uint64_t row = 0; row < dim; ++row) {
            double re = 0.0, im = 0.0;
            for (uint64_t c = 0; c < dim; ++c) {
                const c64 m = g.matrix[row * dim + c];
                re += m.x * in[c].x - m.y * in[c].y;
                im += m.x * in[c].y + m.y * in[c].x;
            }
            out[row] = make_double2(re, im);
        }
        for (uint64_t c = 0; c < dim; ++c) r->state[idx(c)] = out[c];
    }
    r->st_launches++;
    r->st_passes++;
    return QB_OK;
}

int k_probabilities(Reg* r, int target, double* p0, double* p1) {
    const uint64_t n = 1ull << r->n_local;
    double a = 0.0, b = 0.0;
    for (uint64_t i = 0; i < n; ++i) {
        const double p = r->state[i].x * r->state[i].x + r->state[i].y * r->state[i].y;
        if (target >= 0 && ((i >> target) & 1)) b += p; else a += p;
    }
    *p0 = a;
    *p1 = b;
    return QB_OK;
}

int k_collapse(Reg* r, uint32_t t, int outcome, double total_prob) {
    const uint64_t n = 1ull << r->n_local;
    const double s = sqrt(total_prob);
    for (uint64_t i = 0; i < n; ++i) {
        c64& a = r->state[i];
        if ((int)((i >> t) & 1) == outcome) a = make_double2(a.x / s, a.y / s);
        else a = make_double2(0.0, 0.0);
    }
    return QB_OK;
}

int k_scale_or_zero_all(Reg*, bool, double) { return set_error(QB_ERR_UNSUPPORTED, -1, "host stand-in: single rank only"); }

int k_diff(Reg* a, Reg* b, double* max_abs, double* l2sq) {
    const uint64_t n = 1ull << a->n_local;
    double m = 0.0, l = 0.0;
    for (uint64_t i = 0; i < n; ++i) {
        const double dx = a->state[i].x - b->state[i].x, dy = a->state[i].y - b->state[i].y;
        const double d2 = dx * dx + dy * dy;
        l += d2;
        if (sqrt(d2) > m) m = sqrt(d2);
    }
    *max_abs = m;
    *l2sq = l;
    return QB_OK;
}

int ensure_dmat(Reg*, size_t) { return QB_OK; }

// raw images without header or checksum: enough for the API layer's save / load plumbing
int state_save(Reg* r, const char* path) {
    FILE* f = fopen(path, "wb");
    if (!f) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "cannot open %s", path);
    const size_t n = (size_t)1 << r->n_local;
    const bool ok = fwrite(r->state, sizeof(c64), n, f) == n;
    fclose(f);
    return ok ? QB_OK : set_error(QB_ERR_INVALID_ARGUMENT, -1, "short write");
}
int state_load(Reg* r, const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "cannot open %s", path);
    const size_t n = (size_t)1 << r->n_local;
    const bool ok = fread(r->state, sizeof(c64), n, f) == n;
    fclose(f);
    return ok ? QB_OK : set_error(QB_ERR_INVALID_ARGUMENT, -1, "short read");
}

int drain_tile_events(Reg*) { return QB_OK; }

int run_fused(Reg* r, std::vector<Gate>& gates, std::vector<Gate>* hold) {   // program order (the planner only reorders commuting gates)
    size_t n = gates.size();
    if (hold && n > 8) {   // stand-in for "the last planned pass": the trailing gates come back to the queue
        n -= 5;
        for (size_t i = n; i < gates.size(); ++i) hold->push_back(gates[i]);
    }
    for (size_t i = 0; i < n; ++i) QB_TRY(k_apply_gate_unfused(r, gates[i]));
    return QB_OK;
}

int64_t plan_debug(uint32_t, std::vector<Gate>&, uint32_t, int32_t*, int32_t*) { return set_error(QB_ERR_UNSUPPORTED, -1, "host stand-in"); }
int64_t rounds_debug(uint32_t, std::vector<Gate>&, unsigned char*, uint64_t) { return set_error(QB_ERR_UNSUPPORTED, -1, "host stand-in"); }
int jit_drain() { return QB_OK; }
void jit_global_stats(int64_t* c, int64_t* f, int64_t* p, int64_t* ms) { *c = *f = *p = *ms = 0; }

int shard_unique_id(void*) { return set_error(QB_ERR_UNSUPPORTED, -1, "host stand-in: single rank only"); }
int shard_init(Reg*, const void*) { return set_error(QB_ERR_UNSUPPORTED, -1, "host stand-in: single rank only"); }
void shard_destroy(Reg*) {}
int shard_run(Reg*, std::vector<Gate>&) { return set_error(QB_ERR_UNSUPPORTED, -1, "host stand-in: single rank only"); }
int shard_probabilities(Reg*, uint32_t, double*, double*) { return set_error(QB_ERR_UNSUPPORTED, -1, "host stand-in"); }
int shard_collapse(Reg*, uint32_t, int, double) { return set_error(QB_ERR_UNSUPPORTED, -1, "host stand-in"); }
int shard_canonicalize(Reg*) { return QB_OK; }
void shard_reset_placement(Reg* r) { for (uint32_t q = 0; q < r->n; ++q) r->phys_of[q] = q; }
int shard_reduce_diff(Reg*, double*, double*) { return QB_OK; }
int64_t shard_plan_debug(uint32_t, int, int, std::vector<Gate>&, int32_t*, uint64_t, uint32_t*) { return set_error(QB_ERR_UNSUPPORTED, -1, "host stand-in"); }

}  // namespace qb

// the specialised-kernel entry points live in qb_jit.cu (not part of the host build)
extern "C" int64_t qb_jit_source_debug(const void*, uint64_t, int32_t, int32_t, int32_t, int32_t, int32_t, const uint8_t*, char*, uint64_t) { return qb::set_error(QB_ERR_UNSUPPORTED, -1, "host stand-in"); }
extern "C" int64_t qb_jit_compile_debug(const char*, double*, void*, uint64_t) { return qb::set_error(QB_ERR_UNSUPPORTED, -1, "host stand-in"); }
extern "C" int qb_jit_drain(void) { return QB_OK; }

namespace qb {
int state_load_check(Reg*, const char* path) {   // the host stand-in's files have no header: existence is all there is to check
    FILE* f = fopen(path, "rb");
    if (!f) return set_error(QB_ERR_IO, -1, "cannot open %s", path);
    fclose(f);
    return QB_OK;
}
}  // namespace qb

// in-process multi-GPU registers (qb_group.cu) do not exist in the host build: QB_NUM_GPUS is ignored
namespace qb {
Group* group_create(int) { return nullptr; }
void group_destroy(Group*) {}
Group* group_of(Reg*) { return nullptr; }
int group_world(const Group*) { return 1; }
Reg* group_member(Group*, int) { return nullptr; }
void group_set_member(Group*, int, Reg*) {}
void group_barrier(Group*) {}
int group_requested_world() { return 1; }
int group_run(Group*, const std::function<int(Reg*, int)>&) { return set_error(QB_ERR_UNSUPPORTED, -1, "host stand-in"); }
}  // namespace qb
