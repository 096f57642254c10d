// qb_tile2.cu -- K2r: register-blocked fused tile pass (the fast path for tiles of 2^12 amplitudes).
//
// ncu on the first tile kernel (qb_tile.cu, profiles/r1_ncu_tile_kernel.csv) showed it instruction-issue
// bound: every gate was one shared-memory round trip with per-amplitude index arithmetic (~57 instructions
// per amplitude per gate).  Here a CTA still owns one 2^12-amplitude tile per pass, but each of its 256
// threads keeps 16 amplitudes in REGISTERS, addressed by 4 "register bits" R of the tile index; the other 8
// tile bits are the thread id.  A pass is a sequence of ROUNDS:
//   * round 0 loads straight from HBM into registers (R within tile bits 5..11, so a warp still reads 512
//     contiguous bytes per instruction), the last round stores straight back: a pass whose dense gates touch
//     at most 4 tile bits never touches shared memory at all;
//   * between rounds the registers are transposed through a swizzled shared-memory tile (conflict-free
//     128-bit accesses) into a new choice of R;
//   * inside a round every gate whose dense target bits are in R is applied to the registers with compile-time
//     register indices (no index arithmetic per amplitude); controls may be register bits, thread bits or
//     bits outside the tile; diagonal gates whose bits are all thread/outer bits cost ONE complex multiply per
//     thread: they accumulate into a per-thread phase that is applied once at the end of the round.
// Special cases with fewer FP64 instructions: real matrices (H), X (register swap), phase -1 (sign flip).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>

#include "qb_common.h"
#include "qb_device.cuh"
#include "qb_jit.h"
#include "qb_tile.h"
#include "qb_tile2_kernel.cuh"

namespace qb {

// ================================================================================================
// host: split a pass into rounds and encode it
// ================================================================================================
struct LocalOp {
    const Gate* g;
    uint32_t dense;  // tile-local bits that are dense targets
    uint32_t reads;  // tile-local bits that are controls / diagonal index bits
};

static bool is_real(const Gate& g) {
    for (const c64& e : g.matrix)
        if (e.y != 0.0) return false;
    return true;
}
static bool is_scaled_hadamard(const Gate& g) {   // s * [[1, 1], [1, -1]] with real s (operation::hadamard, operation.rs:350)
    if (g.k != 1 || !is_real(g)) return false;
    const double s = g.matrix[0].x;
    return s != 0.0 && g.matrix[1].x == s && g.matrix[2].x == s && g.matrix[3].x == -s;
}
static bool is_x(const Gate& g) {
    return g.k == 1 && g.matrix[0].x == 0.0 && g.matrix[0].y == 0.0 && g.matrix[3].x == 0.0 && g.matrix[3].y == 0.0 &&
           g.matrix[1].x == 1.0 && g.matrix[1].y == 0.0 && g.matrix[2].x == 1.0 && g.matrix[2].y == 0.0;
}

// `mode`: the round-selection strategy (0 plain scan, 1 low tile bits first in middle rounds, 2 roll-out search); encode_rounds
// tries all three and keeps the shortest program.
template <class Prog>
static bool encode_rounds_with(const std::vector<Gate>& gates, const std::vector<size_t>& gate_idx, const std::vector<uint32_t>& tile_pos,
                               uint32_t lc, Prog* prog, int mode) {
    const size_t max_sel = 4;   // selector bits of a merged per-thread phase table
    const uint32_t T = (uint32_t)tile_pos.size();
    if (T != RT) return false;
    int8_t local_of[64];
    uint64_t tile_mask = 0;
    for (int b = 0; b < 64; ++b) local_of[b] = -1;
    for (uint32_t j = 0; j < T; ++j) { local_of[tile_pos[j]] = (int8_t)j; tile_mask |= 1ull << tile_pos[j]; }
    prog->lc = lc;
    prog->h = T - lc;
    for (uint32_t j = 0; j < prog->h; ++j) prog->hp[j] = (uint8_t)tile_pos[lc + j];

    // outer (non-tile) global bit -> its rank among the non-tile bits = its bit position in blockIdx.x
    int8_t outer_ix[64];
    {
        int nx = 0;
        for (int b = 0; b < 64; ++b) outer_ix[b] = ((tile_mask >> b) & 1) ? (int8_t)-1 : (int8_t)nx++;
    }
    auto to_tile_space = [&](uint64_t gmask) {
        uint64_t o = 0;
        for (uint64_t m = gmask; m; m &= m - 1) o |= 1ull << outer_ix[__builtin_ctzll(m)];
        return o;
    };
    std::vector<LocalOp> ops;
    for (size_t gi : gate_idx) {
        const Gate& g = gates[gi];
        LocalOp lo{&g, 0, 0};
        for (uint64_t m = g.ctrl_mask & tile_mask; m; m &= m - 1) lo.reads |= 1u << local_of[__builtin_ctzll(m)];
        for (uint32_t t : g.targets) {
            if (local_of[t] < 0) { if (g.kind == GK_DENSE) return false; continue; }
            if (g.kind == GK_DENSE) lo.dense |= 1u << local_of[t]; else lo.reads |= 1u << local_of[t];
        }
        ops.push_back(lo);
    }
    const size_t max_ops = sizeof(prog->ops) / sizeof(ROp), max_mats = sizeof(prog->mats) / sizeof(c64);
    const size_t max_rounds = sizeof(prog->rounds) / sizeof(RoundDesc);
    std::vector<char> done(ops.size(), 0);
    size_t remaining = ops.size(), n_ops = 0, n_mats = 4, n_rounds = 0;
    prog->mats[0] = prog->mats[3] = make_double2(1.0, 0.0);
    prog->mats[1] = prog->mats[2] = make_double2(0.0, 0.0);
    // register bits of the HBM-facing rounds: tile bits 5..11, so that a warp's lanes are tile bits 0..4 = one contiguous 512-byte
    // segment per access.  Experimental (QB_TILE_IO_LOW=4 or 3, not yet measured): bits 4 / 3 may be register bits there too -- a
    // warp access then covers two 256-byte / four 128-byte pieces (still whole lines), and fewer passes need an extra round whose only
    // purpose is the transpose into a storable layout (bench, 3 steps at 33 qubits: 201 -> 193 / 189 rounds)
    const char* io_env = getenv("QB_TILE_IO_LOW");
    const uint32_t io_low = (io_env && (io_env[0] == '3' || io_env[0] == '4') && !io_env[1]) ? (uint32_t)(io_env[0] - '0') : 5u;
    const uint32_t io_allowed = 0xfffu & ~((1u << io_low) - 1u);
    // Unconditional constant factors (the deferred 1/sqrt(2) of every Hadamard, the common factor d0 of a one-qubit diagonal
    // diag(d0, d1) on a register bit, global phases) commute with everything: they are multiplied together on the host
    // over the whole pass and applied once, in its last round, instead of costing a table op and 16 complex multiplies
    // in every round.
    c64 pass_scalar = make_double2(1.0, 0.0);
    auto scale_pass = [&](c64 f) { pass_scalar = make_double2(pass_scalar.x * f.x - pass_scalar.y * f.y, pass_scalar.x * f.y + pass_scalar.y * f.x); };
    auto emit_round = [&](uint32_t R, const std::vector<size_t>& members) -> bool {
        if (n_rounds >= max_rounds) return false;
        RoundDesc& rd = prog->rounds[n_rounds];
        memset(&rd, 0, sizeof(rd));
        uint32_t ri = 0, reg_of[12];
        for (uint32_t b = 0; b < RT; ++b) reg_of[b] = 99;
        for (uint32_t b = 0; b < RT; ++b)
            if ((R >> b) & 1) { rd.R[ri] = (uint8_t)b; reg_of[b] = ri++; }
        // thread bits: lanes 0..2 must hit distinct (bit mod 3) classes for conflict-free swizzled accesses;
        // the HBM-facing rounds have R within 5..11, so taking the free bits in ascending order gives lanes = bits 0..4.
        std::vector<uint32_t> free_bits;
        for (uint32_t b = 0; b < RT; ++b)
            if (!((R >> b) & 1)) free_bits.push_back(b);
        std::vector<uint32_t> order;
        bool cls[3] = {false, false, false};
        for (uint32_t b : free_bits)
            if (order.size() < 3 && !cls[b % 3]) { order.push_back(b); cls[b % 3] = true; }
        for (uint32_t b : free_bits)
            if (std::find(order.begin(), order.end(), b) == order.end()) order.push_back(b);
        if ((R & ~io_allowed) == 0) order = free_bits;  // bits 0,1,2 first anyway (distinct classes) and lanes = 0..4
        for (int i = 0; i < 8; ++i) rd.tb[i] = (uint8_t)order[i];
        for (uint32_t j = 0; j < 16; ++j) {
            uint32_t l = 0;
            for (uint32_t i = 0; i < 4; ++i)
                if ((j >> i) & 1) l |= 1u << rd.R[i];
            rd.sro[j] = (uint16_t)(l ^ (((l >> 3) ^ (l >> 6) ^ (l >> 9)) & 7u));
        }
        rd.op_begin = (uint16_t)n_ops;
        // the round's ops before merging: unbounded here, the capacity of the program is checked on what is left after
        // the three merging passes below (a round of 150 one-qubit phases shrinks to a handful of table ops)
        std::vector<ROp> raw;
        for (size_t oi : members) {
            const LocalOp& lo = ops[oi];
            const Gate& g = *lo.g;
            if (n_mats + g.matrix.size() > max_mats) return false;
            ROp op;
            memset(&op, 0, sizeof(op));
            op.mat = (uint32_t)n_mats;
            op.omask = to_tile_space(g.ctrl_mask & ~tile_mask);
            for (uint64_t m = g.ctrl_mask & tile_mask; m; m &= m - 1) {
                const uint32_t lb = (uint32_t)local_of[__builtin_ctzll(m)];
                if (reg_of[lb] != 99) op.rmask |= (uint16_t)(1u << reg_of[lb]); else op.tmask |= (uint16_t)(1u << lb);
            }
            std::vector<c64> mat = g.matrix;
            if (g.kind == GK_DENSE) {
                if (g.k == 1) {
                    op.a = (uint8_t)reg_of[local_of[g.targets[0]]];
                    const uint8_t pred = op.rmask ? 16 : 0;
                    op.code = (uint8_t)((is_real(g) ? RC_D1R : RC_D1C) + pred + op.a);
                    if (is_scaled_hadamard(g) && !op.rmask && !op.tmask && !op.omask) { op.code = (uint8_t)(RC_H + op.a); rd.has_tph = 1; }
                } else {
                    op.code = RC_D2;
                    op.a = (uint8_t)reg_of[local_of[g.targets[0]]];
                    op.b = (uint8_t)reg_of[local_of[g.targets[1]]];
                }
            } else if (g.k == 0) {
                if (op.rmask == 0) { op.code = RC_PH_T; rd.has_tph = 1; }
                else op.code = (mat[0].x == -1.0 && mat[0].y == 0.0) ? RC_PH_NEG : RC_PH_R;
            } else {
                op.code = RC_DG;
                op.k = (uint8_t)g.k;
                bool any_reg = false;
                for (uint32_t i = 0; i < g.k; ++i) {
                    const uint32_t t = g.targets[i];
                    if (local_of[t] < 0) op.src[i] = (uint8_t)(0x80 | outer_ix[t]);
                    else if (reg_of[local_of[t]] != 99) { op.src[i] = (uint8_t)reg_of[local_of[t]]; any_reg = true; }
                    else op.src[i] = (uint8_t)(0x40 | local_of[t]);
                }
                if (!any_reg && op.rmask == 0) rd.has_tph = 1;
                if (g.k == 1 && any_reg) { op.a = op.src[0]; op.code = (uint8_t)((op.rmask ? RC_DG1R_P : RC_DG1R) + op.a); }
                if (g.k == 1 && !any_reg && op.rmask == 0) op.code = RC_DG1T;
                // diag(d0, d1) on a register bit = d0 on the whole thread (thread-uniform: merges into the phase tables)
                // times the relative phase d1/d0 on the 8 registers with the bit set (joins the phase chains of that
                // bit): half the register traffic of multiplying all 16 registers.
                if (op.code >= RC_DG1R && op.code < RC_DG1R + 4 && (mat[0].x != 0.0 || mat[0].y != 0.0)) {
                    const c64 d0 = mat[0], d1 = mat[1];
                    const double den = d0.x * d0.x + d0.y * d0.y;
                    const c64 rel = make_double2((d1.x * d0.x + d1.y * d0.y) / den, (d1.y * d0.x - d1.x * d0.y) / den);
                    ROp u = op;               // thread-uniform part, same thread / outer controls
                    u.code = RC_PH_T;
                    u.a = 0; u.k = 0; u.src[0] = u.src[1] = u.src[2] = 0;
                    u.mat = (uint32_t)n_mats;
                    prog->mats[n_mats++] = d0;
                    raw.push_back(u);
                    rd.has_tph = 1;
                    ROp r2 = u;               // relative phase on the registers with the bit set
                    r2.rmask = (uint16_t)(1u << op.a);
                    r2.code = (rel.x == -1.0 && rel.y == 0.0) ? RC_PH_NEG : RC_PH_R;
                    r2.mat = (uint32_t)n_mats;
                    prog->mats[n_mats++] = rel;
                    raw.push_back(r2);
                    continue;
                }
            }
            memcpy(prog->mats + n_mats, mat.data(), mat.size() * sizeof(c64));
            n_mats += mat.size();
            raw.push_back(op);
        }
        // Chains of phases whose only register-bit control is the same single bit b (the controlled-phase cascade
        // of a QFT: CP(c_i, t) for many c_i) accumulate into one per-thread factor and touch the registers once.
        // Diagonal ops commute, so only an op that WRITES register bit b ends a chain.
        {
            std::vector<ROp>& in = raw;
            std::vector<ROp> out;
            auto single_bit = [](const ROp& o) -> int {
                if ((o.code != RC_PH_R && o.code != RC_PH_NEG) || __builtin_popcount(o.rmask) != 1) return -1;
                return __builtin_ctz(o.rmask);
            };
            auto writes_bit = [](const ROp& o, int b) {
                const int c = o.code;
                if (c == RC_D2) return o.a == b || o.b == b;
                const bool d1 = c < RC_D2 || (c >= RC_H && c < RC_H + 4);
                return d1 && (c & 3) == b && !(c >= RC_DG1R && c < RC_DG1R + 4) && !(c >= RC_DG1R_P && c < RC_DG1R_P + 4);
            };
            // Diagonal ops commute with everything that does not write one of their bits, and nobody writes thread bits
            // inside a round: a phase on register bit b may be deferred up to the next op that writes b.  So the phases
            // of each bit are collected (however they interleave with other bits' phases and with dense gates elsewhere)
            // and emitted as one group right before that writer, or at the end of the round.  The accumulator is the
            // single per-thread `tph`, hence one group at a time: ACC ... ACC FLUSH.
            // Groups of fewer than three phases do not pay for a flush: they stay where they are (deferring them would
            // also separate dense gates that the layer packing below wants adjacent).
            // The UNCONDITIONAL phases of a group (no thread / outer control: the relative phase of an RZ, a Z) are first
            // multiplied together on the host into the group's first such op; the others are dropped.
            std::vector<int> group_of(in.size(), -1);
            std::vector<size_t> group_size;
            std::vector<int> group_rep;   // per group: index of the op that carries the product of the unconditional phases
            {
                int open_group[4] = {-1, -1, -1, -1};
                for (size_t i = 0; i < in.size(); ++i) {
                    const int b = single_bit(in[i]);
                    if (b >= 0) {
                        if (open_group[b] < 0) { open_group[b] = (int)group_size.size(); group_size.push_back(0); group_rep.push_back(-1); }
                        const int gid = open_group[b];
                        if (in[i].tmask == 0 && in[i].omask == 0) {
                            if (group_rep[gid] < 0) {
                                group_rep[gid] = (int)i;
                            } else {   // fold into the representative and drop
                                c64& acc = prog->mats[in[group_rep[gid]].mat];
                                const c64 f = prog->mats[in[i].mat];
                                acc = make_double2(acc.x * f.x - acc.y * f.y, acc.x * f.y + acc.y * f.x);
                                in[group_rep[gid]].code = RC_PH_R;   // the product is a general phase (maybe -1 again: fixed below)
                                group_of[i] = -2;                    // dropped
                                continue;
                            }
                        }
                        group_of[i] = gid;
                        group_size[gid]++;
                        continue;
                    }
                    for (int wb = 0; wb < 4; ++wb)
                        if (writes_bit(in[i], wb)) open_group[wb] = -1;
                }
            }
            std::vector<ROp> pending[4];
            auto emit_group = [&](int b) {
                std::vector<ROp>& grp = pending[b];
                if (grp.empty()) return;
                for (ROp a : grp) {
                    a.code = RC_PH_ACC;   // the -1 of RC_PH_NEG is already the coefficient at `mat`
                    a.rmask = 0;
                    out.push_back(a);
                }
                ROp f;
                memset(&f, 0, sizeof(f));
                f.code = (uint8_t)(RC_PH_FLUSH + b);
                out.push_back(f);
                grp.clear();
            };
            for (size_t i = 0; i < in.size(); ++i) {
                const int b = single_bit(in[i]);
                if (b >= 0) {
                    if (group_of[i] == -2) continue;   // folded into its group's representative
                    if (in[i].tmask == 0 && in[i].omask == 0) {   // a representative: normalise the product
                        const c64 f = prog->mats[in[i].mat];
                        if (f.x == 1.0 && f.y == 0.0 && group_size[group_of[i]] < 3) continue;   // the phases cancelled
                        if (f.y == 0.0 && f.x == -1.0) in[i].code = RC_PH_NEG;
                    }
                    if (group_size[group_of[i]] >= 3) pending[b].push_back(in[i]); else out.push_back(in[i]);
                    continue;
                }
                for (int wb = 0; wb < 4; ++wb)
                    if (writes_bit(in[i], wb)) emit_group(wb);
                out.push_back(in[i]);
            }
            for (int b = 0; b < 4; ++b) emit_group(b);
            // Second pass: every thread-uniform phase (RC_PH_T, RC_DG1T, RC_PH_ACC) is a function of a few thread / outer
            // bits.  They all commute, so they are multiplied together on the host into tables over <= 4 selector bits:
            // one table op replaces several ops (each op costs ~25 instructions of decode for 4 FP64 instructions of work).
            struct PhaseFn { std::vector<uint8_t> sel; std::vector<c64> tab; };
            auto to_phase = [&](const ROp& o, PhaseFn* f) -> bool {
                if (o.code != RC_PH_T && o.code != RC_PH_ACC && o.code != RC_DG1T) return false;
                std::vector<uint8_t> ctl;
                for (uint32_t m = o.tmask; m; m &= m - 1) ctl.push_back((uint8_t)(0x40 | __builtin_ctz(m)));
                for (uint64_t m = o.omask; m; m &= m - 1) ctl.push_back((uint8_t)(0x80 | __builtin_ctzll(m)));
                f->sel = ctl;
                int dsel = -1;
                if (o.code == RC_DG1T) {
                    for (size_t i = 0; i < ctl.size(); ++i) if (ctl[i] == o.src[0]) return false;  // cannot happen (validated targets)
                    dsel = (int)f->sel.size();
                    f->sel.push_back(o.src[0]);
                }
                if (f->sel.size() > 4) return false;   // a single phase never has more (k <= 3 targets + a split-off bit)
                const size_t nctl = ctl.size();
                f->tab.assign(1u << f->sel.size(), make_double2(1.0, 0.0));
                for (uint32_t x = 0; x < f->tab.size(); ++x) {
                    if ((x & ((1u << nctl) - 1u)) != ((1u << nctl) - 1u)) continue;   // a control is 0: factor 1
                    f->tab[x] = dsel < 0 ? prog->mats[o.mat] : prog->mats[o.mat + ((x >> dsel) & 1u)];
                }
                return true;
            };
            auto merge_into = [&](std::vector<PhaseFn>* tabs, const PhaseFn& f) {
                for (PhaseFn& t : *tabs) {
                    std::vector<uint8_t> uni = t.sel;
                    for (uint8_t sb : f.sel) if (std::find(uni.begin(), uni.end(), sb) == uni.end()) uni.push_back(sb);
                    if (uni.size() > max_sel) continue;
                    std::vector<c64> nt(1u << uni.size());
                    auto sub = [&](const PhaseFn& g, uint32_t x) {
                        uint32_t ix = 0;
                        for (size_t i = 0; i < g.sel.size(); ++i) {
                            const size_t pos = std::find(uni.begin(), uni.end(), g.sel[i]) - uni.begin();
                            ix |= ((x >> pos) & 1u) << i;
                        }
                        return g.tab[ix];
                    };
                    for (uint32_t x = 0; x < nt.size(); ++x) {
                        const c64 a = sub(t, x), b = sub(f, x);
                        nt[x] = make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
                    }
                    t.sel = uni;
                    t.tab = nt;
                    return;
                }
                tabs->push_back(f);
            };
            bool overflow = false;
            auto emit_tabs = [&](std::vector<PhaseFn>* tabs, uint8_t code, std::vector<ROp>* dst) {
                for (const PhaseFn& t : *tabs) {
                    if (n_mats + t.tab.size() > max_mats) { overflow = true; return; }
                    ROp o;
                    memset(&o, 0, sizeof(o));
                    o.code = code;
                    o.k = (uint8_t)t.sel.size();
                    for (size_t i = 0; i < t.sel.size(); ++i) {
                        if (i < 3) o.src[i] = t.sel[i];
                        else o.b = t.sel[i];
                    }
                    o.mat = (uint32_t)n_mats;
                    memcpy(prog->mats + n_mats, t.tab.data(), t.tab.size() * sizeof(c64));
                    n_mats += t.tab.size();
                    dst->push_back(o);
                }
                tabs->clear();
            };
            std::vector<ROp> merged, tail;   // tail: unmerged writers of tph, moved behind the last chain flush (they commute)
            std::vector<PhaseFn> tph_tabs, acc_tabs;
            double h_scalar = 1.0;
            bool any_h = false;
            for (const ROp& o : out) {
                PhaseFn f;
                if (o.code >= RC_PH_FLUSH && o.code < RC_PH_FLUSH + 4) { emit_tabs(&acc_tabs, RC_TAB_ACC, &merged); merged.push_back(o); continue; }
                if (o.code >= RC_H && o.code < RC_H + 4) { h_scalar *= prog->mats[o.mat].x; any_h = true; merged.push_back(o); continue; }
                if (!to_phase(o, &f)) {
                    const bool writes_tph = o.code == RC_PH_T || o.code == RC_DG1T ||
                                            (o.code == RC_DG && o.rmask == 0 && !(o.src[0] < 4 && o.k > 0) && !(o.k > 1 && o.src[1] < 4) && !(o.k > 2 && o.src[2] < 4));
                    (writes_tph ? tail : merged).push_back(o);
                    continue;
                }
                if (o.code != RC_PH_ACC && f.sel.empty()) { scale_pass(f.tab[0]); continue; }   // a constant: deferred to the end of the pass
                merge_into(o.code == RC_PH_ACC ? &acc_tabs : &tph_tabs, f);
            }
            if (any_h) scale_pass(make_double2(h_scalar, 0.0));   // the deferred 1/sqrt(2) factors of the round's Hadamards
            // the per-thread phase is applied at the end of the round only if something is left in it then: chains flush
            // and reset it, and every other writer has been moved behind the last flush (`tail`, the merged tables)
            rd.has_tph = (!tph_tabs.empty() || !tail.empty()) ? 1 : 0;
            for (const ROp& o : tail) merged.push_back(o);
            emit_tabs(&tph_tabs, RC_TAB, &merged);
            // Third pass: bundle real one-qubit gates (H, X, real rotations; no register-bit or outer control) on DISTINCT
            // register bits into layer ops.  A gate may be hoisted into the open layer past ops that do not touch its
            // register bit (it writes only that bit; nobody writes thread bits inside a round).
            {
                auto reg_bits = [](const ROp& o) -> uint32_t {
                    const int c = o.code;
                    if (c < RC_D2) return (1u << (c & 3)) | o.rmask;                  // D1C / D1R / DG1R and predicated forms
                    if (c == RC_D2) return (1u << o.a) | (1u << o.b) | o.rmask;
                    if (c == RC_PH_NEG || c == RC_PH_R) return o.rmask;
                    if (c >= RC_PH_FLUSH && c < RC_PH_FLUSH + 4) return 1u << (c - RC_PH_FLUSH);
                    if (c >= RC_H && c < RC_H + 4) return 1u << (c - RC_H);
                    if (c == RC_DG) {
                        uint32_t m = o.rmask;
                        for (int i = 0; i < o.k && i < 3; ++i) if (o.src[i] < 4) m |= 1u << o.src[i];
                        return m;
                    }
                    return 0;   // PH_T, DG1T, TAB, TAB_ACC, PH_ACC: per-thread phase only
                };
                std::vector<ROp> packed;
                int open = -1;            // index in `packed` of the open layer
                uint32_t touched = 0;     // register bits touched by ops appended after the open layer
                struct Slot { bool used, h; uint16_t tm; c64 c0, c1; };
                Slot slots[4];
                auto close_layer = [&]() -> bool {
                    if (open < 0) return true;
                    int cnt = 0;
                    for (int b = 0; b < 4; ++b) cnt += slots[b].used;
                    ROp& L = packed[open];
                    if (cnt >= 2) {
                        if (n_mats + 8 > max_mats) return false;
                        const uint32_t mat = (uint32_t)n_mats;
                        uint32_t present = 0, hmask = 0;
                        uint16_t tm[4] = {0, 0, 0, 0};
                        for (int b = 0; b < 4; ++b) {
                            prog->mats[n_mats + 2 * b] = slots[b].c0;
                            prog->mats[n_mats + 2 * b + 1] = slots[b].c1;
                            if (!slots[b].used) continue;
                            present |= 1u << b;
                            if (slots[b].h) hmask |= 1u << b;
                            tm[b] = slots[b].tm;
                        }
                        n_mats += 8;
                        memset(&L, 0, sizeof(L));
                        L.code = RC_LAYER;
                        L.k = (uint8_t)(present | (hmask << 4));
                        L.mat = mat;
                        L.tmask = tm[0];
                        L.rmask = tm[1];
                        L.src[0] = (uint8_t)(tm[2] & 0xff); L.src[1] = (uint8_t)(tm[2] >> 8);
                        L.src[2] = (uint8_t)(tm[3] & 0xff); L.b = (uint8_t)(tm[3] >> 8);
                    }   // a layer with a single gate stays the plain op it was created from
                    open = -1;
                    return true;
                };
                for (const ROp& o : merged) {
                    const int c = o.code;
                    const bool is_h = c >= RC_H && c < RC_H + 4;
                    const bool layerable = (is_h || (c >= RC_D1R && c < RC_D1R + 4)) && o.omask == 0;
                    if (!layerable) {
                        packed.push_back(o);
                        if (open >= 0) touched |= reg_bits(o);
                        continue;
                    }
                    const int bit = c & 3;
                    if (open >= 0 && !slots[bit].used && !((touched >> bit) & 1u)) {
                        // joins the open layer
                    } else {
                        if (!close_layer()) return false;
                        open = (int)packed.size();
                        packed.push_back(o);       // placeholder: rewritten by close_layer if the layer gets >= 2 gates
                        touched = 0;
                        for (int b = 0; b < 4; ++b) slots[b] = Slot{false, false, 0, make_double2(1.0, 0.0), make_double2(0.0, 1.0)};
                    }
                    slots[bit].used = true;
                    slots[bit].h = is_h;
                    slots[bit].tm = o.tmask;
                    if (!is_h) {
                        slots[bit].c0 = make_double2(prog->mats[o.mat + 0].x, prog->mats[o.mat + 1].x);   // p, q
                        slots[bit].c1 = make_double2(prog->mats[o.mat + 2].x, prog->mats[o.mat + 3].x);   // r, s
                    }
                }
                if (!close_layer()) return false;
                merged.swap(packed);
            }
            if (overflow || !acc_tabs.empty()) return false;
            if (rd.op_begin + merged.size() > max_ops) return false;
            n_ops = rd.op_begin;
            for (const ROp& o : merged) prog->ops[n_ops++] = o;
        }
        rd.op_end = (uint16_t)n_ops;
        ++n_rounds;
        return true;
    };

    // ---- choosing the rounds -------------------------------------------------------------------------------------
    // One round = one scan of the remaining ops in program order with the same read / write commutation rule as the pass
    // planner; `preset` reserves register bits in advance.  Returns the register bits used (before filling up to four).
    auto pick_round = [&](std::vector<char>& dn, uint32_t preset, bool first_round, std::vector<size_t>* members) -> uint32_t {
        uint32_t R = preset, blocked_w = 0, blocked_r = 0;
        const uint32_t allowed = first_round ? io_allowed : 0xfffu;
        for (size_t i = 0; i < ops.size(); ++i) {
            if (dn[i]) continue;
            const LocalOp& lo = ops[i];
            bool ok = !(lo.dense & (blocked_w | blocked_r)) && !(lo.reads & blocked_w);
            if (ok && (lo.dense & ~allowed)) ok = false;
            if (ok && __builtin_popcount(R | lo.dense) > 4) ok = false;
            if (!ok) { blocked_w |= lo.dense; blocked_r |= lo.reads; continue; }
            R |= lo.dense;
            dn[i] = 1;
            if (members) members->push_back(i);
        }
        return R;
    };
    // fill R up to 4 bits with free bits, highest first (keeps HBM-facing rounds within bits 5..11)
    // (`avoid`: bits read by the round's diagonal ops -- as thread bits their phases merge into the per-thread tables)
    auto fill_R = [&](uint32_t R, bool first_round, uint32_t avoid = 0) {
        const uint32_t allowed = first_round ? io_allowed : 0xfffu;
        for (int pref = 0; pref < 2; ++pref)
            for (int b = RT - 1; b >= 0 && __builtin_popcount(R) < 4; --b)
                if (!((R >> b) & 1) && ((allowed >> b) & 1) && (pref == 1 || !((avoid >> b) & 1))) R |= 1u << b;
        return R;
    };
    // low tile bits (0..4) that are ready: they can never be register bits of the HBM-facing first and last rounds, so
    // serving them early tends to leave the high bits for a last round that can store straight to HBM
    auto low_seed = [&](const std::vector<char>& dn) {
        uint32_t R = 0, bw = 0, br = 0;
        for (size_t i = 0; i < ops.size(); ++i) {
            if (dn[i]) continue;
            const LocalOp& lo = ops[i];
            bool ok = !(lo.dense & (bw | br)) && !(lo.reads & bw);
            if (ok && (lo.dense & io_allowed)) ok = false;
            if (ok && __builtin_popcount(R | lo.dense) > 4) ok = false;
            if (!ok) { bw |= lo.dense; br |= lo.reads; continue; }
            R |= lo.dense;
        }
        return R;
    };
    // rounds the plain scan would still need from this state on, including the empty round that has to be appended when
    // the last computing round holds low bits
    auto rollout = [&](std::vector<char> dn, size_t rem, bool first_round, uint32_t last) {
        int rounds = 0;
        while (rem) {
            std::vector<size_t> m;
            const uint32_t R = pick_round(dn, 0, first_round, &m);
            rem -= m.size();
            last = fill_R(R, first_round);
            first_round = false;
            ++rounds;
        }
        if (last == 0 || (last & ~io_allowed)) ++rounds;
        return rounds;
    };

    bool first = true;
    uint32_t last_R = 0;
    while (remaining) {
        uint32_t preset = 0;
        if (mode == 1 && !first) preset = low_seed(done);
        if (mode == 2) {
            // roll-out policy: every candidate choice of register bits is completed with the plain scan; the choice with
            // the fewest rounds overall wins (ties: the one that takes more ops now)
            const uint32_t allowed = first ? io_allowed : 0xfffu;
            uint32_t pending = 0;
            for (size_t i = 0; i < ops.size(); ++i)
                if (!done[i]) pending |= ops[i].dense;
            std::vector<uint32_t> cands{0u};
            if (!first) cands.push_back(low_seed(done));
            {
                std::vector<char> dn = done;
                const uint32_t g0 = pick_round(dn, 0, first, nullptr);
                for (uint32_t out_b = g0; out_b; out_b &= out_b - 1)
                    for (uint32_t in_b = pending & allowed & ~g0; in_b; in_b &= in_b - 1)
                        cands.push_back((g0 & ~(out_b & -out_b)) | (in_b & -in_b));
            }
            int best_rounds = 1 << 30;
            size_t best_members = 0;
            for (uint32_t c : cands) {
                if (__builtin_popcount(c) > 4 || (c & ~allowed)) continue;
                std::vector<char> dn = done;
                std::vector<size_t> m;
                const uint32_t R = fill_R(pick_round(dn, c, first, &m), first);
                if (m.empty()) continue;
                const int total = 1 + rollout(dn, remaining - m.size(), false, R);
                if (total < best_rounds || (total == best_rounds && m.size() > best_members)) {
                    best_rounds = total;
                    best_members = m.size();
                    preset = c;
                }
            }
        }
        std::vector<size_t> members;
        uint32_t R = pick_round(done, preset, first, &members);
        remaining -= members.size();
        // A diagonal op costs 8-32 FP64 instructions plus an op slot per thread when one of its bits is a register bit of the
        // round, and next to nothing (a factor in a merged per-thread table) when they are all thread / outer bits.  It may run
        // in any later round as long as no gate after it writes one of its bits, so it is handed on to the next round when
        // that cannot delay anybody: there is a next round (dense gates are left, or this round holds low bits and an
        // HBM-facing round follows anyway) and no later gate of this round or of the rest writes its bits.
        {
            uint32_t diag_reads = 0;
            for (size_t i : members)
                if (!ops[i].dense) diag_reads |= ops[i].reads;
            // when an HBM-facing round has to follow anyway (this one holds low bits), the free slots go to low bits: the
            // high bits then stay thread bits and phases on them are free
            R = fill_R(R, first, diag_reads);
            bool more_dense = false;
            for (size_t i = 0; i < ops.size() && !more_dense; ++i) more_dense = !done[i] && ops[i].dense;
            if (more_dense || (R & ~io_allowed)) {
                std::vector<char> here(ops.size(), 0);
                for (size_t i : members) here[i] = 1;
                std::vector<size_t> kept;
                for (size_t i : members) {
                    bool hand_on = !ops[i].dense && (ops[i].reads & R);
                    for (size_t j = i + 1; j < ops.size() && hand_on; ++j)
                        if ((here[j] || !done[j]) && (ops[j].dense & ops[i].reads)) hand_on = false;
                    if (hand_on) { done[i] = 0; here[i] = 0; ++remaining; } else kept.push_back(i);
                }
                members.swap(kept);
            }
        }
        if (!emit_round(R, members)) return false;
        last_R = R;
        first = false;
    }
    if (n_rounds == 0 || (last_R & ~io_allowed)) {  // the storing round must have its register bits within 5..11
        if (!emit_round(0xf00, {})) return false;
    }
    // Per-thread phase tables travel forward: a merged table (RC_TAB) of round r is a function of thread / tile-index bits
    // only, so it commutes with every dense gate of round r + 1 unless one of its selector bits becomes a register bit there.
    // Moving it to the end of round r + 1 (and on, as long as that holds) leaves many rounds with nothing in their per-thread
    // phase: the 16 complex multiplies that apply it (64 FP64 instructions per thread and round) are not issued at all, and the
    // table costs one multiply per thread where it lands.  QB_TILE_DEFER_TPH=0 keeps every table in its round.
    {
        const char* dt = getenv("QB_TILE_DEFER_TPH");
        if (!(dt && dt[0] == '0') && n_rounds > 1) {
            std::vector<std::vector<ROp>> per(n_rounds);
            for (size_t r = 0; r < n_rounds; ++r)
                per[r].assign(prog->ops + prog->rounds[r].op_begin, prog->ops + prog->rounds[r].op_end);
            auto writes_tph_at_end = [](const std::vector<ROp>& v) {
                size_t from = 0;   // the chains' flushes reset the accumulator: only what follows the last one is applied at the round's end
                for (size_t i = 0; i < v.size(); ++i)
                    if (v[i].code >= RC_PH_FLUSH && v[i].code < RC_PH_FLUSH + 4) from = i + 1;
                for (size_t i = from; i < v.size(); ++i) {
                    const ROp& o = v[i];
                    if (o.code == RC_TAB || o.code == RC_PH_T || o.code == RC_DG1T) return true;
                    if (o.code == RC_DG && o.rmask == 0) {
                        bool reg = false;
                        for (int k = 0; k < o.k && k < 3; ++k) reg = reg || o.src[k] < 4;
                        if (!reg) return true;
                    }
                }
                return false;
            };
            for (size_t r = 0; r + 1 < n_rounds; ++r) {
                uint32_t Rn = 0;
                for (int i = 0; i < 4; ++i) Rn |= 1u << prog->rounds[r + 1].R[i];
                std::vector<ROp> keep;
                for (const ROp& o : per[r]) {
                    bool movable = o.code == RC_TAB && o.omask == 0;
                    const uint8_t sel[4] = {o.src[0], o.src[1], o.src[2], o.b};
                    for (int i = 0; movable && i < o.k && i < 4; ++i)
                        if (!(sel[i] & 0x80) && ((Rn >> (sel[i] & 0x3f)) & 1u)) movable = false;
                    if (movable) per[r + 1].push_back(o); else keep.push_back(o);
                }
                per[r].swap(keep);
            }
            n_ops = 0;
            for (size_t r = 0; r < n_rounds; ++r) {
                prog->rounds[r].op_begin = (uint16_t)n_ops;
                for (const ROp& o : per[r]) prog->ops[n_ops++] = o;
                prog->rounds[r].op_end = (uint16_t)n_ops;
                prog->rounds[r].has_tph = writes_tph_at_end(per[r]) ? 1 : 0;
            }
        }
    }
    if (pass_scalar.x != 1.0 || pass_scalar.y != 0.0) {   // the pass's constant factor, applied in the last round:
        RoundDesc& last = prog->rounds[n_rounds - 1];
        bool absorbed = false;   // ... by a per-thread table that round already has (its entries are scaled), ...
        for (uint32_t oi = last.op_begin; oi < last.op_end && !absorbed; ++oi) {
            const ROp& t = prog->ops[oi];
            if (t.code != RC_TAB) continue;
            for (uint32_t e = 0; e < (1u << t.k); ++e) {
                c64& v = prog->mats[t.mat + e];
                v = make_double2(v.x * pass_scalar.x - v.y * pass_scalar.y, v.x * pass_scalar.y + v.y * pass_scalar.x);
            }
            absorbed = true;
        }
        if (!absorbed) {   // ... or as one table op without selectors
            if (n_ops >= max_ops || n_mats + 1 > max_mats) return false;
            ROp o;
            memset(&o, 0, sizeof(o));
            o.code = RC_TAB;
            o.mat = (uint32_t)n_mats;
            prog->mats[n_mats++] = pass_scalar;
            prog->ops[n_ops++] = o;
            last.op_end = (uint16_t)n_ops;
            last.has_tph = 1;
        }
    }
    prog->n_rounds = (uint32_t)n_rounds;
    prog->n_mats = (uint32_t)n_mats;
    auto goff_host = [&](uint32_t x) {
        uint64_t o = x & ((1u << lc) - 1u);
        for (uint32_t i = 0; i < prog->h; ++i) o |= (uint64_t)((x >> (lc + i)) & 1u) << prog->hp[i];
        return o;
    };
    for (uint32_t j = 0; j < 16; ++j) {
        uint32_t lf = 0, ll = 0;
        for (uint32_t i = 0; i < 4; ++i)
            if ((j >> i) & 1) { lf |= 1u << prog->rounds[0].R[i]; ll |= 1u << prog->rounds[n_rounds - 1].R[i]; }
        prog->gro_first[j] = goff_host(lf);
        prog->gro_last[j] = goff_host(ll);
    }
    return true;
}

// Encodes the pass with every round-selection strategy and keeps the program with the fewest rounds (then the fewest ops).
// Only for big registers (`thorough`): below ~2^26 amplitudes a sweep is cheaper than the extra encoding time.
static bool thorough_encoding(uint32_t n_local) {
    const char* env = getenv("QB_PLAN_LOOKAHEAD");   // same switch as the pass planner's look-ahead: "0" never, "1" always (tests)
    return env ? env[0] == '1' : n_local >= 26;
}

template <class Prog>
static bool encode_rounds(const std::vector<Gate>& gates, const std::vector<size_t>& gate_idx, const std::vector<uint32_t>& tile_pos,
                          uint32_t lc, Prog* prog, bool thorough) {
    static thread_local Prog other;
    auto n_ops = [](const Prog& p) { return p.rounds[p.n_rounds - 1].op_end; };
    bool have = encode_rounds_with(gates, gate_idx, tile_pos, lc, prog, 0);
    for (int mode = 1; thorough && mode <= 2; ++mode) {
        if (!encode_rounds_with(gates, gate_idx, tile_pos, lc, &other, mode)) continue;
        if (!have || other.n_rounds < prog->n_rounds || (other.n_rounds == prog->n_rounds && n_ops(other) < n_ops(*prog))) {
            *prog = other;
            have = true;
        }
    }
    return have;
}

template <class Prog>
static int launch_rpass(Reg* r, const Prog& prog) {
    static std::atomic<uint64_t> attr_devs{0};  // per-device function attribute (bit d = set on device d)
    if (!((attr_devs.load() >> (r->device & 63)) & 1)) {
        QB_CUDA(cudaFuncSetAttribute(k2r_tile_pass<Prog>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(c64) << RT)));
        attr_devs.fetch_or(1ull << (r->device & 63));
    }
    const uint64_t tiles = 1ull << (r->n_local - RT);
    if (tiles > 0x7fffffffull) return set_error(QB_ERR_UNSUPPORTED, (int64_t)tiles, "too many tiles for one grid");
    bool jitted = false;   // the pass's specialised kernel, if it is ready (qb_jit.cu); else the interpreter kernel below
    if (r->xchg_last_pass && r->xchg_request) {
        // sharded registers: the last pass before a global<->local exchange stores every amplitude where it lives after the
        // exchange, in the destination rank's other buffer (specialised kernels only; qb_shard.cu flips the buffers afterwards)
        QB_TRY(jit_try_launch(r, make_view(prog), nullptr, &jitted, r->xchg_request));
        r->xchg_done = jitted;
    }
    if (!jitted) QB_TRY(jit_try_launch(r, make_view(prog), nullptr, &jitted));
    if (!jitted) {
        if (r->profile) QB_TRY(tile_profile_begin(r));
        k2r_tile_pass<Prog><<<(unsigned)tiles, RTHREADS, sizeof(c64) << RT, r->stream>>>(r->state, prog);
        QB_CUDA(cudaGetLastError());
        if (r->profile) QB_TRY(tile_profile_end(r));
    }
    r->st_launches++;
    r->st_passes++;
    r->st_rounds += prog.n_rounds;
    r->st_bytes += (int64_t)(32ull << r->n_local);
    return QB_OK;
}

// Sharded registers, experimental (QB_SHARD_REMOTE=1): the pass runs on the virtual register formed by all shards; see
// k2r_tile_pass_remote.  `bases[p]` = shard of rank p as mapped into this process.
template <class Prog>
static int launch_rpass_remote(Reg* r, const Prog& prog, uint32_t n_virtual, c64* const* bases) {
    static std::atomic<uint64_t> attr_devs{0};
    if (!((attr_devs.load() >> (r->device & 63)) & 1)) {
        QB_CUDA(cudaFuncSetAttribute(k2r_tile_pass_remote<Prog>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(c64) << RT)));
        attr_devs.fetch_or(1ull << (r->device & 63));
    }
    const uint32_t g = n_virtual - r->n_local;
    RemoteArgs ra;
    memset(&ra, 0, sizeof(ra));
    for (uint32_t p = 0; p < (1u << g); ++p) ra.base[p] = bases[p];
    ra.n_local = r->n_local;
    ra.tile_or = remote_fixed_bits(prog.hp, prog.h, r->n_local, g, (uint32_t)r->rank) << (n_virtual - RT - g);
    uint32_t in_tile = 0;   // rank bits among the tile bits: 1 - 2^-in_tile of the tile's amplitudes live on peers
    for (uint32_t j = 0; j < prog.h; ++j) in_tile += prog.hp[j] >= r->n_local;
    const uint64_t tiles = 1ull << (r->n_local - RT);
    if (tiles > 0x7fffffffull) return set_error(QB_ERR_UNSUPPORTED, (int64_t)tiles, "too many tiles for one grid");
    bool jitted = false;   // the specialised kernel of this pass (same generator, amplitudes addressed through `ra.base`)
    QB_TRY(jit_try_launch(r, make_view(prog), &ra, &jitted));
    if (!jitted) {
        if (r->profile) QB_TRY(tile_profile_begin(r));
        k2r_tile_pass_remote<Prog><<<(unsigned)tiles, RTHREADS, sizeof(c64) << RT, r->stream>>>(ra, prog);
        QB_CUDA(cudaGetLastError());
        if (r->profile) QB_TRY(tile_profile_end(r));
    }
    r->st_launches++;
    r->st_passes++;
    r->st_rounds += prog.n_rounds;
    r->st_bytes += (int64_t)(32ull << r->n_local);
    // NVLink accounting: what this rank stores into peers' shards (it loads as much from them)
    r->st_swap_bytes += (int64_t)((16ull << r->n_local) - ((16ull << r->n_local) >> in_tile));
    r->st_remote_passes += in_tile > 0;
    return QB_OK;
}

// Planner introspection (no GPU): serialises the round program of one pass.  Layout (little endian):
//   u32 lc, h, n_rounds, n_ops, n_mats; u8 hp[8]; RoundDesc[n_rounds]; ROp[n_ops]; c64[n_mats]; u64 gro_first[16], gro_last[16]
// Returns the number of bytes written, 0 if the pass cannot go through the register-blocked kernel, -1 if `cap` is too small.
int64_t encode_rounds_debug(uint32_t n_local, const std::vector<Gate>& gates, const std::vector<size_t>& gate_idx,
                            const std::vector<uint32_t>& tile_pos, uint32_t lc, unsigned char* out, uint64_t cap) {
    if (tile_pos.size() != RT) return 0;
    static thread_local RProgLarge prog;
    memset(&prog, 0, sizeof(prog));
    if (!encode_rounds(gates, gate_idx, tile_pos, lc, &prog, thorough_encoding(n_local))) return 0;
    const uint32_t n_rounds = prog.n_rounds, n_ops = prog.rounds[n_rounds - 1].op_end;
    uint32_t n_mats = 4;
    for (uint32_t i = 0; i < n_ops; ++i) {   // the encoder does not keep the table count: recover an upper bound per op
        const ROp& o = prog.ops[i];
        uint32_t len = 4;
        if (o.code == RC_D2) len = 16;
        else if (o.code == RC_DG) len = 1u << o.k;
        else if (o.code == RC_TAB || o.code == RC_TAB_ACC) len = 1u << o.k;
        else if (o.code == RC_LAYER) len = 8;
        n_mats = std::max(n_mats, o.mat + len);
    }
    n_mats = std::min<uint32_t>(n_mats, sizeof(prog.mats) / sizeof(c64));
    const uint64_t need = 20 + 8 + (uint64_t)n_rounds * sizeof(RoundDesc) + (uint64_t)n_ops * sizeof(ROp) + (uint64_t)n_mats * sizeof(c64) +
                          sizeof(prog.gro_first) + sizeof(prog.gro_last);
    if (need > cap) return -1;
    unsigned char* w = out;
    const uint32_t hdr[5] = {prog.lc, prog.h, n_rounds, n_ops, n_mats};
    memcpy(w, hdr, 20); w += 20;
    memcpy(w, prog.hp, 8); w += 8;
    memcpy(w, prog.rounds, n_rounds * sizeof(RoundDesc)); w += n_rounds * sizeof(RoundDesc);
    memcpy(w, prog.ops, n_ops * sizeof(ROp)); w += n_ops * sizeof(ROp);
    memcpy(w, prog.mats, n_mats * sizeof(c64)); w += n_mats * sizeof(c64);
    memcpy(w, prog.gro_first, sizeof(prog.gro_first)); w += sizeof(prog.gro_first);
    memcpy(w, prog.gro_last, sizeof(prog.gro_last)); w += sizeof(prog.gro_last);
    return (int64_t)(w - out);
}

bool pass_encodes_v2(uint32_t n_local, const std::vector<Gate>& gates, const std::vector<size_t>& gate_idx,
                     const std::vector<uint32_t>& tile_pos, uint32_t lc) {
    if (tile_pos.size() != RT || n_local < RT) return false;
    static thread_local RProgLarge prog;
    return encode_rounds(gates, gate_idx, tile_pos, lc, &prog, thorough_encoding(n_local));
}

// returns QB_OK and *handled = true if the pass went through the register-blocked kernel
int run_pass_v2(Reg* r, const std::vector<Gate>& gates, const std::vector<size_t>& gate_idx, const std::vector<uint32_t>& tile_pos,
                uint32_t lc, bool* handled) {
    *handled = false;
    if (tile_pos.size() != RT || r->n_local < RT) return QB_OK;
    size_t mats = 0;
    for (size_t gi : gate_idx) mats += gates[gi].matrix.size();
    if (gate_idx.size() <= 20 && mats <= 150) {
        RProgSmall small;
        if (encode_rounds(gates, gate_idx, tile_pos, lc, &small, thorough_encoding(r->n_local))) {
            *handled = true;
            return launch_rpass(r, small);
        }
    }
    static thread_local RProgLarge large;
    if (!encode_rounds(gates, gate_idx, tile_pos, lc, &large, thorough_encoding(r->n_local))) return QB_OK;
    *handled = true;
    return launch_rpass(r, large);
}

// Sharded registers (qb_shard.cu, QB_SHARD_REMOTE=1): one pass over the virtual register of n_virtual = n_local + g bits.
// `gates` carry PHYSICAL positions in that register (rank bits included, nothing resolved per rank), `tile_pos` may contain
// rank bits.  *handled = false (and nothing launched) if the register-blocked encoder cannot take the pass.
int run_pass_remote(Reg* r, uint32_t n_virtual, const std::vector<Gate>& gates, const std::vector<size_t>& gate_idx,
                    const std::vector<uint32_t>& tile_pos, uint32_t lc, c64* const* bases, bool* handled) {
    *handled = false;
    if (tile_pos.size() != RT || r->n_local < RT || n_virtual <= r->n_local || n_virtual - r->n_local > 3) return QB_OK;
    size_t mats = 0;
    for (size_t gi : gate_idx) mats += gates[gi].matrix.size();
    if (gate_idx.size() <= 20 && mats <= 150) {
        RProgSmall small;
        if (encode_rounds(gates, gate_idx, tile_pos, lc, &small, thorough_encoding(n_virtual))) {
            *handled = true;
            return launch_rpass_remote(r, small, n_virtual, bases);
        }
    }
    static thread_local RProgLarge large;
    if (!encode_rounds(gates, gate_idx, tile_pos, lc, &large, thorough_encoding(n_virtual))) return QB_OK;
    *handled = true;
    return launch_rpass_remote(r, large, n_virtual, bases);
}

}  // namespace qb
