// qb_tile.cu -- fused tile passes: many gates per HBM sweep.
//
// The reference applies one gate per full traversal of the state (src/register.rs:160-188, plus a
// clone at :155).  Here queued gates are planned into PASSES.  A pass picks a set of T index bits
// (the low `lc` bits, contiguous in memory, plus up to T-5 arbitrary high bits); every CTA stages
// one 2^T-amplitude tile of the state in shared memory (coalesced >= 512-byte segments), applies
// every gate of the pass whose dense target bits lie inside the tile, and writes the tile back:
// one read and one write of the state for the whole group of gates.
//   * control qubits never have to be in the tile: a control on a non-tile bit is a per-tile
//     predicate on the tile's base address;
//   * diagonal gates never constrain the tile either: their non-tile bits select a sub-table.
// The pass program travels in kernel-parameter (constant-bank) space, so there is no device
// allocation or copy per pass.
#include <stdio.h>
#include <stdlib.h>

#include <functional>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <unordered_map>

#include "qb_common.h"
#include "qb_device.cuh"
#include "qb_tile.h"

namespace qb {

// ------------------------------------------------------------------------------------------------
// device-side program
// ------------------------------------------------------------------------------------------------
enum TileOpType : uint8_t { TOP_DENSE1 = 0, TOP_DENSE2 = 1, TOP_DIAG = 2 };

struct TileOp {
    uint8_t type;
    uint8_t k;         // DIAG: number of index bits (<= 3)
    uint8_t b[3];      // DENSE: local bit of matrix bit j; DIAG: local bit, or 0x80 | global position (outer bit)
    uint8_t pad[3];
    uint32_t lctrl;    // controls inside the tile (local bit mask)
    uint32_t mat;      // offset of the matrix / diagonal in `mats` (c64 units)
    uint64_t octrl;    // controls outside the tile (global bit mask, tested against the tile base)
};
static_assert(sizeof(TileOp) == 24, "TileOp layout");

template <int NOPS, int NMAT>
struct PassProgram {
    uint32_t tile_bits;  // T' = min(T, n_local)
    uint32_t lc;         // contiguous low bits
    uint32_t h;          // high tile bits (T' = lc + h)
    uint32_t n_ops;
    uint8_t hp[8];       // ascending global positions of the high tile bits
    TileOp ops[NOPS];
    c64 mats[NMAT];
};
typedef PassProgram<16, 112> PassSmall;    // <= 4 KB of parameter space
typedef PassProgram<160, 1280> PassLarge;  // ~24 KB (CUDA >= 12.1 allows 32,764 bytes)

constexpr int TILE_THREADS = 256;

// ------------------------------------------------------------------------------------------------
// K2: the tile-pass kernel
// ------------------------------------------------------------------------------------------------
template <class Prog>
__global__ void __launch_bounds__(TILE_THREADS) k2_tile_pass(c64* __restrict__ st, const __grid_constant__ Prog prog) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c64* tile = reinterpret_cast<c64*>(smem_raw);
    __shared__ uint64_t seg_off[128];

    const uint32_t T = prog.tile_bits, lc = prog.lc, h = prog.h;
    const uint32_t tile_n = 1u << T;
    const uint32_t tid = threadIdx.x;

    // base address of this tile: deposit blockIdx into the non-tile bit positions
    uint64_t base = (uint64_t)blockIdx.x << lc;
    for (uint32_t j = 0; j < h; ++j) base = insert_zero_bit(base, prog.hp[j]);
    if (tid < (1u << h)) {
        uint64_t o = 0;
        for (uint32_t j = 0; j < h; ++j) o |= (uint64_t)((tid >> j) & 1) << prog.hp[j];
        seg_off[tid] = o;
    }
    __syncthreads();
    const uint32_t lmask = (1u << lc) - 1;

    // ---- stage the tile: consecutive threads read consecutive amplitudes (>= 512 B per warp) ----
#pragma unroll 8
    for (uint32_t l = tid; l < tile_n; l += TILE_THREADS)
        tile[l] = ld128(st + (base + (l & lmask) + seg_off[l >> lc]));
    __syncthreads();

    // ---- apply the pass program in shared memory ---------------------------------------------
    for (uint32_t oi = 0; oi < prog.n_ops; ++oi) {
        const TileOp& op = prog.ops[oi];
        if ((base & op.octrl) != op.octrl) continue;  // uniform per CTA
        const uint32_t lctrl = op.lctrl;
        if (op.type == TOP_DENSE1) {
            const uint32_t b = op.b[0];
            const c64 m00 = prog.mats[op.mat + 0], m01 = prog.mats[op.mat + 1];
            const c64 m10 = prog.mats[op.mat + 2], m11 = prog.mats[op.mat + 3];
#pragma unroll 4
            for (uint32_t p = tid; p < (tile_n >> 1); p += TILE_THREADS) {
                const uint32_t l0 = (uint32_t)insert_zero_bit(p, b);
                if ((l0 & lctrl) != lctrl) continue;
                const uint32_t l1 = l0 | (1u << b);
                const c64 a0 = tile[l0], a1 = tile[l1];
                tile[l0] = cmul2(m00, a0, m01, a1);
                tile[l1] = cmul2(m10, a0, m11, a1);
            }
        } else if (op.type == TOP_DENSE2) {
            const uint32_t b0 = op.b[0], b1 = op.b[1];
            const uint32_t lo = b0 < b1 ? b0 : b1, hi = b0 < b1 ? b1 : b0;
            const c64* M = prog.mats + op.mat;
            for (uint32_t p = tid; p < (tile_n >> 2); p += TILE_THREADS) {
                const uint32_t l00 = (uint32_t)insert_zero_bit(insert_zero_bit(p, lo), hi);
                if ((l00 & lctrl) != lctrl) continue;
                const uint32_t i1 = l00 | (1u << b0), i2 = l00 | (1u << b1), i3 = i1 | (1u << b1);
                const c64 a0 = tile[l00], a1 = tile[i1], a2 = tile[i2], a3 = tile[i3];
                c64 r0 = cmul2(M[0], a0, M[1], a1);   r0 = cfma(M[2], a2, r0);   r0 = cfma(M[3], a3, r0);
                c64 r1 = cmul2(M[4], a0, M[5], a1);   r1 = cfma(M[6], a2, r1);   r1 = cfma(M[7], a3, r1);
                c64 r2 = cmul2(M[8], a0, M[9], a1);   r2 = cfma(M[10], a2, r2);  r2 = cfma(M[11], a3, r2);
                c64 r3 = cmul2(M[12], a0, M[13], a1); r3 = cfma(M[14], a2, r3);  r3 = cfma(M[15], a3, r3);
                tile[l00] = r0; tile[i1] = r1; tile[i2] = r2; tile[i3] = r3;
            }
        } else {  // TOP_DIAG, k <= 3 index bits, each local or outer
            uint32_t oidx = 0;
            uint32_t lb[3] = {0, 0, 0};
            uint32_t lsel = 0;  // which matrix bits are local
            for (uint32_t j = 0; j < op.k; ++j) {
                const uint32_t bj = op.b[j];
                if (bj & 0x80) oidx |= (uint32_t)((base >> (bj & 0x7f)) & 1) << j;
                else { lb[j] = bj; lsel |= 1u << j; }
            }
            const c64* D = prog.mats + op.mat;
            // the (at most 8) factors this tile can see, resolved to registers (uniform loads)
            c64 f[8];
#pragma unroll
            for (uint32_t x = 0; x < 8; ++x) f[x] = D[(oidx | (x & lsel)) & ((1u << op.k) - 1)];
#pragma unroll 4
            for (uint32_t l = tid; l < tile_n; l += TILE_THREADS) {
                if ((l & lctrl) != lctrl) continue;
                const uint32_t x = (((l >> lb[0]) & 1) | (((l >> lb[1]) & 1) << 1) | (((l >> lb[2]) & 1) << 2)) & lsel;
                const c64 lo4 = (x & 2) ? ((x & 1) ? f[3] : f[2]) : ((x & 1) ? f[1] : f[0]);
                const c64 hi4 = (x & 2) ? ((x & 1) ? f[7] : f[6]) : ((x & 1) ? f[5] : f[4]);
                tile[l] = cmul((x & 4) ? hi4 : lo4, tile[l]);
            }
        }
        __syncthreads();
    }

    // ---- write the tile back ---------------------------------------------------------------------
#pragma unroll 8
    for (uint32_t l = tid; l < tile_n; l += TILE_THREADS)
        st128(st + (base + (l & lmask) + seg_off[l >> lc]), tile[l]);
}

// ------------------------------------------------------------------------------------------------
// planner
// ------------------------------------------------------------------------------------------------
constexpr uint32_t MIN_LC = 5;       // >= 512-byte contiguous segments
constexpr uint32_t MAX_DIAG_K = 3;

bool tile_supported(const Gate& g) {
    if (g.kind == GK_DIAG) return g.k <= MAX_DIAG_K;
    return g.k >= 1 && g.k <= 2;
}

// Greedy list scheduling.  Gates are scanned in program order; a gate joins the current pass if no
// earlier unscheduled gate shares a qubit with it (so program order is preserved per qubit) and its
// dense target bits fit the tile; otherwise its qubits become blocked for the rest of this pass.
// `sink` receives every pass as soon as it is final (planning later passes never changes an earlier one), so that the
// caller can launch pass k while pass k+1 is being planned.
static void plan_passes(const std::vector<Gate>& gates, uint32_t n_local, uint32_t T_opt, int max_ops, int max_mats,
                        const std::function<void(PlannedPass&&)>& sink) {
    const uint32_t T = std::min(T_opt, n_local);
    // QB_PLAN_MIN_LC=4 (experiment): 256-byte contiguous segments, one more arbitrary high bit per tile
    const char* lc_env = getenv("QB_PLAN_MIN_LC");
    const bool lc4 = lc_env && lc_env[0] == '4' && T == 12;
    const uint32_t lc_min = std::min(lc4 ? 4u : MIN_LC, T);
    const uint32_t h_max = std::min<uint32_t>(T - lc_min, lc4 ? 8 : 7);
    std::vector<char> done(gates.size(), 0);
    size_t remaining = gates.size();
    // One pass = one scan of the remaining gates in program order.  A gate WRITES its dense target bits and READS its
    // control / diagonal bits; two gates commute unless one writes a bit the other touches (diagonal gates and controls
    // commute among themselves).  `preset` = high tile bits fixed in advance; the scan adds further high bits while slots
    // are free.  With `picked` the scan commits (marks the gates done), otherwise it only counts.
    struct PassScan { uint64_t high; uint32_t nh; int count; };
    auto scan_pass = [&](size_t from, uint64_t preset, size_t window, std::vector<size_t>* picked) -> PassScan {
        uint64_t high = preset, blocked_w = 0, blocked_r = 0;
        uint32_t nh = (uint32_t)__builtin_popcountll(preset);
        int n_ops = 0, n_mats = 0;
        size_t seen = 0;
        for (size_t i = from; i < gates.size() && seen < window; ++i) {
            if (done[i]) continue;
            ++seen;
            const Gate& g = gates[i];
            const uint64_t wm = g.kind == GK_DENSE ? g.target_mask() : 0;
            const uint64_t rm = g.ctrl_mask | (g.kind == GK_DIAG ? g.target_mask() : 0);
            if ((wm & (blocked_w | blocked_r)) || (rm & blocked_w)) { blocked_w |= wm; blocked_r |= rm; continue; }
            bool ok = tile_supported(g);
            uint64_t need = 0;
            if (ok && g.kind == GK_DENSE) {
                for (uint32_t t : g.targets)
                    if (t >= lc_min && !((high >> t) & 1)) need |= 1ull << t;
                if (nh + (uint32_t)__builtin_popcountll(need) > h_max) ok = false;
            }
            const int mats = (int)g.matrix.size();
            if (ok && (n_ops + 1 > max_ops || n_mats + mats > max_mats)) ok = false;
            if (!ok) { blocked_w |= wm; blocked_r |= rm; continue; }
            high |= need;
            nh += (uint32_t)__builtin_popcountll(need);
            n_ops++;
            n_mats += mats;
            if (picked) { picked->push_back(i); done[i] = 1; --remaining; }
        }
        return PassScan{high, nh, n_ops};
    };
    // Look-ahead over the choice of high tile bits (big registers only: a sweep must cost more than planning it).  The
    // plain scan takes high bits first come, first served; alternatives -- the most frequent dense targets of the next
    // gates, and the greedy set with one bit exchanged -- are scored by the number of gates the pass would take.
    const char* la_env = getenv("QB_PLAN_LOOKAHEAD");   // "0": never, "1": always (tests), unset: registers of >= 28 local qubits
    const bool lookahead = h_max > 0 && (la_env ? la_env[0] == '1' : n_local >= 28);
    const size_t kWindow = 1024;
    // score of a candidate set of high bits: the gates its pass would take, plus -- for registers of >= 30 local qubits,
    // where a sweep costs far more than planning -- the gates of the plain pass that would follow (two-step look-ahead:
    // commit tentatively, scan again, roll back)
    const bool two_step = n_local >= 30 || (la_env && la_env[0] == '1');
    size_t first = 0;
    auto score = [&](uint64_t preset) -> PassScan {
        if (!two_step) return scan_pass(first, preset, kWindow, nullptr);
        std::vector<size_t> tentative;
        PassScan r = scan_pass(first, preset, kWindow, &tentative);
        size_t next = first;
        while (next < gates.size() && done[next]) ++next;
        if (next < gates.size()) r.count += scan_pass(next, 0, kWindow, nullptr).count;
        for (size_t ix : tentative) done[ix] = 0;
        remaining += tentative.size();
        return r;
    };
    while (remaining) {
        while (done[first]) ++first;
        PlannedPass pass;
        uint64_t preset = 0;
        if (lookahead) {
            const PassScan g0 = score(0);
            int best = g0.count;
            if (g0.count > 0) {
                // dense-target frequency among the upcoming gates
                uint32_t freq[64] = {0};
                std::vector<uint64_t> cands;
                size_t seen = 0;
                const size_t marks[4] = {48, 96, 192, 384};
                auto top_bits = [&](uint32_t k) {
                    uint64_t m = 0;
                    for (uint32_t c = 0; c < k; ++c) {
                        int arg = -1;
                        for (int bpos = (int)lc_min; bpos < 64; ++bpos)
                            if (!((m >> bpos) & 1) && freq[bpos] && (arg < 0 || freq[bpos] > freq[arg])) arg = bpos;
                        if (arg < 0) break;
                        m |= 1ull << arg;
                    }
                    return m;
                };
                size_t mi = 0;
                for (size_t i = first; i < gates.size() && mi < 4; ++i) {
                    if (done[i]) continue;
                    if (gates[i].kind == GK_DENSE)
                        for (uint32_t t : gates[i].targets)
                            if (t >= lc_min) freq[t]++;
                    if (++seen == marks[mi]) { cands.push_back(top_bits(h_max)); ++mi; }
                }
                if (mi == 0) cands.push_back(top_bits(h_max));
                uint64_t centre = g0.high;
                // hill climbing over one-bit exchanges of the best set so far; registers of >= 30 local qubits, whose sweeps cost
                // ~100 ms, search every candidate bit and climb longer (bench, 2 steps at 33 qubits: 45 -> 41 sweeps, same rounds and ops;
                // ~0.26 s of planning per 4096 gates instead of 0.10 s, hidden behind the seconds of GPU work they represent)
                const int max_climbs = two_step ? 6 : 3;
                const uint32_t pool_bits = two_step ? 64 : 12;
                for (int climb = 0; climb < max_climbs; ++climb) {
                    const uint64_t pool = top_bits(pool_bits) & ~centre;
                    for (uint64_t out_b = centre; out_b; out_b &= out_b - 1)
                        for (uint64_t in_b = pool; in_b; in_b &= in_b - 1)
                            cands.push_back((centre & ~(out_b & -out_b)) | (in_b & -in_b));
                    const int before = best;
                    for (uint64_t c : cands) {
                        if (c == 0 || (uint32_t)__builtin_popcountll(c) > h_max) continue;
                        const PassScan r = score(c);
                        if (r.count > best) { best = r.count; preset = c; centre = r.high; }
                    }
                    cands.clear();
                    if (best == before) break;
                }
            }
        }
        const PassScan fin = scan_pass(first, preset, (size_t)-1, &pass.gate_idx);
        const uint64_t high = fin.high;
        const uint32_t nh = fin.nh;
        if (pass.gate_idx.empty()) {  // the first remaining gate cannot go through the tile kernel
            pass.unfused = true;
            pass.gate_idx.push_back(first);
            done[first] = 1;
            --remaining;
            sink(std::move(pass));
            continue;
        }
        // tile = lowest lc bits + the chosen high bits that are not already among them
        uint32_t lc = T - nh;
        for (;;) {
            uint32_t above = 0;
            for (uint32_t b = lc; b < 64; ++b) above += (uint32_t)((high >> b) & 1);
            if (lc + above == T) break;
            lc = T - above;
        }
        pass.lc = lc;
        for (uint32_t b = 0; b < lc; ++b) pass.tile_pos.push_back(b);
        for (uint32_t b = lc; b < 64; ++b)
            if ((high >> b) & 1) pass.tile_pos.push_back(b);
        sink(std::move(pass));
    }
}

static void plan_passes(const std::vector<Gate>& gates, uint32_t n_local, uint32_t T_opt, int max_ops, int max_mats,
                        std::vector<PlannedPass>* out) {
    plan_passes(gates, n_local, T_opt, max_ops, max_mats, [out](PlannedPass&& p) { out->push_back(std::move(p)); });
}

// ------------------------------------------------------------------------------------------------
// Second-tier planner: beam search over the passes' high tile bits.
//
// plan_passes above is a one-pass-at-a-time heuristic (it has to be: a pass is launched as soon as the next one is planned).
// When the SAME gate list comes back (plan cache, run_fused) the search may look at the whole list: `beam` partial plans are
// kept per depth; each is extended by its best few choices of high bits (greedy set, most frequent dense targets of the
// next 48..384 gates, hill climbing over one-bit exchanges, every candidate scored by the gates it schedules now plus what
// the plain pass after it would take) and the `beam` states with the most gates done survive.  On the BASELINE circuit at 33
// qubits: 21 passes -> 19.  Same commutation rule and capacity limits as plan_passes, so whatever it returns is a valid plan.
// ------------------------------------------------------------------------------------------------
namespace {
struct BeamPlanner {
    const std::vector<Gate>& gates;
    uint32_t T, lc_min, h_max;
    int max_ops, max_mats;
    std::vector<uint64_t> wm, rm;   // per gate: written bits (dense targets), read bits (controls, diagonal targets)
    std::vector<char> ok_gate;      // tile_supported

    BeamPlanner(const std::vector<Gate>& g, uint32_t n_local, uint32_t T_opt, int mo, int mm) : gates(g), max_ops(mo), max_mats(mm) {
        T = std::min(T_opt, n_local);
        lc_min = std::min(MIN_LC, T);
        h_max = std::min<uint32_t>(T - lc_min, 7);
        wm.resize(g.size()); rm.resize(g.size()); ok_gate.resize(g.size());
        for (size_t i = 0; i < g.size(); ++i) {
            wm[i] = g[i].kind == GK_DENSE ? g[i].target_mask() : 0;
            rm[i] = g[i].ctrl_mask | (g[i].kind == GK_DIAG ? g[i].target_mask() : 0);
            ok_gate[i] = tile_supported(g[i]);
        }
    }
    struct Scan { uint64_t high; uint32_t nh; int count; };
    Scan scan(const std::vector<char>& done, uint64_t preset, size_t window, std::vector<size_t>* picked) const {
        uint64_t high = preset, bw = 0, br = 0;
        uint32_t nh = (uint32_t)__builtin_popcountll(preset);
        int n_ops = 0, n_mats = 0;
        size_t seen = 0;
        const uint64_t low = (1ull << lc_min) - 1ull;
        for (size_t i = 0; i < gates.size() && seen < window; ++i) {
            if (done[i]) continue;
            ++seen;
            if ((wm[i] & (bw | br)) || (rm[i] & bw)) { bw |= wm[i]; br |= rm[i]; continue; }
            bool ok = ok_gate[i];
            const uint64_t need = wm[i] & ~low & ~high;
            if (ok && nh + (uint32_t)__builtin_popcountll(need) > h_max) ok = false;
            const int mats = (int)gates[i].matrix.size();
            if (ok && (n_ops + 1 > max_ops || n_mats + mats > max_mats)) ok = false;
            if (!ok) { bw |= wm[i]; br |= rm[i]; continue; }
            high |= need;
            nh += (uint32_t)__builtin_popcountll(need);
            n_ops++;
            n_mats += mats;
            if (picked) picked->push_back(i);
        }
        return Scan{high, nh, n_ops};
    }
    // gates this choice schedules now + what the plain pass after it would take
    int score(const std::vector<char>& done, uint64_t preset, uint64_t* high_out) const {
        std::vector<size_t> picked;
        const Scan a = scan(done, preset, (size_t)-1, &picked);
        if (high_out) *high_out = a.high;
        std::vector<char> d2 = done;
        for (size_t i : picked) d2[i] = 1;
        return a.count + scan(d2, 0, (size_t)-1, nullptr).count;
    }
    uint64_t top_bits(const uint32_t* freq, uint32_t k) const {
        uint64_t m = 0;
        for (uint32_t c = 0; c < k; ++c) {
            int arg = -1;
            for (int b = (int)lc_min; b < 64; ++b)
                if (!((m >> b) & 1) && freq[b] && (arg < 0 || freq[b] > freq[arg])) arg = b;
            if (arg < 0) break;
            m |= 1ull << arg;
        }
        return m;
    }
    // the best `keep` distinct choices of high bits for this state
    void options(const std::vector<char>& done, size_t keep, std::vector<std::pair<int, uint64_t>>* out) const {
        uint32_t freq[64] = {0};
        std::vector<uint64_t> cands{0ull};
        size_t seen = 0, mi = 0;
        const size_t marks[4] = {48, 96, 192, 384};
        for (size_t i = 0; i < gates.size() && mi < 4; ++i) {
            if (done[i]) continue;
            for (uint64_t m = wm[i] >> lc_min << lc_min; m; m &= m - 1) freq[__builtin_ctzll(m)]++;
            if (++seen == marks[mi]) { cands.push_back(top_bits(freq, h_max)); ++mi; }
        }
        if (mi == 0) cands.push_back(top_bits(freq, h_max));
        std::vector<std::pair<int, uint64_t>> scored;   // (score, preset)
        auto consider = [&](uint64_t c) -> int {
            if ((uint32_t)__builtin_popcountll(c) > h_max) return -1;
            for (const auto& e : scored) if (e.second == c) return e.first;
            const int sc = score(done, c, nullptr);
            scored.push_back({sc, c});
            return sc;
        };
        for (uint64_t c : cands) consider(c);
        // hill climbing from every seed; the options handed back are the LOCAL OPTIMA the climbs end in (distinct ones): the top of
        // the scored list would be one optimum and its one-bit neighbours, and the beam would lose its diversity
        const std::vector<std::pair<int, uint64_t>> seeds = scored;
        std::vector<std::pair<int, uint64_t>> optima;
        for (const auto& seed : seeds) {
            uint64_t centre_preset = seed.second, centre_high = 0;
            int best = score(done, centre_preset, &centre_high);
            for (int climb = 0; climb < 6; ++climb) {
                const uint64_t pool = top_bits(freq, 64) & ~centre_high;
                uint64_t bp = centre_preset;
                int bs = best;
                for (uint64_t ib = pool; ib; ib &= ib - 1) {
                    if ((uint32_t)__builtin_popcountll(centre_high) < h_max) {   // a free slot: add without giving one up
                        const uint64_t c = centre_high | (ib & -ib);
                        const int sc = consider(c);
                        if (sc > bs) { bs = sc; bp = c; }
                    }
                    for (uint64_t ob = centre_high; ob; ob &= ob - 1) {
                        const uint64_t c = (centre_high & ~(ob & -ob)) | (ib & -ib);
                        const int sc = consider(c);
                        if (sc > bs) { bs = sc; bp = c; }
                    }
                }
                if (bs == best) break;
                best = bs;
                centre_preset = bp;
                score(done, bp, &centre_high);
            }
            bool dup = false;
            for (const auto& o : optima) dup = dup || o.second == centre_preset;
            if (!dup) optima.push_back({best, centre_preset});
        }
        std::sort(optima.begin(), optima.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
        for (size_t i = 0; i < optima.size() && out->size() < keep; ++i) out->push_back(optima[i]);
    }
    PlannedPass make_pass(uint64_t high, uint32_t nh, std::vector<size_t>&& idx) const {
        PlannedPass pass;
        pass.gate_idx = std::move(idx);
        uint32_t lc = T - nh;
        for (;;) {
            uint32_t above = 0;
            for (uint32_t b = lc; b < 64; ++b) above += (uint32_t)((high >> b) & 1);
            if (lc + above == T) break;
            lc = T - above;
        }
        pass.lc = lc;
        for (uint32_t b = 0; b < lc; ++b) pass.tile_pos.push_back(b);
        for (uint32_t b = lc; b < 64; ++b)
            if ((high >> b) & 1) pass.tile_pos.push_back(b);
        return pass;
    }
    struct State { std::vector<char> done; size_t n_done; std::vector<PlannedPass> passes; };
    bool run(size_t beam, size_t per_state, std::vector<PlannedPass>* out) const {
        std::vector<State> states(1);
        states[0].done.assign(gates.size(), 0);
        states[0].n_done = 0;
        for (size_t depth = 0; depth < 4 * gates.size() + 8; ++depth) {
            for (const State& st : states)
                if (st.n_done == gates.size()) { *out = st.passes; return true; }
            std::vector<State> next;
            for (const State& st : states) {
                std::vector<std::pair<int, uint64_t>> opts;
                options(st.done, per_state, &opts);
                bool any = false;
                for (const auto& o : opts) {
                    std::vector<size_t> picked;
                    const Scan sc = scan(st.done, o.second, (size_t)-1, &picked);
                    if (picked.empty()) continue;
                    any = true;
                    State ns;
                    ns.done = st.done;
                    for (size_t i : picked) ns.done[i] = 1;
                    ns.n_done = st.n_done + picked.size();
                    ns.passes = st.passes;
                    ns.passes.push_back(make_pass(sc.high, sc.nh, std::move(picked)));
                    next.push_back(std::move(ns));
                }
                if (!any) {   // the first remaining gate cannot go through the tile kernel: a pass of its own
                    State ns = st;
                    size_t first = 0;
                    while (ns.done[first]) ++first;
                    PlannedPass pass;
                    pass.unfused = true;
                    pass.gate_idx.push_back(first);
                    ns.done[first] = 1;
                    ns.n_done++;
                    ns.passes.push_back(std::move(pass));
                    next.push_back(std::move(ns));
                }
            }
            std::sort(next.begin(), next.end(), [](const State& a, const State& b) { return a.n_done > b.n_done; });
            std::vector<State> kept;
            for (State& ns : next) {
                bool dup = false;
                for (const State& k : kept) if (k.n_done == ns.n_done && k.done == ns.done) { dup = true; break; }
                if (!dup) kept.push_back(std::move(ns));
                if (kept.size() >= beam) break;
            }
            states.swap(kept);
            if (states.empty()) return false;
            if (getenv("QB_PLAN_BEAM_TRACE")) {
                fprintf(stderr, "beam depth %zu:", depth + 1);
                for (const State& k : states) fprintf(stderr, " %zu", k.n_done);
                fprintf(stderr, "\n");
            }
        }
        return false;
    }
};

// The second-tier plan for a gate list, with the encoder in the loop.  The beam search is allowed more gates per pass than the
// shared-memory kernel's program holds (the register-blocked encoder merges the phase cascades of a QFT pass, hundreds of
// gates, into a handful of ops); a pass the encoder still refuses is cut in two sweeps at launch.  So the plan's cost is its
// number of SWEEPS, and when the first plan contains such a pass the search is repeated with smaller caps (the gates that no
// longer fit move on to the next pass instead of paying a sweep of their own) and the cheapest plan wins.
size_t plan_sweeps(const std::vector<Gate>& gates, uint32_t n_local, const std::vector<PlannedPass>& plan) {
    size_t sweeps = 0;
    for (const PlannedPass& p : plan) {
        bool one = p.unfused || p.gate_idx.size() == 1 || p.tile_pos.size() != 12;
        if (!one) one = pass_encodes_v2(n_local, gates, p.gate_idx, p.tile_pos, p.lc);
        sweeps += one ? 1 : 2;
    }
    return sweeps;
}
bool beam_plan(const std::vector<Gate>& gates, uint32_t n_local, uint32_t tile_bits, bool big, int max_ops, int max_mats, size_t beam,
               std::vector<PlannedPass>* out) {
    const int caps[7][2] = {{big ? 320 : max_ops, big ? 1000 : max_mats}, {288, 900}, {256, 800}, {224, 700}, {192, 600}, {160, 500}, {max_ops, max_mats}};
    bool have = false;
    size_t best = 0;
    for (int c = 0; c < (big ? 7 : 1); ++c) {
        if (c > 0 && caps[c][0] >= caps[c - 1][0]) continue;
        BeamPlanner bp(gates, n_local, tile_bits, caps[c][0], caps[c][1]);
        std::vector<PlannedPass> cand;
        if (!bp.run(beam, beam, &cand)) continue;
        const size_t sw = big ? plan_sweeps(gates, n_local, cand) : cand.size();
        if (getenv("QB_PLAN_BEAM_TRACE")) fprintf(stderr, "beam_plan: caps %d/%d -> %zu passes, %zu sweeps\n", caps[c][0], caps[c][1], cand.size(), sw);
        if (!have || sw < best) { *out = cand; best = sw; have = true; }
        if (sw == cand.size()) break;   // every pass of this plan is one sweep: smaller caps cannot do better
    }
    return have;
}

// ---- plan cache: gate-list hash -> passes (process-wide; plans carry gate INDICES, so they fit any list with the same hash) ----
struct CachedPlan {
    std::vector<PlannedPass> passes;
    size_t n_gates = 0;
    int tier = 0;        // 0: the streaming planner's plan; 1: the beam search has looked at it
    uint64_t seen = 0;
};
std::mutex& g_plan_mu = *new std::mutex;
std::unordered_map<uint64_t, CachedPlan>& g_plans = *new std::unordered_map<uint64_t, CachedPlan>;

uint64_t plan_key(const std::vector<Gate>& gates, uint32_t n_local, uint32_t tile_bits) {
    uint64_t h = 0xcbf29ce484222325ull ^ ((uint64_t)n_local << 32) ^ tile_bits;
    auto mix = [&](const void* p, size_t n) {
        const unsigned char* b = (const unsigned char*)p;
        for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 0x100000001b3ull;
    };
    for (const char* name : {"QB_PLAN_LOOKAHEAD", "QB_PLAN_MIN_LC", "QB_PLAN_BEAM"}) {   // the planner's switches are part of the key
        const char* v = getenv(name);
        mix(v ? v : "-", v ? strlen(v) + 1 : 2);
    }
    for (const Gate& g : gates) {
        mix(&g.k, sizeof g.k);
        mix(&g.kind, sizeof g.kind);
        mix(&g.ctrl_mask, sizeof g.ctrl_mask);
        mix(g.targets.data(), g.targets.size() * sizeof(uint32_t));
        mix(g.matrix.data(), g.matrix.size() * sizeof(c64));
    }
    return h;
}
}  // namespace

template <class Prog>
static bool encode_pass(const std::vector<Gate>& gates, const PlannedPass& pass, Prog* prog) {
    const uint32_t T = (uint32_t)pass.tile_pos.size();
    prog->tile_bits = T;
    prog->lc = pass.lc;
    prog->h = T - pass.lc;
    for (uint32_t j = 0; j < prog->h; ++j) prog->hp[j] = (uint8_t)pass.tile_pos[pass.lc + j];
    int8_t local_of[64];
    uint64_t tile_mask = 0;
    for (int b = 0; b < 64; ++b) local_of[b] = -1;
    for (uint32_t j = 0; j < T; ++j) { local_of[pass.tile_pos[j]] = (int8_t)j; tile_mask |= 1ull << pass.tile_pos[j]; }
    uint32_t n_ops = 0, n_mats = 0;
    for (size_t gi : pass.gate_idx) {
        const Gate& g = gates[gi];
        TileOp op;
        memset(&op, 0, sizeof(op));
        op.k = (uint8_t)g.k;
        op.mat = n_mats;
        op.octrl = g.ctrl_mask & ~tile_mask;
        uint32_t lctrl = 0;
        for (uint64_t m = g.ctrl_mask & tile_mask; m; m &= m - 1) lctrl |= 1u << local_of[__builtin_ctzll(m)];
        op.lctrl = lctrl;
        if (g.kind == GK_DIAG) {
            op.type = TOP_DIAG;
            for (uint32_t j = 0; j < g.k; ++j) {
                const uint32_t t = g.targets[j];
                op.b[j] = local_of[t] >= 0 ? (uint8_t)local_of[t] : (uint8_t)(0x80 | t);
            }
        } else {
            op.type = g.k == 1 ? TOP_DENSE1 : TOP_DENSE2;
            for (uint32_t j = 0; j < g.k; ++j) {
                if (local_of[g.targets[j]] < 0) return false;  // planner bug guard
                op.b[j] = (uint8_t)local_of[g.targets[j]];
            }
        }
        if (n_ops >= sizeof(prog->ops) / sizeof(TileOp) || n_mats + g.matrix.size() > sizeof(prog->mats) / sizeof(c64))
            return false;
        memcpy(prog->mats + n_mats, g.matrix.data(), g.matrix.size() * sizeof(c64));
        n_mats += (uint32_t)g.matrix.size();
        prog->ops[n_ops++] = op;
    }
    prog->n_ops = n_ops;
    return true;
}

template <class Prog>
static int launch_pass(Reg* r, const Prog& prog) {
    const size_t smem = sizeof(c64) << prog.tile_bits;
    static std::atomic<uint64_t> attr_devs{0};  // per-device function attribute (bit d = set on device d)
    if (!((attr_devs.load() >> (r->device & 63)) & 1)) {
        QB_CUDA(cudaFuncSetAttribute(k2_tile_pass<Prog>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 << 13));
        attr_devs.fetch_or(1ull << (r->device & 63));
    }
    const uint64_t tiles = 1ull << (r->n_local - prog.tile_bits);
    if (tiles > 0x7fffffffull) return set_error(QB_ERR_UNSUPPORTED, (int64_t)tiles, "too many tiles for one grid");
    if (r->profile) QB_TRY(tile_profile_begin(r));
    k2_tile_pass<Prog><<<(unsigned)tiles, TILE_THREADS, smem, r->stream>>>(r->state, prog);
    QB_CUDA(cudaGetLastError());
    if (r->profile) QB_TRY(tile_profile_end(r));
    r->st_launches++;
    r->st_passes++;
    r->st_bytes += (int64_t)(32ull << r->n_local);
    return QB_OK;
}

int tile_profile_begin(Reg* r) {
    if (r->tile_events_used + 2 > 1024) QB_TRY(drain_tile_events(r));
    while (r->tile_events.size() < r->tile_events_used + 2) {
        cudaEvent_t e;
        QB_CUDA(cudaEventCreate(&e));
        r->tile_events.push_back(e);
    }
    QB_CUDA(cudaEventRecord(r->tile_events[r->tile_events_used], r->stream));
    return QB_OK;
}

int tile_profile_end(Reg* r) {
    QB_CUDA(cudaEventRecord(r->tile_events[r->tile_events_used + 1], r->stream));
    r->tile_events_used += 2;
    r->st_tile_launches++;
    return QB_OK;
}

int drain_tile_events(Reg* r) {
    if (!r->tile_events_used) return QB_OK;
    QB_CUDA(cudaEventSynchronize(r->tile_events[r->tile_events_used - 1]));
    for (size_t i = 0; i + 1 < r->tile_events_used; i += 2) {
        float ms = 0.f;
        QB_CUDA(cudaEventElapsedTime(&ms, r->tile_events[i], r->tile_events[i + 1]));
        r->st_tile_ns += (int64_t)(ms * 1e6);
    }
    r->tile_events_used = 0;
    return QB_OK;
}

int run_fused(Reg* r, std::vector<Gate>& gates, std::vector<Gate>* hold) {
    const int max_ops = (int)(sizeof(PassLarge::ops) / sizeof(TileOp));
    const int max_mats = (int)(sizeof(PassLarge::mats) / sizeof(c64));
    static thread_local PassLarge large;  // host staging; thread_local: different handles may run on different threads
    // A pass neither encoder can take (the second-tier planner is allowed more gates per pass than the shared-memory kernel's
    // program holds, counting on the register-blocked encoder's merging) is cut in two: same tile, same order, two sweeps.
    std::function<int(const PlannedPass&)> launch = [&](const PlannedPass& p) -> int {
        if (p.unfused) return k_apply_gate_unfused(r, gates[p.gate_idx[0]]);
        // a pass with a single gate is cheaper through the streaming kernels (they skip untouched halves)
        if (p.gate_idx.size() == 1 && r->force_path != 2 && r->force_path != 4) {
            const Gate& g = gates[p.gate_idx[0]];
            if (g.kind == GK_DIAG || g.k == 1) return k_apply_gate_unfused(r, g);
        }
        if (r->force_path != 2) {  // register-blocked kernel for 2^12 tiles; force path 2 pins the shared-memory kernel
            bool handled = false;
            QB_TRY(run_pass_v2(r, gates, p.gate_idx, p.tile_pos, p.lc, &handled));
            if (handled) return QB_OK;
        }
        size_t mats = 0;
        for (size_t gi : p.gate_idx) mats += gates[gi].matrix.size();
        if (p.gate_idx.size() <= sizeof(PassSmall::ops) / sizeof(TileOp) && mats <= sizeof(PassSmall::mats) / sizeof(c64)) {
            PassSmall small;
            if (!encode_pass(gates, p, &small)) return set_error(QB_ERR_UNSUPPORTED, -1, "internal: pass encoding failed");
            return launch_pass(r, small);
        }
        if (p.gate_idx.size() <= sizeof(PassLarge::ops) / sizeof(TileOp) && mats <= sizeof(PassLarge::mats) / sizeof(c64)) {
            if (!encode_pass(gates, p, &large)) return set_error(QB_ERR_UNSUPPORTED, -1, "internal: pass encoding failed");
            return launch_pass(r, large);
        }
        if (p.gate_idx.size() < 2) return set_error(QB_ERR_UNSUPPORTED, -1, "internal: pass encoding failed");
        PlannedPass a = p, b = p;
        const size_t half = p.gate_idx.size() / 2;
        a.gate_idx.assign(p.gate_idx.begin(), p.gate_idx.begin() + (long)half);
        b.gate_idx.assign(p.gate_idx.begin() + (long)half, p.gate_idx.end());
        const bool last = r->xchg_last_pass;   // only the second half may carry an exchange
        r->xchg_last_pass = false;
        QB_TRY(launch(a));
        r->xchg_last_pass = last;
        return launch(b);
    };
    // Plan cache.  A gate list that has been planned before (same gates, same register shape) is not planned again: its passes
    // are launched straight from the cache.  The second time a list is seen -- it is evidently being re-applied -- the beam
    // search above looks for a shorter plan, once, and the cache keeps the better one.  QB_PLAN_BEAM=0 disables the search,
    // QB_PLAN_BEAM=k sets the beam width (default 4).
    const char* cache_env = getenv("QB_PLAN_CACHE");
    const bool cacheable = !hold && gates.size() >= 32 && !(cache_env && cache_env[0] == '0');
    uint64_t key = 0;
    if (cacheable) {
        key = plan_key(gates, r->n_local, r->tile_bits);
        std::vector<PlannedPass> plan;
        bool found = false, upgrade = false;
        {
            std::lock_guard<std::mutex> lk(g_plan_mu);
            auto it = g_plans.find(key);
            if (it != g_plans.end() && it->second.n_gates == gates.size()) {
                found = true;
                it->second.seen++;
                const char* be = getenv("QB_PLAN_BEAM");
                const char* bm = getenv("QB_PLAN_BEAM_MIN_QUBITS");   // tests lower it to sizes the oracle can follow
                const uint32_t beam_min = bm ? (uint32_t)atoi(bm) : 28u;
                upgrade = it->second.tier == 0 && it->second.seen >= 2 && r->n_local >= beam_min && gates.size() >= 64 && !(be && be[0] == '0');
                if (upgrade) it->second.tier = 1;   // one attempt, whoever gets here first
                plan = it->second.passes;
            }
        }
        if (found) {
            if (upgrade) {
                const char* be = getenv("QB_PLAN_BEAM");
                const size_t beam = be && atoi(be) > 0 ? (size_t)atoi(be) : 4;
                // more gates per pass than the shared-memory kernel's program holds: the register-blocked encoder merges the
                // phase cascades of a QFT pass (hundreds of gates) into a handful of ops; what it cannot take is cut in two
                const bool big = r->tile_bits == 12 && r->n_local >= 12 && r->force_path == 0;
                std::vector<PlannedPass> better;
                if (beam_plan(gates, r->n_local, r->tile_bits, big, max_ops, max_mats, beam, &better) &&
                    (big ? plan_sweeps(gates, r->n_local, better) < plan_sweeps(gates, r->n_local, plan) : better.size() < plan.size())) {
                    plan = better;
                    std::lock_guard<std::mutex> lk(g_plan_mu);
                    g_plans[key].passes = plan;
                }
            }
            for (size_t i = 0; i < plan.size(); ++i) {
                r->xchg_last_pass = i + 1 == plan.size() && r->xchg_request != nullptr;
                const int rc = launch(plan[i]);
                r->xchg_last_pass = false;
                QB_TRY(rc);
            }
            return QB_OK;
        }
    }
    // Every pass is launched as soon as the NEXT one has been planned: the GPU starts after one pass's worth of planning instead
    // of the whole queue's (0.26 s per 4096 gates on big registers), and the last pass is known when planning ends -- with `hold`
    // (the queue overflowed, more gates are coming) its gates go back to the queue instead, unless it is the only pass.
    int rc = QB_OK;
    bool have_prev = false, launched_any = false;
    PlannedPass prev;
    std::vector<PlannedPass> record;
    plan_passes(gates, r->n_local, r->tile_bits, max_ops, max_mats, [&](PlannedPass&& p) {
        if (rc == QB_OK && have_prev) { rc = launch(prev); launched_any = true; }
        if (cacheable && have_prev) record.push_back(prev);
        prev = std::move(p);
        have_prev = true;
    });
    if (rc != QB_OK) return rc;
    if (have_prev) {
        if (hold && launched_any) {
            for (size_t gi : prev.gate_idx) hold->push_back(gates[gi]);
        } else {
            if (cacheable) {
                record.push_back(prev);
                std::lock_guard<std::mutex> lk(g_plan_mu);
                if (g_plans.size() > 4096) g_plans.clear();   // bound the cache: plans are cheap to recompute
                CachedPlan& c = g_plans[key];
                c.passes = record;
                c.n_gates = gates.size();
                c.tier = 0;
                c.seen = 1;
            }
            r->xchg_last_pass = r->xchg_request != nullptr;   // sharded registers: this pass may carry the exchange that follows it
            const int rc_last = launch(prev);
            r->xchg_last_pass = false;
            QB_TRY(rc_last);
        }
    }
    return QB_OK;
}

// Introspection entry points plan with the streaming planner, or -- QB_PLAN_DEBUG_BEAM=k -- with the beam search of width k, so
// that the CPU tests can replay second-tier plans against the oracle too.
static void plan_for_debug(const std::vector<Gate>& gates, uint32_t n, uint32_t tile_bits, std::vector<PlannedPass>* passes) {
    const int max_ops = (int)(sizeof(PassLarge::ops) / sizeof(TileOp)), max_mats = (int)(sizeof(PassLarge::mats) / sizeof(c64));
    const char* be = getenv("QB_PLAN_DEBUG_BEAM");
    if (be && atoi(be) > 0) {
        const bool big = tile_bits == 12 && n >= 12;
        if (beam_plan(gates, n, tile_bits, big, max_ops, max_mats, (size_t)atoi(be), passes)) return;
        passes->clear();
    }
    plan_passes(gates, n, tile_bits, max_ops, max_mats, passes);
}

// The plan for a register of n index bits with 2^12 tiles (qb_shard.cu plans the virtual register of all shards with it).
void plan_register_passes(const std::vector<Gate>& gates, uint32_t n, std::vector<PlannedPass>* out) {
    plan_passes(gates, n, 12, (int)(sizeof(PassLarge::ops) / sizeof(TileOp)), (int)(sizeof(PassLarge::mats) / sizeof(c64)), out);
}

int64_t plan_debug(uint32_t n, std::vector<Gate>& gates, uint32_t tile_bits, int32_t* pass_of_gate, int32_t* order_in_pass) {
    std::vector<PlannedPass> passes;
    plan_for_debug(gates, n, tile_bits, &passes);
    for (size_t p = 0; p < passes.size(); ++p)
        for (size_t o = 0; o < passes[p].gate_idx.size(); ++o) {
            const Gate& g = gates[passes[p].gate_idx[o]];
            if (g.tag >= 0) { pass_of_gate[g.tag] = (int32_t)p; order_in_pass[g.tag] = (int32_t)o; }
        }
    return (int64_t)passes.size();
}

// Planner introspection (no GPU): every pass of the plan for an n-qubit register, serialised as
//   u32 kind (0 = register-blocked round program follows, 1 = gates applied one by one / by the shared-memory kernel),
//   u32 n_gates, i64 tags[n_gates], u8 tile_pos[12] (kind 0), then the round program (encode_rounds_debug).
int64_t rounds_debug(uint32_t n, std::vector<Gate>& gates, unsigned char* out, uint64_t cap) {
    std::vector<PlannedPass> passes;
    plan_for_debug(gates, n, 12, &passes);
    unsigned char* w = out;
    auto room = [&](uint64_t bytes) { return (uint64_t)(w - out) + bytes <= cap; };
    for (const PlannedPass& p : passes) {
        const uint32_t ng = (uint32_t)p.gate_idx.size();
        if (!room(8 + 8ull * ng + 12)) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "buffer too small");
        unsigned char* kind_at = w;
        uint32_t kind = 1;
        memcpy(w, &kind, 4); w += 4;
        memcpy(w, &ng, 4); w += 4;
        for (size_t gi : p.gate_idx) { const int64_t tag = gates[gi].tag; memcpy(w, &tag, 8); w += 8; }
        if (p.unfused || n < 12) continue;
        // same dispatch rule as run_fused: single diagonal / one-qubit gates go through the streaming kernels
        if (p.gate_idx.size() == 1) {
            const Gate& g = gates[p.gate_idx[0]];
            if (g.kind == GK_DIAG || g.k == 1) continue;
        }
        unsigned char* tp_at = w;
        for (uint32_t j = 0; j < 12; ++j) { const unsigned char b = j < p.tile_pos.size() ? (unsigned char)p.tile_pos[j] : 0; *w++ = b; }
        const int64_t got = encode_rounds_debug(n, gates, p.gate_idx, p.tile_pos, p.lc, w, cap - (uint64_t)(w - out));
        if (got < 0) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "buffer too small");
        if (got == 0) { w = tp_at; continue; }   // falls back to the shared-memory kernel: report as a plain gate list
        kind = 0;
        memcpy(kind_at, &kind, 4);
        w += got;
    }
    return (int64_t)(w - out);
}

}  // namespace qb
