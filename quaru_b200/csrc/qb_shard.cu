// qb_shard.cu -- one register sharded over the P = 2^g GPUs of one NVLink/NVSwitch box.
//
// No reference equivalent: Quaru is single-process.  Rank r holds the amplitudes
// whose top g PHYSICAL index bits equal r; a logical->physical qubit placement (Reg::phys_of)
// decides which logical qubits are currently "global" (rank bits).
//   * gates whose dense targets are all local run through the single-GPU kernels, no communication;
//   * controls and diagonal gates on global qubits need no exchange either: the rank's bits are
//     constants (rank predicate / sub-table of the diagonal);
//   * a dense target on a global qubit triggers ONE global<->local exchange that swaps every global
//     qubit that is needed sooner than the local qubit it is traded against (look-ahead over the queued
//     gates).  Preferred transport: a PEER-MEMORY kernel -- every rank maps its peers' shards (CUDA IPC) and one
//     kernel swaps the blocks in place over NVLink/NVSwitch with 256-bit loads/stores on peer pointers, each rank of
//     a pair handling half of the block (no staging buffer, no extra copy, all peers concurrently), bracketed by two
//     tiny NCCL all-gathers as barriers.  Fallback (QB_SHARD_TRANSPORT=nccl or IPC unavailable): pairwise
//     ncclSend/ncclRecv of contiguous runs through a staging buffer carved from spare HBM;
//   * measurement: per-rank (p0, p1) -> ncclAllGather -> the same adjacent-pair tree on every rank.
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch already loaded, else the system's),
// so a single-GPU process never needs it.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "qb_jit.h"
#include "qb_shard.h"
#include "qb_tile.h"
#include "qb_device.cuh"

namespace qb {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static std::mutex g_nccl_mutex;

static int load_nccl() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);   // registers may be created from several host threads
    if (g_nccl.handle) return QB_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return set_error(QB_ERR_NCCL, -1, "cannot load libnccl.so.2: %s", dlerror());
#define QB_SYM(field, name)                                                                       \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                                    \
    if (!g_nccl.field) return set_error(QB_ERR_NCCL, -1, "libnccl.so.2 lacks symbol %s", name)
    QB_SYM(GetUniqueId, "ncclGetUniqueId");
    QB_SYM(CommInitRank, "ncclCommInitRank");
    QB_SYM(CommDestroy, "ncclCommDestroy");
    QB_SYM(Send, "ncclSend");
    QB_SYM(Recv, "ncclRecv");
    QB_SYM(AllGather, "ncclAllGather");
    QB_SYM(GroupStart, "ncclGroupStart");
    QB_SYM(GroupEnd, "ncclGroupEnd");
    QB_SYM(GetErrorString, "ncclGetErrorString");
#undef QB_SYM
    g_nccl.handle = h;
    return QB_OK;
}

#define QB_NCCL(expr)                                                                                     \
    do {                                                                                                  \
        ncclResult_t _r = (expr);                                                                         \
        if (_r != ncclSuccess)                                                                            \
            return set_error(QB_ERR_NCCL, (int64_t)_r, "NCCL error %d (%s) in %s", (int)_r,               \
                             g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?", #expr);             \
    } while (0)

struct Shard {
    ncclComm_t comm = nullptr;
    uint32_t g = 0;               // log2(world)
    c64* staging = nullptr;       // receive staging for exchanges
    uint64_t staging_amps = 0;
    double* d_gather = nullptr;   // 2 + 2*world doubles
    std::vector<uint32_t> logical_at;  // physical position -> logical qubit
    c64* peer[8] = {nullptr};          // peer-mapped shards (CUDA IPC), index = rank; peer[rank] = own state
    bool p2p = false;                  // peer-memory exchange available
    bool dry = false;                  // planner introspection: record exchanges, move no data
    std::vector<int32_t> events;       // dry run: [gate index, g', V..., J...] per exchange
    int64_t cur_gate = -1;
    int64_t dry_passes = 0;            // dry run: state sweeps the local planner would need for the epochs so far
    // second state buffer for exchanges carried by a tile pass's stores (see shard_run_reordered); nullptr if it does not fit
    c64* alt = nullptr;
    c64* peer_alt[8] = {nullptr};      // every rank's second buffer as mapped here; peer_alt[rank] = alt
    bool fused_ok = false;             // all ranks have it, mapped everywhere
    // placement memory (see shard_run): gate-list hash -> the placement the list was first run from
    struct Seen { uint64_t hash; std::vector<uint32_t> phys_of; };
    std::vector<Seen> seen;
};

// dry run only: how many sweeps the fused local path would need for this batch of mapped gates
static int64_t count_passes(Reg* r, std::vector<Gate>& batch) {
    std::vector<int32_t> pass_of(batch.size()), order(batch.size());
    for (size_t i = 0; i < batch.size(); ++i) batch[i].tag = (int64_t)i;
    return plan_debug(r->n_local, batch, 12, pass_of.data(), order.data());
}

int shard_unique_id(void* out128) {
    if (!out128) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "out128 is NULL");
    QB_TRY(load_nccl());
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    QB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, 128);
    return QB_OK;
}

// Maps buffer `mine` of every rank into this process: CUDA IPC handles all-gathered through NCCL (SPMD), or the peers' plain
// device pointers after cudaDeviceEnablePeerAccess (in-process registers).  COLLECTIVE: every rank calls it, also one whose
// `mine` is nullptr (it publishes a handle nobody can open).  *opened = number of ranks whose buffer is reachable from here.
static int map_peer_buffers(Reg* r, c64* mine, c64* const* group_ptrs, c64** table, int* opened) {
    Shard* s = r->shard;
    *opened = 0;
    if (r->group) {
        bool ok = mine != nullptr;
        for (int p = 0; p < r->world; ++p) {
            table[p] = group_ptrs[p];
            ok = ok && group_ptrs[p] != nullptr;
            if (p == r->rank) continue;
            const cudaError_t e = cudaDeviceEnablePeerAccess(group_member(r->group, p)->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
            cudaGetLastError();
        }
        *opened = ok ? r->world : 0;
        return QB_OK;
    }
    cudaIpcMemHandle_t h, all[8];
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    bool ok = mine != nullptr && cudaIpcGetMemHandle(&h, mine) == cudaSuccess;
    if (!ok) memset(&h, 0, sizeof(h));
    cudaGetLastError();
    char* dbuf = nullptr;
    QB_CUDA(cudaMalloc((void**)&dbuf, 64 * (r->world + 1)));
    const bool moved = cudaMemcpyAsync(dbuf, &h, 64, cudaMemcpyHostToDevice, r->stream) == cudaSuccess &&
                       g_nccl.AllGather(dbuf, dbuf + 64, 64, ncclChar, s->comm, r->stream) == ncclSuccess &&
                       cudaMemcpyAsync(all, dbuf + 64, 64 * r->world, cudaMemcpyDeviceToHost, r->stream) == cudaSuccess &&
                       cudaStreamSynchronize(r->stream) == cudaSuccess;
    cudaFree(dbuf);
    if (!moved) return set_error(QB_ERR_NCCL, -1, "exchange of the shards' IPC handles failed");
    if (ok) {
        for (int p = 0; p < r->world; ++p) {
            if (p == r->rank) { table[p] = mine; ++*opened; continue; }
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) break;
            table[p] = (c64*)ptr;
            ++*opened;
        }
    }
    cudaGetLastError();   // a failed IPC call must not poison later launches
    return QB_OK;
}

// all ranks agree on a yes / no (a two-double all-gather)
static int agree_all(Reg* r, bool mine, bool* all) {
    Shard* s = r->shard;
    double h[2] = {mine ? 1.0 : 0.0, 0.0};
    std::vector<double> flags(2 * r->world);
    QB_CUDA(cudaMemcpyAsync(s->d_gather, h, sizeof(h), cudaMemcpyHostToDevice, r->stream));
    QB_NCCL(g_nccl.AllGather(s->d_gather, s->d_gather + 2, 2, ncclDouble, s->comm, r->stream));
    QB_CUDA(cudaMemcpyAsync(flags.data(), s->d_gather + 2, 2 * r->world * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    QB_CUDA(cudaStreamSynchronize(r->stream));
    *all = true;
    for (int p = 0; p < r->world; ++p) *all = *all && flags[2 * p] == 1.0;
    return QB_OK;
}

static void unmap_peer_buffers(Reg* r, c64** table) {
    for (int p = 0; p < 8; ++p) {
        if (table[p] && p != r->rank && !r->group) cudaIpcCloseMemHandle(table[p]);
        table[p] = nullptr;
    }
    cudaGetLastError();
}

int shard_init(Reg* r, const void* unique_id) {
    QB_TRY(load_nccl());
    Shard* s = new Shard();
    r->shard = s;
    while ((1 << s->g) < r->world) ++s->g;
    ncclUniqueId id;
    memcpy(&id, unique_id, 128);
    QB_NCCL(g_nccl.CommInitRank(&s->comm, r->world, id, r->rank));
    // staging: up to 1/8 of the shard, at most 4 GiB, at most half of what is still free
    size_t free_b = 0, total_b = 0;
    QB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t shard_amps = 1ull << r->n_local;
    uint64_t amps = std::max<uint64_t>(shard_amps >> 3, (uint64_t)(r->world - 1));
    amps = std::min<uint64_t>(amps, (4ull << 30) / sizeof(c64));
    while (amps > (uint64_t)r->world && amps * sizeof(c64) > free_b / 2) amps >>= 1;
    amps = std::max<uint64_t>(amps, (uint64_t)r->world);
    QB_CUDA(cudaMalloc((void**)&s->staging, amps * sizeof(c64)));
    s->staging_amps = amps;
    QB_CUDA(cudaMalloc((void**)&s->d_gather, (2 + 2 * r->world) * sizeof(double)));
    s->logical_at.resize(r->n);
    for (uint32_t q = 0; q < r->n; ++q) s->logical_at[q] = q;
    // peer-memory transport: every rank maps every peer's shard
    const char* tr = getenv("QB_SHARD_TRANSPORT");
    if (!(tr && !strcmp(tr, "nccl")) && r->world <= 8) {
        c64* gp[8] = {nullptr};
        if (r->group) {
            group_barrier(r->group);   // every rank has allocated its shard
            for (int p = 0; p < r->world; ++p) gp[p] = group_member(r->group, p)->state;
        }
        int opened = 0;
        QB_TRY(map_peer_buffers(r, r->state, gp, s->peer, &opened));
        QB_TRY(agree_all(r, opened == r->world, &s->p2p));   // all ranks must agree on the transport
        if (!s->p2p) {   // agreed fallback to NCCL send/recv: unmap whatever was mapped, then let nobody run ahead and
                         // free a shard that a slower peer still has mapped
            unmap_peer_buffers(r, s->peer);
            QB_NCCL(g_nccl.AllGather(s->d_gather, s->d_gather + 2, 2, ncclDouble, s->comm, r->stream));
            QB_CUDA(cudaStreamSynchronize(r->stream));
        }
    }
    // Second state buffer (exchanges carried by a pass's stores): only where it fits beside the shard and only if every rank
    // got and mapped one.  Measured (33 qubits): 2 GPUs 1799 -> 1893 gates/s, 4 GPUs 2971 -> 2772 -- with more than one peer
    // the pass's posted stores reach ~480 GB/s per direction against the exchange kernel's ~630 -- so it is the default on 2
    // GPUs only; QB_SHARD_FUSED_EXCHANGE=1 / 0 forces it on / off.
    const char* fe = getenv("QB_SHARD_FUSED_EXCHANGE");
    const bool want_fused = fe ? fe[0] == '1' : r->world == 2;
    if (s->p2p && want_fused && r->n_local >= 12) {
        const size_t shard_bytes = sizeof(c64) << r->n_local;
        QB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        if (free_b > shard_bytes + (6ull << 30) && cudaMalloc((void**)&s->alt, shard_bytes) != cudaSuccess) s->alt = nullptr;
        cudaGetLastError();
        c64* gp[8] = {nullptr};
        if (r->group) {
            group_barrier(r->group);   // every rank has tried
            for (int p = 0; p < r->world; ++p) gp[p] = group_member(r->group, p)->shard->alt;
            group_barrier(r->group);
        }
        int opened = 0;
        QB_TRY(map_peer_buffers(r, s->alt, gp, s->peer_alt, &opened));
        QB_TRY(agree_all(r, opened == r->world, &s->fused_ok));
        if (!s->fused_ok) {
            unmap_peer_buffers(r, s->peer_alt);
            QB_NCCL(g_nccl.AllGather(s->d_gather, s->d_gather + 2, 2, ncclDouble, s->comm, r->stream));   // nobody frees a buffer a peer still maps
            QB_CUDA(cudaStreamSynchronize(r->stream));
            if (s->alt) cudaFree(s->alt);
            s->alt = nullptr;
        }
    }
    return QB_OK;
}

// Back to the identity placement without moving data (the caller is about to overwrite the state).
void shard_reset_placement(Reg* r) {
    for (uint32_t q = 0; q < r->n; ++q) r->phys_of[q] = q;
    if (r->shard)
        for (uint32_t q = 0; q < r->n; ++q) r->shard->logical_at[q] = q;
}

void shard_destroy(Reg* r) {
    Shard* s = r->shard;
    if (!s) return;
    unmap_peer_buffers(r, s->peer);
    unmap_peer_buffers(r, s->peer_alt);
    // nobody frees its shard before every peer has unmapped it; keyed on the transport all ranks agreed on, so
    // that every rank makes the same collective calls
    if (s->p2p && s->comm && s->d_gather && g_nccl.AllGather) {
        g_nccl.AllGather(s->d_gather, s->d_gather + 2, 2, ncclDouble, s->comm, r->stream);
        cudaStreamSynchronize(r->stream);
    }
    if (s->alt) cudaFree(s->alt);
    if (s->staging) cudaFree(s->staging);
    if (s->d_gather) cudaFree(s->d_gather);
    if (s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
    delete s;
    r->shard = nullptr;
}

// ------------------------------------------------------------------------------------------------
// K9: peer-memory exchange kernel.  For every peer p (xor-distance d = 1 .. 2^g'-1) this rank swaps its block
// a = bits(p) with the peer's block b = bits(me): element e of the block pair is handled by the LOWER rank if e is in
// the first half of the block and by the higher rank otherwise, so each amplitude is moved exactly once and both
// directions of every NVLink are loaded equally.  A thread moves two amplitudes per access (256-bit).
// ------------------------------------------------------------------------------------------------
struct P2PArgs {
    c64* peer[8];          // by xor-distance index d-1: the peer's shard
    uint64_t my_off[8];    // offset of block a (the one I give away) inside my shard, without the run base
    uint64_t peer_off;     // offset of block b (the one every peer gives me) inside THEIR shard
    uint8_t lower[8];      // 1 if this rank is the lower rank of the pair
    uint32_t npeer, gp;
    uint32_t V[3];         // ascending local bit positions being traded
    uint32_t log_run;      // log2(run length) = V[0]
    uint64_t block_amps;   // amplitudes per block = 2^(n_local - gp)
};

__global__ void __launch_bounds__(256) k9_exchange_p2p(c64* __restrict__ mine, P2PArgs a) {
    const uint64_t half_vecs = a.block_amps >> 2;                 // 2-amp vectors in half a block
    const uint64_t total = half_vecs * a.npeer;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
        const uint32_t d = (uint32_t)(w / half_vecs);             // which peer
        uint64_t e = (w - (uint64_t)d * half_vecs) << 1;          // element inside my half of the block
        if (!a.lower[d]) e += a.block_amps >> 1;                  // the higher rank takes the second half
        // element e of a block -> offset: low log_run bits are contiguous, the rest is deposited around the traded bits
        uint64_t x = e >> a.log_run << a.log_run;
        const uint64_t within = e - x;
        for (uint32_t i = 0; i < a.gp; ++i) x = insert_zero_bit(x, a.V[i]);
        const uint64_t lo = x + within + a.my_off[d], ro = x + within + a.peer_off;
        c64* rp = a.peer[d];
        const c64x2 L = ld256(mine + lo), R = ld256(rp + ro);
        st256(mine + lo, R);
        st256(rp + ro, L);
    }
}

// ------------------------------------------------------------------------------------------------
// the exchange: swap local physical positions V[i] with rank bits J[i]
// ------------------------------------------------------------------------------------------------
static int barrier_all(Reg* r);
static int exchange_p2p(Reg* r, const std::vector<uint32_t>& Vs, const std::vector<uint32_t>& Js, uint32_t mybits) {
    Shard* s = r->shard;
    const uint32_t gp = (uint32_t)Vs.size();
    P2PArgs a;
    memset(&a, 0, sizeof(a));
    a.npeer = (1u << gp) - 1;
    a.gp = gp;
    for (uint32_t i = 0; i < gp; ++i) a.V[i] = Vs[i];
    a.log_run = Vs[0];
    a.block_amps = 1ull << (r->n_local - gp);
    auto block_off = [&](uint32_t bits) {
        uint64_t o = 0;
        for (uint32_t i = 0; i < gp; ++i)
            if ((bits >> i) & 1) o |= 1ull << Vs[i];
        return o;
    };
    a.peer_off = block_off(mybits);
    for (uint32_t d = 1; d <= a.npeer; ++d) {
        const uint32_t ab = mybits ^ d;
        int p = r->rank;
        for (uint32_t i = 0; i < gp; ++i) { p &= ~(1 << Js[i]); p |= (int)((ab >> i) & 1) << Js[i]; }
        a.peer[d - 1] = s->peer[p];
        a.my_off[d - 1] = block_off(ab);
        a.lower[d - 1] = r->rank < p;
    }
    if (a.block_amps < 4 || Vs[0] < 1) return -1;  // tiny registers / a traded bit 0 (no 2-amplitude vectors): NCCL path
    QB_TRY(barrier_all(r));           // every rank has finished the kernels that precede the exchange
    const uint64_t total = (a.block_amps >> 2) * a.npeer;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, r->device);
    const unsigned grid = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)sms * 8);
    k9_exchange_p2p<<<grid, 256, 0, r->stream>>>(r->state, a);
    QB_CUDA(cudaGetLastError());
    r->st_launches++;
    r->st_p2p++;
    QB_TRY(barrier_all(r));           // nobody touches its shard before every peer has finished writing into it
    return QB_OK;
}


static int exchange(Reg* r, std::vector<uint32_t> V, std::vector<uint32_t> J) {
    Shard* s = r->shard;
    const uint32_t gp = (uint32_t)V.size();
    if (!gp) return QB_OK;
    // sort the pairs by local position
    std::vector<size_t> ord(gp);
    for (size_t i = 0; i < gp; ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return V[a] < V[b]; });
    std::vector<uint32_t> Vs(gp), Js(gp);
    for (size_t i = 0; i < gp; ++i) { Vs[i] = V[ord[i]]; Js[i] = J[ord[i]]; }

    if (s->dry) {
        s->events.push_back((int32_t)s->cur_gate);
        s->events.push_back((int32_t)gp);
        for (uint32_t v : Vs) s->events.push_back((int32_t)v);
        for (uint32_t j : Js) s->events.push_back((int32_t)j);
        for (uint32_t i = 0; i < gp; ++i) {
            const uint32_t pl = Vs[i], pg = r->n_local + Js[i];
            const uint32_t ql = s->logical_at[pl], qg = s->logical_at[pg];
            s->logical_at[pl] = qg; s->logical_at[pg] = ql;
            r->phys_of[qg] = pl; r->phys_of[ql] = pg;
        }
        return QB_OK;
    }
    uint32_t mybits = 0;
    for (uint32_t i = 0; i < gp; ++i) mybits |= ((uint32_t)(r->rank >> Js[i]) & 1u) << i;
    const uint32_t npeer = (1u << gp) - 1;
    bool done_p2p = false;
    if (s->p2p) {
        int rc = exchange_p2p(r, Vs, Js, mybits);
        if (rc > 0 || rc < -1) return rc;
        done_p2p = rc == QB_OK;
    }
    const uint64_t run_len = 1ull << Vs[0];
    const uint64_t n_runs = 1ull << (r->n_local - gp - Vs[0]);
    uint64_t chunk = s->staging_amps / npeer;
    if (chunk == 0) return set_error(QB_ERR_OOM, -1, "exchange staging buffer too small");
    uint64_t c2 = 1;
    while (c2 * 2 <= chunk) c2 *= 2;
    chunk = std::min(c2, run_len);
    const uint64_t chunks_per_run = run_len / chunk;

    auto block_off = [&](uint32_t a) {
        uint64_t o = 0;
        for (uint32_t i = 0; i < gp; ++i)
            if ((a >> i) & 1) o |= 1ull << Vs[i];
        return o;
    };
    auto peer_of = [&](uint32_t a) {
        int p = r->rank;
        for (uint32_t i = 0; i < gp; ++i) {
            p &= ~(1 << Js[i]);
            p |= (int)((a >> i) & 1) << Js[i];
        }
        return p;
    };

    for (uint64_t o = 0; o < n_runs && !done_p2p; ++o) {
        uint64_t base = o << Vs[0];
        for (uint32_t i = 0; i < gp; ++i) {
            const uint64_t lo = base & ((1ull << Vs[i]) - 1);
            base = ((base >> Vs[i]) << (Vs[i] + 1)) | lo;
        }
        for (uint64_t c = 0; c < chunks_per_run; ++c) {
            QB_NCCL(g_nccl.GroupStart());
            for (uint32_t d = 1; d <= npeer; ++d) {
                const uint32_t a = mybits ^ d;
                const uint64_t off = base + block_off(a) + c * chunk;
                const int peer = peer_of(a);
                QB_NCCL(g_nccl.Send(r->state + off, chunk * 2, ncclDouble, peer, s->comm, r->stream));
                QB_NCCL(g_nccl.Recv(s->staging + (uint64_t)(d - 1) * chunk, chunk * 2, ncclDouble, peer, s->comm, r->stream));
            }
            QB_NCCL(g_nccl.GroupEnd());
            for (uint32_t d = 1; d <= npeer; ++d) {
                const uint32_t a = mybits ^ d;
                const uint64_t off = base + block_off(a) + c * chunk;
                QB_CUDA(cudaMemcpyAsync(r->state + off, s->staging + (uint64_t)(d - 1) * chunk, chunk * sizeof(c64),
                                        cudaMemcpyDeviceToDevice, r->stream));
            }
        }
    }
    const uint64_t sent = (uint64_t)npeer * (1ull << (r->n_local - gp)) * sizeof(c64);
    r->st_swap_bytes += (int64_t)sent;
    r->st_swaps++;
    // bookkeeping of the placement
    for (uint32_t i = 0; i < gp; ++i) {
        const uint32_t pl = Vs[i], pg = r->n_local + Js[i];
        const uint32_t ql = s->logical_at[pl], qg = s->logical_at[pg];
        s->logical_at[pl] = qg; s->logical_at[pg] = ql;
        r->phys_of[qg] = pl; r->phys_of[ql] = pg;
    }
    return QB_OK;
}

// ------------------------------------------------------------------------------------------------
// mapping a (logical) gate onto this rank's shard
// ------------------------------------------------------------------------------------------------
enum MapResult { MAP_LOCAL, MAP_SKIP, MAP_NEEDS_SWAP };

static MapResult map_gate(const Reg* r, const Gate& g, Gate* out) {
    const uint32_t nl = r->n_local;
    // The exchange decision must be identical on every rank (an exchange is a collective): look at the dense
    // targets FIRST, and only then at this rank's control bits.
    if (g.kind == GK_DENSE)
        for (uint32_t j = 0; j < g.k; ++j)
            if (r->phys_of[g.targets[j]] >= nl) return MAP_NEEDS_SWAP;
    // controls
    uint64_t ctrl = 0;
    for (uint64_t m = g.ctrl_mask; m; m &= m - 1) {
        const uint32_t p = r->phys_of[__builtin_ctzll(m)];
        if (p >= nl) {
            if (!((r->rank >> (p - nl)) & 1)) return MAP_SKIP;  // control is 0 on this rank
        } else {
            ctrl |= 1ull << p;
        }
    }
    out->ctrl_mask = ctrl;
    out->kind = g.kind;
    out->tag = g.tag;
    if (g.kind == GK_DENSE) {
        out->targets.resize(g.k);
        for (uint32_t j = 0; j < g.k; ++j) out->targets[j] = r->phys_of[g.targets[j]];
        out->k = g.k;
        out->matrix = g.matrix;
        return MAP_LOCAL;
    }
    // diagonal: global target bits are fixed by the rank -> sub-table over the local targets
    std::vector<uint32_t> keep;
    uint32_t fixed = 0;
    for (uint32_t j = 0; j < g.k; ++j) {
        const uint32_t p = r->phys_of[g.targets[j]];
        if (p >= nl) fixed |= ((uint32_t)(r->rank >> (p - nl)) & 1u) << j;
        else keep.push_back(j);
    }
    out->k = (uint32_t)keep.size();
    out->targets.resize(out->k);
    out->matrix.resize(1ull << out->k);
    for (uint32_t j = 0; j < out->k; ++j) out->targets[j] = r->phys_of[g.targets[keep[j]]];
    for (uint64_t x = 0; x < (1ull << out->k); ++x) {
        uint32_t idx = fixed;
        for (uint32_t j = 0; j < out->k; ++j)
            if ((x >> j) & 1) idx |= 1u << keep[j];
        out->matrix[x] = g.matrix[idx];
    }
    // a sub-table that is all ones is the identity on this rank
    bool ident = true;
    for (const c64& d : out->matrix)
        if (!(d.x == 1.0 && d.y == 0.0)) { ident = false; break; }
    return ident ? MAP_SKIP : MAP_LOCAL;
}

// Chooses and performs the exchange needed before gates[i] (a dense gate with a global target).
static int choose_swap(Reg* r, const std::vector<Gate>& gates, size_t i, const std::vector<char>* done, std::vector<uint32_t>* Vout,
                       std::vector<uint32_t>* Jout) {
    Shard* s = r->shard;
    const uint32_t nl = r->n_local, n = r->n;
    const size_t INF = (size_t)-1;
    // next use as a DENSE target, per logical qubit, from gate i on (gates already executed out of order do not count)
    std::vector<size_t> next_use(n, INF);
    for (size_t j = gates.size(); j-- > i;)
        if (gates[j].kind == GK_DENSE && !(done && (*done)[j]))
            for (uint32_t t : gates[j].targets) next_use[t] = j;
    // candidates to evict: the top local positions (contiguous runs stay long), widened if needed
    const uint32_t window = std::min<uint32_t>(nl, 10);
    std::vector<uint32_t> cand;
    for (uint32_t p = nl - window; p < nl; ++p) cand.push_back(p);
    std::sort(cand.begin(), cand.end(), [&](uint32_t a, uint32_t b) {
        const size_t ua = next_use[s->logical_at[a]], ub = next_use[s->logical_at[b]];
        if (ua != ub) return ua > ub;  // furthest next use first
        return a > b;                  // then the highest position (longest runs)
    });
    // global qubits ordered by urgency
    std::vector<uint32_t> globals;
    for (uint32_t j = 0; j < s->g; ++j) globals.push_back(j);
    std::sort(globals.begin(), globals.end(), [&](uint32_t a, uint32_t b) {
        const size_t ua = next_use[s->logical_at[nl + a]], ub = next_use[s->logical_at[nl + b]];
        if (ua != ub) return ua < ub;
        return a < b;
    });
    std::vector<uint32_t> V, J;
    size_t ci = 0;
    for (uint32_t j : globals) {
        const size_t ug = next_use[s->logical_at[nl + j]];
        if (ug == INF || ci >= cand.size()) break;
        const size_t uv = next_use[s->logical_at[cand[ci]]];
        if (uv <= ug && ug != i) break;        // no gain from trading these two
        if (uv == i) {                         // the victim itself is needed by this gate: widen the search
            bool found = false;
            for (uint32_t p = nl; p-- > 0 && !found;) {
                if (std::find(V.begin(), V.end(), p) != V.end()) continue;
                if (next_use[s->logical_at[p]] > i) { V.push_back(p); J.push_back(j); found = true; }
            }
            if (!found) return set_error(QB_ERR_UNSUPPORTED, (int64_t)i, "gate touches more qubits than one shard holds");
            continue;
        }
        V.push_back(cand[ci++]);
        J.push_back(j);
    }
    if (V.empty()) return set_error(QB_ERR_UNSUPPORTED, (int64_t)i, "internal: no exchange candidate");
    *Vout = V;
    *Jout = J;
    return QB_OK;
}

static int swap_for(Reg* r, const std::vector<Gate>& gates, size_t i, const std::vector<char>* done = nullptr) {
    std::vector<uint32_t> V, J;
    QB_TRY(choose_swap(r, gates, i, done, &V, &J));
    return exchange(r, V, J);
}

// The exchange has been carried by the stores of the epoch's last tile pass (every amplitude already sits in the destination
// rank's second buffer): wait for every peer's stores, make the second buffers the state, update the placement.
static int finish_fused_exchange(Reg* r, std::vector<uint32_t> V, std::vector<uint32_t> J) {
    Shard* s = r->shard;
    const uint32_t gp = (uint32_t)V.size();
    QB_TRY(barrier_all(r));
    std::swap(r->state, s->alt);
    for (int p = 0; p < 8; ++p) std::swap(s->peer[p], s->peer_alt[p]);
    r->st_swap_bytes += (int64_t)(((1ull << gp) - 1) * (1ull << (r->n_local - gp)) * sizeof(c64));
    r->st_swaps++;
    r->st_fused_exchanges++;
    for (uint32_t i = 0; i < gp; ++i) {
        const uint32_t pl = V[i], pg = r->n_local + J[i];
        const uint32_t ql = s->logical_at[pl], qg = s->logical_at[pg];
        s->logical_at[pl] = qg; s->logical_at[pg] = ql;
        r->phys_of[qg] = pl; r->phys_of[ql] = pg;
    }
    return QB_OK;
}

// Commutation-aware scheduler (the default; QB_SHARD_REORDER=0 selects the first scheduler below): instead of cutting the gate
// list at the first gate that needs an exchange, that gate is DEFERRED and the scan goes on; later gates that are local and
// commute with everything deferred (same read / write rule as the pass planner, on logical qubits) still join the current
// epoch.  Epochs get longer, so the local planner fuses more gates per sweep and fewer exchanges are needed.  Every decision
// depends only on the placement and on the gates' qubit masks, never on this rank's bits, so all ranks stay in step.
static int shard_run_reordered(Reg* r, std::vector<Gate>& gates) {
    Shard* s = r->shard;
    std::vector<char> done(gates.size(), 0);
    size_t remaining = gates.size(), first_open = 0;
    while (remaining) {
        std::vector<Gate> batch;
        uint64_t blocked_w = 0, blocked_r = 0;
        size_t first_deferred = (size_t)-1;
        while (first_open < gates.size() && done[first_open]) ++first_open;
        for (size_t i = first_open; i < gates.size(); ++i) {
            if (done[i]) continue;
            const Gate& g = gates[i];
            const uint64_t wm = g.kind == GK_DENSE ? g.target_mask() : 0;
            const uint64_t rm = g.ctrl_mask | (g.kind == GK_DIAG ? g.target_mask() : 0);
            bool defer = (wm & (blocked_w | blocked_r)) || (rm & blocked_w);
            Gate m;
            MapResult res = MAP_SKIP;
            if (!defer) {
                res = map_gate(r, g, &m);
                if (res == MAP_NEEDS_SWAP) {
                    defer = true;
                    if (first_deferred == (size_t)-1) first_deferred = i;
                }
            }
            if (defer) { blocked_w |= wm; blocked_r |= rm; continue; }
            if (s->dry) s->events.push_back(-1 - (int32_t)(g.tag >= 0 ? g.tag : (int64_t)i));   // run record
            if (res == MAP_LOCAL) batch.push_back(std::move(m));
            done[i] = 1;
            --remaining;
        }
        if (remaining && first_deferred == (size_t)-1) return set_error(QB_ERR_UNSUPPORTED, -1, "internal: scheduler made no progress");
        // The exchange that follows this epoch depends on the placement and the gate list only: it is chosen BEFORE the epoch's
        // passes are launched, so that the last of them can carry it in its stores (every amplitude goes straight to where it
        // lives after the exchange, in the destination rank's second buffer: posted writes over NVLink, overlapped with the
        // pass's compute, instead of a separate pull-and-push sweep of the exchange kernel).
        std::vector<uint32_t> V, J;
        if (remaining) {
            s->cur_gate = gates[first_deferred].tag >= 0 ? gates[first_deferred].tag : (int64_t)first_deferred;
            QB_TRY(choose_swap(r, gates, first_deferred, &done, &V, &J));
        }
        bool fused = false;
        if (!batch.empty() && !s->dry) {
            XchgArgs xa;
            if (remaining && s->fused_ok && V.size() <= 3) {
                std::vector<size_t> ord(V.size());
                for (size_t i = 0; i < ord.size(); ++i) ord[i] = i;
                std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return V[a] < V[b]; });
                xa.gp = (int)V.size();
                for (size_t i = 0; i < ord.size(); ++i) { xa.V[i] = (uint8_t)V[ord[i]]; xa.J[i] = (uint8_t)J[ord[i]]; }
                for (int p = 0; p < r->world; ++p) xa.out[p] = s->peer_alt[p];
                xa.rank = (uint32_t)r->rank;
                r->xchg_request = &xa;
                r->xchg_done = false;
            }
            const int rc = run_local(r, batch);
            r->xchg_request = nullptr;
            QB_TRY(rc);
            fused = r->xchg_done;
            r->xchg_done = false;
        }
        if (!batch.empty() && s->dry) s->dry_passes += count_passes(r, batch);
        if (!remaining) break;
        if (fused) QB_TRY(finish_fused_exchange(r, V, J)); else QB_TRY(exchange(r, V, J));
        Gate probe;
        if (map_gate(r, gates[first_deferred], &probe) == MAP_NEEDS_SWAP)
            return set_error(QB_ERR_UNSUPPORTED, (int64_t)first_deferred, "internal: gate still global after exchange");
    }
    return QB_OK;
}

// First scheduler (QB_SHARD_REORDER=0; also the fallback of the remote scheduler): the gate list is cut at every gate that needs an
// exchange.  Every multi-GPU number of round 1 was measured with it.
static int shard_run_default(Reg* r, std::vector<Gate>& gates) {
    std::vector<Gate> batch;
    for (size_t i = 0; i < gates.size(); ++i) {
        Gate m;
        MapResult res = map_gate(r, gates[i], &m);
        if (res == MAP_NEEDS_SWAP) {
            if (!batch.empty() && !r->shard->dry) QB_TRY(run_local(r, batch));
            if (!batch.empty() && r->shard->dry) r->shard->dry_passes += count_passes(r, batch);
            batch.clear();
            r->shard->cur_gate = gates[i].tag >= 0 ? gates[i].tag : (int64_t)i;
            QB_TRY(swap_for(r, gates, i));
            res = map_gate(r, gates[i], &m);
            if (res == MAP_NEEDS_SWAP) return set_error(QB_ERR_UNSUPPORTED, (int64_t)i, "internal: gate still global after exchange");
        }
        if (res == MAP_LOCAL) batch.push_back(std::move(m));
    }
    if (!batch.empty() && !r->shard->dry) QB_TRY(run_local(r, batch));
    if (!batch.empty() && r->shard->dry) r->shard->dry_passes += count_passes(r, batch);
    return QB_OK;
}

// Experimental (QB_SHARD_REMOTE=1, off by default until it has run on hardware): NO exchanges.  The P shards are planned as
// one virtual register of n = n_local + g index bits: rank bits may become tile bits of a pass like any other high bit, and
// such a pass runs through k2r_tile_pass_remote (qb_tile2_kernel.cuh), which loads and stores the tile's amplitudes that live
// on peers through the CUDA-IPC mappings -- the transfer a global<->local exchange would do up front happens inside the
// compute kernel, tile by tile, overlapped with the (compute-bound) gate work of the other resident CTAs, and the qubit
// placement never changes.  Passes without a rank bit in their tile go through the ordinary local path (rank bits of
// controls and diagonals resolved per rank, as always).  Every decision depends on the placement and the gate list only,
// so all ranks launch the same sequence of passes and barriers.  Whatever the register-blocked encoder cannot take falls
// back to the exchange scheduler for the rest of the list.
static const int32_t kRemoteMark = -(1 << 30);   // dry-run event: [kRemoteMark, n_gates, tags..., 12 tile positions]
static int barrier_all(Reg* r);
static int shard_run_default(Reg* r, std::vector<Gate>& gates);
static int shard_run_remote(Reg* r, std::vector<Gate>& gates) {
    Shard* s = r->shard;
    const uint32_t nl = r->n_local, n = r->n;
    bool ok = (s->dry || s->p2p) && nl >= 12 && r->fusion && r->tile_bits == 12 && r->force_path == 0;
    for (const Gate& g : gates) ok = ok && tile_supported(g);
    if (!ok) return shard_run_default(r, gates);
    std::vector<Gate> v(gates.size());   // the gates at their physical positions in the virtual register
    for (size_t i = 0; i < gates.size(); ++i) {
        const Gate& g = gates[i];
        Gate& o = v[i];
        o.kind = g.kind;
        o.k = g.k;
        o.matrix = g.matrix;
        o.tag = (int64_t)i;
        o.targets.resize(g.k);
        for (uint32_t j = 0; j < g.k; ++j) o.targets[j] = r->phys_of[g.targets[j]];
        o.ctrl_mask = 0;
        for (uint64_t m = g.ctrl_mask; m; m &= m - 1) o.ctrl_mask |= 1ull << r->phys_of[__builtin_ctzll(m)];
    }
    std::vector<PlannedPass> passes;
    plan_register_passes(v, n, &passes);
    std::vector<char> done(gates.size(), 0);
    std::vector<Gate> batch;   // mapped gates of consecutive passes whose tiles are local
    bool dirty = false;        // a remote pass has been launched and no barrier has followed it yet
    auto tag_of = [&](size_t gi) { return (int32_t)(gates[gi].tag >= 0 ? gates[gi].tag : (int64_t)gi); };
    auto flush_local = [&]() -> int {
        if (batch.empty()) return QB_OK;
        if (s->dry) {
            s->dry_passes += count_passes(r, batch);
        } else {
            if (dirty) { QB_TRY(barrier_all(r)); dirty = false; }   // peers may still be storing into this shard
            QB_TRY(run_local(r, batch));
        }
        batch.clear();
        return QB_OK;
    };
    auto fall_back = [&]() -> int {   // the exchange scheduler takes over for everything that has not run yet
        QB_TRY(flush_local());
        if (dirty && !s->dry) { QB_TRY(barrier_all(r)); dirty = false; }
        std::vector<Gate> rest;
        for (size_t i = 0; i < gates.size(); ++i)
            if (!done[i]) { rest.push_back(gates[i]); if (rest.back().tag < 0) rest.back().tag = (int64_t)i; }
        return shard_run_default(r, rest);
    };
    for (const PlannedPass& p : passes) {
        bool remote = false;
        for (uint32_t b : p.tile_pos) remote = remote || b >= nl;
        if (!remote) {
            std::vector<Gate> mapped;
            for (size_t gi : p.gate_idx) {
                Gate m;
                const MapResult res = map_gate(r, gates[gi], &m);
                if (res == MAP_NEEDS_SWAP) return fall_back();   // a gate the tile kernels cannot take with a dense rank-bit target
                if (res == MAP_LOCAL) mapped.push_back(std::move(m));
            }
            for (size_t gi : p.gate_idx) {
                done[gi] = 1;
                if (s->dry) s->events.push_back(-1 - tag_of(gi));
            }
            for (Gate& m : mapped) batch.push_back(std::move(m));
            continue;
        }
        QB_TRY(flush_local());
        bool handled = false;
        if (s->dry) {
            std::vector<unsigned char> buf(64 << 10);
            handled = encode_rounds_debug(n, v, p.gate_idx, p.tile_pos, p.lc, buf.data(), buf.size()) > 0;
        } else {
            QB_TRY(barrier_all(r));   // every peer has finished the kernels that wrote what this pass reads
            dirty = false;
            QB_TRY(run_pass_remote(r, n, v, p.gate_idx, p.tile_pos, p.lc, s->peer, &handled));
            dirty = handled;
        }
        if (!handled) return fall_back();
        if (s->dry) {
            s->dry_passes++;
            s->events.push_back(kRemoteMark);
            s->events.push_back((int32_t)p.gate_idx.size());
            for (size_t gi : p.gate_idx) s->events.push_back(tag_of(gi));
            for (uint32_t j = 0; j < 12; ++j) s->events.push_back((int32_t)p.tile_pos[j]);
        }
        for (size_t gi : p.gate_idx) done[gi] = 1;
    }
    QB_TRY(flush_local());
    if (dirty && !s->dry) QB_TRY(barrier_all(r));   // nobody reads or measures its shard while a peer is still storing into it
    return QB_OK;
}

// ------------------------------------------------------------------------------------------------
// Placement memory.  The exchange schedulers leave the qubits wherever the last exchange put them, so a circuit that is
// applied again (every step of an iterative algorithm, every shot, every step of the bench) would start from another
// placement each time: other physical gate positions, hence other pass programs -- none of the specialised kernels
// compiled for the first application (qb_jit.cu) would ever be reused.  shard_run therefore remembers, per gate-list hash,
// the placement the list was first run from, and when the same list comes back it first moves the qubits to that placement:
// ONE exchange that brings the displaced global qubits home (only the rank bits that differ) and one fused pass of local
// SWAPs.  From there every scheduling decision repeats (they depend on the placement and the gate list only), and so does
// every pass program.  All ranks hold the same maps, so all take the same decisions.  QB_SHARD_PLACEMENT_MEMORY=0 disables it.
// ------------------------------------------------------------------------------------------------
static uint64_t gate_list_hash(const std::vector<Gate>& gates) {
    uint64_t h = 0xcbf29ce484222325ull;
    auto mix = [&](const void* p, size_t n) {
        const unsigned char* b = (const unsigned char*)p;
        for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 0x100000001b3ull;
    };
    for (const Gate& g : gates) {
        mix(&g.k, sizeof g.k);
        mix(&g.kind, sizeof g.kind);
        mix(&g.ctrl_mask, sizeof g.ctrl_mask);
        mix(g.targets.data(), g.targets.size() * sizeof(uint32_t));
        mix(g.matrix.data(), g.matrix.size() * sizeof(c64));
    }
    return h;
}

static void push_local_swap(std::vector<Gate>* batch, uint32_t pa, uint32_t pb);
static int shard_move_to(Reg* r, const std::vector<uint32_t>& target) {   // target: logical qubit -> physical position
    Shard* s = r->shard;
    const uint32_t nl = r->n_local, n = r->n;
    if (r->phys_of == target) return QB_OK;
    std::vector<uint32_t> want_at(n);   // physical position -> the logical qubit that belongs there
    for (uint32_t q = 0; q < n; ++q) want_at[target[q]] = q;
    // (1) rank bits: every displaced home qubit that is local now goes back in ONE exchange
    std::vector<uint32_t> V, J;
    for (uint32_t j = 0; j < s->g; ++j) {
        const uint32_t q = want_at[nl + j];
        if (s->logical_at[nl + j] == q) continue;
        if (r->phys_of[q] >= nl) { V.clear(); J.clear(); break; }   // it sits on another rank bit: general path below
        V.push_back(r->phys_of[q]);
        J.push_back(j);
    }
    bool rank_bits_ok = true;
    if (!V.empty()) QB_TRY(exchange(r, V, J));
    for (uint32_t j = 0; j < s->g; ++j) rank_bits_ok = rank_bits_ok && s->logical_at[nl + j] == want_at[nl + j];
    if (!rank_bits_ok) {
        // rank bits permuted among themselves: go through the identity placement (shard_canonicalize) expressed in the target's
        // qubit names -- rename q -> target[q], restore the identity there, rename back
        std::vector<uint32_t> inv(n);
        for (uint32_t q = 0; q < n; ++q) inv[target[q]] = q;
        std::vector<uint32_t> phys_v(n), at_v(n);
        for (uint32_t q = 0; q < n; ++q) phys_v[target[q]] = r->phys_of[q];
        for (uint32_t p = 0; p < n; ++p) at_v[p] = target[s->logical_at[p]];
        r->phys_of = phys_v;
        s->logical_at = at_v;
        const int rc = shard_canonicalize(r);
        for (uint32_t q = 0; q < n; ++q) phys_v[q] = r->phys_of[target[q]];
        for (uint32_t p = 0; p < n; ++p) at_v[p] = inv[s->logical_at[p]];
        r->phys_of = phys_v;
        s->logical_at = at_v;
        return rc;
    }
    // (2) local positions: SWAP gates on physical positions, fused by the planner into (usually) one pass
    std::vector<Gate> batch;
    for (uint32_t p = 0; p < nl; ++p) {
        const uint32_t q = want_at[p];
        if (s->logical_at[p] == q) continue;
        const uint32_t cur = r->phys_of[q];
        push_local_swap(&batch, p, cur);
        const uint32_t other = s->logical_at[p];
        s->logical_at[p] = q; s->logical_at[cur] = other;
        r->phys_of[q] = p; r->phys_of[other] = cur;
    }
    if (!batch.empty()) QB_TRY(run_local(r, batch));
    if (r->phys_of != target) return set_error(QB_ERR_UNSUPPORTED, -1, "internal: placement restore failed");
    return QB_OK;
}

int shard_run(Reg* r, std::vector<Gate>& gates) {
    const char* remote = getenv("QB_SHARD_REMOTE");
    if (remote && !strcmp(remote, "1")) return shard_run_remote(r, gates);   // never changes the placement
    Shard* s = r->shard;
    const char* pm = getenv("QB_SHARD_PLACEMENT_MEMORY");
    if (!s->dry && r->jit != 0 && r->n_local >= r->jit_min_qubits && gates.size() >= 16 && !(pm && pm[0] == '0') && r->n_local >= 2 * s->g) {
        const uint64_t h = gate_list_hash(gates);
        bool found = false;
        for (size_t i = 0; i < s->seen.size() && !found; ++i)
            if (s->seen[i].hash == h) {
                found = true;
                const Shard::Seen hit = s->seen[i];
                s->seen.erase(s->seen.begin() + (long)i);
                s->seen.push_back(hit);                      // most recently used last
                QB_TRY(shard_move_to(r, hit.phys_of));
            }
        if (!found) {
            if (s->seen.size() >= 16) s->seen.erase(s->seen.begin());
            s->seen.push_back(Shard::Seen{h, r->phys_of});
        }
    }
    const char* reorder = getenv("QB_SHARD_REORDER");   // "0": the first scheduler (cuts the gate list at every exchange)
    if (reorder && !strcmp(reorder, "0")) return shard_run_default(r, gates);
    return shard_run_reordered(r, gates);
}

// Dry run of the placement logic for tests: which exchanges would a rank perform for this gate list?
int64_t shard_plan_debug(uint32_t n, int world, int rank, std::vector<Gate>& gates, int32_t* events, uint64_t capacity,
                         uint32_t* final_phys_of) {
    Reg r;
    Shard s;
    r.n = n;
    r.world = world;
    r.rank = rank;
    while ((1 << s.g) < world) ++s.g;
    if (s.g > n) return set_error(QB_ERR_INVALID_ARGUMENT, world, "more rank bits than qubits");
    r.n_local = n - s.g;
    r.phys_of.resize(n);
    s.logical_at.resize(n);
    for (uint32_t q = 0; q < n; ++q) r.phys_of[q] = s.logical_at[q] = q;
    s.dry = true;
    r.shard = &s;
    int rc = shard_run(&r, gates);
    r.shard = nullptr;
    if (rc != QB_OK) return rc;
    if (s.events.size() > capacity) return set_error(QB_ERR_INVALID_ARGUMENT, (int64_t)s.events.size(), "event buffer too small");
    if (getenv("QB_SHARD_PLAN_STATS")) {   // introspection: trailing record [INT32_MIN, sweeps of the local planner]
        s.events.push_back(INT32_MIN);
        s.events.push_back((int32_t)s.dry_passes);
        if (s.events.size() > capacity) return set_error(QB_ERR_INVALID_ARGUMENT, (int64_t)s.events.size(), "event buffer too small");
    }
    for (size_t i = 0; i < s.events.size(); ++i) events[i] = s.events[i];
    if (final_phys_of)
        for (uint32_t q = 0; q < n; ++q) final_phys_of[q] = r.phys_of[q];
    return (int64_t)s.events.size();
}

// ------------------------------------------------------------------------------------------------
// measurement / reductions across ranks
// ------------------------------------------------------------------------------------------------
static int allgather2(Reg* r, double a, double b, std::vector<double>* out) {
    Shard* s = r->shard;
    double h[2] = {a, b};
    QB_CUDA(cudaMemcpyAsync(s->d_gather, h, sizeof(h), cudaMemcpyHostToDevice, r->stream));
    QB_NCCL(g_nccl.AllGather(s->d_gather, s->d_gather + 2, 2, ncclDouble, s->comm, r->stream));
    out->resize(2 * r->world);
    QB_CUDA(cudaMemcpyAsync(out->data(), s->d_gather + 2, 2 * r->world * sizeof(double), cudaMemcpyDeviceToHost, r->stream));
    QB_CUDA(cudaStreamSynchronize(r->stream));
    return QB_OK;
}

static int barrier_all(Reg* r) {   // stream-ordered barrier across ranks: a two-double all-gather
    Shard* s = r->shard;
    QB_NCCL(g_nccl.AllGather(s->d_gather, s->d_gather + 2, 2, ncclDouble, s->comm, r->stream));
    return QB_OK;
}

int shard_probabilities(Reg* r, uint32_t logical_target, double* p0, double* p1) {
    double a = 0.0, b = 0.0;
    if (logical_target == UINT32_MAX) {
        QB_TRY(k_probabilities(r, -1, &a, &b));
    } else {
        const uint32_t p = r->phys_of[logical_target];
        if (p < r->n_local) {
            QB_TRY(k_probabilities(r, (int)p, &a, &b));
        } else {
            double nrm, z;
            QB_TRY(k_probabilities(r, -1, &nrm, &z));
            if ((r->rank >> (p - r->n_local)) & 1) b = nrm; else a = nrm;
        }
    }
    std::vector<double> all;
    QB_TRY(allgather2(r, a, b, &all));
    // adjacent-pair tree over the rank index: the continuation of the on-device tree
    std::vector<double> v0(r->world), v1(r->world);
    for (int i = 0; i < r->world; ++i) { v0[i] = all[2 * i]; v1[i] = all[2 * i + 1]; }
    for (int len = r->world / 2; len >= 1; len /= 2)
        for (int i = 0; i < len; ++i) { v0[i] = v0[2 * i] + v0[2 * i + 1]; v1[i] = v1[2 * i] + v1[2 * i + 1]; }
    *p0 = v0[0];
    *p1 = v1[0];
    return QB_OK;
}

int shard_collapse(Reg* r, uint32_t logical_target, int outcome, double total_prob) {
    const uint32_t p = r->phys_of[logical_target];
    if (p < r->n_local) return k_collapse(r, p, outcome, total_prob);
    const int bit = (r->rank >> (p - r->n_local)) & 1;
    return k_scale_or_zero_all(r, bit == outcome, total_prob);
}

int shard_reduce_diff(Reg* r, double* max_abs, double* l2sq) {
    std::vector<double> all;
    QB_TRY(allgather2(r, *max_abs, *l2sq, &all));
    double mx = 0.0, l2 = 0.0;
    for (int i = 0; i < r->world; ++i) {
        if (all[2 * i] != all[2 * i] || mx != mx) mx = NAN; else mx = std::max(mx, all[2 * i]);
        l2 += all[2 * i + 1];
    }
    *max_abs = mx;
    *l2sq = l2;
    return QB_OK;
}

// A local bit permutation expressed as SWAP gates on physical positions (fused by the planner).
static void push_local_swap(std::vector<Gate>* batch, uint32_t pa, uint32_t pb) {
    Gate g;
    g.kind = GK_DENSE;
    g.k = 2;
    g.targets = {pa, pb};
    g.matrix.assign(16, make_double2(0.0, 0.0));
    g.matrix[0] = g.matrix[1 * 4 + 2] = g.matrix[2 * 4 + 1] = g.matrix[15] = make_double2(1.0, 0.0);
    batch->push_back(std::move(g));
}

// Restore the identity placement (needed before raw state access).  W = the logical qubits that belong on
// the rank bits.  (1) local permutation: the top g local positions hold non-W qubits; (2) exchange them with
// the rank bits: every W qubit is now local; (3) local permutation: top position nl-g+j holds logical nl+j;
// (4) exchange again: rank bit j holds nl+j; (5) local permutation to the identity.  Steps (1)-(4) are skipped
// when the rank bits are already home.  Local permutations are SWAP gates, fused by the planner.
int shard_canonicalize(Reg* r) {
    Shard* s = r->shard;
    if (!s) return QB_OK;
    const uint32_t nl = r->n_local, g = s->g, n = r->n;
    bool identity = true;
    for (uint32_t q = 0; q < n; ++q) identity = identity && r->phys_of[q] == q;
    if (identity || g == 0) return QB_OK;
    std::vector<Gate> batch;
    auto local_swap = [&](uint32_t pa, uint32_t pb) {
        if (pa == pb) return;
        push_local_swap(&batch, pa, pb);
        const uint32_t qa = s->logical_at[pa], qb = s->logical_at[pb];
        s->logical_at[pa] = qb; s->logical_at[pb] = qa;
        r->phys_of[qa] = pb; r->phys_of[qb] = pa;
    };
    auto run_batch = [&]() -> int {
        if (batch.empty()) return QB_OK;
        int rc = run_local(r, batch);
        batch.clear();
        return rc;
    };
    bool home = true;
    for (uint32_t j = 0; j < g; ++j) home = home && s->logical_at[nl + j] == nl + j;
    if (!home) {
        if (nl < 2 * g) return set_error(QB_ERR_UNSUPPORTED, -1, "canonicalize needs n_local >= 2 log2(world)");
        std::vector<uint32_t> V, J;
        for (uint32_t j = 0; j < g; ++j) { V.push_back(nl - g + j); J.push_back(j); }
        // (1)
        uint32_t below = 0;
        for (uint32_t p = nl - g; p < nl; ++p) {
            if (s->logical_at[p] < nl) continue;  // already a non-W qubit
            while (below < nl - g && s->logical_at[below] >= nl) ++below;
            local_swap(p, below++);
        }
        QB_TRY(run_batch());
        QB_TRY(exchange(r, V, J));  // (2)
        // (3)
        for (uint32_t j = 0; j < g; ++j) {
            const uint32_t cur = r->phys_of[nl + j];
            if (cur >= nl) return set_error(QB_ERR_UNSUPPORTED, nl + j, "internal: qubit %u still global", nl + j);
            local_swap(nl - g + j, cur);
        }
        QB_TRY(run_batch());
        QB_TRY(exchange(r, V, J));  // (4)
    }
    // (5)
    for (uint32_t p = 0; p < nl; ++p) {
        if (s->logical_at[p] == p) continue;
        const uint32_t cur = r->phys_of[p];
        if (cur >= nl) return set_error(QB_ERR_UNSUPPORTED, p, "internal: qubit %u left global", p);
        local_swap(p, cur);
    }
    QB_TRY(run_batch());
    for (uint32_t q = 0; q < n; ++q)
        if (r->phys_of[q] != q) return set_error(QB_ERR_UNSUPPORTED, q, "internal: canonicalize failed for qubit %u", q);
    return QB_OK;
}

}  // namespace qb
