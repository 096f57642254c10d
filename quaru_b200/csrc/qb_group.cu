// qb_group.cu -- one register sharded over P GPUs from ONE process, behind the stock constructor.
//
// The reference's `Register::new(&[bool])` (src/register.rs:41-56) has no notion of devices, and a drop-in must keep its
// signature: with QB_NUM_GPUS = P (2, 4 or 8) in the environment qb_create builds the same sharded register that
// qb_create_sharded builds with one process per GPU -- same placement logic, schedulers, exchange kernel and measurement tree
// (qb_shard.cu) -- but drives its P shards from P worker threads of the calling process, one per device:
//   * rank p lives on CUDA device p (its own stream, its own gate queue and planner state, as in the SPMD mode);
//   * peers' shards are reached through plain device pointers after cudaDeviceEnablePeerAccess (no CUDA IPC: a process cannot
//     open its own handles); the stream-ordered barriers and the probability all-gather still go through NCCL, whose
//     communicator is created by the P threads with ncclCommInitRank on one shared id;
//   * every C-ABI call on the handle is executed by all P workers on their shard (the calls are collective, exactly as the
//     header requires of SPMD callers) and returns rank 0's result; reads and writes of amplitudes use the caller's buffer
//     directly, each rank transferring the range it owns.
// The handle the caller holds is rank 0's Reg.
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "qb_common.h"
#include "qb_shard.h"

namespace qb {

const char* last_error_str();
int64_t last_error_payload();

struct Group {
    int world = 0;
    std::vector<Reg*> members;
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<int(Reg*, int)> job;
    uint64_t epoch = 0;
    int pending = 0;
    bool quit = false;
    std::vector<int> rc;
    std::vector<std::string> msg;
    std::vector<int64_t> payload;
    // a plain barrier for the workers (used while the shards are being set up)
    std::mutex bmu;
    std::condition_variable bcv;
    int bcount = 0;
    uint64_t bgen = 0;
};

static thread_local bool t_in_worker = false;

Group* group_of(Reg* r) { return (r && r->group && !t_in_worker) ? r->group : nullptr; }
int group_world(const Group* g) { return g->world; }
Reg* group_member(Group* g, int rank) { return g->members[rank]; }
void group_set_member(Group* g, int rank, Reg* r) { g->members[rank] = r; }

void group_barrier(Group* g) {
    std::unique_lock<std::mutex> lk(g->bmu);
    const uint64_t gen = g->bgen;
    if (++g->bcount == g->world) {
        g->bcount = 0;
        g->bgen++;
        g->bcv.notify_all();
    } else {
        g->bcv.wait(lk, [&] { return g->bgen != gen; });
    }
}

static void worker(Group* g, int rank) {
    t_in_worker = true;
    cudaSetDevice(rank);
    uint64_t seen = 0;
    for (;;) {
        std::function<int(Reg*, int)> job;
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_go.wait(lk, [&] { return g->quit || g->epoch != seen; });
            if (g->quit) return;
            seen = g->epoch;
            job = g->job;
        }
        int rc;
        try {
            rc = job(g->members[rank], rank);
        } catch (...) {
            rc = translate_exception();
        }
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->rc[rank] = rc;
            if (rc != QB_OK) { g->msg[rank] = last_error_str(); g->payload[rank] = last_error_payload(); }
            if (--g->pending == 0) g->cv_done.notify_all();
        }
    }
}

// Runs fn(member, rank) on every worker; returns the first failure in rank order (re-raised on the calling thread's error
// record) or QB_OK.
int group_run(Group* g, const std::function<int(Reg*, int)>& fn) {
    std::unique_lock<std::mutex> lk(g->mu);
    g->job = fn;
    g->pending = g->world;
    for (int p = 0; p < g->world; ++p) g->rc[p] = QB_OK;
    g->epoch++;
    g->cv_go.notify_all();
    g->cv_done.wait(lk, [&] { return g->pending == 0; });
    for (int p = 0; p < g->world; ++p)
        if (g->rc[p] != QB_OK) return set_error(g->rc[p], g->payload[p], "%s", g->msg[p].c_str());
    return QB_OK;
}

Group* group_create(int world) {
    Group* g = new Group();
    g->world = world;
    g->members.assign(world, nullptr);
    g->rc.assign(world, QB_OK);
    g->msg.resize(world);
    g->payload.assign(world, -1);
    for (int p = 0; p < world; ++p) g->threads.emplace_back(worker, g, p);
    return g;
}

void group_destroy(Group* g) {
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->quit = true;
    }
    g->cv_go.notify_all();
    for (std::thread& t : g->threads) t.join();
    delete g;
}

// QB_NUM_GPUS: 0 / unset / 1 = a single-GPU register
int group_requested_world() {
    const char* env = getenv("QB_NUM_GPUS");
    if (!env) return 1;
    const int p = atoi(env);
    return (p == 2 || p == 4 || p == 8) ? p : 1;
}

}  // namespace qb
