// qb_jit.cu -- K2j runtime: program hash -> compiled kernel.  NVRTC (dlopen libnvrtc.so.12) turns the generated source
// (qb_jit_gen.cu) into an sm_100a cubin on a small thread pool; the launching thread loads finished cubins into its device's
// primary context through the driver API (dlopen libcuda.so.1: cuModuleLoadData / cuLaunchKernel on the handle's stream).
// The cache is process-wide and keyed by the program bytes + generator options, so a circuit that is applied again (every
// step of an iterative algorithm, every shot of a sampling loop, every step of the bench) finds its kernels ready.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>

#include "qb_jit.h"
#include "qb_tile.h"

namespace qb {

namespace {

// ---- NVRTC / driver bindings ------------------------------------------------------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    int (*DestroyProgram)(nvrtcProgram*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};
Nvrtc* nvrtc() {
    static Nvrtc* inst = [] {
        Nvrtc* n = new Nvrtc();
        const char* env = getenv("QB_NVRTC_LIB");
        const char* names[] = {env, "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"};
        void* h = nullptr;
        for (const char* nm : names)
            if (nm && nm[0] && (h = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
        if (!h) { n->why = "libnvrtc.so.12 not found (set QB_NVRTC_LIB)"; return n; }
#define QB_SYM(field, name) *(void**)(&n->field) = dlsym(h, name); if (!n->field) { n->why = std::string("missing symbol ") + name; return n; }
        QB_SYM(CreateProgram, "nvrtcCreateProgram") QB_SYM(CompileProgram, "nvrtcCompileProgram") QB_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
        QB_SYM(GetCUBIN, "nvrtcGetCUBIN") QB_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize") QB_SYM(GetProgramLog, "nvrtcGetProgramLog")
        QB_SYM(DestroyProgram, "nvrtcDestroyProgram") QB_SYM(GetErrorString, "nvrtcGetErrorString")
#undef QB_SYM
        n->ok = true;
        return n;
    }();
    return inst;
}

typedef struct CUmod_st* CUmodule;
typedef struct CUfunc_st* CUfunction;
struct Driver {
    int (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    int (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    int (*FuncSetAttribute)(CUfunction, int, int) = nullptr;
    int (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
    int (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
    int (*GetErrorString)(int, const char**) = nullptr;
    bool ok = false;
    std::string why;
};
Driver* driver() {
    static Driver* inst = [] {
        Driver* d = new Driver();
        void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { d->why = "libcuda.so.1 not found"; return d; }
#define QB_SYM(field, name) *(void**)(&d->field) = dlsym(h, name); if (!d->field) { d->why = std::string("missing symbol ") + name; return d; }
        QB_SYM(ModuleLoadData, "cuModuleLoadData") QB_SYM(ModuleGetFunction, "cuModuleGetFunction") QB_SYM(FuncSetAttribute, "cuFuncSetAttribute")
        QB_SYM(LaunchKernel, "cuLaunchKernel") QB_SYM(OccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor")
        QB_SYM(GetErrorString, "cuGetErrorString")
#undef QB_SYM
        d->ok = true;
        return d;
    }();
    return inst;
}
int fail_cu(int e, const char* what) {
    const char* msg = nullptr;
    if (driver()->GetErrorString) driver()->GetErrorString(e, &msg);
    return set_error(QB_ERR_CUDA, e, "%s: driver error %d (%s)", what, e, msg ? msg : "?");
}

// ---- cache ---------------------------------------------------------------------------------------------------------------
struct Key {
    uint64_t a, b;
    bool operator==(const Key& o) const { return a == o.a && b == o.b; }
};
struct KeyHash {
    size_t operator()(const Key& k) const { return (size_t)(k.a ^ (k.b * 0x9e3779b97f4a7c15ull)); }
};
struct Entry {
    std::atomic<int> state{0};   // 0 pending, 1 cubin ready, -1 failed
    std::string source;          // released after compilation
    std::string cubin, log;
    double compile_ms = 0;
    JitOpts opts;
    CUfunction fn[16] = {};      // per device ordinal (primary contexts), loaded by the launching thread
    int per_sm[16] = {};
    std::mutex load_mu;
};

// The pool's workers are detached and spend their idle time waiting on g_cv: none of this state may be destroyed at process
// exit (destroying a condition variable with waiters blocks forever in glibc), so it lives on the heap and is never freed.
std::mutex& g_mu = *new std::mutex;
std::condition_variable& g_cv = *new std::condition_variable;
std::condition_variable& g_idle_cv = *new std::condition_variable;
std::unordered_map<Key, std::shared_ptr<Entry>, KeyHash>& g_cache = *new std::unordered_map<Key, std::shared_ptr<Entry>, KeyHash>;
std::deque<std::shared_ptr<Entry>>& g_jobs = *new std::deque<std::shared_ptr<Entry>>;
std::vector<std::thread>& g_workers = *new std::vector<std::thread>;
std::atomic<int64_t> g_compiled{0}, g_failed{0}, g_pending{0}, g_compile_us{0};

void fnv(const void* data, size_t len, uint64_t* a, uint64_t* b) {
    const unsigned char* p = (const unsigned char*)data;
    for (size_t i = 0; i < len; ++i) {
        *a = (*a ^ p[i]) * 0x100000001b3ull;
        *b = (*b ^ p[i]) * 0x9e3779b97f4a7c15ull;
        *b ^= *b >> 29;
    }
}
Key key_of(const ProgView& p, const JitOpts& o) {
    uint64_t a = 0xcbf29ce484222325ull, b = 0x84222325cbf29ce4ull;
    const uint32_t hdr[12] = {p.lc, p.h, p.n_rounds, p.n_mats, p.n_ops, (uint32_t)o.min_blocks, (uint32_t)o.persistent, (uint32_t)o.prefetch,
                              (uint32_t)o.remote | ((uint32_t)o.pipelined << 8), (uint32_t)o.xchg_bits,
                              (uint32_t)o.xchg_V[0] | ((uint32_t)o.xchg_V[1] << 8) | ((uint32_t)o.xchg_V[2] << 16),
                              (uint32_t)o.xchg_J[0] | ((uint32_t)o.xchg_J[1] << 8) | ((uint32_t)o.xchg_J[2] << 16)};
    fnv(hdr, sizeof(hdr), &a, &b);
    fnv(p.hp, p.h, &a, &b);   // only the entries in use are initialised
    fnv(p.gro_first, 128, &a, &b);
    fnv(p.gro_last, 128, &a, &b);
    fnv(p.rounds, p.n_rounds * sizeof(RoundDesc), &a, &b);
    fnv(p.ops, p.n_ops * sizeof(ROp), &a, &b);
    fnv(p.mats, p.n_mats * sizeof(c64), &a, &b);
    return Key{a, b};
}

void compile(Entry* e) {
    const auto t0 = std::chrono::steady_clock::now();
    Nvrtc* n = nvrtc();
    int state = -1;
    if (!n->ok) {
        e->log = n->why;
    } else {
        nvrtcProgram prog = nullptr;
        int rc = n->CreateProgram(&prog, e->source.c_str(), "qb_jit_pass.cu", 0, nullptr, nullptr);
        if (rc == 0) {
            const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
            rc = n->CompileProgram(prog, 4, opts);
            size_t ls = 0;
            if (n->GetProgramLogSize(prog, &ls) == 0 && ls > 1) { e->log.resize(ls); n->GetProgramLog(prog, &e->log[0]); }
            if (rc == 0) {
                size_t cs = 0;
                if (n->GetCUBINSize(prog, &cs) == 0 && cs > 0) {
                    e->cubin.resize(cs);
                    if (n->GetCUBIN(prog, &e->cubin[0]) == 0) state = 1;
                }
            } else if (e->log.empty()) {
                e->log = n->GetErrorString(rc);
            }
            n->DestroyProgram(&prog);
        } else {
            e->log = n->GetErrorString(rc);
        }
    }
    const auto us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
    e->compile_ms = us / 1e3;
    g_compile_us += us;
    if (state == 1) { g_compiled++; std::string().swap(e->source); } else g_failed++;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        e->state.store(state);
        g_pending--;
    }
    g_idle_cv.notify_all();
}

void worker() {
    for (;;) {
        std::shared_ptr<Entry> job;
        {
            std::unique_lock<std::mutex> lk(g_mu);
            g_cv.wait(lk, [] { return !g_jobs.empty(); });
            job = g_jobs.front();
            g_jobs.pop_front();
        }
        compile(job.get());
    }
}

void ensure_workers() {   // g_mu held
    if (!g_workers.empty()) return;
    int n = (int)std::thread::hardware_concurrency() / 2;
    if (const char* env = getenv("QB_JIT_THREADS")) n = atoi(env);
    n = n < 1 ? 1 : (n > 16 ? 16 : n);
    for (int i = 0; i < n; ++i) {
        g_workers.emplace_back(worker);
        g_workers.back().detach();   // the pool lives as long as the process
    }
}

std::shared_ptr<Entry> lookup_or_submit(const ProgView& p, const JitOpts& o, bool* fresh) {
    const Key k = key_of(p, o);
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(k);
    *fresh = it == g_cache.end();
    if (!*fresh) return it->second;
    std::shared_ptr<Entry> e = std::make_shared<Entry>();
    e->opts = o;
    e->source = jit_generate(p, o);
    g_cache.emplace(k, e);
    g_pending++;
    ensure_workers();
    g_jobs.push_back(e);
    g_cv.notify_one();
    return e;
}

int load_on_device(Entry* e, int device) {
    std::lock_guard<std::mutex> lk(e->load_mu);
    if (e->fn[device]) return QB_OK;
    Driver* d = driver();
    if (!d->ok) return set_error(QB_ERR_UNSUPPORTED, -1, "JIT kernels need the CUDA driver API: %s", d->why.c_str());
    QB_CUDA(cudaFree(nullptr));   // the runtime's primary context is current on this thread from here on
    CUmodule mod = nullptr;
    int rc = d->ModuleLoadData(&mod, e->cubin.data());
    if (rc) return fail_cu(rc, "cuModuleLoadData(jit pass)");
    CUfunction fn = nullptr;
    rc = d->ModuleGetFunction(&fn, mod, "qb_jit_pass");
    if (rc) return fail_cu(rc, "cuModuleGetFunction(qb_jit_pass)");
    // the prefetching ring: three tile buffers + the two done[] flags, 512 threads
    const unsigned smem = e->opts.pipelined ? 3u * 65536u + 128u : (unsigned)(sizeof(c64) << RT);
    const int threads = e->opts.pipelined ? 512 : RTHREADS;
    rc = d->FuncSetAttribute(fn, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, (int)smem);
    if (rc) return fail_cu(rc, "cuFuncSetAttribute(max dynamic shared memory)");
    int per_sm = 0;
    rc = d->OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem);
    if (rc) return fail_cu(rc, "cuOccupancyMaxActiveBlocksPerMultiprocessor");
    e->per_sm[device] = per_sm < 1 ? 1 : per_sm;
    e->fn[device] = fn;
    return QB_OK;
}

}  // namespace

void jit_global_stats(int64_t* compiled, int64_t* failed, int64_t* pending, int64_t* compile_ms) {
    *compiled = g_compiled.load();
    *failed = g_failed.load();
    *pending = g_pending.load();
    *compile_ms = g_compile_us.load() / 1000;
}

int jit_drain() {
    std::unique_lock<std::mutex> lk(g_mu);
    g_idle_cv.wait(lk, [] { return g_pending.load() == 0; });
    return QB_OK;
}

int jit_try_launch(Reg* r, const ProgView& p, const RemoteArgs* remote, bool* launched, const XchgArgs* xchg) {
    *launched = false;
    if (r->jit == 0 || r->n_local < r->jit_min_qubits || r->device < 0 || r->device >= 16) return QB_OK;
    JitOpts o;
    o.min_blocks = r->jit_min_blocks;
    o.persistent = r->jit_persistent;
    o.prefetch = r->jit_persistent ? r->jit_prefetch : 0;
    o.remote = remote != nullptr;
    o.pipelined = (r->jit_pipelined && !remote && !xchg) ? 1 : 0;
    if (o.pipelined) o.persistent = o.prefetch = 0;
    if (xchg) {
        o.xchg_bits = xchg->gp;
        for (int i = 0; i < xchg->gp; ++i) { o.xchg_V[i] = xchg->V[i]; o.xchg_J[i] = xchg->J[i]; }
        o.persistent = o.prefetch = 0;
    }
    // An exchanging pass is a collective decision: whether it runs must not depend on how far this rank's compiler got, so its
    // kernel is always waited for, and a failure is an error on every rank alike.
    const bool must_wait = r->jit == 2 || xchg != nullptr;
    bool fresh = false;
    std::shared_ptr<Entry> e = lookup_or_submit(p, o, &fresh);
    if (fresh) r->st_jit_compiles++;
    if (e->state.load() == 0) {
        if (!must_wait) { r->st_jit_misses++; return QB_OK; }   // asynchronous mode: the interpreter runs this pass
        std::unique_lock<std::mutex> lk(g_mu);
        g_idle_cv.wait(lk, [&] { return e->state.load() != 0; });
    }
    if (e->state.load() < 0) {
        if (must_wait) return set_error(QB_ERR_UNSUPPORTED, -1, "JIT compilation failed: %.600s", e->log.c_str());
        r->st_jit_misses++;
        return QB_OK;
    }
    int rc = load_on_device(e.get(), r->device);
    if (rc != QB_OK) {
        if (must_wait) return rc;
        r->st_jit_misses++;
        return QB_OK;
    }
    const uint64_t tiles = 1ull << (r->n_local - RT);
    if (tiles > 0x7fffffffull) return set_error(QB_ERR_UNSUPPORTED, (int64_t)tiles, "too many tiles for one grid");
    uint32_t tile_or = 0, n_tiles = (uint32_t)tiles;
    unsigned grid = (unsigned)tiles;
    if (o.persistent || o.pipelined) {
        int sms = 148;
        QB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, r->device));
        grid = (unsigned)std::min<uint64_t>(tiles, (uint64_t)sms * (uint64_t)e->per_sm[r->device]);
    }
    RemoteArgs ra;
    void* args[5];
    c64* st = r->state;
    uint32_t rank = 0;
    if (xchg) {   // (st, out table, rank, tile_or, n_tiles)
        memset(&ra, 0, sizeof(ra));
        for (int q = 0; q < 8; ++q) ra.base[q] = xchg->out[q];
        ra.n_local = r->n_local;
        rank = xchg->rank;
        args[0] = &st;
        args[1] = &ra;
        args[2] = &rank;
        args[3] = &tile_or;
        args[4] = &n_tiles;
    } else {
        if (remote) {
            ra = *remote;
            tile_or = ra.tile_or;
            args[0] = &ra;
        } else {
            args[0] = &st;
        }
        args[1] = &tile_or;
        args[2] = &n_tiles;
    }
    if (r->profile) QB_TRY(tile_profile_begin(r));
    const unsigned block = o.pipelined ? 512u : (unsigned)RTHREADS;
    const unsigned smem = o.pipelined ? 3u * 65536u + 128u : (unsigned)(sizeof(c64) << RT);
    rc = driver()->LaunchKernel(e->fn[r->device], grid, 1, 1, block, 1, 1, smem, (void*)r->stream, args, nullptr);
    if (rc) return fail_cu(rc, "cuLaunchKernel(qb_jit_pass)");
    if (r->profile) QB_TRY(tile_profile_end(r));
    r->st_jit_launches++;
    *launched = true;
    return QB_OK;
}

// ---- introspection (no GPU): source of the kernel a serialised round program would get, and an NVRTC compile of it ------
template <class Prog>
static bool deserialise(const unsigned char* bytes, uint64_t len, Prog* prog, uint32_t* n_ops_out) {
    memset(prog, 0, sizeof(*prog));
    if (len < 28) return false;
    uint32_t hdr[5];
    memcpy(hdr, bytes, 20);
    const uint32_t n_rounds = hdr[2], n_ops = hdr[3], n_mats = hdr[4];
    if (n_rounds == 0 || n_rounds > sizeof(prog->rounds) / sizeof(RoundDesc) || n_ops > sizeof(prog->ops) / sizeof(ROp) ||
        n_mats > sizeof(prog->mats) / sizeof(c64))
        return false;
    const uint64_t need = 28 + (uint64_t)n_rounds * sizeof(RoundDesc) + (uint64_t)n_ops * sizeof(ROp) + (uint64_t)n_mats * sizeof(c64) + 256;
    if (len != need) return false;
    prog->lc = hdr[0];
    prog->h = hdr[1];
    prog->n_rounds = n_rounds;
    prog->n_mats = n_mats;
    const unsigned char* q = bytes + 20;
    memcpy(prog->hp, q, 8); q += 8;
    memcpy(prog->rounds, q, n_rounds * sizeof(RoundDesc)); q += n_rounds * sizeof(RoundDesc);
    memcpy(prog->ops, q, n_ops * sizeof(ROp)); q += n_ops * sizeof(ROp);
    memcpy(prog->mats, q, n_mats * sizeof(c64)); q += n_mats * sizeof(c64);
    memcpy(prog->gro_first, q, 128); q += 128;
    memcpy(prog->gro_last, q, 128);
    *n_ops_out = n_ops;
    return true;
}

}  // namespace qb

using namespace qb;

extern "C" {

int64_t qb_jit_source_debug(const void* program, uint64_t len, int32_t min_blocks, int32_t persistent, int32_t prefetch, int32_t remote,
                            int32_t xchg_bits, const uint8_t* xchg_vj, char* out, uint64_t capacity) try {
    if (!program || (capacity && !out)) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "NULL argument");
    static thread_local RProgLarge prog;
    uint32_t n_ops = 0;
    if (!deserialise((const unsigned char*)program, len, &prog, &n_ops)) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "not a serialised round program");
    JitOpts o;
    o.min_blocks = min_blocks;
    o.persistent = persistent == 1;
    o.pipelined = persistent == 2;   // `persistent` = 2 selects the prefetching ring
    o.prefetch = prefetch;
    o.remote = remote;
    if (xchg_bits < 0 || xchg_bits > 3 || (xchg_bits && !xchg_vj)) return set_error(QB_ERR_INVALID_ARGUMENT, xchg_bits, "0..3 exchanged bits");
    o.xchg_bits = xchg_bits;
    for (int i = 0; i < xchg_bits; ++i) { o.xchg_V[i] = xchg_vj[i]; o.xchg_J[i] = xchg_vj[3 + i]; }
    const std::string src = jit_generate(make_view(prog), o);
    if (src.size() + 1 > capacity) return set_error(QB_ERR_INVALID_ARGUMENT, (int64_t)src.size() + 1, "buffer too small");
    memcpy(out, src.c_str(), src.size() + 1);
    return (int64_t)src.size();
} QB_CATCH_ALL

// Compiles `source` for sm_100a with NVRTC (no GPU needed).  Returns the cubin size, or a negative error with the compiler log
// in qb_last_error(); *ms = compile time.
int64_t qb_jit_compile_debug(const char* source, double* ms, void* cubin_out, uint64_t capacity) try {
    if (!source) return set_error(QB_ERR_INVALID_ARGUMENT, -1, "NULL argument");
    Entry e;
    e.source = source;
    g_pending++;
    compile(&e);
    if (ms) *ms = e.compile_ms;
    if (e.state.load() != 1) return set_error(QB_ERR_UNSUPPORTED, -1, "NVRTC: %.900s", e.log.c_str());
    if (cubin_out && e.cubin.size() <= capacity) memcpy(cubin_out, e.cubin.data(), e.cubin.size());
    return (int64_t)e.cubin.size();
} QB_CATCH_ALL

int qb_jit_drain(void) try { return jit_drain(); } QB_CATCH_ALL

}  // extern "C"
