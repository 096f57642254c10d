// qb_stateio.cu -- raw state files: checkpoint / restore of a register's amplitudes.
//
// The reference keeps its state in a public host array (`pub state: Array2<c64>`, src/register.rs:23-37) that drivers
// can serialise however they like; a device-resident 137 GB state needs the engine to stream it.  One file holds one
// rank's shard: a 64-byte header followed by 2^n_local amplitudes as interleaved little-endian f64 (re, im) -- the
// memory image of the reference's `Complex<f64>` vector.  The copy is double-buffered through pinned host memory so
// the device->host DMA of chunk i+1 overlaps the file write of chunk i (and the reverse on load).
//
// Integrity: the header stores a position-weighted 64-bit checksum of the payload's raw words,
// sum_j word_j * (2j+1) mod 2^64, computed on the device (K11 below) from the HBM image on save and again after the
// upload on load.  Integer addition is associative, so the value does not depend on the reduction order, and an odd
// weight makes every single-bit flip visible (2^b * (2j+1) != 0 mod 2^64).  sum|a|^2 of the shard is stored too, for
// tools that read the header; it is informational (a low-order mantissa flip does not move it).
#include <errno.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "qb_common.h"

namespace qb {

namespace {

struct StateFileHeader {  // 64 bytes, little-endian
    char magic[8];        // "QBSTATE1"
    uint32_t version;     // 1
    uint32_t dtype;       // 0 = complex f64 (the only state type)
    uint32_t n;           // logical qubits of the whole register
    uint32_t n_local;     // index bits in this file
    uint32_t world;       // ranks the register was sharded over
    uint32_t rank;        // which shard this is (amplitudes [rank << n_local, (rank+1) << n_local))
    uint64_t count;       // amplitudes in the file = 2^n_local
    double norm2;         // sum |a|^2 over the shard (informational)
    uint64_t checksum;    // sum_j word_j * (2j+1) mod 2^64 over the payload's 2*count 64-bit words
    uint8_t reserved[8];
};
static_assert(sizeof(StateFileHeader) == 64, "header layout");

const char kMagic[8] = {'Q', 'B', 'S', 'T', 'A', 'T', 'E', '1'};
const size_t kChunkBytes = 64u << 20;  // per pinned buffer

struct PinnedPair {
    void* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    ~PinnedPair() {
        for (int i = 0; i < 2; ++i) {
            if (buf[i]) cudaFreeHost(buf[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
        }
    }
    int init(size_t bytes) {
        for (int i = 0; i < 2; ++i) {
            QB_CUDA(cudaMallocHost(&buf[i], bytes));
            QB_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        return QB_OK;
    }
};

struct File {
    FILE* f = nullptr;
    ~File() {
        if (f) fclose(f);
    }
};

// K11: position-weighted checksum of the shard's raw 64-bit words.  HBM-bound streaming read (one 128-bit load per
// amplitude, grid-stride over a grid of SM-count multiples); integer adds only, so any reduction order is exact.
__global__ void __launch_bounds__(256) k11_checksum(const ulonglong2* __restrict__ amps, uint64_t count, unsigned long long* out) {
    unsigned long long acc = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const ulonglong2 v = amps[i];
        acc += v.x * (4 * i + 1) + v.y * (4 * i + 3);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

int device_checksum(Reg* r, uint64_t* out) {
    unsigned long long* d = nullptr;
    QB_CUDA(cudaMalloc(&d, sizeof *d));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, r->device);
    const uint64_t count = 1ull << r->n_local;
    const unsigned blocks = (unsigned)std::min<uint64_t>((count + 255) / 256, (uint64_t)sms * 8);
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof *d, r->stream);
    if (e == cudaSuccess) {
        k11_checksum<<<blocks, 256, 0, r->stream>>>(reinterpret_cast<const ulonglong2*>(r->state), count, d);
        r->st_launches++;
        e = cudaGetLastError();
    }
    unsigned long long h = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, r->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r->stream);
    cudaFree(d);
    QB_CUDA(e);
    *out = h;
    return QB_OK;
}

int io_error(const char* what, const char* path) {
    const int e = errno;
    return set_error(QB_ERR_IO, e, "%s %s: %s", what, path, strerror(e));
}

}  // namespace

int state_save(Reg* r, const char* path) {
    double p0 = 0, p1 = 0;
    QB_TRY(k_probabilities(r, -1, &p0, &p1));  // shard-local: no collective, every rank writes its own file
    StateFileHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, kMagic, 8);
    h.version = 1;
    h.n = r->n;
    h.n_local = r->n_local;
    h.world = (uint32_t)r->world;
    h.rank = (uint32_t)r->rank;
    h.count = 1ull << r->n_local;
    h.norm2 = p0;
    QB_TRY(device_checksum(r, &h.checksum));

    File out;
    out.f = fopen(path, "wb");
    if (!out.f) return io_error("cannot create", path);
    if (fwrite(&h, sizeof h, 1, out.f) != 1) return io_error("cannot write", path);

    const size_t total = (size_t)h.count * sizeof(c64);
    const size_t chunk = std::min(kChunkBytes, total);
    PinnedPair pin;
    QB_TRY(pin.init(chunk));
    const size_t n_chunks = (total + chunk - 1) / chunk;
    const char* src = reinterpret_cast<const char*>(r->state);
    auto len_of = [&](size_t i) { return std::min(chunk, total - i * chunk); };
    QB_CUDA(cudaMemcpyAsync(pin.buf[0], src, len_of(0), cudaMemcpyDeviceToHost, r->stream));
    QB_CUDA(cudaEventRecord(pin.ev[0], r->stream));
    for (size_t i = 0; i < n_chunks; ++i) {
        if (i + 1 < n_chunks) {  // buffer (i+1)&1 was written to the file in iteration i-1: free again
            QB_CUDA(cudaMemcpyAsync(pin.buf[(i + 1) & 1], src + (i + 1) * chunk, len_of(i + 1), cudaMemcpyDeviceToHost, r->stream));
            QB_CUDA(cudaEventRecord(pin.ev[(i + 1) & 1], r->stream));
        }
        QB_CUDA(cudaEventSynchronize(pin.ev[i & 1]));
        if (fwrite(pin.buf[i & 1], 1, len_of(i), out.f) != len_of(i)) {
            const int rc = io_error("cannot write", path);
            cudaStreamSynchronize(r->stream);  // the next chunk's DMA targets a buffer we are about to free
            return rc;
        }
    }
    FILE* f = out.f;
    out.f = nullptr;
    if (fclose(f) != 0) return io_error("cannot close", path);
    return QB_OK;
}

// Opens `path` and validates its header against the handle; nothing of the register is touched.
static int open_and_check(Reg* r, const char* path, File* in, StateFileHeader* h) {
    in->f = fopen(path, "rb");
    if (!in->f) return io_error("cannot open", path);
    if (fread(h, sizeof *h, 1, in->f) != 1) return set_error(QB_ERR_IO, 0, "%s: shorter than a state-file header", path);
    if (memcmp(h->magic, kMagic, 8) != 0 || h->version != 1 || h->dtype != 0)
        return set_error(QB_ERR_IO, 0, "%s: not a version-1 complex-f64 state file", path);
    if (h->n != r->n || h->world != (uint32_t)r->world || h->rank != (uint32_t)r->rank || h->n_local != r->n_local ||
        h->count != (1ull << r->n_local))
        return set_error(QB_ERR_INVALID_ARGUMENT, h->n,
                         "%s holds shard %u/%u of a %u-qubit register; this handle is shard %d/%d of %u qubits", path, h->rank,
                         h->world, h->n, r->rank, r->world, r->n);
    return QB_OK;
}

// qb_load_state calls this BEFORE it drops the gate queue and resets the qubit labels / placement: a missing file or a file
// of another register leaves the handle exactly as it was.
int state_load_check(Reg* r, const char* path) {
    File in;
    StateFileHeader h;
    return open_and_check(r, path, &in, &h);
}

int state_load(Reg* r, const char* path) {
    File in;
    StateFileHeader h;
    QB_TRY(open_and_check(r, path, &in, &h));

    const size_t total = (size_t)h.count * sizeof(c64);
    const size_t chunk = std::min(kChunkBytes, total);
    PinnedPair pin;
    QB_TRY(pin.init(chunk));
    const size_t n_chunks = (total + chunk - 1) / chunk;
    char* dst = reinterpret_cast<char*>(r->state);
    auto len_of = [&](size_t i) { return std::min(chunk, total - i * chunk); };
    int rc = QB_OK;
    for (size_t i = 0; i < n_chunks && rc == QB_OK; ++i) {
        if (i >= 2) QB_CUDA(cudaEventSynchronize(pin.ev[i & 1]));  // the upload that last used this buffer is done
        if (fread(pin.buf[i & 1], 1, len_of(i), in.f) != len_of(i)) {
            rc = set_error(QB_ERR_IO, 0, "%s: truncated (expected %zu payload bytes)", path, total);
            break;
        }
        QB_CUDA(cudaMemcpyAsync(dst + i * chunk, pin.buf[i & 1], len_of(i), cudaMemcpyHostToDevice, r->stream));
        QB_CUDA(cudaEventRecord(pin.ev[i & 1], r->stream));
    }
    QB_CUDA(cudaStreamSynchronize(r->stream));
    if (rc != QB_OK) return rc;  // the register now holds a partial image; the caller must reset or reload it
    if (fgetc(in.f) != EOF) return set_error(QB_ERR_IO, 0, "%s: trailing bytes after the payload", path);
    uint64_t sum = 0;
    QB_TRY(device_checksum(r, &sum));
    if (sum != h.checksum)
        return set_error(QB_ERR_IO, 0, "%s: payload checksum %016llx does not match the header's %016llx", path,
                         (unsigned long long)sum, (unsigned long long)h.checksum);
    return QB_OK;
}

}  // namespace qb
