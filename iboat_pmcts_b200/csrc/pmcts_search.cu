// pmcts_b200 -- the leaf-batched forest search as one native call (pmcts_forest_search) and the multi-GPU
// exchange of its update blocks (pmcts_comm_*, pmcts_allgather_updates).
//
// What the reference spreads over one process per scenario and a Manager().dict() proxy
// (code/solver/forest.py:45-81, code/solver/worker.py:147-186, 348-378) is here a loop over waves on one host
// thread per GPU: the host runtime (pmcts_forest_host.cpp) picks the leaves, K2 (pmcts_b200.cu) rolls them out for
// all local scenarios in one launch, the per-leaf histograms come back packed behind the leaves' paths -- the
// "update block" -- and are filed in the worker trees and in the replicated master.  With several ranks the blocks
// are all-gathered on the device (NCCL over NVLink) before the copy back, so every rank files the same blocks in
// the same order.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "pmcts_internal.h"

namespace {

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    pmcts::set_last_error(buf);
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(PMCTS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                    \
    } while (0)
#define LIB_TRY(expr)            \
    do {                         \
        const int rc_ = (expr);  \
        if (rc_) return rc_;     \
    } while (0)
// (a pmcts_forest_* call reports through its own error string: hand it over)
#define FOREST_TRY(expr)                                                  \
    do {                                                                  \
        const int rc_ = (expr);                                           \
        if (rc_) return fail(rc_, "%s", pmcts_forest_last_error());       \
    } while (0)

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// ---- NCCL, loaded on first use ----------------------------------------------------------------------------
struct NcclId {
    char internal[128];
};
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string error;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // (a process that imported torch already holds torch's own libnccl.so.2: dlopen returns that one)
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.error = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
            return;
        }
        api.GetUniqueId = (int (*)(NcclId *))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(void **, int, NcclId, int))dlsym(api.handle, "ncclCommInitRank");
        api.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(api.handle, "ncclAllGather");
        api.CommDestroy = (int (*)(void *))dlsym(api.handle, "ncclCommDestroy");
        api.GetErrorString = (const char *(*)(int))dlsym(api.handle, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy || !api.GetErrorString)
            api.error = "libnccl.so.2 lacks one of ncclGetUniqueId / ncclCommInitRank / ncclAllGather / ncclCommDestroy";
    });
    return &api;
}

#define NCCL_TRY(api, expr)                                                                          \
    do {                                                                                             \
        const int r_ = (expr);                                                                       \
        if (r_ != 0) return fail(PMCTS_ERR_CUDA, "%s failed: %s", #expr, (api)->GetErrorString(r_)); \
    } while (0)

// ---- packing of a wave's histograms --------------------------------------------------------------------------
// leaf_bins[count] uint32 (what K2's atomics produce) -> counts of kBytes bytes each, written behind the leaves'
// paths in the update block.  A count is at most `replicas`, which the driver checks against the width.
template <typename T>
__global__ void pack_bins_kernel(const uint32_t *__restrict__ bins, T *__restrict__ out, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = (T)bins[i];
}

}  // namespace

struct pmcts_comm {
    void *nccl = nullptr;
    int world = 1, rank = 0, device = 0;
};

extern "C" {

size_t pmcts_search_block_bytes(int n_local, int n_leaves, int max_depth, int replicas) {
    if (n_local < 1 || n_leaves < 1 || max_depth < 1 || replicas < 1) return 0;
    const size_t n = (size_t)n_local * n_leaves;
    return align16(n * sizeof(int32_t)) + align16(n * max_depth) + align16(n * 12 * (replicas <= 255 ? 1 : 2));
}

int pmcts_comm_unique_id(uint8_t id[128]) {
    if (!id) return fail(PMCTS_ERR_ARG, "id is NULL");
    NcclApi *api = nccl_api();
    if (!api->error.empty()) return fail(PMCTS_ERR_CUDA, "%s", api->error.c_str());
    NcclId nid;
    NCCL_TRY(api, api->GetUniqueId(&nid));
    std::memcpy(id, nid.internal, 128);
    return PMCTS_OK;
}

int pmcts_comm_create(pmcts_ctx *ctx, int world, int rank, const uint8_t id[128], pmcts_comm **out) {
    if (!ctx || !id || !out) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(PMCTS_ERR_ARG, "bad world / rank");
    NcclApi *api = nccl_api();
    if (!api->error.empty()) return fail(PMCTS_ERR_CUDA, "%s", api->error.c_str());
    int prev = -1;
    cudaGetDevice(&prev);
    const int dev = pmcts_get_device(ctx);
    CUDA_TRY(cudaSetDevice(dev));
    NcclId nid;
    std::memcpy(nid.internal, id, 128);
    void *comm = nullptr;
    const int r = api->CommInitRank(&comm, world, nid, rank);
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    if (r != 0) return fail(PMCTS_ERR_CUDA, "ncclCommInitRank failed: %s", api->GetErrorString(r));
    pmcts_comm *c = new pmcts_comm();
    c->nccl = comm;
    c->world = world;
    c->rank = rank;
    c->device = dev;
    *out = c;
    return PMCTS_OK;
}

void pmcts_comm_destroy(pmcts_comm *comm) {
    if (!comm) return;
    NcclApi *api = nccl_api();
    if (comm->nccl && api->CommDestroy) api->CommDestroy(comm->nccl);
    delete comm;
}

int pmcts_allgather_updates(pmcts_ctx *ctx, pmcts_comm *comm, const void *send, size_t bytes, void *recv) {
    if (!ctx || !comm || !send || !recv) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (bytes == 0) return PMCTS_OK;
    NcclApi *api = nccl_api();
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != comm->device) CUDA_TRY(cudaSetDevice(comm->device));
    const int r = api->AllGather(send, recv, bytes, /* ncclUint8 */ 1, comm->nccl, (cudaStream_t)pmcts_get_stream(ctx));
    if (prev >= 0 && prev != comm->device) cudaSetDevice(prev);
    if (r != 0) return fail(PMCTS_ERR_CUDA, "ncclAllGather failed: %s", api->GetErrorString(r));
    return PMCTS_OK;
}

}  // extern "C"

namespace {

// ---- the search driver ------------------------------------------------------------------------------------------
// Update block of one rank and one wave (sized for the largest wave, B leaves per tree):
//   [ prefix_len int32 [S_loc][B] | paths int8 [S_loc][B][D] | counts uintN [S_loc][B][12] ]
// The head (lengths + paths) is what the rollout launch reads; the tail is written by pack_bins_kernel.
struct Layout {
    int S_loc = 0, B = 0, D = 0, elem = 1;
    size_t o_path = 0, o_bins = 0, bytes = 0;
    void set(int s_loc, int b, int d, int elem_bytes) {
        S_loc = s_loc; B = b; D = d; elem = elem_bytes;
        o_path = align16((size_t)S_loc * B * sizeof(int32_t));
        o_bins = o_path + align16((size_t)S_loc * B * D);
        bytes = o_bins + align16((size_t)S_loc * B * 12 * elem);
    }
};

struct Wave {
    int index = -1, b = 0;
    size_t slot_off = 0;       // offset of this wave in slots_w / slots_m
    // host side (pinned when there is a device)
    char *h_send = nullptr;    // this rank's block: lengths + paths written by the selection (+ counts on the hook path)
    char *h_recv = nullptr;    // [world] blocks as they come back
    // device side
    char *d_send = nullptr, *d_recv = nullptr;
    uint32_t *d_bins = nullptr;  // [S_loc][B][12] what K2 adds to
    cudaEvent_t done = nullptr;
    size_t cap_send = 0, cap_recv = 0, cap_dsend = 0, cap_drecv = 0, cap_bins = 0;
};

// Staging buffers of the driver, kept with the ctx between searches (grow-only): a search of a few milliseconds
// would otherwise spend a third of its time in cudaMallocHost / cudaMalloc / cudaFree.
struct SearchCache {
    Wave ring[2];
    double *d_stats = nullptr;
    size_t cap_stats = 0;
    cudaStream_t comm = nullptr;
    char *seq_host = nullptr, *seq_dev = nullptr;  // the sequential search's per-iteration arrays
    size_t cap_seq = 0, cap_seq_dev = 0;
};

void free_search_cache(void *p) {
    SearchCache *c = (SearchCache *)p;
    for (Wave &w : c->ring) {
        if (w.h_send) cudaFreeHost(w.h_send);
        if (w.h_recv) cudaFreeHost(w.h_recv);
        if (w.d_send) cudaFree(w.d_send);
        if (w.d_recv) cudaFree(w.d_recv);
        if (w.d_bins) cudaFree(w.d_bins);
        if (w.done) cudaEventDestroy(w.done);
    }
    if (c->d_stats) cudaFree(c->d_stats);
    if (c->comm) cudaStreamDestroy(c->comm);
    if (c->seq_host) cudaFreeHost(c->seq_host);
    if (c->seq_dev) cudaFree(c->seq_dev);
    delete c;
}

template <typename T>
int grow_device(T *&p, size_t &cap, size_t need) {
    if (need <= cap) return PMCTS_OK;
    if (p) CUDA_TRY(cudaFree(p));
    p = nullptr;
    cap = 0;
    CUDA_TRY(cudaMalloc((void **)&p, need + need / 2));
    cap = need + need / 2;
    return PMCTS_OK;
}
int grow_pinned(char *&p, size_t &cap, size_t need) {
    if (need <= cap) return PMCTS_OK;
    if (p) CUDA_TRY(cudaFreeHost(p));
    p = nullptr;
    cap = 0;
    CUDA_TRY(cudaMallocHost((void **)&p, need + need / 2));
    cap = need + need / 2;
    return PMCTS_OK;
}

// the recorded iteration of the sequential search
struct SeqGraph {
    cudaGraphExec_t exec = nullptr;
    int kernels = 0;
    ~SeqGraph() { if (exec) cudaGraphExecDestroy(exec); }
};

struct Driver {
    pmcts_ctx *ctx;
    pmcts_forest *f;
    const pmcts_search_params *p;
    const int32_t *scen_slots;
    const double *root, *dest;
    double tmin;
    pmcts_search_report rep{};
    int S = 0, S_loc = 0, D = 0, B = 0, world = 1, rank = 0, depth = 1;
    Layout lay;
    std::vector<Wave> host_ring;  // hook path: plain host buffers owned by this search
    Wave *ring = nullptr;         // device path: the ctx's cached buffers
    std::vector<int32_t> local_ids, leaf_node_scratch, ids_scratch;
    std::vector<uint32_t> bins32;
    double *d_stats = nullptr;
    cudaStream_t main = nullptr, comm = nullptr;
    bool device = false, dev_gather = false;

    ~Driver() { release(); }

    void release() {
        for (Wave &w : host_ring) {
            delete[] w.h_send;
            delete[] w.h_recv;
            w = Wave();
        }
        host_ring.clear();
    }

    int setup() {
        const pmcts::ForestInfo info = pmcts::forest_info(f);
        S = info.n_scen;
        S_loc = info.n_local;
        D = info.max_depth;
        B = p->n_leaves;
        world = std::max(1, p->world);
        rank = p->rank;
        depth = p->pipeline >= 2 ? 2 : 1;
        device = ctx != nullptr;
        if (!device && !p->rollout_hook) return fail(PMCTS_ERR_NO_DEVICE, "pmcts_forest_search needs a ctx: the search has no CPU path");
        if (B < 1 || p->replicas < 1 || p->budget < 1) return fail(PMCTS_ERR_ARG, "n_leaves, replicas and budget must be positive");
        if (p->replicas > 65535) return fail(PMCTS_ERR_ARG, "at most 65535 replicas per leaf (16-bit counts in the update block)");
        if (S != S_loc * world) return fail(PMCTS_ERR_ARG, "%d scenarios do not split into %d ranks of %d", S, world, S_loc);
        if (rank < 0 || rank >= world) return fail(PMCTS_ERR_ARG, "rank %d outside the world of %d", rank, world);
        if (!p->slots_w || !p->slots_m) return fail(PMCTS_ERR_ARG, "slots_w / slots_m are NULL");
        if (world > 1 && !p->exchange && !(device && (p->comm || p->peer)))
            return fail(PMCTS_ERR_ARG, "world > 1 needs a pmcts_peer / pmcts_comm (device) or an exchange callback (host)");
        if (device && !scen_slots) return fail(PMCTS_ERR_ARG, "scen_slots is NULL");
        local_ids.resize(S_loc);
        for (int i = 0; i < S_loc; ++i) {
            local_ids[i] = pmcts::forest_local_id(f, i);
            if (local_ids[i] != rank * S_loc + i)
                return fail(PMCTS_ERR_ARG, "rank %d must own the scenarios %d .. %d in order", rank, rank * S_loc, (rank + 1) * S_loc - 1);
        }
        lay.set(S_loc, B, D, p->replicas <= 255 ? 1 : 2);
        if (lay.bytes != pmcts_search_block_bytes(S_loc, B, D, p->replicas)) return fail(PMCTS_ERR_ARG, "update block layout out of step");
        leaf_node_scratch.resize((size_t)S_loc * B);
        ids_scratch.resize(S_loc);
        bins32.resize((size_t)S_loc * B * 12);
        dev_gather = device && world > 1 && !p->exchange;
        if (dev_gather && p->peer && pmcts::peer_block_bytes(p->peer) < lay.bytes)
            return fail(PMCTS_ERR_ARG, "the peer windows hold blocks of %zu bytes, this search sends %zu (pmcts_search_block_bytes)",
                        pmcts::peer_block_bytes(p->peer), lay.bytes);
        if (device) {
            main = (cudaStream_t)pmcts_get_stream(ctx);
            void **slot = pmcts::ctx_attachment(ctx, free_search_cache);
            if (!*slot) *slot = new SearchCache();
            SearchCache *cache = (SearchCache *)*slot;
            ring = cache->ring;
            if (depth > 1 && !cache->comm) {
                // (above the rollout stream: the tail of a wave is little work, but at equal priority its blocks wait for
                //  block slots until the next wave's rollout kernel has none of its own left to place)
                int least = 0, greatest = 0;
                CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
                CUDA_TRY(cudaStreamCreateWithPriority(&cache->comm, cudaStreamNonBlocking, greatest));
            }
            comm = cache->comm;
            LIB_TRY(grow_device(cache->d_stats, cache->cap_stats, (size_t)S_loc * 8 * sizeof(double)));
            d_stats = cache->d_stats;
            CUDA_TRY(cudaMemsetAsync(d_stats, 0, (size_t)S_loc * 8 * sizeof(double), main));
            for (int k = 0; k < depth; ++k) {
                Wave &w = ring[k];
                LIB_TRY(grow_pinned(w.h_send, w.cap_send, lay.bytes));
                LIB_TRY(grow_pinned(w.h_recv, w.cap_recv, lay.bytes * world));
                LIB_TRY(grow_device(w.d_send, w.cap_dsend, lay.bytes));
                {
                    char *p = (char *)w.d_bins;
                    LIB_TRY(grow_device(p, w.cap_bins, (size_t)S_loc * B * 12 * sizeof(uint32_t)));
                    w.d_bins = (uint32_t *)p;
                }
                if (dev_gather && !p->peer) LIB_TRY(grow_device(w.d_recv, w.cap_drecv, lay.bytes * world));
                if (!w.done) CUDA_TRY(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
                w.index = -1;
                std::memset(w.h_send, 0, lay.bytes);
            }
        } else {
            host_ring.resize(depth);
            ring = host_ring.data();
            for (Wave &w : host_ring) {
                w.h_send = new char[lay.bytes];
                w.h_recv = new char[lay.bytes * world];
                std::memset(w.h_send, 0, lay.bytes);
            }
        }
        return PMCTS_OK;
    }

    // selection of wave `index` on the host, then its rollouts enqueued on the device
    int launch(Wave &w, int index, int b, size_t slot_off) {
        w.index = index;
        w.b = b;
        w.slot_off = slot_off;
        int32_t *plen = (int32_t *)w.h_send;
        int8_t *path = (int8_t *)(w.h_send + lay.o_path);
        double t0 = now_ms();
        FOREST_TRY(pmcts_forest_select(f, b, D, leaf_node_scratch.data(), plen, path));
        rep.ms_select += now_ms() - t0;
        const size_t n_counts = (size_t)S_loc * b * 12;
        if (!device) {
            // test instrumentation: the stand-in fills the histograms; they are packed here as the device would
            std::fill(bins32.begin(), bins32.begin() + n_counts, 0u);
            const int rc = p->rollout_hook(p->hook_user, index, S_loc, local_ids.data(), b, D, plen, path, bins32.data());
            if (rc) return fail(PMCTS_ERR_ARG, "rollout_hook returned %d", rc);
            pack_host(w.h_send + lay.o_bins, bins32.data(), n_counts);
            return PMCTS_OK;
        }
        // lengths + paths up in one copy (the block head, sized for this wave)
        const size_t head = lay.o_path + (size_t)S_loc * b * D;
        CUDA_TRY(cudaMemcpyAsync(w.d_send, w.h_send, head, cudaMemcpyHostToDevice, main));
        CUDA_TRY(cudaMemsetAsync(w.d_bins, 0, n_counts * sizeof(uint32_t), main));
        pmcts_noise nz{};
        nz.mode = PMCTS_NOISE_PHILOX;
        nz.seed = p->seed;
        nz.id0 = 0;
        const int roll_flags = (p->roll_flags & PMCTS_ROLL_REPLAY_FULL_PREFIX) | PMCTS_ROLL_PREFIX_ACTIONS | PMCTS_ROLL_ACCUMULATE |
                               (depth > 1 ? PMCTS_ROLL_DEFER_REDO : 0);
        for (int c0 = 0; c0 < S_loc; c0 += 16) {  // (a launch takes at most 16 scenarios)
            const int nc = std::min(16, S_loc - c0);
            nz.stream = (uint64_t)index * (uint64_t)S + (uint64_t)(rank * S_loc + c0);
            LIB_TRY(pmcts_rollout_forest(ctx, nc, scen_slots + c0, b, p->replicas, root, dest, tmin,
                                         (const double *)(w.d_send + lay.o_path + (size_t)c0 * b * D),
                                         (const int32_t *)w.d_send + (size_t)c0 * b, D, 0, &nz, nullptr, nullptr, nullptr,
                                         nullptr, nullptr, w.d_bins + (size_t)c0 * b * 12, d_stats + (size_t)c0 * 8, roll_flags));
        }
        // pack + (all-gather) + copy back: behind the fp64 finisher of this wave.  With two waves in flight that
        // is another stream, so the next wave's main kernel does not queue behind this wave's finisher.
        cudaStream_t tail = main;
        if (depth > 1) {
            LIB_TRY(pmcts_set_stream(ctx, comm));
            const int rc = pmcts_join(ctx, 0);
            if (rc) {
                pmcts_set_stream(ctx, main);
                return rc;
            }
            tail = comm;
        }
        const unsigned grid = (unsigned)((n_counts + 255) / 256);
        if (lay.elem == 1)
            pack_bins_kernel<uint8_t><<<grid, 256, 0, tail>>>(w.d_bins, (uint8_t *)(w.d_send + lay.o_bins), n_counts);
        else
            pack_bins_kernel<uint16_t><<<grid, 256, 0, tail>>>(w.d_bins, (uint16_t *)(w.d_send + lay.o_bins), n_counts);
        pmcts::count_launches(ctx, 1);
        cudaError_t ke = cudaGetLastError();
        int rc = PMCTS_OK;
        const size_t tail_bytes = (size_t)S_loc * b * 12 * lay.elem;
        if (ke != cudaSuccess) {
            rc = fail(PMCTS_ERR_CUDA, "pack_bins_kernel failed: %s", cudaGetErrorString(ke));
        } else if (dev_gather && p->peer) {
            // the block goes into every rank's window; this stream waits for the other ranks' and copies them down in place
            const void *gathered = nullptr;
            rc = pmcts_peer_put(ctx, p->peer, w.d_send, lay.bytes);
            if (!rc) rc = pmcts_peer_wait(ctx, p->peer, &gathered);
            if (!rc && cudaMemcpy2DAsync(w.h_recv, lay.bytes, gathered, pmcts::peer_block_bytes(p->peer), lay.bytes, world,
                                         cudaMemcpyDeviceToHost, tail) != cudaSuccess)
                rc = fail(PMCTS_ERR_CUDA, "copy of the gathered blocks failed");
        } else if (dev_gather) {
            rc = pmcts_allgather_updates(ctx, p->comm, w.d_send, lay.bytes, w.d_recv);
            if (!rc && cudaMemcpyAsync(w.h_recv, w.d_recv, lay.bytes * world, cudaMemcpyDeviceToHost, tail) != cudaSuccess)
                rc = fail(PMCTS_ERR_CUDA, "copy of the gathered blocks failed");
        } else {
            // one rank, or a host exchange: only this rank's counts come back (the host has the paths)
            char *dst = (world > 1 ? w.h_send : w.h_recv) + lay.o_bins;
            if (cudaMemcpyAsync(dst, w.d_send + lay.o_bins, tail_bytes, cudaMemcpyDeviceToHost, tail) != cudaSuccess)
                rc = fail(PMCTS_ERR_CUDA, "copy of the wave's histograms failed");
        }
        if (!rc && cudaEventRecord(w.done, tail) != cudaSuccess) rc = fail(PMCTS_ERR_CUDA, "cudaEventRecord failed");
        if (depth > 1) pmcts_set_stream(ctx, main);
        return rc;
    }

    void pack_host(char *out, const uint32_t *bins, size_t count) const {
        if (lay.elem == 1) for (size_t i = 0; i < count; ++i) ((uint8_t *)out)[i] = (uint8_t)bins[i];
        else for (size_t i = 0; i < count; ++i) ((uint16_t *)out)[i] = (uint16_t)bins[i];
    }
    void unpack_host(const char *in, uint32_t *bins, size_t count) const {
        if (lay.elem == 1) for (size_t i = 0; i < count; ++i) bins[i] = ((const uint8_t *)in)[i];
        else for (size_t i = 0; i < count; ++i) bins[i] = ((const uint16_t *)in)[i];
    }

    // wave's results: wait, (host exchange), back-up of the local trees, master update with every rank's block
    int complete(Wave &w) {
        const int b = w.b;
        const size_t n_counts = (size_t)S_loc * b * 12;
        double t0 = now_ms();
        if (device) CUDA_TRY(cudaEventSynchronize(w.done));
        rep.ms_wait += now_ms() - t0;
        if (world > 1 && !dev_gather) {
            t0 = now_ms();
            const int rc = p->exchange(p->exchange_user, w.h_send, lay.bytes, w.h_recv);
            if (rc) return fail(PMCTS_ERR_ARG, "exchange callback returned %d", rc);
            rep.ms_exchange += now_ms() - t0;
        } else if (world == 1 && !device) {
            std::memcpy(w.h_recv, w.h_send, lay.bytes);
        } else if (world == 1) {
            // (device, one rank: the counts were copied straight into h_recv; lengths and paths are in h_send)
            std::memcpy(w.h_recv, w.h_send, lay.o_bins);
        }
        // local trees first (their histograms are this rank's block), then the master, rank by rank
        t0 = now_ms();
        unpack_host(w.h_recv + (size_t)rank * lay.bytes + lay.o_bins, bins32.data(), n_counts);
        FOREST_TRY(pmcts_forest_backup(f, b, bins32.data(), p->slots_w + w.slot_off + (size_t)rank * S_loc * b));
        rep.ms_backup += now_ms() - t0;
        t0 = now_ms();
        for (int r = 0; r < world; ++r) {
            const char *blk = w.h_recv + (size_t)r * lay.bytes;
            if (world > 1) unpack_host(blk + lay.o_bins, bins32.data(), n_counts);  // (one rank: bins32 still holds it)
            for (int i = 0; i < S_loc; ++i) ids_scratch[i] = r * S_loc + i;
            FOREST_TRY(pmcts_forest_master_apply(f, S_loc, ids_scratch.data(), b, D, (const int32_t *)blk,
                                                 (const int8_t *)(blk + lay.o_path),
                                                 p->slots_m + w.slot_off + (size_t)r * S_loc * b, bins32.data()));
        }
        rep.ms_master += now_ms() - t0;
        rep.waves += 1;
        rep.leaves += (int64_t)S_loc * b;
        rep.rollouts += (int64_t)S_loc * b * p->replicas;
        w.index = -1;
        return PMCTS_OK;
    }

    int run() {
        const double t_start = now_ms();
        LIB_TRY(setup());
        if (device) LIB_TRY(pmcts_set_precision(ctx, PMCTS_PREC_FAST));
        const int64_t launches0 = device ? pmcts_launch_count(ctx) : 0;
        int64_t done = 0;
        size_t slot_off = 0;
        int index = 0;
        int rc = PMCTS_OK;
        while (done < p->budget && !rc) {
            Wave &w = ring[index % depth];
            if (w.index >= 0) rc = complete(w);  // the wave that used this slot `depth` waves ago
            if (rc) break;
            const int b = (int)std::min<int64_t>(B, p->budget - done);
            rc = launch(w, index, b, slot_off);
            slot_off += (size_t)S * b;
            done += b;
            ++index;
        }
        // drain, oldest first
        for (int k = 0; k < depth && !rc; ++k) {
            Wave &w = ring[(index + k) % depth];
            if (w.index >= 0) rc = complete(w);
        }
        if (rc) pmcts::forest_abandon_pending(f);  // waves selected and never filed: the forest stays searchable
        if (device) {
            if (depth > 1) pmcts_set_stream(ctx, main);
            if (rc) {
                pmcts_synchronize(ctx);  // nothing of this search may still be running when its buffers go
                if (comm) cudaStreamSynchronize(comm);
            } else {
                std::vector<double> st((size_t)S_loc * 8);
                cudaError_t e = cudaMemcpyAsync(st.data(), d_stats, st.size() * sizeof(double), cudaMemcpyDeviceToHost, main);
                if (e == cudaSuccess) e = cudaStreamSynchronize(main);
                if (e != cudaSuccess) rc = fail(PMCTS_ERR_CUDA, "reading the search statistics failed: %s", cudaGetErrorString(e));
                for (int i = 0; i < S_loc && !rc; ++i) {
                    rep.arrived += (int64_t)st[(size_t)i * 8 + 3];
                    rep.transitions += (int64_t)st[(size_t)i * 8 + 4];
                }
            }
        }
        if (device) rep.launches += pmcts_launch_count(ctx) - launches0;
        rep.ms_total = now_ms() - t_start;
        return rc;
    }
};

}  // namespace

extern "C" int pmcts_forest_search(pmcts_ctx *ctx, pmcts_forest *f, const int32_t *scen_slots, const double root[3],
                                   const double dest[2], double tmin, const pmcts_search_params *params,
                                   pmcts_search_report *report) try {
    if (!f || !params || !root || !dest) return fail(PMCTS_ERR_ARG, "NULL argument");
    Driver d{ctx, f, params, scen_slots, root, dest, tmin};
    int prev = -1;
    if (ctx) {
        cudaGetDevice(&prev);
        const int dev = pmcts_get_device(ctx);
        if (prev != dev) CUDA_TRY(cudaSetDevice(dev));
    }
    const int rc = d.run();
    d.release();
    if (ctx && prev >= 0) cudaSetDevice(prev);
    if (report) *report = d.rep;
    return rc;
} catch (const std::bad_alloc &) {
    return fail(PMCTS_ERR_ARG, "out of host memory");
}

// Tree.uct_search natively: the host runtime walks the tree and keeps the reference's random streams
// (pmcts::forest_sequential); every iteration's rollout is one small launch with host arrays.
extern "C" int pmcts_forest_search_sequential(pmcts_ctx *ctx, pmcts_forest *f, int local_index, int scen_slot,
                                              const double root[3], const double dest[2], double tmin, int64_t budget,
                                              int frequency, int roll_flags, pmcts_mt_state *py_state,
                                              pmcts_mt_state *np_state, pmcts_search_report *report) try {
    if (!ctx) return fail(PMCTS_ERR_NO_DEVICE, "pmcts_forest_search_sequential needs a ctx: the search has no CPU path");
    if (!f || !root || !dest || !py_state || !np_state) return fail(PMCTS_ERR_ARG, "NULL argument");
    const pmcts::ForestInfo info = pmcts::forest_info(f);
    const int n_instants = info.max_depth + 1;
    pmcts_search_report rep{};
    const double t0 = now_ms();
    const int64_t launches0 = pmcts_launch_count(ctx);
    LIB_TRY(pmcts_set_precision(ctx, PMCTS_PREC_FAST));
    const int flags = roll_flags & PMCTS_ROLL_REPLAY_FULL_PREFIX;
    // One block up, one launch on device arrays, one block down per iteration: the rollout's inputs (normals, path in
    // degrees, its length) and outputs (transitions, bin, status) sit side by side in page-locked host memory and in
    // a device mirror, both kept with the ctx.
    void **slot = pmcts::ctx_attachment(ctx, free_search_cache);
    if (!*slot) *slot = new SearchCache();
    SearchCache *cache = (SearchCache *)*slot;
    const size_t o_path = align16((size_t)n_instants * sizeof(double));
    const size_t o_plen = o_path + align16((size_t)info.max_depth * sizeof(double));
    const size_t in_bytes = o_plen + 16;
    const size_t o_steps = in_bytes, o_bin = o_steps + 16, o_status = o_bin + 16, total = o_status + 16;
    int prev = -1;
    cudaGetDevice(&prev);
    const int dev = pmcts_get_device(ctx);
    if (prev != dev) CUDA_TRY(cudaSetDevice(dev));
    struct Restore {
        int prev, dev;
        ~Restore() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
    } restore{prev, dev};
    LIB_TRY(grow_pinned(cache->seq_host, cache->cap_seq, total));
    LIB_TRY(grow_device(cache->seq_dev, cache->cap_seq_dev, total));
    char *h = cache->seq_host, *d = cache->seq_dev;
    cudaStream_t st = (cudaStream_t)pmcts_get_stream(ctx);
    // Results: the kernels store the three numbers of the rollout straight into the page-locked block (it is mapped
    // into the device's address space), so nothing is copied back; if the block cannot be mapped they go to the device
    // mirror and one copy brings them down.
    char *out = nullptr;
    {
        void *mapped = nullptr;
        if (cudaHostGetDevicePointer(&mapped, h, 0) == cudaSuccess && mapped) out = (char *)mapped;
        else cudaGetLastError();
    }
    const bool copy_down = out == nullptr;
    if (copy_down) out = d;
    // The device work of an iteration is the same every time -- same buffers, same root, the path and its length
    // travel in the block -- so the second iteration records it as a graph and the rest replay it: one submission per
    // rollout instead of a copy, a clear and two or three launches.  (Not with per-launch timing on, whose event
    // pairs belong to single launches, and not when the stream cannot be captured: the plain calls stay.)
    const char *env = std::getenv("PMCTS_SEQ_GRAPH");
    bool want_graph = !(env && env[0] == '0') && !pmcts::ctx_timing(ctx) && budget >= 8;
    SeqGraph graph;
    int64_t calls = 0;
    auto enqueue = [&]() -> int {
        CUDA_TRY(cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, st));
        pmcts_noise nz{};
        nz.mode = PMCTS_NOISE_INJECTED;
        nz.z = (const double *)d;
        LIB_TRY(pmcts_rollout_batch(ctx, scen_slot, 1, 1, root, dest, tmin, (const double *)(d + o_path),
                                    (const int32_t *)(d + o_plen), info.max_depth, &nz, nullptr, nullptr,
                                    (int32_t *)(out + o_steps), (uint8_t *)(out + o_status), (int32_t *)(out + o_bin), nullptr,
                                    nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, flags));
        if (copy_down) CUDA_TRY(cudaMemcpyAsync(h + o_steps, d + o_steps, total - o_steps, cudaMemcpyDeviceToHost, st));
        return PMCTS_OK;
    };
    const pmcts::SeqRollout rollout = [&](int, const int8_t *path, int len, const double *z, int *bin, int *steps) -> int {
        std::memcpy(h, z, (size_t)n_instants * sizeof(double));
        double *deg = (double *)(h + o_path);
        for (int j = 0; j < len; ++j) deg[j] = 45.0 * (double)path[j];  // ACTIONS[a] (simulatorTLKT.py:29)
        *(int32_t *)(h + o_plen) = len;
        const int64_t call = calls++;
        if (graph.exec) {
            CUDA_TRY(cudaGraphLaunch(graph.exec, st));
            pmcts::count_launches(ctx, graph.kernels);
        } else if (want_graph && call == 1) {
            // (the first iteration ran plainly: every buffer the launch grows on demand has its size)
            want_graph = false;
            const int64_t l0 = pmcts_launch_count(ctx);
            int rc = PMCTS_OK;
            // (no other host thread's call on this ctx may land between the two ends of the recording)
            const std::unique_lock<std::recursive_mutex> recording = pmcts::ctx_lock(ctx);
            if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
                rc = enqueue();
                cudaGraph_t g = nullptr;
                const cudaError_t e = cudaStreamEndCapture(st, &g);
                if (!rc && e == cudaSuccess && g && cudaGraphInstantiate(&graph.exec, g, 0) == cudaSuccess)
                    graph.kernels = (int)(pmcts_launch_count(ctx) - l0);
                else
                    graph.exec = nullptr;
                if (g) cudaGraphDestroy(g);
            }
            cudaGetLastError();  // (a refused capture leaves its error behind; the plain calls below are the answer to it)
            if (rc < 0 && rc != PMCTS_ERR_CUDA) return rc;
            if (graph.exec) CUDA_TRY(cudaGraphLaunch(graph.exec, st));
            else LIB_TRY(enqueue());
        } else {
            LIB_TRY(enqueue());
        }
        CUDA_TRY(cudaStreamSynchronize(st));
        const uint8_t status = *(volatile uint8_t *)(h + o_status);
        if (status == PMCTS_ST_OUT_OF_GRID || status == PMCTS_ST_PAST_HORIZON || status == PMCTS_ST_DEGENERATE)
            return (int)status;  // what the reference raises on
        *bin = *(volatile int32_t *)(h + o_bin);
        *steps = *(volatile int32_t *)(h + o_steps);
        return 0;
    };
    const int rc = pmcts::forest_sequential(f, local_index, budget, frequency, n_instants, py_state, np_state, rollout, &rep);
    rep.launches = pmcts_launch_count(ctx) - launches0;
    rep.ms_total = now_ms() - t0;
    if (report) *report = rep;
    if (rc < 0 && std::string(pmcts_last_error()).empty()) fail(rc, "%s", pmcts_forest_last_error());
    return rc;
} catch (const std::bad_alloc &) {
    return fail(PMCTS_ERR_ARG, "out of host memory");
}
