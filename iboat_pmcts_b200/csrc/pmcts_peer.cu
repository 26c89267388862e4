// pmcts_b200 -- exchange of the ranks' update blocks over peer memory (pmcts_peer_*).
//
// What travels between the GPUs of a search is small (a wave's histograms narrowed to one byte per count: 0.4 MB per
// rank at 8 scenarios x 4096 leaves) and it travels UNDER the next wave's rollout kernel, whose blocks hold every block
// slot of the chip.  A library all-gather kernel needs a few large blocks resident at once and waits for room until the
// rollout kernel is almost through (measured: 1.0 ms under K2 against 55 us alone, 8 GPUs).  Here every rank owns a
// window in device memory that the other ranks map (CUDA IPC); one small kernel narrows the wave's counters and stores
// them straight into all the windows over NVLink -- the packing pass IS the all-gather -- and its last block raises the
// rank's sequence number in every window.  The consumer's stream then waits on the sequence numbers in its own window
// (stream memory operations: no kernel waits for another GPU) and reads the gathered block in place.
//
// Replaces, for a search over several GPUs, the Manager().dict() round trip of the reference
// (code/solver/worker.py:348-378: every worker pushes its buffer of (leaf, reward) records to the shared dictionary).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

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

constexpr int kMaxWorld = 16;          // ranks of one NVLink domain
constexpr size_t kHeaderBytes = 256;   // sequence numbers, one uint32 per source rank, in front of the data

// Window of rank r as mapped in this process: [sequence numbers | parity 0: world blocks | parity 1: world blocks].
struct Windows {
    unsigned char *base[kMaxWorld];
};

// One 16-byte vector of the outgoing block.  kWidth 1 / 2: counters narrowed to uint8 / uint16 (16 / 8 of them);
// kWidth 0: the block is sent as it is.  `count` = counters (kWidth > 0) or bytes (kWidth 0) in the source.
template <int kWidth>
__device__ __forceinline__ uint4 outgoing_vector(const void *__restrict__ src, size_t v, size_t count) {
    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    if (kWidth == 0) {
        if ((v + 1) * 16 <= count) return ((const uint4 *)src)[v];
        unsigned char b[16] = {0};
        for (size_t i = v * 16; i < count; ++i) b[i - v * 16] = ((const unsigned char *)src)[i];
        memcpy(&out, b, 16);
        return out;
    }
    constexpr int kPer = kWidth == 1 ? 16 : 8;
    const uint32_t *c = (const uint32_t *)src;
    uint32_t x[kPer];
    if ((v + 1) * kPer <= count) {
#pragma unroll
        for (int q = 0; q < kPer / 4; ++q) {
            const uint4 t = ((const uint4 *)c)[v * (kPer / 4) + q];
            x[4 * q] = t.x, x[4 * q + 1] = t.y, x[4 * q + 2] = t.z, x[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < kPer; ++q) x[q] = v * kPer + q < count ? c[v * kPer + q] : 0u;
    }
    if (kWidth == 1) {
        out.x = (x[0] & 255u) | (x[1] & 255u) << 8 | (x[2] & 255u) << 16 | x[3] << 24;
        out.y = (x[4] & 255u) | (x[5] & 255u) << 8 | (x[6] & 255u) << 16 | x[7] << 24;
        out.z = (x[8] & 255u) | (x[9] & 255u) << 8 | (x[10] & 255u) << 16 | x[11] << 24;
        out.w = (x[12] & 255u) | (x[13] & 255u) << 8 | (x[14] & 255u) << 16 | x[15] << 24;
    } else {
        out.x = (x[0] & 65535u) | x[1] << 16;
        out.y = (x[2] & 65535u) | x[3] << 16;
        out.z = (x[4] & 65535u) | x[5] << 16;
        out.w = (x[6] & 65535u) | x[7] << 16;
    }
    return out;
}

// Narrow + store into every rank's window (its own included), then publish: every block fences its stores and counts
// itself; the last one to do so writes this rank's sequence number into all the windows.  A reader that sees the
// number therefore sees the block.  A few blocks striding over the vectors (see backup_kernel: small grids find block
// slots under a resident rollout kernel, large ones queue for them).
template <int kWidth>
__global__ void __launch_bounds__(256)
put_kernel(Windows w, int world, int rank, size_t block_offset, const void *__restrict__ src, size_t count,
           size_t n_vec, unsigned *__restrict__ arrived, uint32_t seq) {
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += (size_t)gridDim.x * blockDim.x) {
        const uint4 val = outgoing_vector<kWidth>(src, v, count);
        for (int q = 0; q < world; ++q) {
            int p = rank + q;  // own window first, then round the ring: the ranks do not all hit one link at a time
            if (p >= world) p -= world;
            ((uint4 *)(w.base[p] + block_offset))[v] = val;
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        last = atomicAdd(arrived, 1u) == gridDim.x - 1;
        if (last) *arrived = 0u;  // (the next launch is ordered after this one by the stream)
    }
    __syncthreads();
    if (last && threadIdx.x < world) {
        __threadfence_system();
        *(volatile uint32_t *)(w.base[threadIdx.x] + 4 * (size_t)rank) = seq;
    }
}

typedef int (*StreamWaitValue32)(cudaStream_t, unsigned long long, uint32_t, unsigned);

}  // namespace

struct pmcts_peer {
    int world = 1, rank = 0, device = 0;
    size_t block_bytes = 0, window_bytes = 0;
    unsigned char *window = nullptr;  // this rank's (cudaMalloc)
    unsigned *arrived = nullptr;
    Windows mapped{};                 // every rank's, as mapped here
    bool connected = false;
    uint32_t seq = 0;                 // of the last put
    StreamWaitValue32 wait_value = nullptr;
};

namespace pmcts {
size_t peer_block_bytes(const pmcts_peer *peer) { return peer ? peer->block_bytes : 0; }
}  // namespace pmcts

extern "C" {

int pmcts_peer_create(pmcts_ctx *ctx, int world, int rank, size_t block_bytes, pmcts_peer **out,
                      uint8_t handle[PMCTS_PEER_HANDLE_BYTES]) {
    if (!ctx || !out || !handle) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
        return fail(PMCTS_ERR_ARG, "world must be 1 .. %d and rank inside it", kMaxWorld);
    if (block_bytes == 0 || block_bytes % 16) return fail(PMCTS_ERR_ARG, "block_bytes must be a positive multiple of 16");
    static_assert(sizeof(cudaIpcMemHandle_t) == PMCTS_PEER_HANDLE_BYTES, "handle size");
    int prev = -1;
    cudaGetDevice(&prev);
    const int dev = pmcts_get_device(ctx);
    CUDA_TRY(cudaSetDevice(dev));
    pmcts_peer *p = new (std::nothrow) pmcts_peer();
    if (!p) return fail(PMCTS_ERR_ARG, "out of host memory");
    p->world = world, p->rank = rank, p->device = dev, p->block_bytes = block_bytes;
    p->window_bytes = kHeaderBytes + 2 * (size_t)world * block_bytes;
    cudaError_t e = cudaMalloc(&p->window, p->window_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&p->arrived, sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(p->window, 0, p->window_bytes);
    if (e == cudaSuccess) e = cudaMemset(p->arrived, 0, sizeof(unsigned));
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->window);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult found = cudaDriverEntryPointSymbolNotFound;
    if (e == cudaSuccess) e = cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &found);
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    if (e != cudaSuccess || found != cudaDriverEntryPointSuccess || !fn) {
        const int rc = fail(PMCTS_ERR_CUDA, "peer window: %s",
                            e != cudaSuccess ? cudaGetErrorString(e) : "the driver has no cuStreamWaitValue32");
        pmcts_peer_destroy(p);
        return rc;
    }
    p->wait_value = (StreamWaitValue32)fn;
    p->mapped.base[rank] = p->window;
    p->connected = world == 1;
    std::memcpy(handle, &h, PMCTS_PEER_HANDLE_BYTES);
    *out = p;
    return PMCTS_OK;
}

int pmcts_peer_connect(pmcts_peer *p, const uint8_t *handles) {
    if (!p || !handles) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (p->connected) return PMCTS_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    CUDA_TRY(cudaSetDevice(p->device));
    int rc = PMCTS_OK;
    for (int r = 0; r < p->world && rc == PMCTS_OK; ++r) {
        if (r == p->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * PMCTS_PEER_HANDLE_BYTES, PMCTS_PEER_HANDLE_BYTES);
        void *ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            rc = fail(PMCTS_ERR_CUDA, "cudaIpcOpenMemHandle of rank %d's window failed: %s", r, cudaGetErrorString(e));
        } else {
            p->mapped.base[r] = (unsigned char *)ptr;
        }
    }
    if (prev >= 0 && prev != p->device) cudaSetDevice(prev);
    p->connected = rc == PMCTS_OK;
    return rc;
}

static int peer_put(pmcts_ctx *ctx, pmcts_peer *p, const void *src, size_t count, int width) {
    if (!ctx || !p || !src) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (!p->connected) return fail(PMCTS_ERR_ARG, "pmcts_peer_connect has not succeeded");
    const size_t bytes = width == 0 ? count : count * (size_t)width;
    if (bytes == 0 || bytes > p->block_bytes)
        return fail(PMCTS_ERR_ARG, "the block (%zu bytes) does not fit the window's (%zu)", bytes, p->block_bytes);
    if ((uintptr_t)src % 16) return fail(PMCTS_ERR_ARG, "the source must be 16-byte aligned");
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != p->device) CUDA_TRY(cudaSetDevice(p->device));
    const uint32_t seq = ++p->seq;
    const size_t offset = kHeaderBytes + (size_t)(seq & 1u) * p->world * p->block_bytes + (size_t)p->rank * p->block_bytes;
    const size_t n_vec = (bytes + 15) / 16;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
    const unsigned grid = (unsigned)std::min<size_t>((n_vec + 255) / 256, (size_t)sms);
    cudaStream_t s = (cudaStream_t)pmcts_get_stream(ctx);
    if (width == 1) put_kernel<1><<<grid, 256, 0, s>>>(p->mapped, p->world, p->rank, offset, src, count, n_vec, p->arrived, seq);
    else if (width == 2) put_kernel<2><<<grid, 256, 0, s>>>(p->mapped, p->world, p->rank, offset, src, count, n_vec, p->arrived, seq);
    else put_kernel<0><<<grid, 256, 0, s>>>(p->mapped, p->world, p->rank, offset, src, count, n_vec, p->arrived, seq);
    const cudaError_t e = cudaGetLastError();
    pmcts::count_launches(ctx, 1);
    if (prev >= 0 && prev != p->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(PMCTS_ERR_CUDA, "put_kernel: %s", cudaGetErrorString(e));
    return PMCTS_OK;
}

int pmcts_peer_put_bins(pmcts_ctx *ctx, pmcts_peer *peer, const uint32_t *bins, int64_t count, int width) {
    if (width != 1 && width != 2) return fail(PMCTS_ERR_ARG, "width must be 1 or 2 bytes");
    if (count <= 0) return fail(PMCTS_ERR_ARG, "count must be positive");
    return peer_put(ctx, peer, bins, (size_t)count, width);
}

int pmcts_peer_put(pmcts_ctx *ctx, pmcts_peer *peer, const void *send, size_t bytes) {
    return peer_put(ctx, peer, send, bytes, 0);
}

int pmcts_peer_wait(pmcts_ctx *ctx, pmcts_peer *p, const void **gathered) {
    if (!ctx || !p) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (p->seq == 0) return fail(PMCTS_ERR_ARG, "nothing has been put");
    cudaStream_t s = (cudaStream_t)pmcts_get_stream(ctx);
    for (int r = 0; r < p->world; ++r) {
        if (r == p->rank) continue;  // (its own block: ordered by the stream)
        const int rc = p->wait_value(s, (unsigned long long)(uintptr_t)(p->window + 4 * (size_t)r), p->seq,
                                     /* CU_STREAM_WAIT_VALUE_GEQ */ 0u);
        if (rc != 0) return fail(PMCTS_ERR_CUDA, "cuStreamWaitValue32 returned %d", rc);
    }
    if (gathered) *gathered = p->window + kHeaderBytes + (size_t)(p->seq & 1u) * p->world * p->block_bytes;
    return PMCTS_OK;
}

int pmcts_peer_poll(pmcts_ctx *ctx, pmcts_peer *p, int *arrived) {
    if (!ctx || !p || !arrived) return fail(PMCTS_ERR_ARG, "NULL argument");
    uint32_t seen[kMaxWorld];
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != p->device) CUDA_TRY(cudaSetDevice(p->device));
    cudaStream_t s = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);  // (not behind whatever waits on ctx's stream)
    if (e == cudaSuccess) e = cudaMemcpyAsync(seen, p->window, sizeof(uint32_t) * p->world, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (s) cudaStreamDestroy(s);
    if (prev >= 0 && prev != p->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(PMCTS_ERR_CUDA, "peer poll: %s", cudaGetErrorString(e));
    int n = 0;
    for (int r = 0; r < p->world; ++r) n += (int32_t)(seen[r] - p->seq) >= 0;
    *arrived = n;
    return PMCTS_OK;
}

void pmcts_peer_destroy(pmcts_peer *p) {
    if (!p) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(p->device);
    for (int r = 0; r < p->world; ++r)
        if (r != p->rank && p->mapped.base[r]) cudaIpcCloseMemHandle(p->mapped.base[r]);
    if (p->window) cudaFree(p->window);
    if (p->arrived) cudaFree(p->arrived);
    if (prev >= 0 && prev != p->device) cudaSetDevice(prev);
    delete p;
}

}  // extern "C"
