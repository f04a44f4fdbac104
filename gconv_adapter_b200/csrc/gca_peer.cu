// Multi-GPU plumbing over NVLink peer memory (one process per GPU; peers mapped with CUDA IPC).
//
// The partitioned adapter exchanges only the r-wide operand of each sparse hop.  Instead of a
// collective library call between the kernels, the kernel that PRODUCES such an operand stores its rows straight into
// every peer's gathered buffer (gca_push: peer-mapped pointers, plain st.global over NVLink / NVSwitch), so the
// transfer overlaps the producer's own streaming and the only thing left between two phases is a cross-GPU barrier:
//
//   k_peer_barrier   one warp per GPU: bump the local sequence number, store it into slot [rank] of every peer's flag
//                    array (st.release.sys after a system fence), spin until all slots of the local array reached it.
//                    Stream order makes the producer's peer stores precede the signal; the waiter's next kernel reads
//                    rows that arrived through its own L2.
//   k_peer_allreduce the 2 d r + d + r + 1 parameter-gradient floats: every rank stores its vector into slot [rank] of
//                    every peer, the same barrier, then every rank adds the slots in RANK order - bitwise identical
//                    results on all ranks, no atomics.
//   k_push_rows      generic copy of a finished local shard to the peers (producers that have no fused push).
//
// Each GPU runs these kernels in its own stream order and every rank issues the same sequence of them, so the sequence
// numbers agree by construction.  No NCCL call sits inside a step: the whole forward + backward is capturable in a CUDA graph.
#include <cstring>

#include "gca_common.cuh"
#include "gca_host.cuh"

namespace gca {
namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// threads 0 .. world-1 of the calling block; every thread of the block must call it (it contains __syncthreads)
__device__ __forceinline__ void peer_barrier_body(const gca_peer_sync& s) {
    __shared__ uint32_t target;
    if (threadIdx.x == 0) {
        target = *s.seq + 1u;
        *s.seq = target;
    }
    __threadfence_system();
    __syncthreads();
    const uint32_t tgt = target;
    if ((int)threadIdx.x < s.world) {
        st_release_sys(s.flags[threadIdx.x] + s.rank, tgt);                        // tell peer t that this rank arrived
        while ((int32_t)(ld_acquire_sys(s.flags[s.rank] + threadIdx.x) - tgt) < 0) { }   // wait for peer t
    }
    __syncthreads();
}

__global__ void __launch_bounds__(32) k_peer_barrier(const gca_peer_sync s) { peer_barrier_body(s); }

__global__ void __launch_bounds__(1024)
k_peer_allreduce(const gca_peer_sync s, const gca_push slots /* dst[h] = peer h's slot array */, const float* __restrict__ src,
                 float* __restrict__ out, int len) {
    // slot [rank] of every GPU <- my vector
    for (int h = 0; h < s.world; ++h) {
        float* dst = slots.dst[h] + (size_t)s.rank * len;
        for (int i = threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
    }
    peer_barrier_body(s);                                                         // fences, signals, waits
    const float* mine = slots.dst[s.rank];
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        float acc = 0.f;
        for (int h = 0; h < s.world; ++h) acc += __ldcg(mine + (size_t)h * len + i);   // rank order: same bits everywhere
        out[i] = acc;
    }
}

__global__ void __launch_bounds__(256)
k_push_rows(const float4* __restrict__ src, size_t n4, const gca_push push) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(src + i);
        for (int h = 0; h < push.count; ++h) reinterpret_cast<float4*>(push.dst[h])[i] = v;
    }
}

inline bool sync_ok(const gca_peer_sync* s) {
    if (!s || s->world < 1 || s->world > GCA_MAX_PEERS || s->rank < 0 || s->rank >= s->world || !s->seq) return false;
    for (int h = 0; h < s->world; ++h)
        if (!s->flags[h]) return false;
    return true;
}

}  // namespace

int launch_push_rows(const float* src, size_t nfloats, const gca_push* push, cudaStream_t st) {
    if (!push || push->count == 0 || nfloats == 0) return GCA_OK;
    if (push->count < 0 || push->count > GCA_MAX_PEERS || (nfloats % 4) != 0) return GCA_ERR_INVALID_ARG;
    const size_t n4 = nfloats / 4;
    size_t grid = (n4 + 255) / 256;
    if (grid > (size_t)(8 * num_sms())) grid = (size_t)(8 * num_sms());
    {
        ProfScope ps("push_rows", st);
        k_push_rows<<<(int)grid, 256, 0, st>>>(reinterpret_cast<const float4*>(src), n4, *push);
    }
    GCA_LAUNCH_OK();
    return GCA_OK;
}

}  // namespace gca

using namespace gca;

extern "C" int gca_peer_barrier(const gca_peer_sync* sync, gca_stream_t stream) {
    if (!sync_ok(sync)) return GCA_ERR_INVALID_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ProfScope ps("peer_barrier", st);
        k_peer_barrier<<<1, 32, 0, st>>>(*sync);
    }
    GCA_LAUNCH_OK();
    return GCA_OK;
}

extern "C" int gca_peer_allreduce(const gca_peer_sync* sync, const gca_push* slots, const float* src, float* out, int32_t len,
                                  gca_stream_t stream) {
    if (!sync_ok(sync) || !slots || slots->count != sync->world || !src || !out || len <= 0) return GCA_ERR_INVALID_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ProfScope ps("peer_allreduce", st);
        k_peer_allreduce<<<1, 1024, 0, st>>>(*sync, *slots, src, out, len);
    }
    GCA_LAUNCH_OK();
    return GCA_OK;
}

extern "C" int gca_push_rows(const float* src, int64_t nfloats, const gca_push* push, gca_stream_t stream) {
    if (!src || nfloats < 0) return GCA_ERR_INVALID_ARG;
    return launch_push_rows(src, (size_t)nfloats, push, static_cast<cudaStream_t>(stream));
}

// ---- peer-mappable allocations (CUDA IPC): what a non-PyTorch client needs to set the above up ----
extern "C" int gca_peer_alloc(size_t bytes, void** ptr) {
    if (!ptr || bytes == 0) return GCA_ERR_INVALID_ARG;
    *ptr = nullptr;
    GCA_CUDA(cudaMalloc(ptr, bytes));
    GCA_CUDA(cudaMemset(*ptr, 0, bytes));
    GCA_CUDA(cudaDeviceSynchronize());
    return GCA_OK;
}
extern "C" int gca_peer_free(void* ptr) {
    if (ptr) GCA_CUDA(cudaFree(ptr));
    return GCA_OK;
}
extern "C" int gca_peer_export(void* ptr, void* handle64) {
    if (!ptr || !handle64) return GCA_ERR_INVALID_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == GCA_IPC_HANDLE_BYTES, "IPC handle size");
    GCA_CUDA(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), ptr));
    return GCA_OK;
}
extern "C" int gca_peer_open(const void* handle64, void** ptr) {
    if (!ptr || !handle64) return GCA_ERR_INVALID_ARG;
    *ptr = nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    GCA_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GCA_OK;
}
extern "C" int gca_peer_close(void* ptr) {
    if (ptr) GCA_CUDA(cudaIpcCloseMemHandle(ptr));
    return GCA_OK;
}
