// Library-level entry points of libgca: version, status strings, device probe, launch counter.
#include <atomic>

#include "gca_common.cuh"

namespace gca {

static std::atomic<int64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
    static int cached = 0;
    if (cached > 0) return cached;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        sms = 148;   // B200
    cached = sms;
    return cached;
}

}  // namespace gca

extern "C" int gca_abi_version(void) { return GCA_ABI_VERSION; }

extern "C" const char* gca_status_string(int status) {
    switch (status) {
        case GCA_OK: return "ok";
        case GCA_ERR_INVALID_ARG: return "invalid argument";
        case GCA_ERR_WORKSPACE: return "workspace too small or misaligned";
        case GCA_ERR_CUDA: return "CUDA error";
        case GCA_ERR_INDEX_RANGE: return "edge_index holds a node id outside [0, N)";
        case GCA_ERR_UNSUPPORTED: return "unsupported shape";
        case GCA_ERR_NO_DEVICE: return "no sm_100 CUDA device";
        default: return "unknown status";
    }
}

extern "C" const char* gca_last_cuda_error(void) { return cudaGetErrorString(gca::tl_last_cuda_error); }

extern "C" int64_t gca_launch_count(void) { return gca::g_launches.load(std::memory_order_relaxed); }
