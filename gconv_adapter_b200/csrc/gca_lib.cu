// Library-level entry points of libgca: version, status strings, device probe, launch counter.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

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

// GCA_DISABLE_TC=1 forces the CUDA-core kernels (A/B runs, debugging).
bool tc_enabled() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("GCA_DISABLE_TC");
        cached = (e && e[0] == '1') ? 0 : 1;
    }
    return cached == 1;
}

// GCA_DISABLE_PDL=1 launches every kernel fully serialised (A/B runs).
bool pdl_enabled() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("GCA_DISABLE_PDL");
        cached = (e && e[0] == '1') ? 0 : 1;
    }
    return cached == 1;
}

struct ProfEntry { const char* name; cudaEvent_t a, b; };
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfEntry> g_prof;

bool prof_enabled() { return g_prof_on.load(std::memory_order_relaxed); }
void prof_begin(const char* name, cudaStream_t st) {
    ProfEntry e{name, nullptr, nullptr};
    if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
    cudaEventRecord(e.a, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(e);
}
void prof_end(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof.empty() && g_prof.back().b) cudaEventRecord(g_prof.back().b, st);
}

}  // namespace gca

extern "C" int gca_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(gca::g_prof_mu);
    for (auto& e : gca::g_prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    gca::g_prof.clear();
    gca::g_prof_on.store(on != 0);
    return GCA_OK;
}

// Synchronises the recorded events and writes {"name": {"launches": n, "ms": total}, ...} (JSON).
extern "C" int gca_profile_report(char* buf, size_t n) {
    if (!buf || n == 0) return GCA_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(gca::g_prof_mu);
    std::map<std::string, std::pair<long, double>> acc;
    std::vector<std::string> order;
    for (auto& e : gca::g_prof) {
        if (cudaEventSynchronize(e.b) != cudaSuccess) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.a, e.b) != cudaSuccess) continue;
        auto it = acc.find(e.name);
        if (it == acc.end()) { order.push_back(e.name); acc[e.name] = {1, ms}; }
        else { it->second.first++; it->second.second += ms; }
    }
    std::string out = "{";
    for (size_t i = 0; i < order.size(); ++i) {
        char tmp[256];
        snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"launches\": %ld, \"ms\": %.6f}", i ? ", " : "", order[i].c_str(),
                 acc[order[i]].first, acc[order[i]].second);
        out += tmp;
    }
    out += "}";
    if (out.size() + 1 > n) return GCA_ERR_WORKSPACE;
    memcpy(buf, out.c_str(), out.size() + 1);
    return GCA_OK;
}

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
