// Library-level pieces of the C ABI: version, error text, device facts.
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace b200ret {

char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

void set_err(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;   // B200; only reached without a device (workspace sizing on a CPU box)
    }
    return cached;
}

// ---- launch accounting + per-kernel CUDA-event timing (bench.py's roofline leg) -------------------------------
namespace {
struct ProfileState {          // searches may run on several host threads (one per shard / stream): counters are atomic,
    std::atomic<bool> enabled{false};      // the event lists are guarded by `lock` (only touched while profiling is on)
    std::atomic<long long> launches{0};
    std::mutex lock;
    std::vector<cudaEvent_t> pool;                 // recycled events
    std::vector<cudaEvent_t> begin[PROF_KINDS], end[PROF_KINDS];
};
ProfileState g_prof;
cudaEvent_t take_event() {
    if (!g_prof.pool.empty()) {
        cudaEvent_t e = g_prof.pool.back();
        g_prof.pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

void count_launches(int n) { g_prof.launches += n; }

void prof_begin(int kind, cudaStream_t stream) {
    if (!g_prof.enabled) return;
    std::lock_guard<std::mutex> guard(g_prof.lock);
    cudaEvent_t e = take_event();
    cudaEventRecord(e, stream);
    g_prof.begin[kind].push_back(e);
}

void prof_end(int kind, cudaStream_t stream) {
    if (!g_prof.enabled) return;
    std::lock_guard<std::mutex> guard(g_prof.lock);
    cudaEvent_t e = take_event();
    cudaEventRecord(e, stream);
    g_prof.end[kind].push_back(e);
}

}  // namespace b200ret

extern "C" int b200ret_profile_enable(int on) {
    b200ret::g_prof.enabled = on != 0;
    return B200RET_OK;
}

extern "C" int b200ret_profile_read(int kind, double* total_ms, int64_t* timed_launches, int64_t* all_launches) {
    using namespace b200ret;
    B200RET_REQUIRE(kind >= 0 && kind < PROF_KINDS, "profile_read: bad kind %d", kind);
    B200RET_CUDA_CHECK(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> guard(g_prof.lock);
    double ms = 0.0;
    const size_t n = g_prof.end[kind].size();
    for (size_t i = 0; i < n; ++i) {
        float t = 0.f;
        B200RET_CUDA_CHECK(cudaEventElapsedTime(&t, g_prof.begin[kind][i], g_prof.end[kind][i]));
        ms += t;
        g_prof.pool.push_back(g_prof.begin[kind][i]);
        g_prof.pool.push_back(g_prof.end[kind][i]);
    }
    g_prof.begin[kind].clear();
    g_prof.end[kind].clear();
    if (total_ms) *total_ms = ms;
    if (timed_launches) *timed_launches = static_cast<int64_t>(n);
    if (all_launches) {
        *all_launches = g_prof.launches.exchange(0);
    }
    return B200RET_OK;
}

extern "C" int b200ret_version(void) { return B200RET_VERSION; }

extern "C" const char* b200ret_last_error(void) { return b200ret::err_buf(); }

extern "C" int b200ret_device_info(int* sm, int* cc_major, int* cc_minor, size_t* smem_optin_bytes) {
    int dev = 0;
    B200RET_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    B200RET_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (sm) *sm = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (smem_optin_bytes) *smem_optin_bytes = prop.sharedMemPerBlockOptin;
    return B200RET_OK;
}
