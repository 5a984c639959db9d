// Library-level pieces of the C ABI: version, error text, device facts.
#include <stdarg.h>

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

}  // namespace b200ret

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
