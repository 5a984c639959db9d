// Shared helpers for the b200ret kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "../../include/b200ret.h"

namespace b200ret {

// Thread-local last-error text behind b200ret_last_error().
char* err_buf();
void set_err(const char* fmt, ...);

#define B200RET_CUDA_CHECK(expr)                                                              \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::b200ret::set_err("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,             \
                               cudaGetErrorString(_e));                                       \
            return B200RET_ECUDA;                                                             \
        }                                                                                     \
    } while (0)

#define B200RET_REQUIRE(cond, ...)                                                            \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            ::b200ret::set_err(__VA_ARGS__);                                                  \
            return B200RET_EINVAL;                                                            \
        }                                                                                     \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace (256-byte aligned slices).
struct Workspace {
    char* base;
    size_t size;
    size_t used;
    Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), used(0) {}
    template <typename T>
    T* take(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        T* p = reinterpret_cast<T*>(base + used);
        used += bytes;
        return p;
    }
    bool ok() const { return used <= size; }
};

int sm_count();

// One-time per-device setup (kernel attributes are per context): `if (once.first()) { cudaFuncSetAttribute(...); once.mark(); }`.
struct PerDeviceOnce {
    std::atomic<bool> done[64] = {};
    static int device() {
        int dev = 0;
        return (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) ? dev : -1;
    }
    // true until the setup has COMPLETED on this device: host threads racing through the first call all do the (idempotent)
    // setup instead of launching before another thread has finished it
    bool first() const {
        const int dev = device();
        return dev < 0 || !done[dev].load(std::memory_order_acquire);
    }
    void mark() {
        const int dev = device();
        if (dev >= 0) done[dev].store(true, std::memory_order_release);
    }
};

// Launch accounting / optional per-kernel CUDA-event timing on the launching stream (see b200ret_profile_*).
constexpr int PROF_KINDS = 4;
constexpr int PROF_SPARSE_SCORE = 0, PROF_SPARSE_SELECT = 1, PROF_DENSE_GEMM = 2, PROF_CSR_SORT = 3;
void count_launches(int n);
void prof_begin(int kind, cudaStream_t stream);
void prof_end(int kind, cudaStream_t stream);

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Order-preserving map fp32 -> u32 (larger float <=> larger key); NaN sorts above +inf.
__device__ __forceinline__ uint32_t float_to_key(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}
// 64-bit candidate key: (score desc, doc id asc) <=> key descending.
__device__ __forceinline__ uint64_t cand_key(float score, int32_t id) {
    return (static_cast<uint64_t>(float_to_key(score)) << 32) | static_cast<uint32_t>(~static_cast<uint32_t>(id));
}
__device__ __forceinline__ float cand_score(uint64_t key) { return key_to_float(static_cast<uint32_t>(key >> 32)); }
__device__ __forceinline__ int32_t cand_id(uint64_t key) { return static_cast<int32_t>(~static_cast<uint32_t>(key)); }

}  // namespace b200ret
