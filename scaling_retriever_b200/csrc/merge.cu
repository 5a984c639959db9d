// Shard merge: G per-shard top-k rows -> one global top-k row per query, plus the fp32->bf16 ingest cast.
//
// The reference never searches shards in parallel (retrieval asserts world_size == 1,
// eval_sparse.py:114 / eval_dense.py:191); this is the reduce step of the doc-range sharded search:
// every rank all-gathers [Q, k] (score, id) rows over NCCL and runs this kernel locally.
#include <cuda_bf16.h>

#include "common.cuh"
#include "topk_select.cuh"

namespace b200ret {

constexpr int MERGE_THREADS = 512;
constexpr int MERGE_MAX_CANDIDATES = 24576;   // 192 KB of keys in shared memory (+ up to 32 KB of winners)

// One CTA per query: gather the G*k keys (padding rows -> key 0, the minimum), radix-select the k-th largest, compact the
// winners into a second array and sort only those (a full bitonic sort of all G*k keys cost 8x the compare-exchanges and
// made the merge 2.4 ms of a 26 ms step on 8 GPUs).
__global__ void __launch_bounds__(MERGE_THREADS) merge_topk_kernel(const float* __restrict__ in_scores,
                                                                   const int64_t* __restrict__ in_ids, int32_t n_shards,
                                                                   int32_t n_queries, int32_t k, float* out_scores,
                                                                   int64_t* out_ids, int32_t* out_counts) {
    extern __shared__ __align__(16) uint64_t skeys[];     // [n_shards * k] candidates, then [next_pow2(k)] winners
    __shared__ uint32_t hist[256];
    __shared__ uint64_t bcast[3];
    __shared__ int n_live, out_pos;
    const int q = blockIdx.x;
    const int total = n_shards * k;
    uint64_t* const wkeys = skeys + total;
    if (threadIdx.x == 0) {
        n_live = 0;
        out_pos = 0;
    }
    __syncthreads();
    // The low key half is the candidate's POSITION i = shard * k + slot, not its id: ids stay 64-bit (no truncation), and
    // because every shard row is already sorted by (score desc, id asc) and shards are ascending doc ranges, the order
    // (score desc, position asc) is the same total order (score desc, id asc) the search kernels use.
    int live = 0;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int g = i / k, j = i - g * k;
        const size_t src = (static_cast<size_t>(g) * n_queries + q) * k + j;
        const int64_t id = in_ids[src];
        skeys[i] = (id >= 0) ? cand_key(in_scores[src], i) : 0ull;
        live += (id >= 0);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) live += __shfl_xor_sync(0xffffffffu, live, off);
    if ((threadIdx.x & 31) == 0 && live) atomicAdd(&n_live, live);
    __syncthreads();
    const int c = n_live;
    const int kept = min(c, k);
    uint64_t kth = 1;                                     // every live key is >= 1
    if (c > k) kth = block_radix_select_kth(skeys, total, k, hist, bcast);
    const int n_sort = next_pow2(max(kept, 1));
    for (int i0 = 0; i0 < total; i0 += blockDim.x) {      // warp-aggregated compaction of the winners
        const int i = i0 + threadIdx.x;
        const uint64_t key = (i < total) ? skeys[i] : 0ull;
        const bool win = key >= kth;
        const unsigned bal = __ballot_sync(0xffffffffu, win);
        if (bal) {
            int base = 0;
            if ((threadIdx.x & 31) == 0) base = atomicAdd(&out_pos, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (win) wkeys[base + __popc(bal & lanemask_lt())] = key;
        }
    }
    for (int i = kept + threadIdx.x; i < n_sort; i += blockDim.x) wkeys[i] = 0;
    __syncthreads();
    block_bitonic_sort_desc(wkeys, n_sort);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const bool lv = i < kept;
        const uint64_t key = lv ? wkeys[i] : 0;
        out_scores[static_cast<size_t>(q) * k + i] = lv ? cand_score(key) : -INFINITY;
        int64_t id = -1;
        if (lv) {
            const int pos = cand_id(key), g = pos / k, j = pos - g * k;
            id = in_ids[(static_cast<size_t>(g) * n_queries + q) * k + j];
        }
        out_ids[static_cast<size_t>(q) * k + i] = id;
    }
    if (threadIdx.x == 0) out_counts[q] = kept;
}

// ---- packed-key exchange path (what the sharded search sends over NVLink: 8 bytes per candidate) --------------------------
// key = (order-preserving score bits << 32) | ~(uint32 global doc id); 0 = padding.  Global ids must be < 2^32 - 1.
__global__ void pack_keys_kernel(const float* __restrict__ scores, const int64_t* __restrict__ ids, int64_t n,
                                 uint64_t* __restrict__ keys) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t id = ids[i];
        keys[i] = (id >= 0) ? ((static_cast<uint64_t>(float_to_key(scores[i])) << 32) | static_cast<uint32_t>(~static_cast<uint32_t>(id)))
                            : 0ull;
    }
}

// One CTA per query row: unpack k keys, count the live ones.
__global__ void unpack_keys_kernel(const uint64_t* __restrict__ keys, int32_t k, float* __restrict__ scores,
                                   int64_t* __restrict__ ids, int32_t* __restrict__ counts) {
    const size_t row = static_cast<size_t>(blockIdx.x) * k;
    int live = 0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const uint64_t key = keys[row + i];
        const bool lv = key != 0ull;
        scores[row + i] = lv ? key_to_float(static_cast<uint32_t>(key >> 32)) : -INFINITY;
        ids[row + i] = lv ? static_cast<int64_t>(static_cast<uint32_t>(~static_cast<uint32_t>(key))) : -1;
        live += lv;
    }
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) live += __shfl_xor_sync(0xffffffffu, live, off);
    if ((threadIdx.x & 31) == 0 && live) atomicAdd(&total, live);
    __syncthreads();
    if (threadIdx.x == 0) counts[blockIdx.x] = total;
}

// One CTA per query: G rows of k packed keys (each sorted descending, zero padded) -> the k largest, sorted descending.
// Keys are unique per query (doc ids are), so the k-th largest key is a total cut.
__global__ void __launch_bounds__(MERGE_THREADS) merge_keys_kernel(const uint64_t* __restrict__ in_keys, int32_t n_shards,
                                                                   int32_t n_queries, int32_t k, uint64_t* __restrict__ out_keys) {
    extern __shared__ __align__(16) uint64_t skeys[];     // [n_shards * k] candidates, then [next_pow2(k)] winners
    __shared__ uint32_t hist[256];
    __shared__ uint64_t bcast[3];
    __shared__ int n_live, out_pos;
    const int q = blockIdx.x;
    const int total = n_shards * k;
    uint64_t* const wkeys = skeys + total;
    if (threadIdx.x == 0) {
        n_live = 0;
        out_pos = 0;
    }
    __syncthreads();
    int live = 0;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int g = i / k, j = i - g * k;
        const uint64_t key = in_keys[(static_cast<size_t>(g) * n_queries + q) * k + j];
        skeys[i] = key;
        live += (key != 0ull);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) live += __shfl_xor_sync(0xffffffffu, live, off);
    if ((threadIdx.x & 31) == 0 && live) atomicAdd(&n_live, live);
    __syncthreads();
    const int c = n_live;
    const int kept = min(c, k);
    uint64_t kth = 1;
    if (c > k) kth = block_radix_select_kth(skeys, total, k, hist, bcast);
    const int n_sort = next_pow2(max(kept, 1));
    for (int i0 = 0; i0 < total; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        const uint64_t key = (i < total) ? skeys[i] : 0ull;
        const bool win = key >= kth;
        const unsigned bal = __ballot_sync(0xffffffffu, win);
        if (bal) {
            int base = 0;
            if ((threadIdx.x & 31) == 0) base = atomicAdd(&out_pos, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (win) wkeys[base + __popc(bal & lanemask_lt())] = key;
        }
    }
    for (int i = kept + threadIdx.x; i < n_sort; i += blockDim.x) wkeys[i] = 0;
    __syncthreads();
    block_bitonic_sort_desc(wkeys, n_sort);
    for (int i = threadIdx.x; i < k; i += blockDim.x) out_keys[static_cast<size_t>(q) * k + i] = (i < kept) ? wkeys[i] : 0ull;
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 4;
    for (int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            const float4 v = *reinterpret_cast<const float4*>(src + i);
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
            __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 packed;
            packed.x = *reinterpret_cast<uint32_t*>(&lo);
            packed.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(dst + i) = packed;
        } else {
            for (int64_t j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
        }
    }
}

}  // namespace b200ret

using namespace b200ret;

extern "C" int b200ret_merge_topk(const float* in_scores, const int64_t* in_ids, int32_t n_shards, int32_t n_queries,
                                  int32_t k, float* out_scores, int64_t* out_ids, int32_t* out_counts, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    B200RET_REQUIRE(n_shards >= 1 && n_queries >= 0 && k >= 1, "merge_topk: bad sizes");
    B200RET_REQUIRE(static_cast<int64_t>(n_shards) * k + next_pow2(k) <= MERGE_MAX_CANDIDATES + B200RET_MAX_K,
                    "merge_topk: %d shards x k=%d exceed the shared-memory bound (merge in passes of <= b200ret_merge_max_shards(k))",
                    n_shards, k);
    if (n_queries == 0) return B200RET_OK;
    B200RET_REQUIRE(in_scores && in_ids && out_scores && out_ids && out_counts, "merge_topk: null pointer");
    const size_t smem = (static_cast<size_t>(n_shards) * k + next_pow2(k)) * sizeof(uint64_t);
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        B200RET_CUDA_CHECK(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (MERGE_MAX_CANDIDATES + B200RET_MAX_K) * (int)sizeof(uint64_t)));
        attr_set.mark();
    }
    merge_topk_kernel<<<n_queries, MERGE_THREADS, smem, stream>>>(in_scores, in_ids, n_shards, n_queries, k, out_scores,
                                                                  out_ids, out_counts);
    count_launches(1);
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}

extern "C" int b200ret_pack_keys(const float* scores, const int64_t* ids, int64_t n, uint64_t* keys, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    B200RET_REQUIRE(n >= 0, "pack_keys: negative size");
    if (n == 0) return B200RET_OK;
    B200RET_REQUIRE(scores && ids && keys, "pack_keys: null pointer");
    pack_keys_kernel<<<sm_count() * 8, 256, 0, stream>>>(scores, ids, n, keys);
    count_launches(1);
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}

extern "C" int b200ret_unpack_keys(const uint64_t* keys, int32_t n_queries, int32_t k, float* out_scores, int64_t* out_ids,
                                   int32_t* out_counts, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    B200RET_REQUIRE(n_queries >= 0 && k >= 1, "unpack_keys: bad sizes");
    if (n_queries == 0) return B200RET_OK;
    B200RET_REQUIRE(keys && out_scores && out_ids && out_counts, "unpack_keys: null pointer");
    unpack_keys_kernel<<<n_queries, 128, 0, stream>>>(keys, k, out_scores, out_ids, out_counts);
    count_launches(1);
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}

extern "C" int32_t b200ret_merge_max_shards(int32_t k) {
    if (k < 1) return 0;
    const int64_t room = (MERGE_MAX_CANDIDATES + B200RET_MAX_K) - next_pow2(k);
    return static_cast<int32_t>(room / k);
}

extern "C" int b200ret_merge_keys(const uint64_t* in_keys, int32_t n_shards, int32_t n_queries, int32_t k, uint64_t* out_keys,
                                  void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    B200RET_REQUIRE(n_shards >= 1 && n_queries >= 0 && k >= 1, "merge_keys: bad sizes");
    B200RET_REQUIRE(n_shards <= b200ret_merge_max_shards(k), "merge_keys: %d shards x k=%d exceed the shared-memory bound (max %d shards per pass)",
                    n_shards, k, b200ret_merge_max_shards(k));
    if (n_queries == 0) return B200RET_OK;
    B200RET_REQUIRE(in_keys && out_keys, "merge_keys: null pointer");
    const size_t smem = (static_cast<size_t>(n_shards) * k + next_pow2(k)) * sizeof(uint64_t);
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        B200RET_CUDA_CHECK(cudaFuncSetAttribute(merge_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (MERGE_MAX_CANDIDATES + B200RET_MAX_K) * (int)sizeof(uint64_t)));
        attr_set.mark();
    }
    merge_keys_kernel<<<n_queries, MERGE_THREADS, smem, stream>>>(in_keys, n_shards, n_queries, k, out_keys);
    count_launches(1);
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}

extern "C" int b200ret_f32_to_bf16(const float* src, void* dst_bf16, int64_t n, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    B200RET_REQUIRE(n >= 0, "f32_to_bf16: negative size");
    if (n == 0) return B200RET_OK;
    B200RET_REQUIRE(src && dst_bf16, "f32_to_bf16: null pointer");
    B200RET_REQUIRE((reinterpret_cast<uintptr_t>(src) % 16 == 0) && (reinterpret_cast<uintptr_t>(dst_bf16) % 8 == 0),
                    "f32_to_bf16: pointers must be 16/8-byte aligned");
    f32_to_bf16_kernel<<<sm_count() * 8, 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst_bf16), n);
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}
