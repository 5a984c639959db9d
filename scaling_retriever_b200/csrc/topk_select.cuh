// Block-level exact top-k over 64-bit candidate keys held in shared memory.
//
// A candidate is (fp32 score, int32 doc id) packed by cand_key() so that key order == (score desc,
// doc id asc).  Keys are unique (doc ids are), which makes the k-th largest key a total cut: the
// result set is deterministic whatever order the candidates were appended in.  This is the GPU
// replacement for np.argpartition in SparseRetrieval.select_topk (reference
// scaling_retriever/indexer.py:315-322) and for faiss's reservoir collector behind
// IndexFlatIP.search (:211); unlike argpartition it resolves boundary ties deterministically
// (lowest doc id wins).
#pragma once
#include "common.cuh"

namespace b200ret {

// MSB-first radix select with 8-bit digits: returns the k-th largest key (1 <= k <= n) of keys[0..n).
// `hist` is 256 shared counters, `bcast` three shared u64 slots.  All threads of the block call it.
// * Common prefix first: one OR-reduction of key ^ keys[0] finds the highest bit in which the keys differ, and the digit
//   windows start THERE instead of at bit 63.  Candidate scores sit in a narrow range (same sign and exponent, often the same
//   leading mantissa bits), so byte-aligned passes from the top spent one or two whole passes putting every key into ONE
//   histogram bin — thousands of serialised shared-memory atomics on the same address per query (the select launches were
//   0.2-0.3 ms each, a per-rank fixed cost that does not shrink with the shard).
// * Early exit: as soon as the bucket that holds the k-th key contains exactly as many keys as are still needed, every key of
//   that bucket is selected and the k-th key is simply the bucket's minimum (one min pass instead of the remaining passes).
// Optional by-product (`m` > 0, `m_lower` a shared-memory slot): a key that at least `m` of the keys reach (1 <= m <= n) —
// the lower edge of the FIRST histogram's bucket that holds the m-th largest key.  It costs one more walk over the 256
// counters; the tau exchange of a sharded search publishes its score (candidates.cu).
__device__ inline uint64_t block_radix_select_kth(const uint64_t* keys, int n, int k, uint32_t* hist,
                                                  uint64_t* bcast, int m = 0, uint64_t* m_lower = nullptr,
                                                  bool first_pass_only = false) {
    // ---- highest differing bit ----
    if (threadIdx.x == 0) bcast[0] = 0;
    __syncthreads();
    {
        const uint64_t k0 = keys[0];
        unsigned long long x = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) x |= keys[i] ^ k0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x |= __shfl_xor_sync(0xffffffffu, x, off);
        if ((threadIdx.x & 31) == 0 && x) atomicOr(reinterpret_cast<unsigned long long*>(bcast), x);
    }
    __syncthreads();
    const uint64_t diff = bcast[0];
    __syncthreads();
    if (diff == 0) {                                                    // all keys equal (n == 1, keys are unique otherwise)
        if (m > 0 && threadIdx.x == 0) *m_lower = keys[0];
        __syncthreads();
        return keys[0];
    }
    int hi = 63 - __clzll(static_cast<long long>(diff));                // block-uniform
    uint64_t mask = (hi == 63) ? 0ull : ~((2ull << hi) - 1ull);         // bits above the first differing one: common to all keys
    uint64_t prefix = keys[0] & mask;
    int remaining = k;
    bool first = true, first_pass = true;
    while (hi >= 0) {
        const int lo = max(0, hi - 7);
        const uint32_t dmask = (1u << (hi - lo + 1)) - 1u;
        if (!first && static_cast<int>(bcast[2]) == remaining) {        // bucket size == keys still needed (block-uniform)
            if (threadIdx.x == 0) bcast[0] = ~0ull;
            __syncthreads();
            unsigned long long lo_key = ~0ull;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const uint64_t key = keys[i];
                if ((key & mask) == prefix) lo_key = min(lo_key, static_cast<unsigned long long>(key));
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) lo_key = min(lo_key, __shfl_xor_sync(0xffffffffu, lo_key, off));
            if ((threadIdx.x & 31) == 0 && lo_key != ~0ull) atomicMin(reinterpret_cast<unsigned long long*>(bcast), lo_key);
            __syncthreads();
            const uint64_t kth = bcast[0];
            __syncthreads();
            return kth;
        }
        first = false;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint64_t key = keys[i];
            if ((key & mask) == prefix) atomicAdd(&hist[static_cast<uint32_t>(key >> lo) & dmask], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // Warp 0 walks the 256 buckets from the top, 8 buckets per lane (lane 0 = highest digits).
            uint32_t c[8];
            uint32_t local = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                c[j] = hist[255 - (threadIdx.x * 8 + j)];
                local += c[j];
            }
            uint32_t incl = local;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
                if (static_cast<int>(threadIdx.x) >= off) incl += v;
            }
            uint32_t above = incl - local;   // candidates in strictly higher buckets than this lane's
            if (above < static_cast<uint32_t>(remaining) && incl >= static_cast<uint32_t>(remaining)) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (above + c[j] >= static_cast<uint32_t>(remaining)) {
                        bcast[0] = static_cast<uint64_t>(255 - (threadIdx.x * 8 + j));
                        bcast[1] = static_cast<uint64_t>(remaining - above);
                        bcast[2] = static_cast<uint64_t>(c[j]);      // size of the chosen bucket
                        break;
                    }
                    above += c[j];
                }
            }
        }
        if (first_pass && m > 0) {
            // By-product for the tau exchange: the bucket of the m-th largest key in this (first) histogram, refined once by a
            // second histogram over that bucket's keys -> a key that at least m keys reach, 16 bits below the first differing
            // bit (8 bits alone left the bound ~5 % low: twice the candidates per round).  Reuses `hist` after warp 0 has walked
            // the first histogram for the main selection (whose result sits in bcast[], untouched here).
            __shared__ uint64_t m_state[2];           // {digit of the m-th key's bucket, keys still needed inside it}
            auto walk = [&](int need) {               // warp 1: bucket holding the need-th largest of the current histogram
                const int l = threadIdx.x - 32;
                uint32_t c[8];
                uint32_t local = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    c[j] = hist[255 - (l * 8 + j)];
                    local += c[j];
                }
                uint32_t incl = local;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
                    if (l >= off) incl += v;
                }
                uint32_t above = incl - local;
                if (above < static_cast<uint32_t>(need) && incl >= static_cast<uint32_t>(need)) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (above + c[j] >= static_cast<uint32_t>(need)) {
                            m_state[0] = static_cast<uint64_t>(255 - (l * 8 + j));
                            m_state[1] = static_cast<uint64_t>(need - above);
                            break;
                        }
                        above += c[j];
                    }
                }
            };
            if (threadIdx.x >= 32 && threadIdx.x < 64) walk(m);
            __syncthreads();
            uint64_t m_prefix = prefix | (m_state[0] << lo);
            const int m_need = static_cast<int>(m_state[1]);
            if (lo > 0) {                             // second level: the next (up to) 8 bits inside that bucket
                const uint64_t m_mask = mask | (static_cast<uint64_t>(dmask) << lo);
                const int hi2 = lo - 1, lo2 = max(0, hi2 - 7);
                const uint32_t dmask2 = (1u << (hi2 - lo2 + 1)) - 1u;
                __syncthreads();
                for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
                __syncthreads();
                for (int i = threadIdx.x; i < n; i += blockDim.x) {
                    const uint64_t key = keys[i];
                    if ((key & m_mask) == m_prefix) atomicAdd(&hist[static_cast<uint32_t>(key >> lo2) & dmask2], 1u);
                }
                __syncthreads();
                if (threadIdx.x >= 32 && threadIdx.x < 64) walk(m_need);
                __syncthreads();
                m_prefix |= m_state[0] << lo2;
            }
            if (threadIdx.x == 0) *m_lower = m_prefix;
        }
        __syncthreads();
        if (first_pass && first_pass_only) return 0;     // the caller only wanted the by-product (block-uniform)
        first_pass = false;
        const uint64_t digit = bcast[0];
        remaining = static_cast<int>(bcast[1]);
        prefix |= digit << lo;
        mask |= static_cast<uint64_t>(dmask) << lo;
        __syncthreads();
        hi = lo - 1;
    }
    return prefix;
}

// In-place bitonic sort, DESCENDING, of keys[0..n_pow2) (n_pow2 a power of two; pad with 0).
__device__ inline void block_bitonic_sort_desc(uint64_t* keys, int n_pow2) {
    for (int size = 2; size <= n_pow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < (n_pow2 >> 1); i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const uint64_t a = keys[lo], b = keys[hi];
                if ((a < b) == desc) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
        }
    }
    __syncthreads();
}

__host__ __device__ inline int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

}  // namespace b200ret
