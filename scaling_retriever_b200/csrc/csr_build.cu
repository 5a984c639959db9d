// Sparse index build on the GPU: COO postings -> CSR posting lists + doc-block skip table.
//
// Replaces the per-posting CPython loop of IndexDictOfArray.add_batch_document
// (reference scaling_retriever/utils/inverted_index.py:67-76) and the list->ndarray conversion of
// IndexDictOfArray.save (:84-88).  The reference appends (doc, value) to the list of term `col` in feed
// order; that is exactly a STABLE sort of the posting stream by term id, so the build is a hand-written
// stable LSD radix sort (<= 8 bits per pass, match-any warp multisplit, no atomics on the output order)
// followed by a boundary scan that turns the sorted term column into term offsets.
//
// HBM-bound integer/byte work: every pass streams 12 B/posting in (coalesced, one posting per lane)
// and 12 B/posting out; the grid is persistent (a multiple of the SM count) so the per-pass digit
// table is tiny (buckets x blocks) and one CTA scans it.
#include <cuda_fp16.h>

#include "common.cuh"

namespace b200ret {

constexpr int SORT_THREADS = 1024;                             // one CTA per SM
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ROUNDS = 8;                                  // postings per lane per tile
constexpr int SORT_TILE = SORT_THREADS * SORT_ROUNDS;           // 8192 postings per tile
#ifndef B200RET_SORT_BITS
#define B200RET_SORT_BITS 9       // digit width per pass: 17-bit term ids sort in 2 passes (9 + 8)
#endif
constexpr int SORT_MAX_BITS = B200RET_SORT_BITS;
constexpr int SORT_MAX_BUCKETS = 1 << SORT_MAX_BITS;
// scatter kernel shared memory: per-warp digit counters (16 bit: a warp holds 256 postings of a tile), digit tables, and the
// tile staged in digit order (row, col, val) so that the global writes are contiguous runs per digit
constexpr size_t SORT_SCATTER_SMEM = static_cast<size_t>(SORT_WARPS) * SORT_MAX_BUCKETS * sizeof(uint16_t) +
                                     3 * SORT_MAX_BUCKETS * sizeof(uint32_t) + 64 * sizeof(uint32_t) +
                                     3 * static_cast<size_t>(SORT_TILE) * sizeof(uint32_t);
static_assert(SORT_SCATTER_SMEM <= 227 * 1024, "scatter tile exceeds shared memory");

struct SortPass {
    const int32_t* src_row;
    const int32_t* src_col;
    const float* src_val;
    int32_t* dst_row;
    int32_t* dst_col;
    float* dst_val;
    int key_is_row;   // digit taken from the row (doc) column instead of the term column
    int shift;
    int bits;
};

__device__ __forceinline__ uint32_t pass_digit(const SortPass& p, int32_t row, int32_t col) {
    uint32_t key = static_cast<uint32_t>(p.key_is_row ? row : col);
    return (key >> p.shift) & ((1u << p.bits) - 1u);
}

// Block b owns postings [b * per_block, min(nnz, (b + 1) * per_block)); per_block is a multiple of SORT_TILE.
__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(SortPass p, int64_t nnz, int64_t per_block,
                                                                 uint32_t* __restrict__ counts) {
    __shared__ uint32_t hist[SORT_MAX_BUCKETS];
    const int buckets = 1 << p.bits;
    for (int i = threadIdx.x; i < buckets; i += SORT_THREADS) hist[i] = 0;
    __syncthreads();
    const int64_t lo = static_cast<int64_t>(blockIdx.x) * per_block;
    const int64_t hi = min(nnz, lo + per_block);
    const int32_t* keys = p.key_is_row ? p.src_row : p.src_col;
    const uint32_t mask = (1u << p.bits) - 1u;
    // 8 independent loads per thread in flight (the loop is bandwidth work; one load per iteration left it latency-bound)
    constexpr int U = 8;
    int64_t i = lo + threadIdx.x;
    for (; i + static_cast<int64_t>(U - 1) * SORT_THREADS < hi; i += static_cast<int64_t>(U) * SORT_THREADS) {
        uint32_t k[U];
#pragma unroll
        for (int u = 0; u < U; ++u) k[u] = static_cast<uint32_t>(__ldg(keys + i + static_cast<int64_t>(u) * SORT_THREADS));
#pragma unroll
        for (int u = 0; u < U; ++u) atomicAdd(&hist[(k[u] >> p.shift) & mask], 1u);
    }
    for (; i < hi; i += SORT_THREADS) atomicAdd(&hist[(static_cast<uint32_t>(__ldg(keys + i)) >> p.shift) & mask], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < buckets; i += SORT_THREADS) counts[static_cast<size_t>(i) * gridDim.x + blockIdx.x] = hist[i];
}

// Exclusive scan of counts[bucket][block] (bucket-major) in place; one CTA, sequential chunks per thread.
__global__ void __launch_bounds__(1024) sort_scan_kernel(uint32_t* __restrict__ counts, int n) {
    __shared__ uint32_t partial[1024];
    const int per = (n + 1023) / 1024;
    const int lo = min(n, static_cast<int>(threadIdx.x) * per);
    const int hi = min(n, lo + per);
    uint32_t sum = 0;
    for (int i = lo; i < hi; ++i) sum += counts[i];
    partial[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 partials.
    for (int off = 1; off < 1024; off <<= 1) {
        uint32_t v = (threadIdx.x >= off) ? partial[threadIdx.x - off] : 0;
        __syncthreads();
        partial[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = partial[threadIdx.x] - sum;
    for (int i = lo; i < hi; ++i) {
        uint32_t c = counts[i];
        counts[i] = run;
        run += c;
    }
}

// One pass of the stable LSD sort.  Per tile of 8192 postings (feed order = warp-major, round-major, lane-minor):
//   1. rank every posting among the postings of the same digit in its warp (ballot multisplit, warp-private counters);
//   2. scan the counters over warps and digits -> position of every posting in the tile's digit-sorted order;
//   3. stage the tile in that order in shared memory;
//   4. copy it out: consecutive threads hold consecutive postings of a digit, whose destinations are consecutive too
//      (global base of the digit + offset inside the tile's run), so the stores are contiguous runs of ~16-32 postings
//      instead of one 4-byte scattered store per posting and array (the previous version of this kernel was bound by L2
//      store transactions: 350 GB/s).
__global__ void __launch_bounds__(SORT_THREADS, 1) sort_scatter_kernel(SortPass p, int64_t nnz, int64_t per_block,
                                                                    const uint32_t* __restrict__ bases) {
    extern __shared__ __align__(16) unsigned char sort_smem[];
    uint16_t* const wcnt = reinterpret_cast<uint16_t*>(sort_smem);                                   // [SORT_WARPS][buckets]
    uint32_t* const gbase = reinterpret_cast<uint32_t*>(wcnt + SORT_WARPS * SORT_MAX_BUCKETS);       // running global offset per digit
    uint32_t* const tbase = gbase + SORT_MAX_BUCKETS;                                                // start of the digit's run inside the tile
    uint32_t* const ttot = tbase + SORT_MAX_BUCKETS;                                                 // postings of the digit in the tile
    uint32_t* const wsum = ttot + SORT_MAX_BUCKETS;                                                  // [32] scan scratch
    int32_t* const s_row = reinterpret_cast<int32_t*>(wsum + 64);
    int32_t* const s_col = s_row + SORT_TILE;
    float* const s_val = reinterpret_cast<float*>(s_col + SORT_TILE);

    const int buckets = 1 << p.bits;
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < buckets; i += SORT_THREADS) gbase[i] = bases[static_cast<size_t>(i) * gridDim.x + blockIdx.x];

    const int64_t lo = static_cast<int64_t>(blockIdx.x) * per_block;
    const int64_t hi = min(nnz, lo + per_block);
    for (int64_t tile = lo; tile < hi; tile += SORT_TILE) {
        for (int i = threadIdx.x; i < SORT_WARPS * buckets / 2; i += SORT_THREADS) reinterpret_cast<uint32_t*>(wcnt)[i] = 0;
        __syncthreads();

        int32_t row[SORT_ROUNDS], col[SORT_ROUNDS];
        float val[SORT_ROUNDS];
        uint32_t rank[SORT_ROUNDS];
        uint16_t* const my_cnt = wcnt + warp * buckets;
        // Warp w owns the contiguous slice [tile + w*256, +256): round-major, lane-minor == feed order.
        const int64_t wbase = tile + static_cast<int64_t>(warp) * (32 * SORT_ROUNDS);
#pragma unroll
        for (int r = 0; r < SORT_ROUNDS; ++r) {
            const int64_t i = wbase + r * 32 + lane;
            const bool valid = i < hi;
            row[r] = valid ? __ldg(p.src_row + i) : 0;
            col[r] = valid ? __ldg(p.src_col + i) : 0;
            val[r] = valid ? __ldg(p.src_val + i) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < SORT_ROUNDS; ++r) {
            const bool valid = (wbase + r * 32 + lane) < hi;
            const uint32_t d = pass_digit(p, row[r], col[r]);
            // lanes holding the same digit: composed from one ballot per digit bit (cheaper than MATCH.ANY at <= 9 bits)
            const unsigned vb = __ballot_sync(0xffffffffu, valid);
            unsigned peers = valid ? vb : ~vb;
            for (int b = 0; b < p.bits; ++b) {
                const bool bit = (d >> b) & 1u;
                const unsigned bal = __ballot_sync(0xffffffffu, bit);
                peers &= bit ? bal : ~bal;
            }
            const unsigned before = __popc(peers & lanemask_lt());
            uint32_t old = 0;
            if (valid && before == 0) {   // lowest lane of each digit group bumps the warp-private counter
                old = my_cnt[d];
                my_cnt[d] = static_cast<uint16_t>(old + __popc(peers));
            }
            old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
            rank[r] = old + before;
            __syncwarp();
        }
        __syncthreads();
        // Per digit: exclusive scan of the counters over warps; total of the digit in the tile.
        uint32_t my_tot = 0;
        if (threadIdx.x < buckets) {
            const int d = threadIdx.x;
            uint32_t run = 0;
#pragma unroll 8
            for (int w = 0; w < SORT_WARPS; ++w) {
                const uint32_t c = wcnt[w * buckets + d];
                wcnt[w * buckets + d] = static_cast<uint16_t>(run);
                run += c;
            }
            my_tot = run;
            ttot[d] = run;
        }
        // Exclusive scan of the totals over digits -> start of every digit's run in the staged tile.
        {
            uint32_t incl = my_tot;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += v;
            }
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                uint32_t t = wsum[lane];
                uint32_t ti = t;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, ti, off);
                    if (lane >= off) ti += v;
                }
                wsum[32 + lane] = ti - t;
            }
            __syncthreads();
            if (threadIdx.x < buckets) tbase[threadIdx.x] = wsum[32 + warp] + incl - my_tot;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SORT_ROUNDS; ++r) {
            if ((wbase + r * 32 + lane) < hi) {
                const uint32_t d = pass_digit(p, row[r], col[r]);
                const uint32_t li = tbase[d] + my_cnt[d] + rank[r];
                s_row[li] = row[r];
                s_col[li] = col[r];
                s_val[li] = val[r];
            }
        }
        __syncthreads();
        const int n_tile = static_cast<int>(min(static_cast<int64_t>(SORT_TILE), hi - tile));
#pragma unroll
        for (int r = 0; r < SORT_ROUNDS; ++r) {
            const int j = r * SORT_THREADS + threadIdx.x;
            if (j < n_tile) {
                const int32_t rw = s_row[j], cl = s_col[j];
                const uint32_t d = pass_digit(p, rw, cl);
                const uint32_t pos = gbase[d] + (static_cast<uint32_t>(j) - tbase[d]);
                p.dst_row[pos] = rw;
                p.dst_col[pos] = cl;
                p.dst_val[pos] = s_val[j];
            }
        }
        __syncthreads();
        if (threadIdx.x < buckets) gbase[threadIdx.x] += my_tot;
    }
}

// term_offsets from the sorted term column: thread i fills offsets (prev_term, cur_term] = i.
__global__ void term_offsets_kernel(const int32_t* __restrict__ sorted_col, int64_t nnz, int32_t n_terms,
                                    int64_t* __restrict__ term_offsets) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nnz; i += stride) {
        const int32_t cur = sorted_col[i];
        const int32_t prev = (i > 0) ? sorted_col[i - 1] : -1;
        for (int32_t t = prev + 1; t <= cur; ++t) term_offsets[t] = i;
        if (i == nnz - 1)
            for (int32_t t = cur + 1; t <= n_terms; ++t) term_offsets[t] = nnz;
    }
}

__global__ void fill_i64_kernel(int64_t* p, int64_t n, int64_t v) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- doc-block skip table ------------------------------------------------------------------------

__device__ __forceinline__ int32_t term_of_posting(const int64_t* __restrict__ term_offsets, int32_t n_terms, int64_t i) {
    // largest t with term_offsets[t] <= i  (upper_bound - 1); empty terms are skipped automatically.
    int32_t lo = 0, hi = n_terms;   // answer in [lo, hi)
    while (hi - lo > 1) {
        int32_t mid = (lo + hi) >> 1;
        if (__ldg(term_offsets + mid) <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void block_table_kernel(const int64_t* __restrict__ term_offsets, const int32_t* __restrict__ doc_ids,
                                   int64_t nnz, int32_t n_terms, int32_t block_docs, int32_t n_blocks,
                                   uint32_t* __restrict__ table, int32_t* __restrict__ status) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const size_t row_len = static_cast<size_t>(n_blocks) + 1;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nnz; i += stride) {
        const int32_t t = term_of_posting(term_offsets, n_terms, i);
        const int64_t beg = __ldg(term_offsets + t), end = __ldg(term_offsets + t + 1);
        const int32_t d = __ldg(doc_ids + i);
        const int32_t b = d / block_docs;
        int32_t b_prev = -1;
        if (i > beg) {
            const int32_t dp = __ldg(doc_ids + i - 1);
            b_prev = dp / block_docs;
            if (dp > d) *status = B200RET_EUNSORTED;
        }
        if (b >= n_blocks || d < 0) { *status = B200RET_EINVAL; continue; }
        uint32_t* row = table + static_cast<size_t>(t) * row_len;
        for (int32_t bb = b_prev + 1; bb <= b; ++bb) row[bb] = static_cast<uint32_t>(i);
        if (i + 1 == end)
            for (int32_t bb = b + 1; bb <= n_blocks; ++bb) row[bb] = static_cast<uint32_t>(i + 1);
    }
}

// Rows of empty terms: every entry is the (shared) list position.  One warp per term.
__global__ void block_table_empty_kernel(const int64_t* __restrict__ term_offsets, int32_t n_terms, int32_t n_blocks,
                                         uint32_t* __restrict__ table) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = lane_id();
    for (int32_t t = blockIdx.x * warps_per_block + (threadIdx.x >> 5); t < n_terms; t += gridDim.x * warps_per_block) {
        const int64_t beg = term_offsets[t];
        if (term_offsets[t + 1] != beg) continue;
        uint32_t* row = table + static_cast<size_t>(t) * (static_cast<size_t>(n_blocks) + 1);
        for (int32_t bb = lane; bb <= n_blocks; bb += 32) row[bb] = static_cast<uint32_t>(beg);
    }
}

// ---- bank-aware posting order inside every (term, doc block) slice ---------------------------------------------------
// The search kernel adds 32 consecutive postings to a warp-private fp32 tile with one LDS + one STS; two postings whose
// doc ids are congruent mod 32 hit the same shared-memory bank and serialise (measured: 2.7 wavefronts per access on
// doc-sorted lists).  Inside a slice the order of postings is free (one doc occurs at most once per list), so the slice
// is rewritten in bank-quantile order (every bank's postings spread evenly over the slice): a window of 32 consecutive
// postings then holds about its even share 32*c/len of each bank instead of runs.  One warp per slice, slice staged in
// shared memory, in place.
// Implementation: posting i of bank b with in-bank rank r (of c) gets the integer key floor((2r+1) * len / (2c)) in
// [0, len) - its quantile position in the slice; keys of one bank are >= len/c >= 1 apart, so a key value holds at most one
// posting per bank.  A counting sort over the key values (shared-memory histogram + warp scan) then yields the permutation:
// one integer division per posting instead of a comparison sort.
constexpr int BANK_MAX_WARPS = 32;
constexpr int BANK_SMALL_CAP = 256;     // slices up to this length are laid out by the many-warp launch
constexpr int BANK_MAX_BLOCK_DOCS = 8192;
// per-warp shared memory for slices of up to `cap` postings: ids, weights, (key, slot), key counters + 32 bank counters
static inline size_t bank_warp_smem(int cap) { return static_cast<size_t>(cap) * 16 + 128; }

// Output: the search-side posting array, (doc id, weight bits) interleaved as 8-byte elements at the SAME positions as the CSR
// (so the skip table addresses both), every slice bank-ordered when `bank_order` is set, copied as is otherwise.
// One posting of the search-side array.  FMT 0: {int32 doc id, fp32 weight} (8 bytes, the parity format).  FMT 1: one 32-bit
// word, fp16(weight) << 16 | (doc id - first doc of the block) — the opt-in compressed format (round-to-nearest-even fp16
// weights, 16-bit block-local doc ids): half the bytes per posting.
template <int FMT>
__device__ __forceinline__ void store_posting(void* out, size_t pos, int32_t doc, float w, int32_t block_first_doc) {
    if (FMT == 0) {
        static_cast<uint2*>(out)[pos] = make_uint2(static_cast<uint32_t>(doc), __float_as_uint(w));
    } else {
        const uint32_t h = __half_as_ushort(__float2half_rn(w));
        static_cast<uint32_t*>(out)[pos] = (h << 16) | static_cast<uint32_t>(doc - block_first_doc);
    }
}

template <int FMT>
__global__ void __launch_bounds__(BANK_MAX_WARPS * 32) posting_layout_kernel(const uint32_t* __restrict__ table,
                                                                              const int32_t* __restrict__ doc_ids,
                                                                              const float* __restrict__ weights, void* __restrict__ out,
                                                                              int32_t n_terms, int32_t n_blocks, int32_t cap,
                                                                              int32_t bank_order, int32_t min_len, int32_t max_len,
                                                                              int32_t block_docs) {
    extern __shared__ __align__(16) unsigned char bank_smem[];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int n_warps = blockDim.x >> 5;
    unsigned char* mine = bank_smem + static_cast<size_t>(warp) * (static_cast<size_t>(cap) * 16 + 128);
    int32_t* s_ids = reinterpret_cast<int32_t*>(mine);
    float* s_w = reinterpret_cast<float*>(mine + static_cast<size_t>(cap) * 4);
    uint32_t* s_key = reinterpret_cast<uint32_t*>(mine + static_cast<size_t>(cap) * 8);     // pass 1: in-bank rank; pass 2: key << 8 | slot
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(mine + static_cast<size_t>(cap) * 12);    // postings per key value -> exclusive offsets
    uint32_t* hist = reinterpret_cast<uint32_t*>(mine + static_cast<size_t>(cap) * 16);
    const size_t row_len = static_cast<size_t>(n_blocks) + 1;
    const int groups = (n_blocks + 31) / 32;
    const long long n_items = static_cast<long long>(n_terms) * groups;
    const long long stride = static_cast<long long>(gridDim.x) * n_warps;
    for (long long item = static_cast<long long>(blockIdx.x) * n_warps + warp; item < n_items; item += stride) {
        const int t = static_cast<int>(item / groups);
        const int b = static_cast<int>(item % groups) * 32 + lane;
        uint32_t beg = 0, end = 0;
        if (b < n_blocks) {
            beg = table[static_cast<size_t>(t) * row_len + b];
            end = table[static_cast<size_t>(t) * row_len + b + 1];
        }
        // this launch handles the slices with min_len < length <= max_len (the shared memory per warp is sized for max_len)
        unsigned todo = __ballot_sync(0xffffffffu, end - beg > static_cast<uint32_t>(min_len) && end - beg <= static_cast<uint32_t>(max_len));
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t s_beg = __shfl_sync(0xffffffffu, beg, j);
            const int len = static_cast<int>(__shfl_sync(0xffffffffu, end, j) - s_beg);
            const int32_t first_doc = (static_cast<int>(item % groups) * 32 + j) * block_docs;     // of the slice's doc block
            if (len == 1 || !bank_order) {                     // nothing to reorder: interleave and copy
                for (int i = lane; i < len; i += 32) store_posting<FMT>(out, s_beg + i, doc_ids[s_beg + i], weights[s_beg + i], first_doc);
                continue;
            }
            hist[lane] = 0;
            for (int i = lane; i < len; i += 32) s_cnt[i] = 0;
            __syncwarp();
            for (int i = lane; i < len; i += 32) {             // stage the slice, rank every posting inside its bank
                const int32_t d = doc_ids[s_beg + i];
                s_ids[i] = d;
                s_w[i] = weights[s_beg + i];
                s_key[i] = atomicAdd(&hist[d & 31], 1u);
            }
            __syncwarp();
            for (int i = lane; i < len; i += 32) {             // quantile key + slot inside the key's bin (<= 1 per bank)
                const uint32_t r = s_key[i];
                const uint32_t c = hist[s_ids[i] & 31];
                const uint32_t key = ((2u * r + 1u) * static_cast<uint32_t>(len)) / (2u * c);   // < len; products < 2^26
                const uint32_t slot = atomicAdd(&s_cnt[key], 1u);
                s_key[i] = (key << 8) | slot;
            }
            __syncwarp();
            uint32_t carry = 0;                                 // exclusive scan of the bin sizes, 32 bins per step
            for (int base = 0; base < len; base += 32) {
                const uint32_t v = (base + lane < len) ? s_cnt[base + lane] : 0u;
                uint32_t incl = v;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += u;
                }
                if (base + lane < len) s_cnt[base + lane] = carry + incl - v;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            __syncwarp();
            for (int i = lane; i < len; i += 32) {
                const uint32_t ks = s_key[i];
                const uint32_t pos = s_cnt[ks >> 8] + (ks & 0xffu);
                store_posting<FMT>(out, s_beg + pos, s_ids[i], s_w[i], first_doc);
            }
            __syncwarp();
        }
    }
}

static int bits_for(int64_t n_values) {   // bits needed to represent values in [0, n_values)
    int b = 0;
    while ((int64_t{1} << b) < n_values) ++b;
    return b < 1 ? 1 : b;
}

struct PassPlan {
    int n;
    int key_is_row[16];
    int shift[16];
    int bits[16];
};

static void plan_bits(PassPlan& plan, int total_bits, int key_is_row) {
    const int passes = (total_bits + SORT_MAX_BITS - 1) / SORT_MAX_BITS;
    int done = 0;
    for (int i = 0; i < passes; ++i) {
        int b = (total_bits - done + (passes - i) - 1) / (passes - i);   // spread evenly, <= SORT_MAX_BITS
        plan.key_is_row[plan.n] = key_is_row;
        plan.shift[plan.n] = done;
        plan.bits[plan.n] = b;
        ++plan.n;
        done += b;
    }
}

static PassPlan make_plan(int32_t n_terms, int32_t n_docs, int sort_docs) {
    PassPlan plan{};
    if (sort_docs) plan_bits(plan, bits_for(n_docs), 1);   // least significant key first (LSD)
    plan_bits(plan, bits_for(n_terms), 0);
    return plan;
}

static int sort_grid() { return sm_count(); }

}  // namespace b200ret

using namespace b200ret;

extern "C" size_t b200ret_csr_build_workspace_bytes(int64_t nnz, int32_t n_terms, int32_t n_docs, int sort_docs) {
    (void)n_terms; (void)n_docs; (void)sort_docs;
    const size_t n = static_cast<size_t>(nnz > 0 ? nnz : 1);
    // two ping-pong (row, col, val) triples + the bucket x block digit table
    size_t bytes = 0;
    for (int i = 0; i < 6; ++i) bytes += align_up(n * 4, 256);
    bytes += align_up(static_cast<size_t>(SORT_MAX_BUCKETS) * sort_grid() * sizeof(uint32_t), 256);
    return bytes + 256;
}

extern "C" int b200ret_csr_build(const int32_t* rows, const int32_t* cols, const float* vals, int64_t nnz,
                                 int32_t n_terms, int32_t n_docs, int sort_docs,
                                 int64_t* term_offsets, int32_t* doc_ids, float* weights,
                                 void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    B200RET_REQUIRE(nnz >= 0 && n_terms > 0 && n_docs >= 0, "csr_build: bad sizes nnz=%lld n_terms=%d n_docs=%d",
                    (long long)nnz, n_terms, n_docs);
    B200RET_REQUIRE(nnz < (int64_t{1} << 32) - SORT_TILE, "csr_build: nnz=%lld exceeds the 32-bit position range",
                    (long long)nnz);
    B200RET_REQUIRE(term_offsets != nullptr, "csr_build: term_offsets is null");
    if (nnz == 0) {
        const int64_t n = static_cast<int64_t>(n_terms) + 1;
        fill_i64_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(term_offsets, n, 0);
        B200RET_CUDA_CHECK(cudaGetLastError());
        return B200RET_OK;
    }
    B200RET_REQUIRE(rows && cols && vals && doc_ids && weights && workspace, "csr_build: null pointer");
    if (workspace_bytes < b200ret_csr_build_workspace_bytes(nnz, n_terms, n_docs, sort_docs)) {
        set_err("csr_build: workspace too small (%zu bytes)", workspace_bytes);
        return B200RET_EWORKSPACE;
    }
    Workspace ws(workspace, workspace_bytes);
    int32_t* a_row = ws.take<int32_t>(nnz);
    int32_t* a_col = ws.take<int32_t>(nnz);
    float* a_val = ws.take<float>(nnz);
    int32_t* b_row = ws.take<int32_t>(nnz);
    int32_t* b_col = ws.take<int32_t>(nnz);
    float* b_val = ws.take<float>(nnz);
    const int grid = sort_grid();
    uint32_t* counts = ws.take<uint32_t>(static_cast<size_t>(SORT_MAX_BUCKETS) * grid);

    int64_t per_block = (nnz + grid - 1) / grid;
    per_block = (per_block + SORT_TILE - 1) / SORT_TILE * SORT_TILE;

    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        B200RET_CUDA_CHECK(cudaFuncSetAttribute(sort_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                static_cast<int>(SORT_SCATTER_SMEM)));
        attr_set.mark();
    }
    const PassPlan plan = make_plan(n_terms, n_docs, sort_docs);
    const int32_t* src_row = rows;
    const int32_t* src_col = cols;
    const float* src_val = vals;
    const int32_t* sorted_col = nullptr;
    for (int i = 0; i < plan.n; ++i) {
        const bool last = (i == plan.n - 1);
        const bool to_a = (i % 2 == 0);
        SortPass p;
        p.src_row = src_row; p.src_col = src_col; p.src_val = src_val;
        p.dst_row = last ? doc_ids : (to_a ? a_row : b_row);
        p.dst_col = to_a ? a_col : b_col;
        p.dst_val = last ? weights : (to_a ? a_val : b_val);
        p.key_is_row = plan.key_is_row[i]; p.shift = plan.shift[i]; p.bits = plan.bits[i];
        const int buckets = 1 << p.bits;
        prof_begin(PROF_CSR_SORT, stream);
        sort_hist_kernel<<<grid, SORT_THREADS, 0, stream>>>(p, nnz, per_block, counts);
        sort_scan_kernel<<<1, 1024, 0, stream>>>(counts, buckets * grid);
        sort_scatter_kernel<<<grid, SORT_THREADS, SORT_SCATTER_SMEM, stream>>>(p, nnz, per_block, counts);
        prof_end(PROF_CSR_SORT, stream);
        count_launches(3);
        B200RET_CUDA_CHECK(cudaGetLastError());
        src_row = p.dst_row; src_col = p.dst_col; src_val = p.dst_val;
        sorted_col = p.dst_col;
    }
    term_offsets_kernel<<<sm_count() * 8, 256, 0, stream>>>(sorted_col, nnz, n_terms, term_offsets);
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}

extern "C" int b200ret_block_table_build(const int64_t* term_offsets, const int32_t* doc_ids, int64_t nnz,
                                         int32_t n_terms, int32_t n_docs, int32_t block_docs,
                                         uint32_t* table, int32_t* status, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    B200RET_REQUIRE(term_offsets && table && status, "block_table_build: null pointer");
    B200RET_REQUIRE(n_terms > 0 && n_docs >= 0 && block_docs > 0 && nnz >= 0, "block_table_build: bad sizes");
    B200RET_REQUIRE(nnz < (int64_t{1} << 32), "block_table_build: nnz exceeds 32-bit positions");
    const int32_t n_blocks = (n_docs + block_docs - 1) / block_docs;
    B200RET_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int32_t), stream));
    block_table_empty_kernel<<<sm_count() * 4, 256, 0, stream>>>(term_offsets, n_terms, n_blocks, table);
    if (nnz > 0) {
        B200RET_REQUIRE(doc_ids != nullptr, "block_table_build: doc_ids is null");
        block_table_kernel<<<sm_count() * 8, 256, 0, stream>>>(term_offsets, doc_ids, nnz, n_terms, block_docs, n_blocks,
                                                               table, status);
    }
    B200RET_CUDA_CHECK(cudaGetLastError());
    int32_t host_status = 0;
    B200RET_CUDA_CHECK(cudaMemcpyAsync(&host_status, status, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    B200RET_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (host_status == B200RET_EUNSORTED) {
        set_err("block_table_build: a posting list is not ascending in doc id (build the CSR with sort_docs=1)");
        return B200RET_EUNSORTED;
    }
    if (host_status != 0) {
        set_err("block_table_build: doc id out of range [0, n_docs)");
        return B200RET_EINVAL;
    }
    return B200RET_OK;
}

static int sparse_layout_impl(int fmt, const uint32_t* table, const int32_t* doc_ids, const float* weights, int64_t nnz,
                              int32_t n_terms, int32_t n_docs, int32_t block_docs, int bank_order, void* postings_out,
                              cudaStream_t stream) {
    B200RET_REQUIRE(table && n_terms > 0 && n_docs >= 0 && nnz >= 0, "sparse_layout: bad arguments");
    B200RET_REQUIRE(block_docs > 0 && block_docs <= BANK_MAX_BLOCK_DOCS && block_docs % 32 == 0,
                    "sparse_layout: block_docs=%d must be a multiple of 32 and <= %d", block_docs, BANK_MAX_BLOCK_DOCS);
    const int32_t n_blocks = (n_docs + block_docs - 1) / block_docs;
    if (n_blocks == 0 || nnz == 0) return B200RET_OK;
    B200RET_REQUIRE(doc_ids && weights && postings_out, "sparse_layout: null pointer");
    B200RET_REQUIRE(reinterpret_cast<uintptr_t>(postings_out) % 8 == 0, "sparse_layout: postings_out must be 8-byte aligned");
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        B200RET_CUDA_CHECK(cudaFuncSetAttribute(posting_layout_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        B200RET_CUDA_CHECK(cudaFuncSetAttribute(posting_layout_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        attr_set.mark();
    }
    // Two launches: the shared memory of a warp is sized for the longest slice it may meet, and almost all slices are
    // short — sizing every warp for block_docs postings left 4 warps per SM and made the layout latency-bound (0.3 s).
    const int caps[2] = {min(BANK_SMALL_CAP, block_docs), block_docs};
    const int mins[2] = {0, caps[0]};
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1 && caps[1] <= caps[0]) break;
        const int warps = max(1, min(BANK_MAX_WARPS, static_cast<int>((220 * 1024) / bank_warp_smem(caps[pass]))));
        const size_t smem = static_cast<size_t>(warps) * bank_warp_smem(caps[pass]);
        if (fmt == 0)
            posting_layout_kernel<0><<<sm_count(), warps * 32, smem, stream>>>(table, doc_ids, weights, postings_out, n_terms, n_blocks,
                                                                              caps[pass], bank_order, mins[pass], caps[pass], block_docs);
        else
            posting_layout_kernel<1><<<sm_count(), warps * 32, smem, stream>>>(table, doc_ids, weights, postings_out, n_terms, n_blocks,
                                                                              caps[pass], bank_order, mins[pass], caps[pass], block_docs);
        count_launches(1);
    }
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}

extern "C" int b200ret_sparse_layout(const uint32_t* table, const int32_t* doc_ids, const float* weights, int64_t nnz,
                                     int32_t n_terms, int32_t n_docs, int32_t block_docs, int bank_order, void* postings_out,
                                     void* stream_) {
    return sparse_layout_impl(0, table, doc_ids, weights, nnz, n_terms, n_docs, block_docs, bank_order, postings_out,
                              static_cast<cudaStream_t>(stream_));
}

extern "C" int b200ret_sparse_layout_f16(const uint32_t* table, const int32_t* doc_ids, const float* weights, int64_t nnz,
                                         int32_t n_terms, int32_t n_docs, int32_t block_docs, int bank_order, void* postings_out,
                                         void* stream_) {
    B200RET_REQUIRE(block_docs <= 32768, "sparse_layout_f16: block-local doc ids take 15 bits (block_docs=%d)", block_docs);
    return sparse_layout_impl(1, table, doc_ids, weights, nnz, n_terms, n_docs, block_docs, bank_order, postings_out,
                              static_cast<cudaStream_t>(stream_));
}
