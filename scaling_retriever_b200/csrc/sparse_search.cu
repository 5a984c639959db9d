// Batched term-at-a-time sparse scoring + exact top-k on sm_100a.
//
// Replaces SparseRetrieval.numba_score_float + select_topk (reference scaling_retriever/indexer.py:324-344,
// :315-322) as driven once per query by _sparse_retrieve_multithreaded (:405-474).
//
// Design:
//   * The reference keeps one fp32 score per document (35 MB/query at 8.8 M docs) in DRAM and does a random
//     read-modify-write per posting.  Here a WARP owns one (query, doc-block) work item: the scores of the
//     block's BLOCK_DOCS documents live in a warp-private shared-memory tile, the warp streams the slice of
//     each query term's posting list that falls into the block (located with the doc-block skip table), and
//     adds `q_w * d_w` with a separate fp32 multiply and add, term after term in query order.  Doc ids are
//     unique inside one posting list, so lanes never collide inside a term, and terms are serialised by
//     __syncwarp() -> no atomics, and every score is bit-identical to the reference's sequential sum.
//   * Work items are handed out block-major (all queries of doc block b before block b+1) through a global
//     counter, so at any time the whole GPU touches the postings of one or two doc blocks: the index is read
//     from HBM once per query batch and re-served from L2 for the other queries.
//   * Posting loads are 128-byte aligned warp rows (one posting per lane), software-pipelined PIPE deep in
//     registers across term boundaries.
//   * After the last term the warp sweeps its tile once (128-bit shared loads, zeroing as it goes) and appends
//     the documents with score > tau[q] to the query's candidate list.  Doc blocks are processed in rounds of
//     doubling size; after each round a select kernel cuts every list back to its k best and raises tau[q] to
//     the k-th score, so only ~k*ln(N) candidates per query ever reach HBM and the Q x N score matrix never
//     exists.  A candidate-list overflow (adversarial doc order) is detected and the affected queries are
//     re-run with a schedule that cannot overflow.
#include "common.cuh"
#include "topk_select.cuh"

namespace b200ret {

constexpr int BLOCK_DOCS = 3072;       // documents per warp-private score tile (12 KB fp32)
constexpr int SCORE_WARPS = 16;        // 16 x 12 KB = 192 KB of the 227 KB shared memory per SM
constexpr int SCORE_THREADS = SCORE_WARPS * 32;
constexpr int PIPE = 8;                // posting rows in flight per warp
constexpr int ROUND0_BLOCKS = 4;       // first round / safe-schedule round size, in doc blocks
constexpr int SELECT_THREADS = 512;

struct ScoreParams {
    const uint32_t* table;     // [n_terms][table_stride]
    size_t table_stride;       // n_blocks + 1
    const int32_t* doc_ids;
    const float* weights;
    const int32_t* q_offsets;
    const int32_t* q_terms;
    const float* q_weights;
    const int32_t* q_list;     // optional subset of query indices (safe re-run), else nullptr
    int32_t n_active;          // number of queries in this launch
    int32_t n_docs;
    int32_t blk_begin, blk_end;
    const float* tau;          // [n_queries] current eligibility bound (score must be > tau)
    uint64_t* cand;            // [n_queries][cap] candidate keys
    int32_t* cand_count;       // [n_queries]
    int32_t cap;
    unsigned long long* item_counter;
    float* dense_out;          // optional [n_active][n_blocks * BLOCK_DOCS]: dump every score instead of selecting
    size_t dense_stride;
};

__global__ void __launch_bounds__(SCORE_THREADS, 1) sparse_score_kernel(const ScoreParams p) {
    extern __shared__ __align__(16) float smem_acc[];
    const unsigned lane = lane_id();
    const unsigned warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    float* acc = smem_acc + warp * BLOCK_DOCS;
    for (int i = lane * 4; i < BLOCK_DOCS; i += 128) *reinterpret_cast<float4*>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    const long long n_items = static_cast<long long>(p.blk_end - p.blk_begin) * p.n_active;
    while (true) {
        long long item = 0;
        if (lane == 0) item = static_cast<long long>(atomicAdd(p.item_counter, 1ULL));
        item = __shfl_sync(FULL, item, 0);
        if (item >= n_items) break;
        const int blk = p.blk_begin + static_cast<int>(item / p.n_active);
        const int qi = static_cast<int>(item % p.n_active);
        const int q = p.q_list ? p.q_list[qi] : qi;
        const int doc_base = blk * BLOCK_DOCS;
        const int qb = p.q_offsets[q], qe = p.q_offsets[q + 1];

        for (int g = qb; g < qe; g += 32) {
            // Lane j holds term g+j of the query: its weight and its posting slice inside this doc block.
            unsigned seg_beg = 0, seg_end = 0;
            float seg_qw = 0.f;
            if (g + static_cast<int>(lane) < qe) {
                const int t = __ldg(p.q_terms + g + lane);
                seg_qw = __ldg(p.q_weights + g + lane);
                const uint32_t* e = p.table + static_cast<size_t>(t) * p.table_stride + blk;
                seg_beg = __ldg(e);
                seg_end = __ldg(e + 1);
            }
            unsigned pending = __ballot_sync(FULL, seg_end > seg_beg);   // non-empty slices, ascending term order

            // Warp-uniform cursor over the rows (32 postings, 128-byte aligned) of the pending slices.
            unsigned cur_row = 0, cur_beg = 0, cur_end = 0;
            float cur_qw = 0.f;
            bool exhausted = false;
            int id[PIPE];
            float w[PIPE], qw[PIPE];

            auto fetch = [&](int s) {
                if (cur_row >= cur_end) {
                    if (pending) {
                        const int j = __ffs(pending) - 1;
                        pending &= pending - 1;
                        cur_beg = __shfl_sync(FULL, seg_beg, j);
                        cur_end = __shfl_sync(FULL, seg_end, j);
                        cur_qw = __shfl_sync(FULL, seg_qw, j);
                        cur_row = cur_beg & ~31u;
                    } else {
                        exhausted = true;
                    }
                }
                id[s] = -1;
                const unsigned pos = cur_row + lane;
                if (!exhausted && pos >= cur_beg && pos < cur_end) {
                    id[s] = __ldg(p.doc_ids + pos);
                    w[s] = __ldg(p.weights + pos);
                }
                qw[s] = cur_qw;
                cur_row += 32;
            };
            auto consume = [&](int s) {
                if (id[s] >= 0) {
                    // reference arithmetic: scores[doc] += q * w  -> fp32 multiply, then fp32 add (no FMA)
                    const float v = __fmul_rn(qw[s], w[s]);
                    float* a = acc + (id[s] - doc_base);
                    *a = __fadd_rn(*a, v);
                }
                __syncwarp();   // orders this row's shared-memory updates before the next row (next term)
            };

#pragma unroll
            for (int s = 0; s < PIPE; ++s) fetch(s);
            while (true) {
                const bool was_exhausted = exhausted;
#pragma unroll
                for (int s = 0; s < PIPE; ++s) {
                    consume(s);
                    fetch(s);
                }
                if (was_exhausted) break;
            }
        }

        if (p.dense_out) {   // verification mode (b200ret_sparse_scores): write the tile out, no selection
            float* dst = p.dense_out + static_cast<size_t>(qi) * p.dense_stride + doc_base;
            for (int i = lane * 4; i < BLOCK_DOCS; i += 128) {
                *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(acc + i);
                *reinterpret_cast<float4*>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncwarp();
            continue;
        }
        // Sweep the tile: emit docs with score > tau[q], zero the tile for the next item.
        const float tq = __ldg(p.tau + q);
        const int limit = min(BLOCK_DOCS, p.n_docs - doc_base);
        uint64_t* cq = p.cand + static_cast<size_t>(q) * p.cap;
        for (int i = lane * 4; i < BLOCK_DOCS; i += 128) {
            const float4 v = *reinterpret_cast<const float4*>(acc + i);
            *reinterpret_cast<float4*>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
            const float m = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
            if (__any_sync(FULL, (m > tq) && (i < limit))) {
                const float vc[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const bool hit = (vc[c] > tq) && (i + c < limit);
                    const unsigned b = __ballot_sync(FULL, hit);
                    if (b) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(p.cand_count + q, __popc(b));
                        base = __shfl_sync(FULL, base, 0);
                        const int pos = base + __popc(b & lanemask_lt());
                        if (hit && pos < p.cap) cq[pos] = cand_key(vc[c], doc_base + i + c);
                    }
                }
            }
        }
        __syncwarp();
    }
}

__global__ void search_init_kernel(float* tau, int32_t* cand_count, int32_t* overflow, int32_t n_queries, float threshold) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_queries) {
        tau[i] = threshold;
        cand_count[i] = 0;
        overflow[i] = 0;
    }
    if (i == 0) overflow[n_queries] = 0;   // any-overflow flag
}

// Re-arm the overflowed queries for the safe re-run and compact their indices into q_list.
__global__ void search_rearm_kernel(float* tau, int32_t* cand_count, int32_t* overflow, int32_t n_queries, float threshold,
                                    int32_t* q_list, int32_t* n_list) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_queries && overflow[i]) {
        tau[i] = threshold;
        cand_count[i] = 0;
        overflow[i] = 0;
        q_list[atomicAdd(n_list, 1)] = i;
    }
}

// One CTA per query: cut the candidate list to its k best keys, raise tau to the k-th score.
// FINAL additionally sorts and writes the output row.
template <bool FINAL>
__global__ void __launch_bounds__(SELECT_THREADS) select_kernel(uint64_t* cand, int32_t* cand_count, int32_t cap, int32_t k,
                                                                float* tau, int32_t* overflow, int32_t n_queries,
                                                                const int32_t* q_list, int64_t doc_id_base,
                                                                float* out_scores, int64_t* out_ids, int32_t* out_counts) {
    extern __shared__ __align__(16) uint64_t skeys[];
    __shared__ uint32_t hist[256];
    __shared__ uint64_t bcast[2];
    __shared__ int out_pos;
    const int q = q_list ? q_list[blockIdx.x] : blockIdx.x;
    uint64_t* cq = cand + static_cast<size_t>(q) * cap;
    int c = cand_count[q];
    if (c > cap) {   // appended past the end: the list lost candidates -> flag for the safe re-run
        if (threadIdx.x == 0) {
            overflow[q] = 1;
            overflow[n_queries] = 1;
        }
        c = cap;
    }
    if (!FINAL && c <= k) return;   // nothing to cut; tau keeps its value (block-uniform exit)

    const int n_sort = FINAL ? next_pow2(max(min(c, k), 1)) : 0;
    for (int i = threadIdx.x; i < c; i += blockDim.x) skeys[i] = cq[i];
    if (threadIdx.x == 0) out_pos = 0;
    __syncthreads();
    int kept = c;
    if (c > k) {
        const uint64_t kth = block_radix_select_kth(skeys, c, k, hist, bcast);
        // Compact the k winners to the front of the global list (their order there is irrelevant).
        for (int i = threadIdx.x; i < c; i += blockDim.x) {
            const uint64_t key = skeys[i];
            if (key >= kth) cq[atomicAdd(&out_pos, 1)] = key;
        }
        kept = k;
        if (threadIdx.x == 0) {
            cand_count[q] = k;
            tau[q] = cand_score(kth);
        }
        __syncthreads();
        if (FINAL) {
            for (int i = threadIdx.x; i < k; i += blockDim.x) skeys[i] = cq[i];
        }
    }
    if (FINAL) {
        __syncthreads();
        for (int i = kept + threadIdx.x; i < n_sort; i += blockDim.x) skeys[i] = 0;
        block_bitonic_sort_desc(skeys, n_sort);
        for (int i = threadIdx.x; i < k; i += blockDim.x) {
            const bool live = i < kept;
            const uint64_t key = live ? skeys[i] : 0;
            out_scores[static_cast<size_t>(q) * k + i] = live ? cand_score(key) : -INFINITY;
            out_ids[static_cast<size_t>(q) * k + i] = live ? static_cast<int64_t>(cand_id(key)) + doc_id_base : -1;
        }
        if (threadIdx.x == 0) out_counts[q] = kept;
    }
}

static int search_cap(int k) { return k + ROUND0_BLOCKS * BLOCK_DOCS; }

struct SearchBuffers {
    uint64_t* cand;
    int32_t* cand_count;
    float* tau;
    int32_t* overflow;    // [n_queries + 1], last = any-overflow
    int32_t* q_list;
    int32_t* n_list;
    unsigned long long* item_counter;
};

static size_t carve(Workspace& ws, int32_t n_queries, int32_t k, SearchBuffers* b) {
    const size_t nq = static_cast<size_t>(n_queries > 0 ? n_queries : 1);
    SearchBuffers tmp;
    tmp.cand = ws.take<uint64_t>(nq * search_cap(k));
    tmp.cand_count = ws.take<int32_t>(nq);
    tmp.tau = ws.take<float>(nq);
    tmp.overflow = ws.take<int32_t>(nq + 1);
    tmp.q_list = ws.take<int32_t>(nq);
    tmp.n_list = ws.take<int32_t>(1);
    tmp.item_counter = ws.take<unsigned long long>(1);
    if (b) *b = tmp;
    return ws.used;
}

}  // namespace b200ret

using namespace b200ret;

extern "C" int32_t b200ret_sparse_block_docs(void) { return BLOCK_DOCS; }

extern "C" size_t b200ret_sparse_search_workspace_bytes(int32_t n_queries, int32_t k) {
    Workspace ws(nullptr, 0);
    return carve(ws, n_queries, k, nullptr) + 256;
}

namespace {

// Runs score+select rounds over all doc blocks for `n_active` queries (all, or the q_list subset).
// `safe` uses fixed rounds of ROUND0_BLOCKS blocks, which cannot overflow a list of capacity cap.
int run_rounds(ScoreParams sp, const SearchBuffers& b, int32_t n_queries, int32_t k, int32_t n_blocks, bool safe,
               int64_t doc_id_base, float* out_scores, int64_t* out_ids, int32_t* out_counts, cudaStream_t stream) {
    const size_t score_smem = static_cast<size_t>(SCORE_WARPS) * BLOCK_DOCS * sizeof(float);
    const size_t select_smem = static_cast<size_t>(sp.cap) * sizeof(uint64_t);
    const int grid = sm_count();
    int blk = 0, size = ROUND0_BLOCKS;
    while (blk < n_blocks) {
        const int end = (n_blocks - blk <= size) ? n_blocks : blk + size;
        sp.blk_begin = blk;
        sp.blk_end = end;
        B200RET_CUDA_CHECK(cudaMemsetAsync(b.item_counter, 0, sizeof(unsigned long long), stream));
        prof_begin(PROF_SPARSE_SCORE, stream);
        sparse_score_kernel<<<grid, SCORE_THREADS, score_smem, stream>>>(sp);
        prof_end(PROF_SPARSE_SCORE, stream);
        count_launches(1);
        if (end < n_blocks) {
            prof_begin(PROF_SPARSE_SELECT, stream);
            select_kernel<false><<<sp.n_active, SELECT_THREADS, select_smem, stream>>>(
                b.cand, b.cand_count, sp.cap, k, b.tau, b.overflow, n_queries, sp.q_list, doc_id_base, nullptr, nullptr, nullptr);
            prof_end(PROF_SPARSE_SELECT, stream);
            count_launches(1);
        }
        B200RET_CUDA_CHECK(cudaGetLastError());
        blk = end;
        if (!safe) size = blk;   // doubling: the next round covers as many docs as all rounds so far
    }
    prof_begin(PROF_SPARSE_SELECT, stream);
    select_kernel<true><<<sp.n_active, SELECT_THREADS, select_smem, stream>>>(
        b.cand, b.cand_count, sp.cap, k, b.tau, b.overflow, n_queries, sp.q_list, doc_id_base, out_scores, out_ids, out_counts);
    prof_end(PROF_SPARSE_SELECT, stream);
    count_launches(1);
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}

}  // namespace

extern "C" int b200ret_sparse_scores(const uint32_t* table, const int32_t* doc_ids, const float* weights,
                                     int32_t n_terms, int32_t n_docs, int32_t block_docs,
                                     const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights,
                                     int32_t n_queries, float* out_scores, void* workspace, size_t workspace_bytes,
                                     void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    B200RET_REQUIRE(block_docs == BLOCK_DOCS, "sparse_scores: table built for block_docs=%d, kernel uses %d", block_docs, BLOCK_DOCS);
    B200RET_REQUIRE(n_queries >= 0 && n_docs >= 0 && n_terms > 0, "sparse_scores: bad sizes");
    if (n_queries == 0 || n_docs == 0) return B200RET_OK;
    B200RET_REQUIRE(table && q_offsets && out_scores && workspace && workspace_bytes >= 256, "sparse_scores: null pointer / workspace < 256 B");
    B200RET_CUDA_CHECK(cudaFuncSetAttribute(sparse_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            SCORE_WARPS * BLOCK_DOCS * (int)sizeof(float)));
    const int32_t n_blocks = (n_docs + BLOCK_DOCS - 1) / BLOCK_DOCS;
    ScoreParams sp{};
    sp.table = table;
    sp.table_stride = static_cast<size_t>(n_blocks) + 1;
    sp.doc_ids = doc_ids;
    sp.weights = weights;
    sp.q_offsets = q_offsets;
    sp.q_terms = q_terms;
    sp.q_weights = q_weights;
    sp.n_active = n_queries;
    sp.n_docs = n_docs;
    sp.blk_begin = 0;
    sp.blk_end = n_blocks;
    sp.item_counter = static_cast<unsigned long long*>(workspace);
    sp.dense_out = out_scores;
    sp.dense_stride = static_cast<size_t>(n_blocks) * BLOCK_DOCS;
    B200RET_CUDA_CHECK(cudaMemsetAsync(sp.item_counter, 0, sizeof(unsigned long long), stream));
    sparse_score_kernel<<<sm_count(), SCORE_THREADS, static_cast<size_t>(SCORE_WARPS) * BLOCK_DOCS * sizeof(float), stream>>>(sp);
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}

extern "C" int b200ret_sparse_search(const uint32_t* table, const int32_t* doc_ids, const float* weights,
                                     int32_t n_terms, int32_t n_docs, int32_t block_docs,
                                     const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights,
                                     int32_t n_queries, int32_t k, float threshold, int64_t doc_id_base,
                                     float* out_scores, int64_t* out_ids, int32_t* out_counts,
                                     void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    B200RET_REQUIRE(block_docs == BLOCK_DOCS, "sparse_search: table built for block_docs=%d, kernel uses %d", block_docs, BLOCK_DOCS);
    B200RET_REQUIRE(k >= 1 && k <= B200RET_MAX_K, "sparse_search: k=%d outside [1, %d]", k, B200RET_MAX_K);
    B200RET_REQUIRE(n_queries >= 0 && n_docs >= 0 && n_terms > 0, "sparse_search: bad sizes");
    if (n_queries == 0) return B200RET_OK;
    B200RET_REQUIRE(table && q_offsets && out_scores && out_ids && out_counts && workspace, "sparse_search: null pointer");
    if (workspace_bytes < b200ret_sparse_search_workspace_bytes(n_queries, k)) {
        set_err("sparse_search: workspace too small (%zu bytes)", workspace_bytes);
        return B200RET_EWORKSPACE;
    }
    static bool attrs_set = false;
    if (!attrs_set) {
        B200RET_CUDA_CHECK(cudaFuncSetAttribute(sparse_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                SCORE_WARPS * BLOCK_DOCS * (int)sizeof(float)));
        const int max_sel = search_cap(B200RET_MAX_K) * (int)sizeof(uint64_t);
        B200RET_CUDA_CHECK(cudaFuncSetAttribute(select_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_sel));
        B200RET_CUDA_CHECK(cudaFuncSetAttribute(select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_sel));
        attrs_set = true;
    }
    Workspace ws(workspace, workspace_bytes);
    SearchBuffers b;
    carve(ws, n_queries, k, &b);
    const int32_t n_blocks = (n_docs + BLOCK_DOCS - 1) / BLOCK_DOCS;

    ScoreParams sp;
    sp.table = table;
    sp.table_stride = static_cast<size_t>(n_blocks) + 1;
    sp.doc_ids = doc_ids;
    sp.weights = weights;
    sp.q_offsets = q_offsets;
    sp.q_terms = q_terms;
    sp.q_weights = q_weights;
    sp.q_list = nullptr;
    sp.n_active = n_queries;
    sp.n_docs = n_docs;
    sp.blk_begin = sp.blk_end = 0;
    sp.tau = b.tau;
    sp.cand = b.cand;
    sp.cand_count = b.cand_count;
    sp.cap = search_cap(k);
    sp.item_counter = b.item_counter;
    sp.dense_out = nullptr;
    sp.dense_stride = 0;

    const int init_grid = (n_queries + 255) / 256;
    search_init_kernel<<<init_grid, 256, 0, stream>>>(b.tau, b.cand_count, b.overflow, n_queries, threshold);
    count_launches(1);
    B200RET_CUDA_CHECK(cudaGetLastError());
    int rc = run_rounds(sp, b, n_queries, k, n_blocks, /*safe=*/false, doc_id_base, out_scores, out_ids, out_counts, stream);
    if (rc != B200RET_OK) return rc;

    // The only host round trip: did any candidate list overflow?
    int32_t any_overflow = 0;
    B200RET_CUDA_CHECK(cudaMemcpyAsync(&any_overflow, b.overflow + n_queries, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    B200RET_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (!any_overflow) return B200RET_OK;

    B200RET_CUDA_CHECK(cudaMemsetAsync(b.n_list, 0, sizeof(int32_t), stream));
    search_rearm_kernel<<<init_grid, 256, 0, stream>>>(b.tau, b.cand_count, b.overflow, n_queries, threshold, b.q_list, b.n_list);
    B200RET_CUDA_CHECK(cudaMemsetAsync(b.overflow + n_queries, 0, sizeof(int32_t), stream));
    int32_t n_list = 0;
    B200RET_CUDA_CHECK(cudaMemcpyAsync(&n_list, b.n_list, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    B200RET_CUDA_CHECK(cudaStreamSynchronize(stream));
    sp.q_list = b.q_list;
    sp.n_active = n_list;
    rc = run_rounds(sp, b, n_queries, k, n_blocks, /*safe=*/true, doc_id_base, out_scores, out_ids, out_counts, stream);
    if (rc != B200RET_OK) return rc;
    B200RET_CUDA_CHECK(cudaMemcpyAsync(&any_overflow, b.overflow + n_queries, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    B200RET_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (any_overflow) {
        set_err("sparse_search: candidate list overflowed under the safe schedule (internal error)");
        return B200RET_EOVERFLOW;
    }
    return B200RET_OK;
}
