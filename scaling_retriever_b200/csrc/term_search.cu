// Gather-sum scoring over semantic-id codes + exact top-k on sm_100a.
//
// Replaces TermEncoderRetriever.get_doc_scores + torch.topk (reference scaling_retriever/indexer.py:621-641, :688):
//     doc_scores[b, n] = sum_l pred_scores[b, doc_encodings[n, l]]         pred_scores [bz, V], doc_encodings [N, L]
//     top_scores, top_idxes = torch.topk(doc_scores, k)
// The reference materialises pred_scores[:, doc_encodings] ([bz, 1e6, L] floats per 1 M-doc slice) and the full [bz, N]
// score matrix in HBM.  Here a CTA owns one query: its score table pred_scores[b, :] sits in shared memory (V * 4 bytes,
// up to 227 KB -> V <= 57,344; larger vocabularies gather through L1/L2), every thread streams the codes of its documents
// with 128-bit loads, sums the L table entries in code order (fp32) and appends the document to the query's candidate list
// only if the sum beats the query's running bound tau (candidates.cuh: rounds of growing doc ranges, radix-select cut to k
// between rounds) — the [bz, N] matrix never exists.  CTAs working on different queries walk the documents in the same
// order, so the code array streams from HBM once per query batch and is re-served from L2.
#include "candidates.cuh"

namespace b200ret {

constexpr int T_THREADS = 512;
constexpr int T_UNIT_DOCS = 4096;                 // round unit
constexpr int T_ROUND0_UNITS = 2;                 // first round / safe schedule: 8192 docs
constexpr int T_MAX_SMEM_VOCAB = (227 * 1024 - 1024) / 4;

struct TermParams {
    const float* pred;        // [n_queries][n_vocab]
    const int32_t* codes;     // [n_docs][code_len]
    int32_t n_vocab, code_len;
    int32_t doc_begin, doc_end;
    int32_t n_active, n_splits;
    const int32_t* q_list;
    const float* tau;
    uint64_t* cand;
    int32_t* cand_count;
    int32_t cap;
    float* dense_out;         // optional [n_queries][dense_stride]: write every score instead of selecting
    size_t dense_stride;
};

template <bool TABLE_IN_SMEM>
__global__ void __launch_bounds__(T_THREADS) term_gather_kernel(const TermParams p) {
    extern __shared__ __align__(16) float s_table[];
    const int n_items = p.n_active * p.n_splits;
    const int span = p.doc_end - p.doc_begin;
    const int per_split = ((span + p.n_splits - 1) / p.n_splits + 31) & ~31;     // whole warps per split
    const int l4 = p.code_len >> 2;
    int loaded_q = -1;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int qi = item / p.n_splits, split = item - qi * p.n_splits;
        const int q = p.q_list ? p.q_list[qi] : qi;
        const float* __restrict__ row = p.pred + static_cast<size_t>(q) * p.n_vocab;
        if (TABLE_IN_SMEM && q != loaded_q) {
            __syncthreads();                                    // everyone is done with the previous table
            for (int i = threadIdx.x; i < p.n_vocab; i += blockDim.x) s_table[i] = row[i];
            loaded_q = q;
            __syncthreads();
        }
        const float tq = p.tau ? __ldg(p.tau + q) : -INFINITY;
        const int d0 = p.doc_begin + split * per_split;
        const int d1 = min(p.doc_end, d0 + per_split);
        for (int base = d0; base < d1; base += blockDim.x) {    // block- and warp-uniform trip count
            const int d = base + threadIdx.x;
            const bool live = d < d1;
            float sum = 0.f;
            if (live) {
                const int4* __restrict__ c4 = reinterpret_cast<const int4*>(p.codes + static_cast<size_t>(d) * p.code_len);
                for (int l = 0; l < l4; ++l) {
                    const int4 c = __ldg(c4 + l);
                    if (TABLE_IN_SMEM) {
                        sum += s_table[c.x];
                        sum += s_table[c.y];
                        sum += s_table[c.z];
                        sum += s_table[c.w];
                    } else {
                        sum += __ldg(row + c.x);
                        sum += __ldg(row + c.y);
                        sum += __ldg(row + c.z);
                        sum += __ldg(row + c.w);
                    }
                }
            }
            if (p.dense_out) {
                if (live) p.dense_out[static_cast<size_t>(q) * p.dense_stride + d] = sum;
                continue;
            }
            const bool hit = live && sum > tq;
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (bal) {                                          // one slot reservation per warp
                int pos = 0;
                if (lane_id() == 0) pos = atomicAdd(p.cand_count + q, __popc(bal));
                pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(bal & lanemask_lt());
                if (hit && pos < p.cap) p.cand[static_cast<size_t>(q) * p.cap + pos] = cand_key(sum, d);
            }
        }
    }
}

static int term_cap(int k) { return (k + max(T_ROUND0_UNITS * T_UNIT_DOCS, (ROUND_GROWTH + 1) * k) + 1) & ~1; }

static int launch_term(TermParams r, int32_t n_vocab, cudaStream_t stream) {
    const bool in_smem = n_vocab <= T_MAX_SMEM_VOCAB;
    const size_t smem = in_smem ? static_cast<size_t>(n_vocab) * sizeof(float) : 0;
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        B200RET_CUDA_CHECK(cudaFuncSetAttribute(term_gather_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                T_MAX_SMEM_VOCAB * static_cast<int>(sizeof(float))));
        attr_set.mark();
    }
    // few queries: split the doc range so that every SM has work; many queries: one CTA keeps one query's table resident
    const int sms = sm_count();
    const int span_warps = (r.doc_end - r.doc_begin + T_THREADS - 1) / T_THREADS;
    r.n_splits = std::max(1, std::min(span_warps, (2 * sms + r.n_active - 1) / std::max(r.n_active, 1)));
    const int n_items = r.n_active * r.n_splits;
    const int ctas_per_sm = in_smem ? std::max(1, static_cast<int>((227 * 1024) / std::max<size_t>(smem + 1024, 1))) : 4;
    const int grid = std::max(1, std::min(n_items, sms * std::min(ctas_per_sm, 4)));
    if (in_smem)
        term_gather_kernel<true><<<grid, T_THREADS, smem, stream>>>(r);
    else
        term_gather_kernel<false><<<grid, T_THREADS, 0, stream>>>(r);
    count_launches(1);
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}

}  // namespace b200ret

using namespace b200ret;

extern "C" size_t b200ret_term_search_workspace_bytes(int32_t n_queries, int32_t k) {
    Workspace ws(nullptr, 0);
    return carve_cand(ws, n_queries, term_cap(k), nullptr) + 256;
}

static int term_check(const float* pred, const int32_t* codes, int32_t n_queries, int32_t n_vocab, int32_t n_docs, int32_t code_len) {
    B200RET_REQUIRE(n_queries >= 0 && n_docs >= 0 && n_vocab > 0, "term: bad sizes");
    B200RET_REQUIRE(code_len >= 4 && code_len % 4 == 0, "term: code_len=%d must be a positive multiple of 4 (the reference asserts 16/32/64/128)", code_len);
    B200RET_REQUIRE(n_queries == 0 || pred, "term: pred is null");
    B200RET_REQUIRE(n_docs == 0 || (codes && reinterpret_cast<uintptr_t>(codes) % 16 == 0), "term: codes null or not 16-byte aligned");
    return B200RET_OK;
}

extern "C" int b200ret_term_scores(const float* pred, const int32_t* codes, int32_t n_queries, int32_t n_vocab, int32_t n_docs,
                                   int32_t code_len, float* out_scores, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int rc = term_check(pred, codes, n_queries, n_vocab, n_docs, code_len);
    if (rc != B200RET_OK) return rc;
    if (n_queries == 0 || n_docs == 0) return B200RET_OK;
    B200RET_REQUIRE(out_scores, "term_scores: out_scores is null");
    TermParams r{};
    r.pred = pred;
    r.codes = codes;
    r.n_vocab = n_vocab;
    r.code_len = code_len;
    r.doc_begin = 0;
    r.doc_end = n_docs;
    r.n_active = n_queries;
    r.dense_out = out_scores;
    r.dense_stride = static_cast<size_t>(n_docs);
    return launch_term(r, n_vocab, stream);
}

extern "C" int b200ret_term_search(const float* pred, const int32_t* codes, int32_t n_queries, int32_t n_vocab, int32_t n_docs,
                                   int32_t code_len, int32_t k, float* out_scores, int64_t* out_ids, int32_t* out_counts,
                                   void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int rc = term_check(pred, codes, n_queries, n_vocab, n_docs, code_len);
    if (rc != B200RET_OK) return rc;
    B200RET_REQUIRE(k >= 1 && k <= B200RET_MAX_K, "term_search: k=%d outside [1, %d]", k, B200RET_MAX_K);
    if (n_queries == 0) return B200RET_OK;
    B200RET_REQUIRE(out_scores && out_ids && out_counts && workspace, "term_search: null pointer");
    if (workspace_bytes < b200ret_term_search_workspace_bytes(n_queries, k)) {
        set_err("term_search: workspace too small (%zu bytes)", workspace_bytes);
        return B200RET_EWORKSPACE;
    }
    Workspace ws(workspace, workspace_bytes);
    CandBuffers b;
    const int cap = term_cap(k);
    carve_cand(ws, n_queries, cap, &b);
    TermParams tp{};
    tp.pred = pred;
    tp.codes = codes;
    tp.n_vocab = n_vocab;
    tp.code_len = code_len;
    tp.tau = b.tau;
    tp.cand = b.cand;
    tp.cand_count = b.cand_count;
    tp.cap = cap;
    const int32_t n_units = (n_docs + T_UNIT_DOCS - 1) / T_UNIT_DOCS;
    auto launch_round = [&](int unit_begin, int unit_end, const int32_t* q_list, int32_t n_active) -> int {
        TermParams r = tp;
        r.doc_begin = unit_begin * T_UNIT_DOCS;
        r.doc_end = std::min(n_docs, unit_end * T_UNIT_DOCS);
        r.q_list = q_list;
        r.n_active = n_active;
        return launch_term(r, n_vocab, stream);
    };
    return run_search(launch_round, b, cap, k, n_queries, n_units, T_ROUND0_UNITS, -INFINITY, 0, out_scores, out_ids, out_counts,
                      stream);
}
