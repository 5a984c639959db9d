// Rank metrics on the GPU over the [Q, k] result rows: reciprocal rank @cut and recall @cuts per query.
//
// Quick-regression replacement for the run.json -> pytrec_eval round trip of the reference (scaling_retriever/utils/metrics.py:
// mrr_k :22-30 = truncate_run to the top `k` by score, then trec_eval's recip_rank; recall_k :32-42 = trec_eval's recall_<k>):
//   recip_rank@c = 1 / (rank of the first relevant doc among the first c rows), 0 if none;
//   recall@c     = (#relevant docs among the first c rows) / (#relevant docs of the query).
// Rows are the search output (sorted by score desc); relevance judgements are CSR-packed per query (rel_ids ascending).
// One warp per query: lane j tests rows j, j + 32, ... against the query's relevant list by binary search.
#include "common.cuh"

namespace b200ret {

constexpr int METRIC_MAX_CUTS = 8;
struct MetricParams {
    const int64_t* ids;
    const int32_t* counts;
    int32_t n_queries, k;
    const int64_t* rel_offsets;
    const int64_t* rel_ids;
    int32_t mrr_cut, n_cuts;
    int32_t cuts[METRIC_MAX_CUTS];
    float* out_rr;
    float* out_recall;
};

__global__ void rank_metrics_kernel(const MetricParams p) {
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= p.n_queries) return;
    const unsigned lane = lane_id();
    const int64_t r0 = p.rel_offsets[q], r1 = p.rel_offsets[q + 1];
    const int n_rel = static_cast<int>(r1 - r0);
    const int live = p.counts ? min(p.counts[q], p.k) : p.k;
    int max_cut = p.mrr_cut;
    for (int c = 0; c < p.n_cuts; ++c) max_cut = max(max_cut, p.cuts[c]);
    const int n = min(live, max_cut);
    int first = 0x7fffffff;
    int hits[METRIC_MAX_CUTS];
#pragma unroll
    for (int c = 0; c < METRIC_MAX_CUTS; ++c) hits[c] = 0;
    for (int j = lane; j < n; j += 32) {
        const int64_t id = p.ids[static_cast<size_t>(q) * p.k + j];
        int64_t lo = r0, hi = r1;                      // binary search in the ascending relevant list
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (p.rel_ids[mid] < id) lo = mid + 1; else hi = mid;
        }
        if (id >= 0 && lo < r1 && p.rel_ids[lo] == id) {
            first = min(first, j);
#pragma unroll
            for (int c = 0; c < METRIC_MAX_CUTS; ++c) hits[c] += (c < p.n_cuts && j < p.cuts[c]) ? 1 : 0;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        first = min(first, __shfl_xor_sync(0xffffffffu, first, off));
#pragma unroll
        for (int c = 0; c < METRIC_MAX_CUTS; ++c) hits[c] += __shfl_xor_sync(0xffffffffu, hits[c], off);
    }
    if (lane == 0) {
        p.out_rr[q] = (first < p.mrr_cut) ? 1.0f / static_cast<float>(first + 1) : 0.f;
        for (int c = 0; c < p.n_cuts; ++c)
            p.out_recall[static_cast<size_t>(q) * p.n_cuts + c] = n_rel > 0 ? static_cast<float>(hits[c]) / static_cast<float>(n_rel) : 0.f;
    }
}

}  // namespace b200ret

using namespace b200ret;

extern "C" int b200ret_rank_metrics(const int64_t* ids, const int32_t* counts, int32_t n_queries, int32_t k,
                                    const int64_t* rel_offsets, const int64_t* rel_ids, int32_t mrr_cut,
                                    const int32_t* recall_cuts_host, int32_t n_cuts, float* out_rr, float* out_recall,
                                    void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    B200RET_REQUIRE(n_queries >= 0 && k >= 1 && mrr_cut >= 1, "rank_metrics: bad sizes");
    B200RET_REQUIRE(n_cuts >= 0 && n_cuts <= METRIC_MAX_CUTS, "rank_metrics: at most %d recall cut-offs", METRIC_MAX_CUTS);
    if (n_queries == 0) return B200RET_OK;
    B200RET_REQUIRE(ids && rel_offsets && out_rr && (n_cuts == 0 || (recall_cuts_host && out_recall)), "rank_metrics: null pointer");
    MetricParams p{};
    p.ids = ids;
    p.counts = counts;
    p.n_queries = n_queries;
    p.k = k;
    p.rel_offsets = rel_offsets;
    p.rel_ids = rel_ids;
    p.mrr_cut = mrr_cut;
    p.n_cuts = n_cuts;
    for (int c = 0; c < n_cuts; ++c) {
        B200RET_REQUIRE(recall_cuts_host[c] >= 1, "rank_metrics: cut-off %d", recall_cuts_host[c]);
        p.cuts[c] = recall_cuts_host[c];
    }
    p.out_rr = out_rr;
    p.out_recall = out_recall;
    const int warps = 8;
    rank_metrics_kernel<<<(n_queries + warps - 1) / warps, warps * 32, 0, stream>>>(p);
    count_launches(1);
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}
