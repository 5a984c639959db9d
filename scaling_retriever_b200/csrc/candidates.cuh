// Per-query candidate lists shared by the sparse and the dense search: append in the scoring kernel's epilogue,
// cut back to the k best (and raise the eligibility bound tau) between rounds, sort + write rows at the end.
//
// Protocol (both searches): documents are scored in rounds over ascending doc-id ranges.  In a round every document
// whose score is > tau[q] is appended to query q's list (capacity cap = k + docs of the first round).  After the
// round, select_kernel keeps the k best under the total order (score desc, doc id asc) and sets tau[q] to the k-th
// score; later documents have larger ids, so a tie with tau can never displace a kept one and `> tau` is exact.
// The docs scored grow ROUND_GROWTH x per round, so about (ROUND_GROWTH - 1) * k new candidates per query and round
// survive tau on exchangeable data; the capacity is k + max(first-round docs, (ROUND_GROWTH + 1) * k), i.e. at least two k
// of head-room over that expectation for every k up to B200RET_MAX_K.  A list that overflows anyway (adversarial doc
// order) is flagged and the query is re-run with fixed rounds of the first-round size, which cannot overflow.  A sharded
// search with the tau exchange (b200ret_round_exchange) grows by exchange_growth(n_shards) per round instead and has a middle
// tier before that: overflowed queries are first re-run with the shard's own bounds and the plain schedule (run_search).
#pragma once
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "topk_select.cuh"

namespace b200ret {

constexpr int SELECT_THREADS = 512;
#ifndef B200RET_ROUND_GROWTH
#define B200RET_ROUND_GROWTH 4
#endif
constexpr int ROUND_GROWTH = B200RET_ROUND_GROWTH;   // total docs scored grow 4x per round (4, 16, 64, ... units)

struct CandBuffers {
    uint64_t* cand;        // [n_queries][cap] keys (cand_key)
    int32_t* cand_count;   // [n_queries] appended so far (may exceed cap -> overflow)
    float* tau;            // [n_queries]
    int32_t* overflow;     // [n_queries + 1]; last = any query overflowed
    int32_t* q_list;       // compacted indices of the queries to re-run
    int32_t* n_list;
    unsigned long long* work_counter;
};

inline size_t carve_cand(Workspace& ws, int32_t n_queries, int32_t cap, CandBuffers* b) {
    const size_t nq = static_cast<size_t>(n_queries > 0 ? n_queries : 1);
    CandBuffers tmp;
    tmp.cand = ws.take<uint64_t>(nq * cap);
    tmp.cand_count = ws.take<int32_t>(nq);
    tmp.tau = ws.take<float>(nq);
    tmp.overflow = ws.take<int32_t>(nq + 1);
    tmp.q_list = ws.take<int32_t>(nq);
    tmp.n_list = ws.take<int32_t>(1);
    tmp.work_counter = ws.take<unsigned long long>(1);
    if (b) *b = tmp;
    return ws.used;
}

// Non-final round count of the geometric schedule (= select launches between rounds) for `n_units` units.
inline int32_t schedule_exchanges(int32_t n_units, int32_t round0_units, int32_t growth) {
    int64_t unit = 0, size = round0_units;
    int32_t selects = 0;
    while (unit < n_units) {
        const int64_t end = (n_units - unit <= size) ? n_units : unit + size;
        if (end < n_units) ++selects;
        unit = end;
        size = unit * (growth - 1);
    }
    return selects;
}
inline int32_t exchange_growth(int32_t n_shards) {
    if (const char* e = getenv("B200RET_TEST_EXCHANGE_GROWTH")) return std::max(2, atoi(e));     // test hook: force list overflows
    return n_shards > 1 ? std::max(ROUND_GROWTH, n_shards + 1) : ROUND_GROWTH;
}

// Kernels + launchers live in candidates.cu (one definition for both searches).
int launch_cand_init(const CandBuffers& b, int32_t n_queries, float threshold, cudaStream_t stream);
int launch_cand_rearm(const CandBuffers& b, int32_t n_queries, float threshold, cudaStream_t stream);
// One CTA per active query: cut the candidate list to its k best keys, raise tau to the k-th score; `final` also sorts
// and writes the output row (ids + doc_id_base, tail padded with (-inf, -1)).
int launch_select(bool final, const CandBuffers& b, int32_t cap, int32_t k, int32_t n_queries, int32_t n_active,
                  const int32_t* q_list, int64_t doc_id_base, float* out_scores, int64_t* out_ids, int32_t* out_counts,
                  cudaStream_t stream, int32_t aux_rank = 0, float* aux = nullptr);
// tau exchange of a sharded search (b200ret_round_exchange): aux <- -inf; after the hook: tau <- max(tau, just below aux)
int launch_aux_init(float* aux, int32_t n_queries, cudaStream_t stream);
int launch_tau_raise(float* tau, const float* aux, int32_t n_queries, cudaStream_t stream);

// Round driver shared by both searches.  `launch_round(unit_begin, unit_end, q_list, n_active)` scores the doc units
// [unit_begin, unit_end) for the active queries and appends candidates; a "unit" is a fixed number of documents
// (a sparse doc block / a dense doc tile) and the candidate capacity must be >= k + round0_units * docs_per_unit.
template <class LaunchRound>
int run_rounds_once(LaunchRound& launch_round, const CandBuffers& b, int32_t cap, int32_t k, int32_t n_queries, int32_t n_active,
                    const int32_t* q_list, int32_t n_units, int32_t round0_units, bool safe, int64_t doc_id_base,
                    float* out_scores, int64_t* out_ids, int32_t* out_counts, cudaStream_t stream,
                    const b200ret_round_exchange* ex = nullptr) {
    int unit = 0, size = round0_units, exchanged = 0;
    const int growth = (ex && !safe) ? ex->growth : ROUND_GROWTH;    // a sharded search with the exchange takes larger steps
    auto exchange = [&]() -> int {     // all shards: MIN of the published bounds, then raise tau (see b200ret.h (3b))
        if (ex->hook(ex->user) != 0) {
            set_err("search: the tau-exchange hook failed");
            return B200RET_EINVAL;
        }
        ++exchanged;
        return launch_tau_raise(b.tau, ex->aux, n_queries, stream);
    };
    if (ex) {
        int rc = launch_aux_init(ex->aux, n_queries, stream);
        if (rc != B200RET_OK) return rc;
    }
    while (unit < n_units) {
        const int end = (n_units - unit <= size) ? n_units : unit + size;
        int rc = launch_round(unit, end, q_list, n_active);
        if (rc != B200RET_OK) return rc;
        if (end < n_units) {
            const bool ex_now = ex && exchanged < ex->n_exchanges;
            rc = launch_select(false, b, cap, k, n_queries, n_active, q_list, doc_id_base, nullptr, nullptr, nullptr, stream,
                               ex_now ? ex->aux_rank : 0, ex_now ? ex->aux : nullptr);
            if (rc != B200RET_OK) return rc;
            if (ex_now && (rc = exchange()) != B200RET_OK) return rc;
        }
        unit = end;
        // geometric schedule: the next round covers (ROUND_GROWTH - 1) x the docs seen so far, so about
        // (ROUND_GROWTH - 1) * k new candidates per query survive tau on exchangeable data (capacity: see the header)
        if (!safe) size = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(unit) * (growth - 1), n_units));
    }
    // a shard smaller than the largest one has fewer rounds: it still takes part in the remaining exchanges (collectives)
    while (ex && exchanged < ex->n_exchanges) {
        const int rc = exchange();
        if (rc != B200RET_OK) return rc;
    }
    return launch_select(true, b, cap, k, n_queries, n_active, q_list, doc_id_base, out_scores, out_ids, out_counts, stream);
}

template <class LaunchRound>
int run_search(LaunchRound& launch_round, const CandBuffers& b, int32_t cap, int32_t k, int32_t n_queries, int32_t n_units,
               int32_t round0_units, float threshold, int64_t doc_id_base, float* out_scores, int64_t* out_ids,
               int32_t* out_counts, cudaStream_t stream, const b200ret_round_exchange* ex = nullptr) {
    if (ex && (ex->aux_rank < 1 || ex->n_exchanges < 0 || ex->growth < 2 || !ex->aux || !ex->hook)) {
        set_err("search: bad round exchange (aux_rank %d, n_exchanges %d, growth %d)", ex->aux_rank, ex->n_exchanges, ex->growth);
        return B200RET_EINVAL;
    }
    int rc = launch_cand_init(b, n_queries, threshold, stream);
    if (rc != B200RET_OK) return rc;
    rc = run_rounds_once(launch_round, b, cap, k, n_queries, n_queries, nullptr, n_units, round0_units, /*safe=*/false,
                             doc_id_base, out_scores, out_ids, out_counts, stream, ex);
    if (rc != B200RET_OK) return rc;

    // The only host round trip: did any candidate list overflow?
    int32_t any_overflow = 0;
    B200RET_CUDA_CHECK(cudaMemcpyAsync(&any_overflow, b.overflow + n_queries, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    B200RET_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (!any_overflow) return B200RET_OK;

    if (ex) {
        // A sharded search took larger steps on the strength of the exchanged bounds and a list overflowed (shards that are
        // not exchangeable: the bound of the emptiest shard holds everybody back).  Middle tier: re-run those queries with
        // this shard's own bounds and the plain geometric schedule (no collectives: the other shards are not involved).
        rc = launch_cand_rearm(b, n_queries, threshold, stream);
        if (rc != B200RET_OK) return rc;
        int32_t n_mid = 0;
        B200RET_CUDA_CHECK(cudaMemcpyAsync(&n_mid, b.n_list, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        B200RET_CUDA_CHECK(cudaStreamSynchronize(stream));
        rc = run_rounds_once(launch_round, b, cap, k, n_queries, n_mid, b.q_list, n_units, round0_units, /*safe=*/false, doc_id_base,
                             out_scores, out_ids, out_counts, stream, nullptr);
        if (rc != B200RET_OK) return rc;
        B200RET_CUDA_CHECK(cudaMemcpyAsync(&any_overflow, b.overflow + n_queries, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        B200RET_CUDA_CHECK(cudaStreamSynchronize(stream));
        if (!any_overflow) return B200RET_OK;
    }

    rc = launch_cand_rearm(b, n_queries, threshold, stream);
    if (rc != B200RET_OK) return rc;
    int32_t n_list = 0;
    B200RET_CUDA_CHECK(cudaMemcpyAsync(&n_list, b.n_list, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    B200RET_CUDA_CHECK(cudaStreamSynchronize(stream));
    rc = run_rounds_once(launch_round, b, cap, k, n_queries, n_list, b.q_list, n_units, round0_units, /*safe=*/true, doc_id_base,
                         out_scores, out_ids, out_counts, stream);
    if (rc != B200RET_OK) return rc;
    B200RET_CUDA_CHECK(cudaMemcpyAsync(&any_overflow, b.overflow + n_queries, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    B200RET_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (any_overflow) {
        set_err("search: candidate list overflowed under the safe schedule (internal error)");
        return B200RET_EOVERFLOW;
    }
    return B200RET_OK;
}

}  // namespace b200ret
