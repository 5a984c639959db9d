/*
 * CPU restatement (plain C + OpenMP) of the reference's sparse retrieval hot path.
 *
 * TEST INFRASTRUCTURE ONLY — built into oracle/liboracle.so and used by tests/ as the checker at sizes where
 * the numpy restatement is too slow, and by bench.py as the timed CPU baseline ("port").  The product package
 * never loads it.
 *
 * Restates, per query (paths relative to the reference checkout):
 *   SparseRetrieval.numba_score_float   scaling_retriever/indexer.py:324-344
 *       scores = zeros(N, f32); for each query term in order: scores[id] += q * w   (fp32 mul, fp32 add)
 *       filtered = argwhere(scores > threshold)
 *   SparseRetrieval.select_topk          scaling_retriever/indexer.py:315-322
 *       top-k of the filtered scores (the reference uses np.argpartition: unordered, ties arbitrary)
 *   IndexDictOfArray.add_batch_document  scaling_retriever/utils/inverted_index.py:67-76  (oracle_build_csr)
 * The reference parallelises with 4 Python threads x numba prange (indexer.py:459, :339); this port runs one
 * query per OpenMP thread with a thread-private score array, which is the same work decomposition without
 * the nested oversubscription.  Build: gcc -O3 -fopenmp -ffp-contract=off -shared -fPIC (see oracle/Makefile);
 * -ffp-contract=off keeps the multiply and the add separate like numba's un-fused float32 code.
 *
 * Output rows are sorted by (score desc, doc id asc) so they can be compared 1:1 with the GPU rows; that
 * final ordering of <= k items is not part of the reference and costs O(k log k) per query.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    float score;
    int32_t id;
} cand_t;

/* (score desc, id asc): returns 1 if a ranks before b */
static inline int before(cand_t a, cand_t b) { return a.score > b.score || (a.score == b.score && a.id < b.id); }

static int cmp_cand(const void* pa, const void* pb) {
    cand_t a = *(const cand_t*)pa, b = *(const cand_t*)pb;
    if (before(a, b)) return -1;
    if (before(b, a)) return 1;
    return 0;
}

/* Quickselect: afterwards c[0..k) are the k best under `before` (unordered), like np.argpartition. */
static void select_k(cand_t* c, int64_t n, int64_t k) {
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        cand_t pivot = c[lo + (hi - lo) / 2];
        int64_t i = lo, j = hi;
        while (i <= j) {
            while (before(c[i], pivot)) ++i;
            while (before(pivot, c[j])) --j;
            if (i <= j) {
                cand_t t = c[i];
                c[i] = c[j];
                c[j] = t;
                ++i;
                --j;
            }
        }
        if (k - 1 <= j) hi = j;
        else if (k - 1 >= i) lo = i;
        else break;
    }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/*
 * term_offsets[n_terms+1], doc_ids[nnz], weights[nnz]: CSR posting lists (any order inside a list).
 * q_offsets[nq+1], q_terms, q_weights: CSR-packed queries.
 * out_scores/out_ids [nq][k], out_counts[nq]: unused slots hold (-inf, -1).
 * Returns 0, or -1 on allocation failure.
 */
int oracle_sparse_search(const int64_t* term_offsets, const int32_t* doc_ids, const float* weights, int32_t n_docs,
                         const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights, int32_t nq, int32_t k,
                         float threshold, float* out_scores, int64_t* out_ids, int32_t* out_counts, int32_t n_threads) {
    int failed = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        float* scores = (float*)malloc(sizeof(float) * (size_t)(n_docs > 0 ? n_docs : 1));
        cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * (size_t)(n_docs > 0 ? n_docs : 1));
        if (!scores || !cand) {
#pragma omp atomic write
            failed = 1;
        }
#pragma omp barrier
        if (!failed) {
#pragma omp for schedule(dynamic, 1)
            for (int32_t q = 0; q < nq; ++q) {
                memset(scores, 0, sizeof(float) * (size_t)n_docs);                  /* np.zeros(size_collection) */
                for (int32_t j = q_offsets[q]; j < q_offsets[q + 1]; ++j) {           /* terms in query order */
                    const int32_t t = q_terms[j];
                    const float qw = q_weights[j];
                    for (int64_t p = term_offsets[t]; p < term_offsets[t + 1]; ++p) {
                        const float prod = qw * weights[p];                           /* fp32 multiply */
                        scores[doc_ids[p]] = scores[doc_ids[p]] + prod;               /* fp32 add      */
                    }
                }
                int64_t h = 0;
                for (int32_t d = 0; d < n_docs; ++d)                                   /* argwhere(scores > threshold) */
                    if (scores[d] > threshold) {
                        cand[h].score = scores[d];
                        cand[h].id = d;
                        ++h;
                    }
                int64_t kept = h;
                if (h > k) {                                                           /* select_topk */
                    select_k(cand, h, k);
                    kept = k;
                }
                qsort(cand, (size_t)kept, sizeof(cand_t), cmp_cand);
                for (int64_t i = 0; i < k; ++i) {
                    out_scores[(size_t)q * k + i] = i < kept ? cand[i].score : -INFINITY;
                    out_ids[(size_t)q * k + i] = i < kept ? (int64_t)cand[i].id : -1;
                }
                out_counts[q] = (int32_t)kept;
            }
        }
        free(scores);
        free(cand);
    }
    return failed ? -1 : 0;
}

/* Dense score vector of ONE query (for bit-exactness checks): scores[n_docs]. */
void oracle_sparse_scores(const int64_t* term_offsets, const int32_t* doc_ids, const float* weights, int32_t n_docs,
                          const int32_t* q_terms, const float* q_weights, int32_t nnz_q, float* scores) {
    memset(scores, 0, sizeof(float) * (size_t)n_docs);
    for (int32_t j = 0; j < nnz_q; ++j) {
        const int32_t t = q_terms[j];
        const float qw = q_weights[j];
        for (int64_t p = term_offsets[t]; p < term_offsets[t + 1]; ++p) {
            const float prod = qw * weights[p];
            scores[doc_ids[p]] = scores[doc_ids[p]] + prod;
        }
    }
}

/*
 * COO (feed order) -> CSR, stable in feed order: what add_batch_document's per-term appends produce.
 * Counting sort by term id.
 */
int oracle_build_csr(const int32_t* rows, const int32_t* cols, const float* vals, int64_t nnz, int32_t n_terms,
                     int64_t* term_offsets, int32_t* doc_ids, float* weights) {
    int64_t* cursor = (int64_t*)calloc((size_t)n_terms + 1, sizeof(int64_t));
    if (!cursor) return -1;
    for (int64_t i = 0; i < nnz; ++i) cursor[cols[i] + 1]++;
    term_offsets[0] = 0;
    for (int32_t t = 0; t < n_terms; ++t) term_offsets[t + 1] = term_offsets[t] + cursor[t + 1];
    for (int32_t t = 0; t < n_terms; ++t) cursor[t] = term_offsets[t];
    for (int64_t i = 0; i < nnz; ++i) {
        const int64_t p = cursor[cols[i]]++;
        doc_ids[p] = rows[i];
        weights[p] = vals[i];
    }
    free(cursor);
    return 0;
}
