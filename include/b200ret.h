/*
 * b200ret.h — C ABI of the B200-native first-stage retrieval engine.
 *
 * The reference (HansiZeng/scaling-retriever) is pure Python: its retrieval hot path calls numba,
 * numpy and faiss-cpu directly and has no FFI of its own.  This header is the seam a maintainer binds
 * (ctypes) underneath the reference's Python classes; every entry point names the
 * reference interface it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the parameter name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are asynchronous
 *     with respect to the host unless stated otherwise.
 *   - Every function returns 0 on success, a negative B200RET_E* code on failure;
 *     b200ret_last_error() returns a thread-local message for the last failure.
 *   - The library never allocates device memory: callers hand in a workspace whose size the matching
 *     *_workspace_bytes() function reports (the Python host side uses torch's caching allocator).
 *   - There is no CPU fallback anywhere: without a CUDA device every compute entry point fails.
 */
#ifndef B200RET_H
#define B200RET_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RET_VERSION 1

#define B200RET_OK 0
#define B200RET_EINVAL (-1)    /* bad argument (null pointer, size, alignment, k too large ...) */
#define B200RET_ECUDA (-2)     /* a CUDA runtime call or kernel launch failed */
#define B200RET_EWORKSPACE (-3)/* workspace too small */
#define B200RET_EUNSORTED (-4) /* posting lists are not ascending in doc id (block table build) */
#define B200RET_EOVERFLOW (-5) /* internal candidate buffer overflow that the safe schedule could not absorb */

/* Largest top-k the select kernels support (shared-memory bound). */
#define B200RET_MAX_K 4096

int b200ret_version(void);
const char* b200ret_last_error(void);

/* Device facts the host side needs for grid sizing / sanity (cudaGetDeviceProperties). */
int b200ret_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin_bytes);

/* Measurement hooks (bench.py): when enabled, the dominant kernels are bracketed with CUDA events on the
 * launching stream.  kind: 0 sparse score kernel, 1 sparse select kernels, 2 dense GEMM+top-k kernel,
 * 3 CSR radix-sort passes.  b200ret_profile_read synchronises the device, returns the summed duration
 * (ms) and count of the timed launches of `kind` since the last read, plus the number of ALL kernels
 * this library launched since the last read with all_launches != NULL (then reset). */
int b200ret_profile_enable(int on);
int b200ret_profile_read(int kind, double* total_ms, int64_t* timed_launches, int64_t* all_launches);

/* ------------------------------------------------------------------------------------------------
 * (1) Sparse index build: COO -> CSR posting lists.
 * Replaces IndexDictOfArray.add_batch_document (scaling_retriever/utils/inverted_index.py:67-76) and
 * the array->numpy conversion of IndexDictOfArray.save (:84-88) / SparseIndexer.index
 * (scaling_retriever/indexer.py:298-304).
 *
 * Input: `nnz` postings in FEED order (the order add_batch_document would have seen them):
 *   rows[i] = document row id, cols[i] = term id in [0, n_terms), vals[i] = fp32 weight.
 * Output (canonical CSR, bit-exact with the reference's per-term arrays):
 *   term_offsets[n_terms + 1] (int64), doc_ids[nnz] (int32), weights[nnz] (fp32) with the postings of
 *   term t at [term_offsets[t], term_offsets[t+1]) in FEED order (stable) when sort_docs == 0, or in
 *   ascending doc-id order when sort_docs != 0 (what the search kernels need; identical to feed order
 *   for the single-rank row-major feed of SparseIndexer.index).
 * Implementation: hand-written stable LSD radix sort (9-bit digits, ballot multisplit per warp, every 8192-posting
 * tile staged in digit order in shared memory and written out as contiguous runs) + boundary scan.
 * ---------------------------------------------------------------------------------------------- */
size_t b200ret_csr_build_workspace_bytes(int64_t nnz, int32_t n_terms, int32_t n_docs, int sort_docs);

int b200ret_csr_build(const int32_t* rows, const int32_t* cols, const float* vals, int64_t nnz,
                      int32_t n_terms, int32_t n_docs, int sort_docs,
                      int64_t* term_offsets, int32_t* doc_ids, float* weights,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Doc-block skip table over a doc-sorted CSR:  table[t * (n_blocks + 1) + b] = absolute position of
 * the first posting of term t whose doc id is >= b * block_docs (b == n_blocks: term_offsets[t+1]).
 * n_blocks = ceil(n_docs / block_docs).  No reference counterpart (the numba kernel walks whole
 * lists, indexer.py:334-341); it lets one warp own a doc block and stream only its slice of a list.
 * `status` is a device int32 the kernel sets to B200RET_EUNSORTED if a list is not ascending; the call
 * synchronises the stream and returns that code.
 * ---------------------------------------------------------------------------------------------- */
int b200ret_block_table_build(const int64_t* term_offsets, const int32_t* doc_ids, int64_t nnz,
                              int32_t n_terms, int32_t n_docs, int32_t block_docs,
                              uint32_t* table, int32_t* status, void* stream);

/* Search-side posting array, after b200ret_block_table_build: postings_out[nnz] holds 8-byte elements
 * {int32 doc id, fp32 weight} at the SAME positions as the doc-sorted CSR (the skip table addresses both), so
 * the search kernel fetches a posting with one 64-bit load.  With bank_order != 0 the postings inside every
 * (term, doc block) slice are additionally permuted into bank-quantile order (bank = doc id mod 32; every
 * bank's postings spread evenly over the slice): 32 consecutive postings then hold about the even share
 * 32*c/len of each bank instead of runs, which halves the shared-memory bank conflicts of the score tile.
 * The multiset of postings per slice is unchanged, so scores are unaffected (a doc occurs once per list).
 * The canonical CSR arrays are only read. */
int b200ret_sparse_layout(const uint32_t* table, const int32_t* doc_ids, const float* weights, int64_t nnz,
                          int32_t n_terms, int32_t n_docs, int32_t block_docs, int bank_order,
                          void* postings_out, void* stream);

/* Opt-in COMPRESSED posting format (north_star: "int32 doc ids, fp32 or fp16 weights"): postings_out[nnz] holds ONE 32-bit
 * word per posting, fp16(weight, round-to-nearest-even) << 16 | (doc id - first doc id of its doc block), at the same CSR
 * positions and in the same bank-quantile order as b200ret_sparse_layout.  4 B/posting instead of 8; searched with the
 * *_f16 entry points below.  Scores are then the reference's arithmetic applied to the fp16-ROUNDED weights (bit-identical to
 * numba_score_float run on weights.astype(float16).astype(float32)); they differ from the fp32-weight scores by up to
 * 2^-11 relative per term, so this is not the parity format.  block_docs <= 32768. */
int b200ret_sparse_layout_f16(const uint32_t* table, const int32_t* doc_ids, const float* weights, int64_t nnz,
                              int32_t n_terms, int32_t n_docs, int32_t block_docs, int bank_order,
                              void* postings_out, void* stream);

/* Doc-block size (documents per warp-private accumulator tile) compiled into the search kernel. */
int32_t b200ret_sparse_block_docs(void);

/* ------------------------------------------------------------------------------------------------
 * (2) Sparse query scoring + top-k for a batch of queries.
 * Replaces SparseRetrieval.numba_score_float (scaling_retriever/indexer.py:324-344) followed by
 * SparseRetrieval.select_topk (:315-322) for every query of _sparse_retrieve_multithreaded (:405-474).
 *
 * Queries are CSR-packed: query q owns q_terms/q_weights[q_offsets[q] .. q_offsets[q+1]) in the order
 * the reference iterates them (ascending term id from torch.nonzero, indexer.py:396-401).
 * Arithmetic is the reference's: per doc, fp32 `score += q_w * d_w` as a separate round-to-nearest
 * multiply and add (no FMA) in query-term order -> scores are bit-identical to numba_score_float.
 * Only docs with score > threshold are eligible (strict, indexer.py:342).
 *
 * Output per query q (row q of out_scores/out_ids, `k` columns): the out_counts[q] = min(k, #eligible)
 * best docs sorted by (score descending, doc id ascending); unused tail slots hold (-inf, -1).
 * Doc ids are LOCAL row ids + doc_id_base (shard offset for multi-GPU use).
 * `n_docs` is the reference's size_collection; `table` / `postings` come from b200ret_block_table_build /
 * b200ret_sparse_layout.
 * ---------------------------------------------------------------------------------------------- */
size_t b200ret_sparse_search_workspace_bytes(int32_t n_queries, int32_t k);

int b200ret_sparse_search(const uint32_t* table, const void* postings,
                          int32_t n_terms, int32_t n_docs, int32_t block_docs,
                          const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights,
                          int32_t n_queries, int32_t k, float threshold, int64_t doc_id_base,
                          float* out_scores, int64_t* out_ids, int32_t* out_counts,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Verification / API-parity entry point: the full fp32 score vector of every query, i.e. the `scores`
 * array inside numba_score_float (indexer.py:332-341) before the threshold filter.
 * out_scores: [n_queries][n_blocks * block_docs] (row stride padded to whole doc blocks; entries past
 * n_docs are 0).  Same kernel, same arithmetic as b200ret_sparse_search.  workspace: >= 256 bytes. */
int b200ret_sparse_scores(const uint32_t* table, const void* postings,
                          int32_t n_terms, int32_t n_docs, int32_t block_docs,
                          const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights,
                          int32_t n_queries, float* out_scores, void* workspace, size_t workspace_bytes,
                          void* stream);

/* The same two calls over the compressed posting array of b200ret_sparse_layout_f16 (same arguments and outputs). */
int b200ret_sparse_search_f16(const uint32_t* table, const void* postings,
                              int32_t n_terms, int32_t n_docs, int32_t block_docs,
                              const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights,
                              int32_t n_queries, int32_t k, float threshold, int64_t doc_id_base,
                              float* out_scores, int64_t* out_ids, int32_t* out_counts,
                              void* workspace, size_t workspace_bytes, void* stream);
int b200ret_sparse_scores_f16(const uint32_t* table, const void* postings,
                              int32_t n_terms, int32_t n_docs, int32_t block_docs,
                              const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights,
                              int32_t n_queries, float* out_scores, void* workspace, size_t workspace_bytes,
                              void* stream);

/* ------------------------------------------------------------------------------------------------
 * (3) Dense flat inner-product search.
 * Replaces faiss.IndexFlatIP.search as called by DenseFlatIndexer.search_knn
 * (scaling_retriever/indexer.py:210-214): exact top-k of Q . D^T, sorted descending.
 * corpus: bf16 [n_docs, dim] row-major; queries: bf16 [n_queries, dim] row-major (bf16 storage,
 * fp32 accumulation on tcgen05 tensor cores).  dim must be a multiple of 64.
 * Output rows as in b200ret_sparse_search; unused tail slots hold (-inf, -1) (faiss pads with -1).
 * ---------------------------------------------------------------------------------------------- */
size_t b200ret_dense_search_workspace_bytes(int32_t n_queries, int32_t n_docs, int32_t dim, int32_t k);

int b200ret_dense_search(const void* corpus_bf16, const void* queries_bf16,
                         int32_t n_docs, int32_t n_queries, int32_t dim, int32_t k,
                         int64_t doc_id_base,
                         float* out_scores, int64_t* out_ids, int32_t* out_counts,
                         void* workspace, size_t workspace_bytes, void* stream);

/* fp32 -> bf16 (round-to-nearest-even) row conversion used when DenseFlatIndexer.index_data
 * (indexer.py:198-208) ingests the fp32 .npy shards written by store_embs (:56-88). */
int b200ret_f32_to_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (3b) Sharded search with a tau exchange between the rounds.
 * A doc-range sharded search runs the rounds of (2)/(3) on every shard.  Alone, a shard can only raise its bound tau[q] to
 * ITS k-th best score, so every shard emits and selects as many candidates per round as a whole corpus would.  With the
 * exchange, after every round each shard also publishes aux[q] = a score that at least ceil(k / n_shards) of its candidates so
 * far reach (a by-product of the round's radix selection, two 8-bit histogram levels deep; -inf if it has fewer
 * candidates); `hook` — provided by the host side, which owns the communicator — all-reduces aux with MIN over the
 * shards on the search's stream, and tau[q] is raised to just below that minimum: every shard holds at least
 * ceil(k / n_shards) documents at or above it, so at least k documents of the corpus do, and a document scoring strictly
 * below it cannot be in the global top-k.  Results are exactly those of the plain entry points after the merge.
 * `hook(user)` must ENQUEUE the all-reduce (no host synchronisation) and return 0; it is called exactly `n_exchanges` times
 * per search on every shard (b200ret_*_exchange_rounds of the LARGEST shard), whatever the shard's own size.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t aux_rank;            /* ceil(k / n_shards) */
    int32_t n_exchanges;         /* hook calls per search, the same on every shard */
    int32_t growth;              /* docs scored grow `growth` x per round (b200ret_exchange_growth(n_shards)) */
    int32_t reserved;
    float* aux;                  /* device [n_queries] */
    int (*hook)(void* user);
    void* user;
} b200ret_round_exchange;

/* Because the exchanged bound follows the documents ALL shards have seen, a shard may take larger steps without emitting more
 * candidates per round than an unsharded search does: the docs scored grow max(4, n_shards + 1) times per round instead of
 * 4 times (about k * n_shards survivors per round over the whole corpus, k per shard, which leaves head-room for the slack
 * of the bound).  Fewer rounds = fewer select launches, kernel tails and all-reduces per shard.  A list that overflows anyway
 * (shards that are not exchangeable) is re-run with the shard's own bounds and the plain schedule, then the safe one. */
int32_t b200ret_exchange_growth(int32_t n_shards);
int32_t b200ret_sparse_exchange_rounds(int32_t n_docs_largest_shard, int32_t n_shards);
int32_t b200ret_dense_exchange_rounds(int32_t n_docs_largest_shard, int32_t n_shards);

int b200ret_sparse_search_sharded(const uint32_t* table, const void* postings,
                                  int32_t n_terms, int32_t n_docs, int32_t block_docs,
                                  const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights,
                                  int32_t n_queries, int32_t k, float threshold, int64_t doc_id_base,
                                  float* out_scores, int64_t* out_ids, int32_t* out_counts,
                                  void* workspace, size_t workspace_bytes, void* stream,
                                  const b200ret_round_exchange* exchange);

int b200ret_dense_search_sharded(const void* corpus_bf16, const void* queries_bf16,
                                 int32_t n_docs, int32_t n_queries, int32_t dim, int32_t k, int64_t doc_id_base,
                                 float* out_scores, int64_t* out_ids, int32_t* out_counts,
                                 void* workspace, size_t workspace_bytes, void* stream,
                                 const b200ret_round_exchange* exchange);

/* ------------------------------------------------------------------------------------------------
 * (4) Shard merge: G per-shard top-k lists (as produced above, gathered with an NCCL all-gather)
 * -> one global top-k per query under the same total order (score desc, doc id asc).
 * in_scores/in_ids: [G, n_queries, k] contiguous; shard g's ids are all smaller than shard g+1's (doc-range
 * shards in rank order), ids are full 64-bit.  G*k is bounded by shared memory (b200ret_merge_max_shards).
 * No reference counterpart: the reference asserts
 * world_size == 1 for retrieval (eval_sparse.py:114, eval_dense.py:191).
 * ---------------------------------------------------------------------------------------------- */
int b200ret_merge_topk(const float* in_scores, const int64_t* in_ids, int32_t n_shards,
                       int32_t n_queries, int32_t k,
                       float* out_scores, int64_t* out_ids, int32_t* out_counts, void* stream);

/* Packed-key form of the same merge — what the sharded search actually exchanges over NVLink (8 bytes per
 * candidate instead of 12, one collective instead of two).  A key is
 *   (order-preserving fp32 score bits << 32) | ~(uint32 GLOBAL doc id),   0 = padding,
 * so that integer order == (score desc, doc id asc); global doc ids must be < 2^32 - 1.
 *   b200ret_pack_keys:   rows (scores, ids; id -1 = padding) -> keys            [n elements]
 *   b200ret_merge_keys:  in_keys [n_shards, n_queries, k] (rows sorted descending, zero padded) ->
 *                        out_keys [n_queries, k], the k largest per query, sorted descending, zero padded;
 *                        n_shards <= b200ret_merge_max_shards(k) per call (shared-memory bound; merge in passes)
 *   b200ret_unpack_keys: keys [n_queries, k] -> rows as b200ret_sparse_search writes them + live counts.  */
int b200ret_pack_keys(const float* scores, const int64_t* ids, int64_t n, uint64_t* keys, void* stream);
int32_t b200ret_merge_max_shards(int32_t k);
int b200ret_merge_keys(const uint64_t* in_keys, int32_t n_shards, int32_t n_queries, int32_t k,
                       uint64_t* out_keys, void* stream);
int b200ret_unpack_keys(const uint64_t* keys, int32_t n_queries, int32_t k,
                        float* out_scores, int64_t* out_ids, int32_t* out_counts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (4b) Gather-sum scoring over semantic-id codes + exact top-k.
 * Replaces TermEncoderRetriever.get_doc_scores + torch.topk (scaling_retriever/indexer.py:621-641, :688):
 *   doc_scores[b, n] = sum_l pred[b, codes[n, l]]   (fp32, summed in code order);   top-k per query, sorted descending.
 * pred: fp32 [n_queries, n_vocab] (the encoder's lex_encode output); codes: int32 [n_docs, code_len] row-major
 * (the reference's doc_encodings LongTensor, narrowed), code_len a multiple of 4 (the reference asserts 16/32/64/128),
 * 16-byte aligned, every code in [0, n_vocab) (callers check).  b200ret_term_scores writes the full [n_queries, n_docs]
 * matrix (get_doc_scores' return value); b200ret_term_search never materialises it.  Output rows as in
 * b200ret_sparse_search (ties: lowest doc row first).
 * ---------------------------------------------------------------------------------------------- */
size_t b200ret_term_search_workspace_bytes(int32_t n_queries, int32_t k);
int b200ret_term_search(const float* pred, const int32_t* codes, int32_t n_queries, int32_t n_vocab, int32_t n_docs,
                        int32_t code_len, int32_t k, float* out_scores, int64_t* out_ids, int32_t* out_counts,
                        void* workspace, size_t workspace_bytes, void* stream);
int b200ret_term_scores(const float* pred, const int32_t* codes, int32_t n_queries, int32_t n_vocab, int32_t n_docs,
                        int32_t code_len, float* out_scores, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (4c) Rank metrics over the result rows (quick regression checks without the run.json -> pytrec_eval round trip of
 * scaling_retriever/utils/metrics.py:22-42): per query, reciprocal rank of the first relevant row among the first
 * `mrr_cut` rows (mrr_k's truncate_run + trec_eval recip_rank) and recall at up to 8 cut-offs (trec_eval recall_<c>).
 * ids/counts: search output rows [n_queries, k] (sorted by score desc; counts NULL = k live rows everywhere);
 * rel_offsets int64 [n_queries + 1] / rel_ids int64 (ascending per query): the relevant row labels of each query.
 * out_rr fp32 [n_queries]; out_recall fp32 [n_queries, n_cuts] (0 for queries without relevant docs).
 * ---------------------------------------------------------------------------------------------- */
int b200ret_rank_metrics(const int64_t* ids, const int32_t* counts, int32_t n_queries, int32_t k,
                         const int64_t* rel_offsets, const int64_t* rel_ids, int32_t mrr_cut,
                         const int32_t* recall_cuts_host, int32_t n_cuts, float* out_rr, float* out_recall, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (5) Result materialisation (HOST pointers only, no device work): write run.json
 * {qid: {external doc id: score}} straight from the [n_queries, row_stride] result arrays.
 * Replaces `res[str(qid)][str(doc_ids[id_])] = float(sc)` (scaling_retriever/indexer.py:429-430; the same
 * loop in eval_dense.py:229-241) followed by json.dump(res) (indexer.py:537-538): byte-identical output for
 * distinct query ids and distinct external ids inside a row (callers check), formatted in parallel.
 *   ids/scores: row q holds counts[q] live entries (counts NULL = row_stride everywhere); a query with
 *   count 0 gets NO key (the reference's defaultdict never sees it).
 *   qid_blob/qid_offsets[n_queries+1]: UTF-8 text of str(qid) per query, every entry followed by one NUL byte
 *   (entry q = qid_blob[qid_offsets[q] .. qid_offsets[q+1] - 1)).
 *   external ids of row label r: docid_blob[docid_offsets[r] .. docid_offsets[r+1] - 1) (same layout), or
 *   str(docid_ints[r]) when the integer table is given instead, or str(r) when both are NULL.
 *   Negative labels index from the end of the table like the reference's Python list (indexer.py:212).
 * ---------------------------------------------------------------------------------------------- */
int b200ret_write_run_json(const char* path_host, const int64_t* ids_host, const float* scores_host,
                           const int32_t* counts_host, int32_t n_queries, int32_t row_stride,
                           const char* qid_blob_host, const int64_t* qid_offsets_host,
                           const char* docid_blob_host, const int64_t* docid_offsets_host,
                           const int64_t* docid_ints_host, int64_t n_doc_ids, int64_t* bytes_written_host);

/* Host helper: multi-threaded memcpy of result rows out of the reusable pinned staging buffers into caller-owned arrays
 * (n_threads <= 0: 8). */
int b200ret_host_copy(void* dst_host, const void* src_host, size_t bytes, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* B200RET_H */
