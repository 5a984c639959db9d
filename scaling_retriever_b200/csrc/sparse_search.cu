// Batched term-at-a-time sparse scoring + exact top-k on sm_100a.
//
// Replaces SparseRetrieval.numba_score_float + select_topk (reference scaling_retriever/indexer.py:324-344,
// :315-322) as driven once per query by _sparse_retrieve_multithreaded (:405-474).
//
// Design:
//   * The reference keeps one fp32 score per document (35 MB/query at 8.8 M docs) in DRAM and does a random
//     read-modify-write per posting.  Here a WARP owns one (query, doc-block) work item: the scores of the
//     block's BD documents live in a warp-private shared-memory tile, the warp streams the slice of each query
//     term's posting list that falls into the block (located with the doc-block skip table), and adds
//     `q_w * d_w` with a separate fp32 multiply and add, term after term in query order.  Doc ids are unique
//     inside one posting list, so lanes never collide inside a term, and terms are serialised by __syncwarp()
//     -> no atomics, and every score is bit-identical to the reference's sequential sum.
//   * Work items are handed out block-major (all queries of doc block b before block b+1) through a global
//     counter, so at any time the whole GPU touches the postings of one or two doc blocks: the index is read
//     from HBM once per query batch and re-served from L2 for the other queries.
//   * The streaming loop walks a warp-uniform cursor over 256-byte aligned rows of 32 postings (one posting per
//     lane, one predicated coalesced 64-bit load of {doc id, weight}), STEP_ROWS rows of one slice per step; the
//     slice descriptors (begin, end, query weight) of a group of 32 query terms sit in a small warp-private
//     shared array.  Steps are loaded in two register batches, double-buffered: the loads of one batch are in flight
//     while the other is accumulated (all posting loads of a warp share one hardware scoreboard, so the order
//     wait -> issue -> accumulate is pinned with a data dependency; see the main loop).  The load pipeline NEVER
//     drains: the fetch side runs ahead of the accumulate side across term groups AND across work items.  Its inputs
//     come from a software pipeline of their own, advanced once per term group: item claim (atomic) -> q_offsets of
//     that item -> term ids / weights of a group -> skip-table entries of the group -> descriptors in shared memory,
//     each stage an asynchronous copy issued one group before its result is needed.
//   * Sweep (end of an item): one pass over the tile that zeroes everything but the hits (score > tau[q]) and records
//     them in per-lane bit masks, one atomic that reserves the item's slots in the query's candidate list, then
//     every lane emits its hits (candidates.cuh: rounds of growing size, radix-select cut to k between rounds,
//     overflow -> safe re-run).
#include <cstdlib>
#include <type_traits>

#include "candidates.cuh"

namespace b200ret {

#ifndef B200RET_SCORE_WARPS     // kernel shape (tuning knobs): warps per CTA, docs per warp tile, blocks in round 0
#define B200RET_SCORE_WARPS 15
#define B200RET_BLOCK_DOCS 3584
#define B200RET_ROUND0_BLOCKS 2
#endif
constexpr int ROUND0_BLOCKS = B200RET_ROUND0_BLOCKS;   // first round / safe-schedule round size, in doc blocks
#ifndef B200RET_LDNC           // posting-load flavour (tuning knob)
#define B200RET_LDNC "ld.global.nc.L2::256B"
#endif
#ifndef B200RET_BATCH_STEPS     // steps per register batch (two batches, double-buffered); 2 and 3 measure equal
#define B200RET_BATCH_STEPS 2
#endif
constexpr int STEP_ROWS = 4;                     // rows (of 32 postings) fetched per pipeline step
// Slice-descriptor layout in the warp's control area (tuning knob): how the cursor reads (begin, end, query weight) of the
// next slice — 0: three arrays, three broadcast LDS.32 (3 data-pipe wavefronts per slice); 1: {begin, end} pairs + weight
// array, LDS.64 + LDS.32 (2 wavefronts); 2: {begin, end, weight, -} records, one LDS.128 (1 wavefront; 128 B more per buffer)
#ifndef B200RET_DESC_MODE
#define B200RET_DESC_MODE 1
#endif
constexpr int DESC_MODE = B200RET_DESC_MODE;
#if defined(B200RET_ADV_MODE) && B200RET_ADV_MODE == 2 && B200RET_DESC_MODE != 1
#error "B200RET_ADV_MODE 2 forms its addresses for the DESC_MODE 1 layout"
#endif
// Row-liveness predicates of a step (tuning knob): 0 = four unsigned compares of rel + 32 r against len (3 adds + 4 setp),
// 1 = one subtraction + compares against immediates (1 sub + 4 setp)
#ifndef B200RET_PRED_MODE
#define B200RET_PRED_MODE 1
#endif
// Tile sweep (tuning knob): 1 = blocks that lie entirely inside the collection skip the per-document bound checks;
// 2 = additionally clear unconditionally and branch only for lanes with a hit
#ifndef B200RET_ADV_MODE        // cursor advance (tuning knob): 1 = fused predicate compares, bit-reversed pending mask;
#define B200RET_ADV_MODE 1      // 2 = additionally addresses formed from the highest-set-bit index (needs DESC_MODE 1)
#endif
#ifndef B200RET_ADDR32          // 1 = posting row address from a 32-bit position (one IMAD.WIDE) instead of 64-bit pointer arithmetic
#define B200RET_ADDR32 0
#endif
#ifndef B200RET_SWEEP_FAST
#define B200RET_SWEEP_FAST 2
#endif
constexpr int DESC_BUF_WORDS = DESC_MODE >= 2 ? 128 : 96;     // one descriptor buffer (two are used alternately)
// mode 3 = mode 2 records, and the staged term ids / query weights of the group after the next live in the unused 4th words
// of the records of buffer 0 / buffer 1 instead of an area of their own (same 1088 bytes per warp as modes 0 and 1)

// Kernel shape: one CTA of WARPS warps per SM, BD docs per warp-private score tile (BD * 4 bytes of shared memory).
constexpr int SCORE_WARPS = B200RET_SCORE_WARPS;
constexpr int BLOCK_DOCS = B200RET_BLOCK_DOCS;   // default: 15 warps x (14 KB scores + 1088 B control area) = 226 KB of 227 KB
constexpr int SCORE_THREADS = SCORE_WARPS * 32;
// Per-warp control area in shared memory (uint32 words): two descriptor buffers (beg[32], end[32], query weight[32]) used
// alternately by consecutive term groups, the staged term ids / weights of the group after the next, and the cold,
// warp-uniform state of the fetch side's input pipeline (every lane writes the same value and reads back its own write).
struct WarpStash {
    int k_q, k_blk, a_q, a_blk, b_q, b_blk, b_g, b_qe, n_q, n_blk;   // identities of the pipeline stages (see the kernel)
    int n_qb, n_qe;          // q_offsets of item n (cp.async destinations)
    unsigned flags;
    unsigned gen;            // groups installed so far; its parity selects the descriptor buffer
    int pad[2];
};
enum : unsigned { K_VALID = 1, K_LAST = 2, K_MARKED = 4, A_VALID = 8, A_LAST = 16, B_VALID = 32, B_LAST = 64, N_VALID = 128 };
constexpr int CTRL_DESC = 0, CTRL_TERMS = 2 * DESC_BUF_WORDS, CTRL_STASH = CTRL_TERMS + (DESC_MODE == 3 ? 0 : 64);   // word offsets
// word index (from the start of the control area) of lane j's staged term id / query weight
__device__ __forceinline__ constexpr int stage_t(int j) { return DESC_MODE == 3 ? 4 * j + 3 : CTRL_TERMS + j; }
__device__ __forceinline__ constexpr int stage_w(int j) { return DESC_MODE == 3 ? DESC_BUF_WORDS + 4 * j + 3 : CTRL_TERMS + 32 + j; }
constexpr int CTRL_WORDS = CTRL_STASH + static_cast<int>(sizeof(WarpStash) / 4);       // 272 words = 1088 bytes per warp (modes 0, 1)
// word index of (begin, end, weight) of slice j inside a descriptor buffer
__device__ __forceinline__ constexpr int desc_beg(int j) { return DESC_MODE == 0 ? j : (DESC_MODE == 1 ? 2 * j : 4 * j); }
__device__ __forceinline__ constexpr int desc_end(int j) { return DESC_MODE == 0 ? 32 + j : (DESC_MODE == 1 ? 2 * j + 1 : 4 * j + 1); }
__device__ __forceinline__ constexpr int desc_qw(int j) { return DESC_MODE >= 2 ? 4 * j + 2 : 64 + j; }
constexpr size_t SCORE_SMEM = static_cast<size_t>(SCORE_WARPS) * (BLOCK_DOCS * sizeof(float) + CTRL_WORDS * sizeof(uint32_t));
static_assert(BLOCK_DOCS % 128 == 0, "tile sweep uses 128-bit accesses by 32 lanes");
static_assert(SCORE_SMEM <= 227 * 1024, "exceeds the shared memory of one SM");

struct ScoreParams {
    const uint32_t* table;     // [n_terms][table_stride]
    size_t table_stride;       // n_blocks + 1
    const void* postings;      // [nnz] postings at the CSR positions: {doc id, fp32 weight} (b200ret_sparse_layout) or the packed
                               // 32-bit fp16 format (b200ret_sparse_layout_f16)
    int32_t format;            // 0 / 1 (host side only: selects the kernel instantiation)
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
    unsigned* item_counter;
    float* dense_out;          // optional [n_active][n_blocks * BLOCK_DOCS]: dump every score instead of selecting
    size_t dense_stride;
};

// End of a work item: emit the docs of the tile with score > tq, zero the tile for the next item.
// (Clearing the tile with a bulk copy of zeros through the async proxy instead of the 128-bit stores was measured slower:
// 126 ms vs 119 ms per step — the copy's latency is exposed once per item.)
__device__ __forceinline__ void sweep_tile(const ScoreParams& p, float* acc, int q, int doc_base, float tq, unsigned lane) {
    constexpr int BD = BLOCK_DOCS;
    const unsigned FULL = 0xffffffffu;
    if (p.dense_out) {   // verification mode (b200ret_sparse_scores): write the tile out, no selection
        float* dst = p.dense_out + static_cast<size_t>(q) * p.dense_stride + doc_base;
        for (int i = lane * 4; i < BD; i += 128) {
            *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(acc + i);
            *reinterpret_cast<float4*>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        // Two phases, ONE counter update per item (its round trip blocks the warp; with one update per 128-doc chunk the warps
        // spent about as long blocked on them as accumulating when the shard is small or tau still low):
        //   1. read the tile, write it back with everything but the hits zeroed, remember the hits in a per-lane bit mask
        //      (4 bits per 128-doc chunk);
        //   2. reserve all slots of the item at once; every lane then emits (and zeroes) its own hits.
        const int limit = min(BD, p.n_docs - doc_base);
        uint64_t* cq = p.cand + static_cast<size_t>(q) * p.cap;
        constexpr int CHUNKS = BD / 128;
        unsigned mask[(CHUNKS + 7) / 8];
#pragma unroll
        for (int wd = 0; wd < (CHUNKS + 7) / 8; ++wd) mask[wd] = 0u;
        auto scan = [&](auto bounded) {
#pragma unroll
            for (int ch = 0; ch < CHUNKS; ++ch) {
                const int i = ch * 128 + lane * 4;
                float4 v = *reinterpret_cast<const float4*>(acc + i);
                unsigned h;
                if (decltype(bounded)::value) {
                    h = ((v.x > tq) && (i < limit) ? 1u : 0u) | ((v.y > tq) && (i + 1 < limit) ? 2u : 0u) |
                        ((v.z > tq) && (i + 2 < limit) ? 4u : 0u) | ((v.w > tq) && (i + 3 < limit) ? 8u : 0u);
                } else if (B200RET_SWEEP_FAST >= 2) {
                    // Once tau has risen almost no document beats it: clear the four slots unconditionally (the store does not
                    // wait for the load) and only a lane whose maximum beats tau takes the branch that puts its hits back.
                    *reinterpret_cast<float4*>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) > tq) {
                        h = (v.x > tq ? 1u : 0u) | (v.y > tq ? 2u : 0u) | (v.z > tq ? 4u : 0u) | (v.w > tq ? 8u : 0u);
                        if (h & 1u) acc[i] = v.x;
                        if (h & 2u) acc[i + 1] = v.y;
                        if (h & 4u) acc[i + 2] = v.z;
                        if (h & 8u) acc[i + 3] = v.w;
                        mask[ch >> 3] |= h << ((ch & 7) * 4);
                    }
                    continue;
                } else {
                    h = (v.x > tq ? 1u : 0u) | (v.y > tq ? 2u : 0u) | (v.z > tq ? 4u : 0u) | (v.w > tq ? 8u : 0u);
                }
                if (!(h & 1u)) v.x = 0.f;
                if (!(h & 2u)) v.y = 0.f;
                if (!(h & 4u)) v.z = 0.f;
                if (!(h & 8u)) v.w = 0.f;
                *reinterpret_cast<float4*>(acc + i) = v;
                mask[ch >> 3] |= h << ((ch & 7) * 4);
            }
        };
        // only the last block of the collection is partial: every other item skips the per-document bound checks (warp-uniform)
        if (B200RET_SWEEP_FAST && limit == BD)
            scan(std::false_type{});
        else
            scan(std::true_type{});
        int mine = 0;
#pragma unroll
        for (int wd = 0; wd < (CHUNKS + 7) / 8; ++wd) mine += __popc(mask[wd]);
        int incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int u = __shfl_up_sync(FULL, incl, off);
            if (lane >= static_cast<unsigned>(off)) incl += u;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        if (total > 0) {                                     // warp-uniform
            int base = 0;
            if (lane == 0) base = atomicAdd(p.cand_count + q, total);
            int pos = __shfl_sync(FULL, base, 0) + incl - mine;
#pragma unroll
            for (int wd = 0; wd < (CHUNKS + 7) / 8; ++wd) {
                unsigned m = mask[wd];
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1;
                    const int idx = (wd * 8 + (bit >> 2)) * 128 + lane * 4 + (bit & 3);
                    const float val = acc[idx];
                    acc[idx] = 0.f;
                    if (pos < p.cap) cq[pos] = cand_key(val, doc_base + idx);
                    ++pos;
                }
            }
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ---- the input pipeline of the fetch side ---------------------------------------------------------------------------
// A "group" is up to 32 consecutive terms of one query against one doc block (lane j <-> term g + j).
//   k : the group being streamed (descriptor buffer gen & 1)
//   a : the next group            — its skip-table entries are in flight into the other descriptor buffer
//   b : the group after that      — its term ids / query weights are in flight into the staging arrays
//   n : the work item after b's   — its q_offsets are in flight into the stash
//   nn: the item after that       — its claim (atomicAdd on the item counter, lane 0 of the CALLER) is in flight
// Every transfer is an asynchronous global->shared copy (cp.async), so nothing is held in registers and each dependent
// load has a whole group's streaming time to complete.  The stages move one group forward per installed group.
struct FetchStages {
    const ScoreParams& p;
    uint32_t* ctrl;
    WarpStash& st;
    unsigned lane;
    unsigned n_items;
    bool want_claim = false;
    __device__ FetchStages(const ScoreParams& p_, uint32_t* ctrl_, unsigned lane_)
        : p(p_), ctrl(ctrl_), st(*reinterpret_cast<WarpStash*>(ctrl_ + CTRL_STASH)), lane(lane_),
          n_items(static_cast<unsigned>(p_.blk_end - p_.blk_begin) * static_cast<unsigned>(p_.n_active)) {}

    // n <- item `it` (block-major item order); the caller claims the following one
    __device__ __forceinline__ void stage_item(unsigned& flags, unsigned it) {
        flags &= ~N_VALID;
        if (it < n_items) {
            flags |= N_VALID;
            st.n_blk = p.blk_begin + static_cast<int>(it / static_cast<unsigned>(p.n_active));
            const int qi = static_cast<int>(it % static_cast<unsigned>(p.n_active));
            const int q = p.q_list ? p.q_list[qi] : qi;
            st.n_q = q;
            if (lane == 0) {
                cp_async4(&st.n_qb, p.q_offsets + q);
                cp_async4(&st.n_qe, p.q_offsets + q + 1);
            }
            want_claim = true;
        }
    }
    // The stash is shared by the 32 lanes of the warp, which all execute this code with the same values.  Every stage reads
    // what it needs into registers, then __syncwarp(), then writes: no lane can overwrite a field another lane has yet to read.

    // b <- the group after b: issue its term-id and query-weight copies (requires n's q_offsets to have landed)
    __device__ __forceinline__ void stage_terms(unsigned& flags, unsigned nn_item) {
        const bool next_group = (flags & (B_VALID | B_LAST)) == B_VALID;      // same item, next 32 terms
        const bool next_item = !next_group && (flags & N_VALID);
        int g = 0, qe = 0, n_q = 0, n_blk = 0;
        if (next_group) {
            g = st.b_g + 32;
            qe = st.b_qe;
        } else if (next_item) {
            n_q = st.n_q;
            n_blk = st.n_blk;
            g = st.n_qb;
            qe = st.n_qe;
        }
        __syncwarp();
        if (next_item) {
            st.b_q = n_q;
            st.b_blk = n_blk;
            st.b_qe = qe;
            flags |= B_VALID;
            stage_item(flags, nn_item);       // re-targets n_q / n_blk / n_qb / n_qe
        } else if (!next_group) {
            flags &= ~B_VALID;
        }
        st.b_g = g;
        flags &= ~B_LAST;
        int32_t* tb_t = reinterpret_cast<int32_t*>(ctrl + stage_t(lane));      // lane-private staging slots
        float* tb_w = reinterpret_cast<float*>(ctrl + stage_w(lane));
        if ((flags & B_VALID) && g + 32 >= qe) flags |= B_LAST;
        if ((flags & B_VALID) && g + static_cast<int>(lane) < qe) {
            cp_async4(tb_t, p.q_terms + g + lane);
            cp_async4(tb_w, p.q_weights + g + lane);
        } else {
            *tb_t = -1;
            *tb_w = 0.f;
        }
    }
    // a <- b: issue the two skip-table copies of every term of the group into descriptor buffer `buf` (requires b's term
    // ids to have landed)
    __device__ __forceinline__ void stage_table(unsigned& flags, uint32_t* buf) {
        const int blk = st.b_blk, bq = st.b_q;
        const int t = static_cast<int32_t>(ctrl[stage_t(lane)]);
        const uint32_t qw_bits = ctrl[stage_w(lane)];
        __syncwarp();
        st.a_q = bq;
        st.a_blk = blk;
        flags = (flags & ~(A_VALID | A_LAST)) | ((flags & B_VALID) ? A_VALID : 0u) | ((flags & B_LAST) ? A_LAST : 0u);
        buf[desc_qw(lane)] = qw_bits;     // query weight
        if ((flags & A_VALID) && t >= 0) {
            const uint32_t* e = p.table + static_cast<size_t>(t) * p.table_stride + blk;
            cp_async4(buf + desc_beg(lane), e);
            cp_async4(buf + desc_end(lane), e + 1);
        } else {
            buf[desc_beg(lane)] = 0;
            buf[desc_end(lane)] = 0;
        }
    }
};

// Warm-up of the pipeline (blocking): claims the first items and runs the stages until group a is loading.
// Returns the index of the latest claimed item (nn).
__device__ __forceinline__ unsigned prime_pipeline(const ScoreParams& p, uint32_t* ctrl, unsigned lane) {
    FetchStages fs(p, ctrl, lane);
    const unsigned FULL = 0xffffffffu;
    unsigned flags = 0, nn = 0;
    fs.st.gen = 0;
    auto claim = [&]() {
        if (lane == 0) nn = atomicAdd(p.item_counter, 1u);
        nn = __shfl_sync(FULL, nn, 0);
    };
    claim();
    fs.stage_item(flags, nn);                 // n <- item 0
    claim();
    cp_async_wait_all();
    __syncwarp();
    fs.stage_terms(flags, nn);                // b <- group 0 of item 0; n <- item 1
    if (fs.want_claim) claim();
    cp_async_wait_all();
    __syncwarp();
    fs.stage_table(flags, ctrl + CTRL_DESC);  // a <- b (descriptor buffer 0 = buffer of gen 0)
    fs.want_claim = false;
    fs.stage_terms(flags, nn);                // b <- the second group
    if (fs.want_claim) claim();
    fs.st.flags = flags;
    __syncwarp();
    return nn;
}

// FMT 0: 8-byte postings {int32 doc id, fp32 weight} — the parity format (bit-identical to numba_score_float).
// FMT 1: 4-byte postings fp16(weight) << 16 | block-local doc id — the opt-in compressed format: same arithmetic on the
//        fp16-rounded weights (bit-identical to the oracle run on those), half the global-load wavefronts per row.
template <int FMT>
__global__ void __launch_bounds__(SCORE_THREADS, 1) sparse_score_kernel(const __grid_constant__ ScoreParams p) {
    constexpr int PBYTES = FMT ? 4 : 8;
    constexpr int BD = BLOCK_DOCS, R = STEP_ROWS;
    extern __shared__ __align__(16) float smem_acc[];
    const unsigned lane = lane_id();
    const unsigned warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    float* const acc = smem_acc + warp * BD;
    uint32_t* const ctrl = reinterpret_cast<uint32_t*>(smem_acc + SCORE_WARPS * BD) + warp * CTRL_WORDS;
    const WarpStash& st = *reinterpret_cast<const WarpStash*>(ctrl + CTRL_STASH);
    // descriptor buffer of the group being streamed: starts at buffer 1 so that the first install flips it to buffer 0
    const uint32_t desc_s01 = 2u * static_cast<uint32_t>(__cvta_generic_to_shared(ctrl + CTRL_DESC)) + DESC_BUF_WORDS * 4u;
    uint32_t desc_s = static_cast<uint32_t>(__cvta_generic_to_shared(ctrl + CTRL_DESC)) + DESC_BUF_WORDS * 4u;
    const uint32_t acc_s = static_cast<uint32_t>(__cvta_generic_to_shared(acc));
    const char* __restrict__ const g_post = static_cast<const char*>(p.postings) + lane * PBYTES;     // lane-private base

    for (int i = lane * 4; i < BD; i += 128) *reinterpret_cast<float4*>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    // the fetch side's input pipeline (FetchStages above): primed here, advanced once per term group in the main loop
    unsigned nn_item = prime_pipeline(p, ctrl, lane);

    // ---- warp-uniform cursor over 256-byte aligned rows of 32 postings; a step covers up to R rows of ONE slice ----
    unsigned c_row = 0, c_beg = 0, c_end = 0, pending = 0;
    float c_qw = 0.f;
    // A step's liveness is (rel + 32*row < len) per lane: rel = position of the lane in row 0 relative to the slice begin
    // (wraps to a huge value before the slice), len = slice length.  The loads leave dead lanes' registers unwritten (no
    // initialisation moves); consume() re-derives the same predicates.  Everything is predicated per lane: uniform
    // branches around the rows past a short slice were measured slower than the dead L1 data-pipe slots they save.
    // advance(): branch-free cursor step; every state change is predicated on "slice exhausted and another one pending".
    // len == 0 on return: the group is exhausted (refill below).
    auto advance = [&](float& qw, unsigned& rel, unsigned& len, const char*& row0) {
        asm volatile(
            "{\n\t"
            ".reg .pred adv, have, take, dead;\n\t"
            ".reg .u32 j, t, a;\n\t"
            "setp.ge.u32 adv, %0, %2;\n\t"                  // c_row >= c_end : current slice exhausted
#if B200RET_ADV_MODE == 0
            "setp.ne.u32 have, %4, 0;\n\t"
            "and.pred take, adv, have;\n\t"
            "not.pred have, have;\n\t"
            "and.pred dead, adv, have;\n\t"                 // exhausted and nothing pending: empty step
            "brev.b32 t, %4;\n\t"
            "clz.b32 j, t;\n\t"                             // index of the lowest pending slice
            "add.u32 t, %4, -1;\n\t"
            "@take and.b32 %4, %4, t;\n\t"
#elif B200RET_ADV_MODE == 1   // `pending` is kept BIT-REVERSED (slice j <-> bit 31 - j): the next slice is the count of leading zeros
            "setp.ne.and.u32 take, %4, 0, adv;\n\t"
            "setp.eq.and.u32 dead, %4, 0, adv;\n\t"         // exhausted and nothing pending: empty step
            "clz.b32 j, %4;\n\t"                            // index of the lowest pending slice (32 when none: loads predicated off)
            "shr.u32 t, 0x80000000, j;\n\t"
            "@take xor.b32 %4, %4, t;\n\t"
#else       // mode 2: bit-reversed `pending`, next slice j = 31 - (highest set bit f); addresses are formed from f directly
            "setp.ne.and.u32 take, %4, 0, adv;\n\t"
            "setp.eq.and.u32 dead, %4, 0, adv;\n\t"
            "bfind.u32 j, %4;\n\t"                          // f (0xffffffff when none: loads predicated off, shift gives 0)
            "shl.b32 t, 1, j;\n\t"
            "@take xor.b32 %4, %4, t;\n\t"
#endif
#if B200RET_DESC_MODE == 0
            "shl.b32 a, j, 2;\n\t"
            "add.u32 a, a, %6;\n\t"
            "@take ld.shared.u32 %1, [a];\n\t"              // c_beg
            "@take ld.shared.u32 %2, [a + 128];\n\t"        // c_end
            "@take ld.shared.f32 %3, [a + 256];\n\t"        // c_qw
#elif B200RET_DESC_MODE == 1 && B200RET_ADV_MODE == 2
            "mad.lo.s32 a, j, -8, %6;\n\t"                  // desc + 8 * (31 - f)
            "@take ld.shared.v2.u32 {%1, %2}, [a + 248];\n\t"   // {c_beg, c_end}
            "mad.lo.s32 a, j, -4, %6;\n\t"
            "@take ld.shared.f32 %3, [a + 380];\n\t"        // c_qw at desc + 256 + 4 * (31 - f)
#elif B200RET_DESC_MODE == 1
            "shl.b32 a, j, 3;\n\t"
            "add.u32 a, a, %6;\n\t"
            "@take ld.shared.v2.u32 {%1, %2}, [a];\n\t"     // {c_beg, c_end}
            "shl.b32 a, j, 2;\n\t"
            "add.u32 a, a, %6;\n\t"
            "@take ld.shared.f32 %3, [a + 256];\n\t"        // c_qw
#else       // modes 2 and 3
            "shl.b32 a, j, 4;\n\t"
            "add.u32 a, a, %6;\n\t"
            "@take ld.shared.v4.b32 {%1, %2, %3, j}, [a];\n\t"  // {c_beg, c_end, c_qw, -}
#endif
            "@take and.b32 %0, %1, 0xffffffe0;\n\t"         // c_row = c_beg rounded down to a row
            "sub.u32 %5, %2, %1;\n\t"
            "@dead mov.u32 %5, 0;\n\t"                      // len
            "}\n"
            : "+r"(c_row), "+r"(c_beg), "+r"(c_end), "+f"(c_qw), "+r"(pending), "=r"(len)
            : "r"(desc_s)
            : "memory");
        qw = c_qw;
#if B200RET_ADDR32
        const unsigned pos = c_row + lane;          // < nnz < 2^32 (csr_build rejects larger indexes): 32-bit sum, one widening multiply-add
        rel = pos - c_beg;
        row0 = static_cast<const char*>(p.postings) + static_cast<size_t>(pos) * PBYTES;
#else
        rel = c_row + lane - c_beg;
        row0 = g_post + static_cast<size_t>(c_row) * PBYTES;     // one 64-bit address per step; rows are 32 postings apart (immediates)
#endif
        c_row += 32u * R;
    };
    // Fetch one step into ring registers.  The cold path (once per term group) installs the next group's descriptors,
    // moves the input pipeline one stage forward, or emits a control record (end of item / no work left).
    // `dep` is a doc id loaded by the batch that is about to be accumulated (always < 2^31): the loads are made to depend on
    // it (dep >> 31 == 0 is added to the lane position), so they cannot be issued before that batch's loads have landed —
    // see the main loop for why this order matters.
    auto fetch = [&](int (&id)[R], float (&w)[R], float& qw, unsigned& rel, unsigned& len, int dep) {
        const char* row0;
        advance(qw, rel, len, row0);
#if B200RET_PRED_MODE == 0
#define B200RET_FETCH_PRED                  \
            "add.u32 t1, t0, 32;\n\t"       \
            "add.u32 t2, t0, 64;\n\t"       \
            "add.u32 t3, t0, 96;\n\t"       \
            "setp.lt.u32 p0, t0, %9;\n\t"   \
            "setp.lt.u32 p1, t1, %9;\n\t"   \
            "setp.lt.u32 p2, t2, %9;\n\t"   \
            "setp.lt.u32 p3, t3, %9;\n\t"
#else       // u = len - rel as a SIGNED count of postings from this lane's row-0 position to the slice end: row r >= 1 is live iff
            // u > 32 r (lanes before the slice begin have rel = -x, u = len + x; lanes past the end have u <= 0); row 0 also
            // needs rel >= 0, i.e. the unsigned rel < len
#define B200RET_FETCH_PRED                  \
            "sub.s32 t1, %9, t0;\n\t"       \
            "setp.lt.u32 p0, t0, %9;\n\t"   \
            "setp.gt.s32 p1, t1, 32;\n\t"   \
            "setp.gt.s32 p2, t1, 64;\n\t"   \
            "setp.gt.s32 p3, t1, 96;\n\t"
#endif
        if (FMT == 0) {
            asm volatile(
                "{\n\t"
                ".reg .pred p0, p1, p2, p3;\n\t"
                ".reg .u32 t0, t1, t2, t3;\n\t"
                "shr.u32 t0, %11, 31;\n\t"
                "add.u32 t0, t0, %8;\n\t"
                B200RET_FETCH_PRED
                "@p0 " B200RET_LDNC ".v2.b32 {%0, %4}, [%10];\n\t"          // one posting = {doc id, weight bits}
                "@p1 " B200RET_LDNC ".v2.b32 {%1, %5}, [%10 + 256];\n\t"
                "@p2 " B200RET_LDNC ".v2.b32 {%2, %6}, [%10 + 512];\n\t"
                "@p3 " B200RET_LDNC ".v2.b32 {%3, %7}, [%10 + 768];\n\t"
                "}\n"
                : "+r"(id[0]), "+r"(id[1]), "+r"(id[2]), "+r"(id[3]), "+f"(w[0]), "+f"(w[1]), "+f"(w[2]), "+f"(w[3])
                : "r"(rel), "r"(len), "l"(row0), "r"(dep));
        } else {     // packed postings: one 32-bit word each; bit 15 of a word (block-local doc ids are < 2^15) is the always-0 dependency
            asm volatile(
                "{\n\t"
                ".reg .pred p0, p1, p2, p3;\n\t"
                ".reg .u32 t0, t1, t2, t3;\n\t"
                "bfe.u32 t0, %11, 15, 1;\n\t"
                "add.u32 t0, t0, %8;\n\t"
                B200RET_FETCH_PRED
                "@p0 " B200RET_LDNC ".b32 %0, [%10];\n\t"
                "@p1 " B200RET_LDNC ".b32 %1, [%10 + 128];\n\t"
                "@p2 " B200RET_LDNC ".b32 %2, [%10 + 256];\n\t"
                "@p3 " B200RET_LDNC ".b32 %3, [%10 + 384];\n\t"
                "}\n"
                : "+r"(id[0]), "+r"(id[1]), "+r"(id[2]), "+r"(id[3]), "+f"(w[0]), "+f"(w[1]), "+f"(w[2]), "+f"(w[3])
                : "r"(rel), "r"(len), "l"(row0), "r"(dep));
        }
    };
    // Accumulate one step.  Its rows belong to ONE posting list, so their doc ids are distinct and the R
    // read-modify-writes are independent: loads, adds and stores are issued R-wide (one latency per step).
    // reference arithmetic: scores[doc] += q * w  -> fp32 multiply, then fp32 add (no FMA).
    // acc_rel_s: shared-memory byte address such that acc_rel_s + 4 * doc_id is the doc's score slot.
    // In two parts: the score-slot addresses (first use of the loaded doc ids: the scoreboard wait for the batch's loads
    // happens here), and the read-modify-writes.
    auto consume_pre = [&](const int (&id)[R], float (&w)[R], uint32_t (&d)[R], uint32_t acc_rel_s) {
        if (FMT == 0) {
            asm volatile(
                "mad.lo.u32 %0, %4, 4, %8;\n\t"
                "mad.lo.u32 %1, %5, 4, %8;\n\t"
                "mad.lo.u32 %2, %6, 4, %8;\n\t"
                "mad.lo.u32 %3, %7, 4, %8;\n"
                : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                : "r"(id[0]), "r"(id[1]), "r"(id[2]), "r"(id[3]), "r"(acc_rel_s));
        } else {     // word = fp16 weight << 16 | block-local doc id: slot address from the low half, fp32 weight from the high half
            asm volatile(
                "{\n\t"
                ".reg .b16 l0, l1, l2, l3, h0, h1, h2, h3;\n\t"
                ".reg .u32 u0, u1, u2, u3;\n\t"
                "mov.b32 {l0, h0}, %8;\n\t"
                "mov.b32 {l1, h1}, %9;\n\t"
                "mov.b32 {l2, h2}, %10;\n\t"
                "mov.b32 {l3, h3}, %11;\n\t"
                "cvt.u32.u16 u0, l0;\n\t"
                "cvt.u32.u16 u1, l1;\n\t"
                "cvt.u32.u16 u2, l2;\n\t"
                "cvt.u32.u16 u3, l3;\n\t"
                "mad.lo.u32 %0, u0, 4, %12;\n\t"
                "mad.lo.u32 %1, u1, 4, %12;\n\t"
                "mad.lo.u32 %2, u2, 4, %12;\n\t"
                "mad.lo.u32 %3, u3, 4, %12;\n\t"
                "cvt.f32.f16 %4, h0;\n\t"
                "cvt.f32.f16 %5, h1;\n\t"
                "cvt.f32.f16 %6, h2;\n\t"
                "cvt.f32.f16 %7, h3;\n\t"
                "}\n"
                : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=f"(w[0]), "=f"(w[1]), "=f"(w[2]), "=f"(w[3])
                : "r"(id[0]), "r"(id[1]), "r"(id[2]), "r"(id[3]), "r"(acc_rel_s));
        }
    };
    auto consume_rest = [&](const uint32_t (&d)[R], const float (&w)[R], float qw, unsigned rel, unsigned len) {
        asm volatile(
            "{\n\t"
            ".reg .pred p0, p1, p2, p3;\n\t"
            ".reg .f32 a0, a1, a2, a3, v0, v1, v2, v3;\n\t"
            ".reg .u32 t1, t2, t3;\n\t"
#if B200RET_PRED_MODE == 0
            "add.u32 t1, %9, 32;\n\t"
            "add.u32 t2, %9, 64;\n\t"
            "add.u32 t3, %9, 96;\n\t"
            "setp.lt.u32 p0, %9, %10;\n\t"
            "setp.lt.u32 p1, t1, %10;\n\t"
            "setp.lt.u32 p2, t2, %10;\n\t"
            "setp.lt.u32 p3, t3, %10;\n\t"
#else
            "sub.s32 t1, %10, %9;\n\t"
            "setp.lt.u32 p0, %9, %10;\n\t"
            "setp.gt.s32 p1, t1, 32;\n\t"
            "setp.gt.s32 p2, t1, 64;\n\t"
            "setp.gt.s32 p3, t1, 96;\n\t"
#endif
            "@p0 ld.shared.f32 a0, [%0];\n\t"
            "@p1 ld.shared.f32 a1, [%1];\n\t"
            "@p2 ld.shared.f32 a2, [%2];\n\t"
            "@p3 ld.shared.f32 a3, [%3];\n\t"
            "mul.rn.f32 v0, %8, %4;\n\t"
            "mul.rn.f32 v1, %8, %5;\n\t"
            "mul.rn.f32 v2, %8, %6;\n\t"
            "mul.rn.f32 v3, %8, %7;\n\t"
            "@p0 add.rn.f32 a0, a0, v0;\n\t"
            "@p1 add.rn.f32 a1, a1, v1;\n\t"
            "@p2 add.rn.f32 a2, a2, v2;\n\t"
            "@p3 add.rn.f32 a3, a3, v3;\n\t"
            "@p0 st.shared.f32 [%0], a0;\n\t"
            "@p1 st.shared.f32 [%1], a1;\n\t"
            "@p2 st.shared.f32 [%2], a2;\n\t"
            "@p3 st.shared.f32 [%3], a3;\n\t"
            "}\n" ::"r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "f"(w[0]), "f"(w[1]), "f"(w[2]), "f"(w[3]), "f"(qw), "r"(rel), "r"(len)
            : "memory");
        __syncwarp();
    };
    // Two register batches of HB steps each, double-buffered: all posting loads of a warp share ONE hardware scoreboard
    // (ptxas gives the other five to the shared-memory loads), so "wait for this step's loads" means "wait for every load in
    // flight".  A ring that fetches one step per consumed step therefore exposes the full load latency again and again (ptxas
    // answers by sinking all loads to the end of the revolution: no overlap at all).  Here the order is
    //     wait(A) -> issue loads of B -> accumulate A -> wait(B) -> issue loads of A -> accumulate B -> ...
    // so every wait finds only loads that have been in flight for a whole batch's accumulate time.
    // Control (one copy of the cold code): group install / item close at the top (before the loads of B are issued), the
    // sweep of a closed item between the two halves: the item's last steps are in A, the next item's first steps in B.
    constexpr int HB = B200RET_BATCH_STEPS;
    int id[2 * HB][R];
    float w[2 * HB][R], qw[2 * HB];
    unsigned rel[2 * HB], len[2 * HB];
#pragma unroll
    for (int s = 0; s < 2 * HB; ++s) {
        len[s] = 0u;
        qw[s] = 0.f;
        rel[s] = 0u;
#pragma unroll
        for (int r = 0; r < R; ++r) {     // dead lanes keep old register contents: make sure they are always doc ids (< 2^31)
            id[s][r] = 0;
            w[s][r] = 0.f;
        }
    }
    FetchStages fs(p, ctrl, lane);
    // tile base of the first item (group a); packed postings carry block-local doc ids, so their base never moves
    uint32_t acc_rel_s = FMT ? acc_s : acc_s - static_cast<uint32_t>(st.a_blk * BD) * 4u;
    bool sweep_due = false, fin = false;
    int sw_q = 0, sw_doc_base = 0;
    float sw_tau = 0.f;
    uint32_t sw_next_acc_rel = 0;
    auto half = [&](int c0, int f0) {     // accumulate batch [c0, c0 + HB) while the loads of batch [f0, f0 + HB) are issued
        uint32_t d[HB][R];
#pragma unroll
        for (int s = 0; s < HB; ++s) consume_pre(id[c0 + s], w[c0 + s], d[s], acc_rel_s);
#pragma unroll
        for (int s = 0; s < HB; ++s) fetch(id[f0 + s], w[f0 + s], qw[f0 + s], rel[f0 + s], len[f0 + s], id[c0][0]);
#pragma unroll
        for (int s = 0; s < HB; ++s)
            if (len[c0 + s] != 0u) consume_rest(d[s], w[c0 + s], qw[c0 + s], rel[c0 + s], len[c0 + s]);
    };
    while (!fin) {
        while (pending == 0u && c_row >= c_end) {             // group exhausted: warp-uniform, once per term group
            unsigned flags = st.flags;
            if ((flags & (K_VALID | K_LAST | K_MARKED)) == (K_VALID | K_LAST)) {   // the item is complete (its last steps are in A)
                if (sweep_due) break;                         // (an item without postings right behind: one sweep per iteration)
                sweep_due = true;
                sw_q = st.k_q;
                sw_doc_base = st.k_blk * BD;
                sw_tau = p.tau ? __ldg(p.tau + sw_q) : 0.f;   // arrives while A is accumulated
                sw_next_acc_rel = FMT ? acc_s : acc_s - static_cast<uint32_t>(((flags & A_VALID) ? st.a_blk : st.k_blk) * BD) * 4u;
                flags |= K_MARKED;
                __syncwarp();                                 // stash: reads above, write below
                fs.st.flags = flags;
            }
            if (!(flags & A_VALID)) {
                fin = true;                                   // A is still accumulated (and swept) in this iteration
                break;
            }
            // install group a (its skip-table entries were requested one group ago), then move the stages forward
            cp_async_wait_all();
            __syncwarp();
            const unsigned gen = st.gen;
            const int a_q = st.a_q, a_blk = st.a_blk;
            const uint32_t* buf = ctrl + CTRL_DESC + (gen & 1u) * DESC_BUF_WORDS;
            pending = __ballot_sync(FULL, buf[desc_end(lane)] > buf[desc_beg(lane)]);   // non-empty slices, ascending term order
            if (B200RET_ADV_MODE) pending = __brev(pending);
            desc_s = desc_s01 - desc_s;                       // the cursor reads slice j's descriptor with 3 broadcast LDS.32
            __syncwarp();                                     // stash: reads above, writes below
            fs.st.k_q = a_q;
            fs.st.k_blk = a_blk;
            flags = (flags & ~(K_VALID | K_LAST | K_MARKED)) | K_VALID | ((flags & A_LAST) ? K_LAST : 0u);
            fs.want_claim = false;
            fs.stage_table(flags, ctrl + CTRL_DESC + ((gen + 1u) & 1u) * DESC_BUF_WORDS);
            fs.stage_terms(flags, __shfl_sync(FULL, nn_item, 0));
            fs.st.gen = gen + 1u;
            fs.st.flags = flags;
            __syncwarp();   // the stash is written by every lane with the same values; order them before the next reads
            if (fs.want_claim && lane == 0) nn_item = atomicAdd(p.item_counter, 1u);   // stays in flight
        }
        half(0, HB);
        if (sweep_due) {
            sweep_tile(p, acc, sw_q, sw_doc_base, sw_tau, lane);
            acc_rel_s = sw_next_acc_rel;
            sweep_due = false;
        }
        half(HB, 0);
    }
}

static int block_docs_of_shape() { return BLOCK_DOCS; }

static int launch_score(const ScoreParams& sp, cudaStream_t stream) {
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        B200RET_CUDA_CHECK(cudaFuncSetAttribute(sparse_score_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                static_cast<int>(SCORE_SMEM)));
        B200RET_CUDA_CHECK(cudaFuncSetAttribute(sparse_score_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                static_cast<int>(SCORE_SMEM)));
        attr_set.mark();
    }
    // The item counter is 32 bits wide (claims run past the end by a few per warp): cut the block range so that one
    // launch hands out < 2^31 (query, block) items.
    unsigned max_items = (1u << 31) - (1u << 20);
    if (const char* e = getenv("B200RET_TEST_MAX_ITEMS")) max_items = std::max(1u, static_cast<unsigned>(atoi(e)));   // test hook
    const int max_blocks = std::max(1, static_cast<int>(max_items / static_cast<unsigned>(std::max(sp.n_active, 1))));
    B200RET_REQUIRE(sp.n_active < (1 << 30), "sparse: %d queries in one launch", sp.n_active);
    for (int b0 = sp.blk_begin; b0 < sp.blk_end; b0 += max_blocks) {
        ScoreParams r = sp;
        r.blk_begin = b0;
        r.blk_end = std::min(sp.blk_end, b0 + max_blocks);
        B200RET_CUDA_CHECK(cudaMemsetAsync(r.item_counter, 0, sizeof(unsigned), stream));
        prof_begin(PROF_SPARSE_SCORE, stream);
        if (r.format == 0)
            sparse_score_kernel<0><<<sm_count(), SCORE_THREADS, SCORE_SMEM, stream>>>(r);
        else
            sparse_score_kernel<1><<<sm_count(), SCORE_THREADS, SCORE_SMEM, stream>>>(r);
        prof_end(PROF_SPARSE_SCORE, stream);
        count_launches(1);
        B200RET_CUDA_CHECK(cudaGetLastError());
    }
    return B200RET_OK;
}

// candidate capacity: the k kept keys + what one round may append — the first round's docs, or for large k the ~(ROUND_GROWTH - 1) * k
// survivors a geometric round is expected to produce, with a two-k margin (candidates.cuh)
static int search_cap(int k) { return (k + std::max(ROUND0_BLOCKS * block_docs_of_shape(), (ROUND_GROWTH + 1) * k) + 1) & ~1; }   // even: lists stay 16-byte aligned

}  // namespace b200ret

using namespace b200ret;

extern "C" int32_t b200ret_sparse_block_docs(void) { return block_docs_of_shape(); }

extern "C" size_t b200ret_sparse_search_workspace_bytes(int32_t n_queries, int32_t k) {
    Workspace ws(nullptr, 0);
    return carve_cand(ws, n_queries, search_cap(k), nullptr) + 256;
}

static int fill_params(ScoreParams& sp, const uint32_t* table, const void* postings, int32_t n_docs,
                       int32_t block_docs, const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights,
                       int32_t n_queries) {
    const int bd = block_docs_of_shape();
    B200RET_REQUIRE(block_docs == bd, "sparse: table built for block_docs=%d, kernel uses %d", block_docs, bd);
    const int32_t n_blocks = (n_docs + bd - 1) / bd;
    sp = ScoreParams{};
    sp.table = table;
    sp.table_stride = static_cast<size_t>(n_blocks) + 1;
    sp.postings = postings;
    sp.q_offsets = q_offsets;
    sp.q_terms = q_terms;
    sp.q_weights = q_weights;
    sp.n_active = n_queries;
    sp.n_docs = n_docs;
    return B200RET_OK;
}

static int sparse_scores_impl(int format, const uint32_t* table, const void* postings,
                              int32_t n_terms, int32_t n_docs, int32_t block_docs,
                              const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights,
                              int32_t n_queries, float* out_scores, void* workspace, size_t workspace_bytes,
                              void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    B200RET_REQUIRE(n_queries >= 0 && n_docs >= 0 && n_terms > 0, "sparse_scores: bad sizes");
    ScoreParams sp;
    int rc = fill_params(sp, table, postings, n_docs, block_docs, q_offsets, q_terms, q_weights, n_queries);
    if (rc != B200RET_OK) return rc;
    sp.format = format;
    if (n_queries == 0 || n_docs == 0) return B200RET_OK;
    B200RET_REQUIRE(table && q_offsets && out_scores && workspace && workspace_bytes >= 256, "sparse_scores: null pointer / workspace < 256 B");
    const int32_t n_blocks = static_cast<int32_t>(sp.table_stride) - 1;
    sp.blk_begin = 0;
    sp.blk_end = n_blocks;
    sp.item_counter = static_cast<unsigned*>(workspace);
    sp.dense_out = out_scores;
    sp.dense_stride = static_cast<size_t>(n_blocks) * block_docs;
    return launch_score(sp, stream);
}

extern "C" int b200ret_sparse_scores(const uint32_t* table, const void* postings, int32_t n_terms, int32_t n_docs, int32_t block_docs,
                                     const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights, int32_t n_queries,
                                     float* out_scores, void* workspace, size_t workspace_bytes, void* stream) {
    return sparse_scores_impl(0, table, postings, n_terms, n_docs, block_docs, q_offsets, q_terms, q_weights, n_queries, out_scores,
                              workspace, workspace_bytes, stream);
}

extern "C" int b200ret_sparse_scores_f16(const uint32_t* table, const void* postings, int32_t n_terms, int32_t n_docs,
                                         int32_t block_docs, const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights,
                                         int32_t n_queries, float* out_scores, void* workspace, size_t workspace_bytes, void* stream) {
    return sparse_scores_impl(1, table, postings, n_terms, n_docs, block_docs, q_offsets, q_terms, q_weights, n_queries, out_scores,
                              workspace, workspace_bytes, stream);
}

static int sparse_search_impl(int format, const uint32_t* table, const void* postings,
                              int32_t n_terms, int32_t n_docs, int32_t block_docs,
                              const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights,
                              int32_t n_queries, int32_t k, float threshold, int64_t doc_id_base,
                              float* out_scores, int64_t* out_ids, int32_t* out_counts,
                              void* workspace, size_t workspace_bytes, void* stream_, const b200ret_round_exchange* exchange = nullptr) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ScoreParams sp;
    int rc = fill_params(sp, table, postings, n_docs, block_docs, q_offsets, q_terms, q_weights, n_queries);
    if (rc != B200RET_OK) return rc;
    sp.format = format;
    B200RET_REQUIRE(k >= 1 && k <= B200RET_MAX_K, "sparse_search: k=%d outside [1, %d]", k, B200RET_MAX_K);
    B200RET_REQUIRE(n_queries >= 0 && n_docs >= 0 && n_terms > 0, "sparse_search: bad sizes");
    if (n_queries == 0) return B200RET_OK;
    B200RET_REQUIRE(table && q_offsets && out_scores && out_ids && out_counts && workspace, "sparse_search: null pointer");
    if (workspace_bytes < b200ret_sparse_search_workspace_bytes(n_queries, k)) {
        set_err("sparse_search: workspace too small (%zu bytes)", workspace_bytes);
        return B200RET_EWORKSPACE;
    }
    Workspace ws(workspace, workspace_bytes);
    CandBuffers b;
    const int cap = search_cap(k);
    carve_cand(ws, n_queries, cap, &b);
    sp.tau = b.tau;
    sp.cand = b.cand;
    sp.cand_count = b.cand_count;
    sp.cap = cap;
    sp.item_counter = reinterpret_cast<unsigned*>(b.work_counter);
    const int32_t n_blocks = static_cast<int32_t>(sp.table_stride) - 1;

    auto launch_round = [&](int blk_begin, int blk_end, const int32_t* q_list, int32_t n_active) -> int {
        ScoreParams r = sp;
        r.blk_begin = blk_begin;
        r.blk_end = blk_end;
        r.q_list = q_list;
        r.n_active = n_active;
        return launch_score(r, stream);
    };
    return run_search(launch_round, b, cap, k, n_queries, n_blocks, ROUND0_BLOCKS, threshold, doc_id_base, out_scores, out_ids,
                      out_counts, stream, exchange);
}

extern "C" int b200ret_sparse_search(const uint32_t* table, const void* postings, int32_t n_terms, int32_t n_docs, int32_t block_docs,
                                     const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights, int32_t n_queries,
                                     int32_t k, float threshold, int64_t doc_id_base, float* out_scores, int64_t* out_ids,
                                     int32_t* out_counts, void* workspace, size_t workspace_bytes, void* stream) {
    return sparse_search_impl(0, table, postings, n_terms, n_docs, block_docs, q_offsets, q_terms, q_weights, n_queries, k, threshold,
                              doc_id_base, out_scores, out_ids, out_counts, workspace, workspace_bytes, stream);
}

extern "C" int b200ret_sparse_search_f16(const uint32_t* table, const void* postings, int32_t n_terms, int32_t n_docs,
                                         int32_t block_docs, const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights,
                                         int32_t n_queries, int32_t k, float threshold, int64_t doc_id_base, float* out_scores,
                                         int64_t* out_ids, int32_t* out_counts, void* workspace, size_t workspace_bytes, void* stream) {
    return sparse_search_impl(1, table, postings, n_terms, n_docs, block_docs, q_offsets, q_terms, q_weights, n_queries, k, threshold,
                              doc_id_base, out_scores, out_ids, out_counts, workspace, workspace_bytes, stream);
}

extern "C" int32_t b200ret_exchange_growth(int32_t n_shards) { return exchange_growth(n_shards); }

extern "C" int32_t b200ret_sparse_exchange_rounds(int32_t n_docs_largest_shard, int32_t n_shards) {
    const int bd = block_docs_of_shape();
    return schedule_exchanges((std::max(n_docs_largest_shard, 0) + bd - 1) / bd, ROUND0_BLOCKS, exchange_growth(n_shards));
}

extern "C" int b200ret_sparse_search_sharded(const uint32_t* table, const void* postings, int32_t n_terms, int32_t n_docs,
                                             int32_t block_docs, const int32_t* q_offsets, const int32_t* q_terms,
                                             const float* q_weights, int32_t n_queries, int32_t k, float threshold, int64_t doc_id_base,
                                             float* out_scores, int64_t* out_ids, int32_t* out_counts, void* workspace,
                                             size_t workspace_bytes, void* stream, const b200ret_round_exchange* exchange) {
    return sparse_search_impl(0, table, postings, n_terms, n_docs, block_docs, q_offsets, q_terms, q_weights, n_queries, k, threshold,
                              doc_id_base, out_scores, out_ids, out_counts, workspace, workspace_bytes, stream, exchange);
}
