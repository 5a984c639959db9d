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
//   * The streaming loop walks a warp-uniform cursor over 128-byte aligned rows of 32 postings (one posting per
//     lane, predicated coalesced loads of the id and the weight), STEP_ROWS rows of one slice per step; the slice
//     descriptors (begin, end, query weight) of a term group sit in a small warp-private shared array.  PIPE_DEPTH
//     steps are kept in flight in a register ring (fetch step i+DEPTH-1, then consume step i), across term
//     boundaries, so posting loads overlap the shared-memory read-modify-writes; the accumulate of one step is
//     branch-free predicated PTX that issues its STEP_ROWS loads, adds and stores side by side.
//   * After the last term the warp sweeps its tile once (128-bit shared loads, zeroing as it goes) and appends
//     the documents with score > tau[q] to the query's candidate list (candidates.cuh: rounds of doubling
//     size, radix-select cut to k between rounds, overflow -> safe re-run).
#include "candidates.cuh"

namespace b200ret {

#ifndef B200RET_SCORE_WARPS     // kernel shape (tuning knobs): warps per CTA, docs per warp tile, blocks in round 0
#define B200RET_SCORE_WARPS 16
#define B200RET_BLOCK_DOCS 3456
#define B200RET_ROUND0_BLOCKS 4
#endif
constexpr int ROUND0_BLOCKS = B200RET_ROUND0_BLOCKS;   // first round / safe-schedule round size, in doc blocks
#ifndef B200RET_LDNC           // posting-load flavour (tuning knob)
#define B200RET_LDNC "ld.global.nc"
#endif
#ifndef B200RET_LOOKAHEAD       // prefetch the next term group's skip-table entries (tuning knob)
#define B200RET_LOOKAHEAD 1
#endif
#ifndef B200RET_FUSED_STEP      // issue step f's posting loads inside step s's accumulate block (tuning knob)
#define B200RET_FUSED_STEP 0     // measured equal to the separate blocks (129.8 vs 129.5 ms per 6,980-query step)
#endif
#ifndef B200RET_SHORT_STEP      // one-row code path for steps whose slice ends inside row 0 (tuning knob)
#define B200RET_SHORT_STEP 0     // measured SLOWER (138.7 vs 128.7 ms): the warp-uniform branch costs more than the dead slots
#endif
#ifndef B200RET_PTX_ADVANCE     // branch-free predicated cursor step in PTX instead of the C++ if/else (tuning knob)
#define B200RET_PTX_ADVANCE 1
#endif
#ifndef B200RET_STEP_ROWS
#define B200RET_STEP_ROWS 4
#endif
constexpr int STEP_ROWS = B200RET_STEP_ROWS;     // rows (of 32 postings) fetched per pipeline step (2 or 4)
#ifndef B200RET_PIPE_DEPTH
#define B200RET_PIPE_DEPTH 5     // 6 spills a few registers at 128 regs/thread and measures slower
#endif
constexpr int PIPE_DEPTH = B200RET_PIPE_DEPTH;   // steps in flight per warp (register ring)

// Kernel shape: one CTA of WARPS warps per SM, BD docs per warp-private score tile (BD * 4 bytes of shared memory).
constexpr int SCORE_WARPS = B200RET_SCORE_WARPS;
constexpr int BLOCK_DOCS = B200RET_BLOCK_DOCS;   // default: 16 warps x (13.5 KB scores + 384 B slice descriptors) = 222 KB of 227 KB
constexpr int SCORE_THREADS = SCORE_WARPS * 32;
constexpr size_t SCORE_SMEM = static_cast<size_t>(SCORE_WARPS) * (BLOCK_DOCS * sizeof(float) + 3 * 32 * sizeof(uint32_t));
static_assert(BLOCK_DOCS % 128 == 0, "tile sweep uses 128-bit accesses by 32 lanes");
static_assert(SCORE_SMEM <= 227 * 1024, "exceeds the shared memory of one SM");

struct ScoreParams {
    const uint32_t* table;     // [n_terms][table_stride]
    size_t table_stride;       // n_blocks + 1
    const uint2* postings;     // [nnz] {doc id, weight bits}, CSR positions (b200ret_sparse_layout)
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
    constexpr int BD = BLOCK_DOCS, R = STEP_ROWS;
    extern __shared__ __align__(16) float smem_acc[];
    const unsigned lane = lane_id();
    const unsigned warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    float* const acc = smem_acc + warp * BD;
    // warp-private slice descriptors of the current term group: beg[32], end[32], query weight[32]
    uint32_t* const desc = reinterpret_cast<uint32_t*>(smem_acc + SCORE_WARPS * BD) + warp * 96;
    const uint2* __restrict__ const g_post = p.postings + lane;     // lane-private base: address = base + row position
    const uint32_t* __restrict__ const g_table = p.table;
    const size_t table_stride = p.table_stride;
    const int n_active = p.n_active;

    for (int i = lane * 4; i < BD; i += 128) *reinterpret_cast<float4*>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    const long long n_items = static_cast<long long>(p.blk_end - p.blk_begin) * n_active;
    long long item = 0;
    if (lane == 0) item = static_cast<long long>(atomicAdd(p.item_counter, 1ULL));
    item = __shfl_sync(FULL, item, 0);
    while (item < n_items) {
        // Claim the next item now; its index is only needed after this one is done (hides the atomic's latency).
        long long next_item = 0;
        if (lane == 0) next_item = static_cast<long long>(atomicAdd(p.item_counter, 1ULL));

        const int blk = p.blk_begin + static_cast<int>(item / n_active);
        const int qi = static_cast<int>(item % n_active);
        const int q = p.q_list ? p.q_list[qi] : qi;
        const int doc_base = blk * BD;
        // shared-memory byte address such that acc_rel_s + 4 * doc_id is the doc's score slot
        const uint32_t acc_rel_s = static_cast<uint32_t>(__cvta_generic_to_shared(acc)) - static_cast<uint32_t>(doc_base) * 4u;
        const int qb = p.q_offsets[q], qe = p.q_offsets[q + 1];

        // Lane j holds term g+j of the query: its weight and its posting slice inside this doc block.  The lookup of
        // the NEXT group of 32 terms (query term -> skip table, two dependent loads) is issued before the current
        // group is streamed, so its latency is paid once per item, not once per group.
        auto lookup = [&](int g, unsigned& beg, unsigned& end, float& qw) {
            beg = 0;
            end = 0;
            qw = 0.f;
            if (g + static_cast<int>(lane) < qe) {
                const int t = __ldg(p.q_terms + g + lane);
                qw = __ldg(p.q_weights + g + lane);
                const uint32_t* e = g_table + static_cast<size_t>(t) * table_stride + blk;
                beg = __ldg(e);
                end = __ldg(e + 1);
            }
        };
#if B200RET_LOOKAHEAD
        unsigned nxt_beg, nxt_end;
        float nxt_qw;
        lookup(qb, nxt_beg, nxt_end, nxt_qw);
#endif
        for (int g = qb; g < qe; g += 32) {
#if B200RET_LOOKAHEAD
            const unsigned seg_beg = nxt_beg, seg_end = nxt_end;
            const float seg_qw = nxt_qw;
            if (g + 32 < qe) lookup(g + 32, nxt_beg, nxt_end, nxt_qw);
#else
            unsigned seg_beg, seg_end;
            float seg_qw;
            lookup(g, seg_beg, seg_end, seg_qw);
#endif
            unsigned pending = __ballot_sync(FULL, seg_end > seg_beg);   // non-empty slices, ascending term order
            // Publish the descriptors: the cursor reads slice j with three broadcast LDS.32 (3 wavefronts) instead of
            // three SHFLs (4 wavefronts each on the same, saturated, L1 data pipe).
            desc[lane] = seg_beg;
            desc[32 + lane] = seg_end;
            desc[64 + lane] = __float_as_uint(seg_qw);
            __syncwarp();

            // Warp-uniform cursor over 256-byte aligned rows of 32 postings; a step covers up to R rows of ONE slice.
            unsigned c_row = 0, c_beg = 0, c_end = 0;
            float c_qw = 0.f;
            const uint32_t desc_s = static_cast<uint32_t>(__cvta_generic_to_shared(desc));
            (void)desc_s;
            // Fetch one step into registers (ids = -1 on dead lanes).  Returns false when nothing is left.
            // Everything is predicated per lane: uniform branches around the rows past a short slice were measured
            // slower (register ring spills, lost overlap) than the dead L1 data-pipe slots they save.
            // A step's liveness is (rel + 32*row < len) per lane: rel = position of the lane in row 0 relative to the
            // slice begin (wraps to a huge value before the slice), len = slice length (0 = empty step).  The loads leave
            // dead lanes' registers unwritten (no initialisation moves); consume() re-derives the same predicates.
            // advance(): the warp-uniform cursor logic of one step (no memory traffic except 3 LDS at a slice change)
#if B200RET_PTX_ADVANCE
            // branch-free cursor step: every state change is predicated on "slice exhausted and another one pending"
            auto advance = [&](float& qw, unsigned& rel, unsigned& len, const uint2*& row0) -> bool {
                asm volatile(
                    "{\n\t"
                    ".reg .pred adv, have, take, dead;\n\t"
                    ".reg .u32 j, t, a;\n\t"
                    "setp.ge.u32 adv, %0, %2;\n\t"                  // c_row >= c_end : current slice exhausted
                    "setp.ne.u32 have, %4, 0;\n\t"
                    "and.pred take, adv, have;\n\t"
                    "not.pred have, have;\n\t"
                    "and.pred dead, adv, have;\n\t"                 // exhausted and nothing pending: empty step
                    "brev.b32 t, %4;\n\t"
                    "clz.b32 j, t;\n\t"                             // index of the lowest pending slice
                    "add.u32 t, %4, -1;\n\t"
                    "@take and.b32 %4, %4, t;\n\t"
                    "shl.b32 a, j, 2;\n\t"
                    "add.u32 a, a, %6;\n\t"
                    "@take ld.shared.u32 %1, [a];\n\t"              // c_beg
                    "@take ld.shared.u32 %2, [a + 128];\n\t"        // c_end
                    "@take ld.shared.f32 %3, [a + 256];\n\t"        // c_qw
                    "@take and.b32 %0, %1, 0xffffffe0;\n\t"         // c_row = c_beg rounded down to a row
                    "sub.u32 %5, %2, %1;\n\t"
                    "@dead mov.u32 %5, 0;\n\t"                      // len
                    "}\n"
                    : "+r"(c_row), "+r"(c_beg), "+r"(c_end), "+f"(c_qw), "+r"(pending), "=r"(len)
                    : "r"(desc_s)
                    : "memory");
                qw = c_qw;
                rel = c_row + lane - c_beg;
                row0 = g_post + c_row;
                c_row += 32u * R;
                return len != 0u;
            };
#else
            auto advance = [&](float& qw, unsigned& rel, unsigned& len, const uint2*& row0) -> bool {
                bool more = true;
                if (c_row >= c_end) {                      // warp-uniform: current slice exhausted
                    if (pending != 0) {
                        const int j = __ffs(pending) - 1;
                        pending &= pending - 1;
                        c_beg = desc[j];
                        c_end = desc[32 + j];
                        c_qw = __uint_as_float(desc[64 + j]);
                        c_row = c_beg & ~31u;
                    } else {
                        more = false;                      // c_row >= c_end stays true: every lane below is dead
                    }
                }
                qw = c_qw;
                rel = c_row + lane - c_beg;
                len = more ? c_end - c_beg : 0u;
                row0 = g_post + c_row;                // one 64-bit address per step; rows are 256 bytes apart (immediates)
                c_row += 32u * R;
                return more;
            };
#endif
            auto fetch = [&](int (&id)[R], float (&w)[R], float& qw, unsigned& rel, unsigned& len) -> bool {
                const uint2* row0;
                const bool more = advance(qw, rel, len, row0);
#if B200RET_SHORT_STEP
                // Half of the query terms are rare: their slice is a handful of postings inside row 0.  A warp-uniform branch
                // keeps the three dead rows' loads out of the L1 data pipe (a fully predicated-off load still takes a slot).
                if (static_cast<int>(rel - lane) + 32 >= static_cast<int>(len)) {
                    asm volatile(
                        "{\n\t"
                        ".reg .pred p0;\n\t"
                        "setp.lt.u32 p0, %2, %3;\n\t"
                        "@p0 " B200RET_LDNC ".v2.b32 {%0, %1}, [%4];\n\t"
                        "}\n"
                        : "=r"(id[0]), "=f"(w[0])
                        : "r"(rel), "r"(len), "l"(row0));
                    return more;
                }
#endif
                asm volatile(
                    "{\n\t"
                    ".reg .pred p0, p1, p2, p3;\n\t"
                    ".reg .u32 t1, t2, t3;\n\t"
                    "add.u32 t1, %8, 32;\n\t"
                    "add.u32 t2, %8, 64;\n\t"
                    "add.u32 t3, %8, 96;\n\t"
                    "setp.lt.u32 p0, %8, %9;\n\t"
                    "setp.lt.u32 p1, t1, %9;\n\t"
                    "setp.lt.u32 p2, t2, %9;\n\t"
                    "setp.lt.u32 p3, t3, %9;\n\t"
                    "@p0 " B200RET_LDNC ".v2.b32 {%0, %4}, [%10];\n\t"          // one posting = {doc id, weight bits}
                    "@p1 " B200RET_LDNC ".v2.b32 {%1, %5}, [%10 + 256];\n\t"
                    "@p2 " B200RET_LDNC ".v2.b32 {%2, %6}, [%10 + 512];\n\t"
                    "@p3 " B200RET_LDNC ".v2.b32 {%3, %7}, [%10 + 768];\n\t"
                    "}\n"
                    : "=r"(id[0]), "=r"(id[1]), "=r"(id[2]), "=r"(id[3]), "=f"(w[0]), "=f"(w[1]), "=f"(w[2]), "=f"(w[3])
                    : "r"(rel), "r"(len), "l"(row0));
                return more;
            };
            // fused(): accumulate step s AND issue the loads of step f in one instruction block, so the posting loads and
            // their predicate arithmetic fill the shared-memory load latency of the accumulate (LDS -> FADD -> STS).
            auto fused = [&](const int (&id)[R], const float (&w)[R], float qw, unsigned rel, unsigned len,
                             int (&fid)[R], float (&fw)[R], unsigned frel, unsigned flen, const uint2* frow0) {
                asm volatile(
                    "{\n\t"
                    ".reg .pred p0, p1, p2, p3, q0, q1, q2, q3;\n\t"
                    ".reg .f32 a0, a1, a2, a3, v0, v1, v2, v3;\n\t"
                    ".reg .u32 d0, d1, d2, d3, t1, t2, t3, u1, u2, u3;\n\t"
                    "add.u32 t1, %17, 32;\n\t"
                    "add.u32 t2, %17, 64;\n\t"
                    "add.u32 t3, %17, 96;\n\t"
                    "setp.lt.u32 p0, %17, %18;\n\t"
                    "setp.lt.u32 p1, t1, %18;\n\t"
                    "setp.lt.u32 p2, t2, %18;\n\t"
                    "setp.lt.u32 p3, t3, %18;\n\t"
                    "mad.lo.u32 d0, %8, 4, %19;\n\t"
                    "mad.lo.u32 d1, %9, 4, %19;\n\t"
                    "mad.lo.u32 d2, %10, 4, %19;\n\t"
                    "mad.lo.u32 d3, %11, 4, %19;\n\t"
                    "@p0 ld.shared.f32 a0, [d0];\n\t"
                    "@p1 ld.shared.f32 a1, [d1];\n\t"
                    "@p2 ld.shared.f32 a2, [d2];\n\t"
                    "@p3 ld.shared.f32 a3, [d3];\n\t"
                    "add.u32 u1, %20, 32;\n\t"
                    "add.u32 u2, %20, 64;\n\t"
                    "add.u32 u3, %20, 96;\n\t"
                    "setp.lt.u32 q0, %20, %21;\n\t"
                    "setp.lt.u32 q1, u1, %21;\n\t"
                    "setp.lt.u32 q2, u2, %21;\n\t"
                    "setp.lt.u32 q3, u3, %21;\n\t"
                    "@q0 " B200RET_LDNC ".v2.b32 {%0, %4}, [%22];\n\t"
                    "@q1 " B200RET_LDNC ".v2.b32 {%1, %5}, [%22 + 256];\n\t"
                    "@q2 " B200RET_LDNC ".v2.b32 {%2, %6}, [%22 + 512];\n\t"
                    "@q3 " B200RET_LDNC ".v2.b32 {%3, %7}, [%22 + 768];\n\t"
                    "mul.rn.f32 v0, %16, %12;\n\t"
                    "mul.rn.f32 v1, %16, %13;\n\t"
                    "mul.rn.f32 v2, %16, %14;\n\t"
                    "mul.rn.f32 v3, %16, %15;\n\t"
                    "@p0 add.rn.f32 a0, a0, v0;\n\t"
                    "@p1 add.rn.f32 a1, a1, v1;\n\t"
                    "@p2 add.rn.f32 a2, a2, v2;\n\t"
                    "@p3 add.rn.f32 a3, a3, v3;\n\t"
                    "@p0 st.shared.f32 [d0], a0;\n\t"
                    "@p1 st.shared.f32 [d1], a1;\n\t"
                    "@p2 st.shared.f32 [d2], a2;\n\t"
                    "@p3 st.shared.f32 [d3], a3;\n\t"
                    "}\n"
                    : "=&r"(fid[0]), "=&r"(fid[1]), "=&r"(fid[2]), "=&r"(fid[3]), "=&f"(fw[0]), "=&f"(fw[1]), "=&f"(fw[2]), "=&f"(fw[3])
                    : "r"(id[0]), "r"(id[1]), "r"(id[2]), "r"(id[3]), "f"(w[0]), "f"(w[1]), "f"(w[2]), "f"(w[3]), "f"(qw), "r"(rel),
                      "r"(len), "r"(acc_rel_s), "r"(frel), "r"(flen), "l"(frow0)
                    : "memory");
                __syncwarp();   // orders this step's shared-memory updates before the next step (possibly the next term)
            };
            // Accumulate one step.  Its rows belong to ONE posting list, so their doc ids are distinct and the R
            // read-modify-writes are independent: loads, adds and stores are issued R-wide (one latency per step).
            // reference arithmetic: scores[doc] += q * w  -> fp32 multiply, then fp32 add (no FMA); dead lanes
            // (id < 0) are predicated off.
            auto consume = [&](const int (&id)[R], const float (&w)[R], float qw, unsigned rel, unsigned len) {
                static_assert(R == 4, "the fetch/accumulate blocks are written for 4 rows per step");
#if B200RET_SHORT_STEP
                if (static_cast<int>(rel - lane) + 32 >= static_cast<int>(len)) {   // same warp-uniform test as in fetch()
                    asm volatile(
                        "{\n\t"
                        ".reg .pred p0;\n\t"
                        ".reg .f32 a0, v0;\n\t"
                        ".reg .u32 d0;\n\t"
                        "setp.lt.u32 p0, %3, %4;\n\t"
                        "mad.lo.u32 d0, %0, 4, %5;\n\t"
                        "@p0 ld.shared.f32 a0, [d0];\n\t"
                        "mul.rn.f32 v0, %2, %1;\n\t"
                        "@p0 add.rn.f32 a0, a0, v0;\n\t"
                        "@p0 st.shared.f32 [d0], a0;\n\t"
                        "}\n" ::"r"(id[0]), "f"(w[0]), "f"(qw), "r"(rel), "r"(len), "r"(acc_rel_s)
                        : "memory");
                    __syncwarp();
                    return;
                }
#endif
                asm volatile(
                    "{\n\t"
                    ".reg .pred p0, p1, p2, p3;\n\t"
                    ".reg .f32 a0, a1, a2, a3, v0, v1, v2, v3;\n\t"
                    ".reg .u32 d0, d1, d2, d3, t1, t2, t3;\n\t"
                    "add.u32 t1, %9, 32;\n\t"
                    "add.u32 t2, %9, 64;\n\t"
                    "add.u32 t3, %9, 96;\n\t"
                    "setp.lt.u32 p0, %9, %10;\n\t"
                    "setp.lt.u32 p1, t1, %10;\n\t"
                    "setp.lt.u32 p2, t2, %10;\n\t"
                    "setp.lt.u32 p3, t3, %10;\n\t"
                    "mad.lo.u32 d0, %0, 4, %11;\n\t"
                    "mad.lo.u32 d1, %1, 4, %11;\n\t"
                    "mad.lo.u32 d2, %2, 4, %11;\n\t"
                    "mad.lo.u32 d3, %3, 4, %11;\n\t"
                    "@p0 ld.shared.f32 a0, [d0];\n\t"
                    "@p1 ld.shared.f32 a1, [d1];\n\t"
                    "@p2 ld.shared.f32 a2, [d2];\n\t"
                    "@p3 ld.shared.f32 a3, [d3];\n\t"
                    "mul.rn.f32 v0, %8, %4;\n\t"
                    "mul.rn.f32 v1, %8, %5;\n\t"
                    "mul.rn.f32 v2, %8, %6;\n\t"
                    "mul.rn.f32 v3, %8, %7;\n\t"
                    "@p0 add.rn.f32 a0, a0, v0;\n\t"
                    "@p1 add.rn.f32 a1, a1, v1;\n\t"
                    "@p2 add.rn.f32 a2, a2, v2;\n\t"
                    "@p3 add.rn.f32 a3, a3, v3;\n\t"
                    "@p0 st.shared.f32 [d0], a0;\n\t"
                    "@p1 st.shared.f32 [d1], a1;\n\t"
                    "@p2 st.shared.f32 [d2], a2;\n\t"
                    "@p3 st.shared.f32 [d3], a3;\n\t"
                    "}\n" ::"r"(id[0]), "r"(id[1]), "r"(id[2]), "r"(id[3]), "f"(w[0]), "f"(w[1]), "f"(w[2]), "f"(w[3]), "f"(qw),
                    "r"(rel), "r"(len), "r"(acc_rel_s)
                    : "memory");
                __syncwarp();   // orders this step's shared-memory updates before the next step (possibly the next term)
            };

            // DEPTH steps in flight in a register ring: fetch step i+DEPTH-1, then consume step i.  The ring indices are
            // compile-time constants after unrolling, so the slots stay in registers.
            constexpr int S = PIPE_DEPTH;
            int id[S][R];
            float w[S][R], qw[S];
            unsigned rel[S], len[S];
            // An empty step (nothing left to fetch) has len == 0; no separate flags are carried through the ring.
#pragma unroll
            for (int s = 0; s < S - 1; ++s) fetch(id[s], w[s], qw[s], rel[s], len[s]);
            bool running = true;
            while (running) {
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const int f = (s + S - 1) % S;                 // slot freed by the previous consume
#if B200RET_FUSED_STEP
                    const uint2* frow0;
                    advance(qw[f], rel[f], len[f], frow0);
                    if (len[s] == 0u) {                            // oldest step is empty: nothing is left at all
                        running = false;
                        break;
                    }
                    fused(id[s], w[s], qw[s], rel[s], len[s], id[f], w[f], rel[f], len[f], frow0);
#else
                    fetch(id[f], w[f], qw[f], rel[f], len[f]);
                    if (len[s] == 0u) {                            // oldest step is empty: nothing is left at all
                        running = false;
                        break;
                    }
                    consume(id[s], w[s], qw[s], rel[s], len[s]);
#endif
                }
            }
            __syncwarp();   // the descriptors are rewritten by the next term group
        }

        if (p.dense_out) {   // verification mode (b200ret_sparse_scores): write the tile out, no selection
            float* dst = p.dense_out + static_cast<size_t>(qi) * p.dense_stride + doc_base;
            for (int i = lane * 4; i < BD; i += 128) {
                *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(acc + i);
                *reinterpret_cast<float4*>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            // Sweep the tile: emit docs with score > tau[q], zero the tile for the next item.
            const float tq = __ldg(p.tau + q);
            const int limit = min(BD, p.n_docs - doc_base);
            uint64_t* cq = p.cand + static_cast<size_t>(q) * p.cap;
            for (int i = lane * 4; i < BD; i += 128) {
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
        }
        __syncwarp();
        item = __shfl_sync(FULL, next_item, 0);
    }
}

static int block_docs_of_shape() { return BLOCK_DOCS; }

static int launch_score(const ScoreParams& sp, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        B200RET_CUDA_CHECK(cudaFuncSetAttribute(sparse_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                static_cast<int>(SCORE_SMEM)));
        attr_set = true;
    }
    B200RET_CUDA_CHECK(cudaMemsetAsync(sp.item_counter, 0, sizeof(unsigned long long), stream));
    prof_begin(PROF_SPARSE_SCORE, stream);
    sparse_score_kernel<<<sm_count(), SCORE_THREADS, SCORE_SMEM, stream>>>(sp);
    prof_end(PROF_SPARSE_SCORE, stream);
    count_launches(1);
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}

static int search_cap(int k) { return k + ROUND0_BLOCKS * block_docs_of_shape(); }

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
    sp.postings = static_cast<const uint2*>(postings);
    sp.q_offsets = q_offsets;
    sp.q_terms = q_terms;
    sp.q_weights = q_weights;
    sp.n_active = n_queries;
    sp.n_docs = n_docs;
    return B200RET_OK;
}

extern "C" int b200ret_sparse_scores(const uint32_t* table, const void* postings,
                                     int32_t n_terms, int32_t n_docs, int32_t block_docs,
                                     const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights,
                                     int32_t n_queries, float* out_scores, void* workspace, size_t workspace_bytes,
                                     void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    B200RET_REQUIRE(n_queries >= 0 && n_docs >= 0 && n_terms > 0, "sparse_scores: bad sizes");
    ScoreParams sp;
    int rc = fill_params(sp, table, postings, n_docs, block_docs, q_offsets, q_terms, q_weights, n_queries);
    if (rc != B200RET_OK) return rc;
    if (n_queries == 0 || n_docs == 0) return B200RET_OK;
    B200RET_REQUIRE(table && q_offsets && out_scores && workspace && workspace_bytes >= 256, "sparse_scores: null pointer / workspace < 256 B");
    const int32_t n_blocks = static_cast<int32_t>(sp.table_stride) - 1;
    sp.blk_begin = 0;
    sp.blk_end = n_blocks;
    sp.item_counter = static_cast<unsigned long long*>(workspace);
    sp.dense_out = out_scores;
    sp.dense_stride = static_cast<size_t>(n_blocks) * block_docs;
    return launch_score(sp, stream);
}

extern "C" int b200ret_sparse_search(const uint32_t* table, const void* postings,
                                     int32_t n_terms, int32_t n_docs, int32_t block_docs,
                                     const int32_t* q_offsets, const int32_t* q_terms, const float* q_weights,
                                     int32_t n_queries, int32_t k, float threshold, int64_t doc_id_base,
                                     float* out_scores, int64_t* out_ids, int32_t* out_counts,
                                     void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ScoreParams sp;
    int rc = fill_params(sp, table, postings, n_docs, block_docs, q_offsets, q_terms, q_weights, n_queries);
    if (rc != B200RET_OK) return rc;
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
    sp.item_counter = b.work_counter;
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
                      out_counts, stream);
}
