// Dense flat inner-product search on sm_100a: tcgen05 (UMMA) bf16 GEMM fed by TMA, accumulators in TMEM, fused top-k.
//
// Replaces faiss.IndexFlatIP.search as called by DenseFlatIndexer.search_knn (reference scaling_retriever/indexer.py:210-214;
// faiss-cpu computes blocked fp32 sgemm + a reservoir top-k on the host).  Here S = Q . D^T is computed tile by tile and
// NEVER written to memory: the epilogue reads each 128 x 256 fp32 accumulator tile from tensor memory, compares every
// score with the query's running bound tau[q] and appends the few survivors to the query's candidate list
// (candidates.cuh: rounds of doubling doc ranges, radix-select cut to k between rounds -> exact top-k).
//
// Kernel shape (one CTA per SM, CTA PAIRS via cta_group::2):
//   * a pair computes a 256 (queries) x 256 (docs) tile: UMMA M=256, N=256, K=16, bf16 x bf16 -> fp32.  Each CTA of the pair
//     owns 128 query rows (its half of A and of the accumulator, 128 TMEM lanes x 256 columns) and TMA-loads half of
//     the doc rows (its half of B); the tensor core reads both halves of B across the pair.
//   * K loop in blocks of 64 (one 128-byte swizzle atom per row), 6-stage TMA -> smem ring (32 KB per stage and CTA).
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only, one thread), warp 2 = TMEM allocator,
//     warps 4-7 = epilogue (thread = one query row = one TMEM lane).  Two accumulator stages (2 x 256 of the 512 TMEM
//     columns): the epilogue of tile i overlaps the MMAs of tile i+1.
//   * persistent: pair p walks tiles p, p + #pairs, ... with the query-block index fastest, so the ~28 pairs working on
//     the same 256 docs share that B tile through L2 and the corpus is streamed from HBM once per query batch.
#include <cuda.h>
#include <cuda_bf16.h>

#include "candidates.cuh"

namespace b200ret {

constexpr int D_BLOCK_M = 128;             // query rows per CTA (256 per pair)
constexpr int D_BLOCK_N = 256;             // doc rows per pair tile (128 loaded by each CTA)
constexpr int D_BLOCK_K = 64;              // bf16 elements per K block = 128 bytes = one swizzle atom row
#ifndef B200RET_DENSE_STAGES
#define B200RET_DENSE_STAGES 6
#endif
#ifndef B200RET_DENSE_EPI_PIPE          // 1 = the epilogue's maximum pass keeps two tcgen05.ld in flight per wait
#define B200RET_DENSE_EPI_PIPE 1
#endif
constexpr int D_STAGES = B200RET_DENSE_STAGES;
constexpr int D_THREADS = 256;
constexpr int D_TMEM_COLS = 512;           // 2 accumulator stages x 256 fp32 columns
constexpr uint32_t D_TILE_BYTES = 128 * D_BLOCK_K * 2;            // one 128-row operand tile: 16 KB
constexpr uint32_t D_STAGE_BYTES = 2 * D_TILE_BYTES;              // A half + B half per CTA
#ifndef B200RET_DENSE_ROUND0_DOCS
#define B200RET_DENSE_ROUND0_DOCS 8192
#endif
constexpr int D_ROUND0_DOCS = B200RET_DENSE_ROUND0_DOCS;        // first round / safe-schedule round size (candidate capacity = k + this)
constexpr int D_UNIT_DOCS = D_BLOCK_N;     // round unit = one doc tile
#ifndef B200RET_DENSE_TILES_PER_PAIR
#define B200RET_DENSE_TILES_PER_PAIR 256
#endif
constexpr int D_MAX_TILES_PER_PAIR = B200RET_DENSE_TILES_PER_PAIR;   // tiles per CTA pair and launch (bounds schedule drift)
constexpr size_t D_SMEM_BYTES = static_cast<size_t>(D_STAGES) * D_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct DenseParams {
    int32_t n_queries;
    int32_t n_docs;        // rows of the corpus shard (TMA bounds); docs >= doc_end are masked in the epilogue
    int32_t doc_begin;     // this round covers doc rows [doc_begin, doc_end)
    int32_t doc_end;
    int32_t k_blocks;      // dim / 64
    int32_t m_blocks;      // ceil(n_active / 256)
    int32_t n_tiles;       // m_blocks * ceil((doc_end - doc_begin) / 256)
    const float* tau;
    uint64_t* cand;
    int32_t* cand_count;
    int32_t cap;
};

// ---- PTX wrappers -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on a barrier that lives in CTA `cta` of the cluster (address given as the local offset of the same barrier)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 r;\n\t"
        "mapa.shared::cluster.u32 r, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [r];\n\t"
        "}\n" ::"r"(local_bar), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMA: 2D tile -> this CTA's shared memory, completion bytes signalled on the LEADER CTA's mbarrier (cta_group::2).
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}

// UMMA shared-memory descriptor: K-major operand tile of 128-byte rows, SWIZZLE_128B, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);   // start address (16-byte units)
    d |= static_cast<uint64_t>(1) << 16;                        // leading byte offset: unused for swizzled K-major
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                // stride byte offset between 8-row groups
    d |= static_cast<uint64_t>(1) << 46;                        // descriptor version (sm_100)
    d |= static_cast<uint64_t>(2) << 61;                        // SWIZZLE_128B
    return d;
}
// UMMA instruction descriptor: D fp32, A/B bf16, both K-major, M = 256 (pair), N = 256.
__device__ __forceinline__ uint32_t umma_instr_desc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit all prior MMAs of this thread; arrive on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t local_bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(local_bar),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
}
__device__ __forceinline__ void tmem_load_32cols(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// two 32-column loads in flight, one wait
__device__ __forceinline__ void tmem_load_2x32cols(uint32_t taddr0, uint32_t taddr1, uint32_t (&a)[32], uint32_t (&b)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%64];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%65];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]),
          "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]), "=r"(a[16]), "=r"(a[17]), "=r"(a[18]),
          "=r"(a[19]), "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]), "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]),
          "=r"(a[28]), "=r"(a[29]), "=r"(a[30]), "=r"(a[31]),
          "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]), "=r"(b[8]), "=r"(b[9]),
          "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15]), "=r"(b[16]), "=r"(b[17]), "=r"(b[18]),
          "=r"(b[19]), "=r"(b[20]), "=r"(b[21]), "=r"(b[22]), "=r"(b[23]), "=r"(b[24]), "=r"(b[25]), "=r"(b[26]), "=r"(b[27]),
          "=r"(b[28]), "=r"(b[29]), "=r"(b[30]), "=r"(b[31])
        : "r"(taddr0), "r"(taddr1)
        : "memory");
}

// ---- the kernel -----------------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(D_THREADS, 1)
dense_search_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_d, const DenseParams p) {
    extern __shared__ unsigned char dense_smem_raw[];
    // 1024-byte alignment: SWIZZLE_128B atoms are 8 rows x 128 B
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dense_smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(D_STAGES) * D_STAGE_BYTES);
    uint64_t* full_bar = bars;                       // [D_STAGES]  TMA bytes landed (used in the leader CTA)
    uint64_t* empty_bar = bars + D_STAGES;           // [D_STAGES]  MMAs reading the stage retired (both CTAs)
    uint64_t* tmem_full_bar = bars + 2 * D_STAGES;   // [2]         accumulator stage complete (both CTAs)
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]         accumulator stage drained by all 8 epilogue warps (leader)
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const uint32_t pair = blockIdx.x >> 1;
    const uint32_t n_pairs = gridDim.x >> 1;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < D_STAGES; ++s) {
            mbar_init(smem_u32(full_bar + s), 1);        // one arrive.expect_tx by the leader's producer
            mbar_init(smem_u32(empty_bar + s), 1);       // one tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(tmem_full_bar + a), 1);   // one tcgen05.commit
            mbar_init(smem_u32(tmem_empty_bar + a), 8);  // 4 epilogue warps x 2 CTAs
        }
        fence_barrier_init();
    }
    if (warp == 2) {   // TMEM allocation: one warp of EACH CTA of the pair
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "n"(D_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer (both CTAs; each loads its 128 query rows and its 128 doc rows per K block) =====
        const uint32_t leader_full0 = map_to_cta(smem_u32(full_bar), 0);
        uint32_t stage = 0, phase = 0;
        for (uint32_t tile = pair; tile < static_cast<uint32_t>(p.n_tiles); tile += n_pairs) {
            const int32_t m_blk = tile % p.m_blocks, n_blk = tile / p.m_blocks;
            const int32_t q_row = m_blk * 256 + cta_rank * 128;
            const int32_t d_row = p.doc_begin + n_blk * D_BLOCK_N + cta_rank * 128;
            for (int32_t kb = 0; kb < p.k_blocks; ++kb) {
                mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
                const uint32_t a_dst = smem_u32(smem + stage * D_STAGE_BYTES);
                const uint32_t b_dst = a_dst + D_TILE_BYTES;
                const uint32_t bar = leader_full0 + stage * 8;
                if (leader) mbar_arrive_expect_tx(smem_u32(full_bar + stage), 2 * D_STAGE_BYTES);   // both CTAs' bytes
                tma_load_2d_pair(a_dst, &map_q, kb * D_BLOCK_K, q_row, bar);
                tma_load_2d_pair(b_dst, &map_d, kb * D_BLOCK_K, d_row, bar);
                if (++stage == D_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0 && leader) {
        // ===== MMA issuer (leader CTA, one thread) =====
        const uint32_t idesc = umma_instr_desc();
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
        for (uint32_t tile = pair; tile < static_cast<uint32_t>(p.n_tiles); tile += n_pairs) {
            mbar_wait(smem_u32(tmem_empty_bar + acc), acc_phase ^ 1);    // epilogue drained this accumulator stage
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * D_BLOCK_N;
            for (int32_t kb = 0; kb < p.k_blocks; ++kb) {
                mbar_wait(smem_u32(full_bar + stage), phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + stage * D_STAGE_BYTES);
                const uint64_t adesc = umma_smem_desc(a_addr);
                const uint64_t bdesc = umma_smem_desc(a_addr + D_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < D_BLOCK_K / 16; ++k)   // advance 32 bytes (16 bf16) inside the swizzle atom per MMA
                    umma_bf16_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                umma_commit_pair(smem_u32(empty_bar + stage));           // frees the stage in both CTAs
                if (++stage == D_STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit_pair(smem_u32(tmem_full_bar + acc));             // accumulator ready in both CTAs
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===== epilogue (4 warps per CTA): thread = query row = TMEM lane =====
        const uint32_t ew = warp & 3u;
        const uint32_t row = ew * 32 + lane;
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t tile = pair; tile < static_cast<uint32_t>(p.n_tiles); tile += n_pairs) {
            const int32_t m_blk = tile % p.m_blocks, n_blk = tile / p.m_blocks;
            const int32_t q = m_blk * 256 + static_cast<int32_t>(cta_rank * 128 + row);
            const bool q_live = q < p.n_queries;
            const float tq = q_live ? __ldg(p.tau + q) : INFINITY;
            const int32_t doc0 = p.doc_begin + n_blk * D_BLOCK_N;
            uint64_t* cq = p.cand + static_cast<size_t>(q_live ? q : 0) * p.cap;
            mbar_wait(smem_u32(tmem_full_bar + acc), acc_phase);
            tc_fence_after();
            // Pass 0 (every tile): row maximum vs tau.  Only if some row of the warp has a hit: count the row's hits, reserve
            // its candidate slots with ONE atomic per row, then emit.  (One atomic per hit serialised ~1 us round trips: with
            // tau still low in the first rounds a tile's epilogue took 20x its MMA time — a fixed cost per shard that
            // dominated small shards.)  tcgen05.ld is warp-collective, so the extra passes are taken warp-uniformly.
            bool any = false;
#if B200RET_DENSE_EPI_PIPE
#pragma unroll 1
            for (int c = 0; c < D_BLOCK_N / 32; c += 2) {
                uint32_t v[32], u[32];
                const uint32_t t0 = tmem_base + ((ew * 32u) << 16) + acc * D_BLOCK_N + c * 32;
                tmem_load_2x32cols(t0, t0 + 32, v, u);
                float m = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j) m = fmaxf(m, fmaxf(__uint_as_float(v[j]), __uint_as_float(u[j])));
                any |= m > tq;
            }
#else
#pragma unroll 1
            for (int c = 0; c < D_BLOCK_N / 32; ++c) {
                uint32_t v[32];
                tmem_load_32cols(tmem_base + ((ew * 32u) << 16) + acc * D_BLOCK_N + c * 32, v);
                float m = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
                any |= m > tq;
            }
#endif
            if (__any_sync(0xffffffffu, any)) {
                const int32_t live_cols = p.doc_end - doc0;      // columns >= live_cols are past the end of the shard
                int cnt = 0;
#pragma unroll 1
                for (int c = 0; c < D_BLOCK_N / 32; ++c) {
                    uint32_t v[32];
                    tmem_load_32cols(tmem_base + ((ew * 32u) << 16) + acc * D_BLOCK_N + c * 32, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) cnt += (__uint_as_float(v[j]) > tq && c * 32 + j < live_cols) ? 1 : 0;
                }
                int pos = 0;
                if (cnt > 0) pos = atomicAdd(p.cand_count + q, cnt);
                if (__any_sync(0xffffffffu, cnt > 0)) {
#pragma unroll 1
                    for (int c = 0; c < D_BLOCK_N / 32; ++c) {
                        uint32_t v[32];
                        tmem_load_32cols(tmem_base + ((ew * 32u) << 16) + acc * D_BLOCK_N + c * 32, v);
#pragma unroll              // static register indices: v[] must not be demoted to local memory
                        for (int j = 0; j < 32; ++j) {
                            const float sc = __uint_as_float(v[j]);
                            if (sc > tq && c * 32 + j < live_cols) {
                                if (pos < p.cap) cq[pos] = cand_key(sc, doc0 + c * 32 + j);
                                ++pos;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(smem_u32(tmem_empty_bar + acc), 0);   // 8 arrivals release the stage
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    __syncwarp();   // single-lane roles rejoin their warp before the warp-aligned cluster barrier
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(D_TMEM_COLS) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// bf16 [rows, dim] row-major -> TMA map with a (64 x 128) box, 128-byte swizzle; out-of-range rows read as zeros.
static int make_map(CUtensorMap* map, const void* base, int64_t rows, int32_t dim) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) {
        set_err("dense_search: cuTensorMapEncodeTiled is not available from the driver");
        return B200RET_ECUDA;
    }
    const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(dim), static_cast<cuuint64_t>(rows)};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(dim) * 2};
    const cuuint32_t box[2] = {D_BLOCK_K, 128};
    const cuuint32_t estride[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_err("dense_search: cuTensorMapEncodeTiled failed (%d) for %lld x %d", static_cast<int>(r), (long long)rows, dim);
        return B200RET_ECUDA;
    }
    return B200RET_OK;
}

static int dense_cap(int k) { return (k + max(D_ROUND0_DOCS, (ROUND_GROWTH + 1) * k) + 1) & ~1; }   // see search_cap (sparse_search.cu)

}  // namespace b200ret

using namespace b200ret;

extern "C" size_t b200ret_dense_search_workspace_bytes(int32_t n_queries, int32_t n_docs, int32_t dim, int32_t k) {
    (void)n_docs; (void)dim;
    Workspace ws(nullptr, 0);
    return carve_cand(ws, n_queries, dense_cap(k), nullptr) + 256;
}

static int dense_search_impl(const void* corpus_bf16, const void* queries_bf16, int32_t n_docs, int32_t n_queries, int32_t dim,
                             int32_t k, int64_t doc_id_base, float* out_scores, int64_t* out_ids, int32_t* out_counts,
                             void* workspace, size_t workspace_bytes, void* stream_, const b200ret_round_exchange* exchange) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    B200RET_REQUIRE(k >= 1 && k <= B200RET_MAX_K, "dense_search: k=%d outside [1, %d]", k, B200RET_MAX_K);
    B200RET_REQUIRE(n_queries >= 0 && n_docs >= 0, "dense_search: bad sizes");
    B200RET_REQUIRE(dim >= D_BLOCK_K && dim % D_BLOCK_K == 0, "dense_search: dim=%d must be a positive multiple of %d", dim, D_BLOCK_K);
    if (n_queries == 0) return B200RET_OK;
    B200RET_REQUIRE(queries_bf16 && out_scores && out_ids && out_counts && workspace, "dense_search: null pointer");
    B200RET_REQUIRE(n_docs == 0 || corpus_bf16, "dense_search: corpus is null");
    B200RET_REQUIRE((reinterpret_cast<uintptr_t>(queries_bf16) % 16 == 0) && (reinterpret_cast<uintptr_t>(corpus_bf16) % 16 == 0),
                    "dense_search: operands must be 16-byte aligned");
    if (workspace_bytes < b200ret_dense_search_workspace_bytes(n_queries, n_docs, dim, k)) {
        set_err("dense_search: workspace too small (%zu bytes)", workspace_bytes);
        return B200RET_EWORKSPACE;
    }
    Workspace ws(workspace, workspace_bytes);
    CandBuffers b;
    const int cap = dense_cap(k);
    carve_cand(ws, n_queries, cap, &b);

    CUtensorMap map_q, map_d;
    int rc = make_map(&map_q, queries_bf16, n_queries, dim);
    if (rc != B200RET_OK) return rc;
    if (n_docs > 0) {
        rc = make_map(&map_d, corpus_bf16, n_docs, dim);
        if (rc != B200RET_OK) return rc;
    } else {
        map_d = map_q;
    }
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        B200RET_CUDA_CHECK(cudaFuncSetAttribute(dense_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                static_cast<int>(D_SMEM_BYTES)));
        attr_set.mark();
    }
    const int grid = (sm_count() / 2) * 2;   // whole pairs

    DenseParams dp{};
    dp.n_queries = n_queries;
    dp.n_docs = n_docs;
    dp.k_blocks = dim / D_BLOCK_K;
    dp.tau = b.tau;
    dp.cand = b.cand;
    dp.cand_count = b.cand_count;
    dp.cap = cap;
    const int32_t n_units = (n_docs + D_UNIT_DOCS - 1) / D_UNIT_DOCS;

    // The safe re-run scores all query blocks again (q_list only narrows the select kernels): rows of queries that did
    // not overflow are recomputed identically, which keeps the GEMM tiling independent of the subset.
    // A round is cut into launches of at most D_MAX_TILES_PER_PAIR tiles per CTA pair.  The tile schedule inside a launch
    // is static (pair p takes tiles p, p + #pairs, ...); over a long launch the pairs drift apart, the ~28 pairs that
    // should share one 256-doc B tile through L2 stop being co-temporal and the corpus is re-read from HBM many times
    // (measured 13x in a 127 ms launch, none in launches <= 10 ms).  Re-synchronising at launch boundaries bounds the
    // drift for ~10 us per launch.
    const int32_t m_blocks = (n_queries + 255) / 256;
    const int32_t units_per_launch = max(1, (D_MAX_TILES_PER_PAIR * (grid / 2)) / m_blocks);
    auto launch_round = [&](int unit_begin, int unit_end, const int32_t* q_list, int32_t n_active) -> int {
        (void)q_list; (void)n_active;
        for (int u = unit_begin; u < unit_end; u += units_per_launch) {
            const int ue = min(unit_end, u + units_per_launch);
            DenseParams r = dp;
            r.doc_begin = u * D_UNIT_DOCS;
            r.doc_end = min(n_docs, ue * D_UNIT_DOCS);
            r.m_blocks = m_blocks;
            r.n_tiles = m_blocks * (ue - u);
            prof_begin(PROF_DENSE_GEMM, stream);
            dense_search_kernel<<<grid, D_THREADS, D_SMEM_BYTES, stream>>>(map_q, map_d, r);
            prof_end(PROF_DENSE_GEMM, stream);
            count_launches(1);
        }
        B200RET_CUDA_CHECK(cudaGetLastError());
        return B200RET_OK;
    };
    return run_search(launch_round, b, cap, k, n_queries, n_units, D_ROUND0_DOCS / D_UNIT_DOCS, -INFINITY, doc_id_base, out_scores,
                      out_ids, out_counts, stream, exchange);
}

extern "C" int b200ret_dense_search(const void* corpus_bf16, const void* queries_bf16, int32_t n_docs, int32_t n_queries, int32_t dim,
                                    int32_t k, int64_t doc_id_base, float* out_scores, int64_t* out_ids, int32_t* out_counts,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    return dense_search_impl(corpus_bf16, queries_bf16, n_docs, n_queries, dim, k, doc_id_base, out_scores, out_ids, out_counts,
                             workspace, workspace_bytes, stream, nullptr);
}

extern "C" int32_t b200ret_dense_exchange_rounds(int32_t n_docs_largest_shard, int32_t n_shards) {
    return schedule_exchanges((max(n_docs_largest_shard, 0) + D_UNIT_DOCS - 1) / D_UNIT_DOCS, D_ROUND0_DOCS / D_UNIT_DOCS,
                              exchange_growth(n_shards));
}

extern "C" int b200ret_dense_search_sharded(const void* corpus_bf16, const void* queries_bf16, int32_t n_docs, int32_t n_queries,
                                            int32_t dim, int32_t k, int64_t doc_id_base, float* out_scores, int64_t* out_ids,
                                            int32_t* out_counts, void* workspace, size_t workspace_bytes, void* stream,
                                            const b200ret_round_exchange* exchange) {
    return dense_search_impl(corpus_bf16, queries_bf16, n_docs, n_queries, dim, k, doc_id_base, out_scores, out_ids, out_counts,
                             workspace, workspace_bytes, stream, exchange);
}
