// Candidate-list kernels shared by the sparse and the dense search (declared in candidates.cuh).
#include "candidates.cuh"

#ifndef B200RET_SELECT_BULK      // 1 = a query's candidate list enters shared memory as ONE bulk asynchronous copy (TMA engine)
#define B200RET_SELECT_BULK 1   //     instead of 8-byte loads by every thread (the select launches are bound by that load's latency)
#endif

namespace b200ret {

namespace {
__device__ __forceinline__ uint32_t sel_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned), completion counted on `bar`
__device__ __forceinline__ void bulk_load_to_smem(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    const uint32_t b = sel_smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sel_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void bar_wait_phase0(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(sel_smem_u32(bar))
        : "memory");
}
}  // namespace

__global__ void cand_init_kernel(float* tau, int32_t* cand_count, int32_t* overflow, int32_t n_queries, float threshold) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_queries) {
        tau[i] = threshold;
        cand_count[i] = 0;
        overflow[i] = 0;
    }
    if (i == 0) overflow[n_queries] = 0;
}

// Re-arm the overflowed queries for the safe re-run and compact their indices into q_list.
__global__ void cand_rearm_kernel(float* tau, int32_t* cand_count, int32_t* overflow, int32_t n_queries, float threshold,
                                         int32_t* q_list, int32_t* n_list) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_queries) return;
    if (overflow[i]) {
        tau[i] = threshold;
        cand_count[i] = 0;
        overflow[i] = 0;
        q_list[atomicAdd(n_list, 1)] = i;
    } else {
        tau[i] = INFINITY;   // finished query: a scoring kernel that visits all queries (dense GEMM) appends nothing for it
    }
}

// One CTA per query: cut the candidate list to its k best keys, raise tau to the k-th score.
// FINAL additionally sorts and writes the output row (ids + doc_id_base, tail padded with (-inf, -1)).
template <bool FINAL>
__global__ void __launch_bounds__(SELECT_THREADS) select_kernel(uint64_t* cand, int32_t* cand_count, int32_t cap, int32_t k,
                                                                       float* tau, int32_t* overflow, int32_t n_queries,
                                                                       const int32_t* q_list, int64_t doc_id_base,
                                                                       float* out_scores, int64_t* out_ids, int32_t* out_counts,
                                                                       int32_t aux_rank, float* aux) {
    extern __shared__ __align__(16) uint64_t skeys[];
    __shared__ uint32_t hist[256];
    __shared__ uint64_t bcast[3];
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
    // tau exchange of a sharded search: this shard's aux_rank-th best score so far (-inf if it has fewer candidates)
    const bool want_aux = !FINAL && aux != nullptr;
    if (want_aux && c < aux_rank) {
        if (threadIdx.x == 0) aux[q] = -INFINITY;
    }
    if (!FINAL && c <= k && !(want_aux && c >= aux_rank)) return;   // nothing to cut or publish (block-uniform exit)

    const int n_sort = FINAL ? next_pow2(max(min(c, k), 1)) : 0;
#if B200RET_SELECT_BULK
    __shared__ __align__(8) uint64_t load_bar;
    if (threadIdx.x == 0) {
        out_pos = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sel_smem_u32(&load_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (c > 0) {     // block-uniform.  cap is even and the list starts 16-byte aligned, so an odd count copies one spare key
        if (threadIdx.x == 0) bulk_load_to_smem(skeys, cq, static_cast<uint32_t>((c + 1) & ~1) * 8u, &load_bar);
        bar_wait_phase0(&load_bar);
    }
    __syncthreads();
#else
    for (int i = threadIdx.x; i < c; i += blockDim.x) skeys[i] = cq[i];
    if (threadIdx.x == 0) out_pos = 0;
    __syncthreads();
#endif
    int kept = c;
    // tau exchange: a score that at least aux_rank of this shard's candidates reach — read off the first histogram of the
    // selection below (no extra pass over the keys); when nothing has to be cut the same call only serves that purpose
    __shared__ uint64_t aux_lower;
    const bool publish = want_aux && c >= aux_rank;     // block-uniform
    if (publish && c <= k) {
        block_radix_select_kth(skeys, c, 1, hist, bcast, aux_rank, &aux_lower, /*first_pass_only=*/true);
        if (threadIdx.x == 0) aux[q] = cand_score(aux_lower);
    }
    if (c > k) {
        const uint64_t kth = block_radix_select_kth(skeys, c, k, hist, bcast, publish ? aux_rank : 0, &aux_lower);
        if (publish && threadIdx.x == 0) aux[q] = cand_score(aux_lower);
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

// Launch helpers (host).  `prof_kind` brackets the select launches for bench.py.
int launch_select(bool final, const CandBuffers& b, int32_t cap, int32_t k, int32_t n_queries, int32_t n_active,
                         const int32_t* q_list, int64_t doc_id_base, float* out_scores, int64_t* out_ids, int32_t* out_counts,
                         cudaStream_t stream, int32_t aux_rank, float* aux) {
    const size_t smem = static_cast<size_t>(cap) * sizeof(uint64_t);
    static PerDeviceOnce attrs_set;
    if (attrs_set.first()) {
        const int max_smem = 200 * 1024;
        B200RET_CUDA_CHECK(cudaFuncSetAttribute(select_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        B200RET_CUDA_CHECK(cudaFuncSetAttribute(select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        attrs_set.mark();
    }
    if (smem > 200 * 1024) {
        set_err("select: candidate capacity %d needs %zu bytes of shared memory", cap, smem);
        return B200RET_EINVAL;
    }
    prof_begin(PROF_SPARSE_SELECT, stream);
    if (final)
        select_kernel<true><<<n_active, SELECT_THREADS, smem, stream>>>(b.cand, b.cand_count, cap, k, b.tau, b.overflow, n_queries,
                                                                      q_list, doc_id_base, out_scores, out_ids, out_counts, 0, nullptr);
    else
        select_kernel<false><<<n_active, SELECT_THREADS, smem, stream>>>(b.cand, b.cand_count, cap, k, b.tau, b.overflow, n_queries,
                                                                       q_list, doc_id_base, nullptr, nullptr, nullptr, aux_rank, aux);
    prof_end(PROF_SPARSE_SELECT, stream);
    count_launches(1);
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}


__global__ void aux_init_kernel(float* aux, int32_t n_queries) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_queries) aux[i] = -INFINITY;
}

// tau <- max(tau, largest float below aux): documents scoring EXACTLY the exchanged bound stay eligible (a tie with a
// document of another shard is decided by the doc id, which this shard cannot judge)
__global__ void tau_raise_kernel(float* tau, const float* aux, int32_t n_queries) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_queries) return;
    const float a = aux[i];
    if (a > -INFINITY && a == a) tau[i] = fmaxf(tau[i], nextafterf(a, -INFINITY));
}

int launch_aux_init(float* aux, int32_t n_queries, cudaStream_t stream) {
    aux_init_kernel<<<(n_queries + 255) / 256, 256, 0, stream>>>(aux, n_queries);
    count_launches(1);
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}

int launch_tau_raise(float* tau, const float* aux, int32_t n_queries, cudaStream_t stream) {
    tau_raise_kernel<<<(n_queries + 255) / 256, 256, 0, stream>>>(tau, aux, n_queries);
    count_launches(1);
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}

int launch_cand_init(const CandBuffers& b, int32_t n_queries, float threshold, cudaStream_t stream) {
    cand_init_kernel<<<(n_queries + 255) / 256, 256, 0, stream>>>(b.tau, b.cand_count, b.overflow, n_queries, threshold);
    count_launches(1);
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}

int launch_cand_rearm(const CandBuffers& b, int32_t n_queries, float threshold, cudaStream_t stream) {
    B200RET_CUDA_CHECK(cudaMemsetAsync(b.n_list, 0, sizeof(int32_t), stream));
    cand_rearm_kernel<<<(n_queries + 255) / 256, 256, 0, stream>>>(b.tau, b.cand_count, b.overflow, n_queries, threshold, b.q_list,
                                                                   b.n_list);
    count_launches(1);
    B200RET_CUDA_CHECK(cudaMemsetAsync(b.overflow + n_queries, 0, sizeof(int32_t), stream));
    B200RET_CUDA_CHECK(cudaGetLastError());
    return B200RET_OK;
}

}  // namespace b200ret
