// Dense flat inner-product search (placeholder until the tcgen05 kernel lands; fails loudly, no fallback).
#include "common.cuh"

using namespace b200ret;

extern "C" size_t b200ret_dense_search_workspace_bytes(int32_t n_queries, int32_t n_docs, int32_t dim, int32_t k) {
    (void)n_queries; (void)n_docs; (void)dim; (void)k;
    return 256;
}

extern "C" int b200ret_dense_search(const void*, const void*, int32_t, int32_t, int32_t, int32_t, int64_t, float*, int64_t*,
                                    int32_t*, void*, size_t, void*) {
    set_err("dense_search: not implemented yet");
    return B200RET_EINVAL;
}
