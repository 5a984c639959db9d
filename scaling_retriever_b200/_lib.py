"""ctypes binding of libb200ret.so — the only way the Python host side reaches the CUDA kernels.

Every prototype below mirrors include/b200ret.h.  Loading fails loudly (RuntimeError) when the library
is absent and cannot be built: there is no CPU fallback and no alternative backend.
"""
import ctypes
import os
import threading

from . import build as _build

_c_i32 = ctypes.c_int32
_c_i64 = ctypes.c_int64
_c_f32 = ctypes.c_float
_c_sz = ctypes.c_size_t
_c_ptr = ctypes.c_void_p
_c_int = ctypes.c_int



class RoundExchange(ctypes.Structure):
    """b200ret_round_exchange (include/b200ret.h (3b)): the tau exchange between the rounds of a sharded search."""
    HOOK = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p)
    _fields_ = [("aux_rank", ctypes.c_int32), ("n_exchanges", ctypes.c_int32), ("growth", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("aux", ctypes.c_void_p), ("hook", HOOK), ("user", ctypes.c_void_p)]


# name -> (restype, argtypes); kept in one table so tests can check the exported symbol set.
PROTOTYPES = {
    "b200ret_version": (_c_int, []),
    "b200ret_last_error": (ctypes.c_char_p, []),
    "b200ret_device_info": (_c_int, [ctypes.POINTER(_c_int), ctypes.POINTER(_c_int), ctypes.POINTER(_c_int),
                                     ctypes.POINTER(_c_sz)]),
    "b200ret_profile_enable": (_c_int, [_c_int]),
    "b200ret_profile_read": (_c_int, [_c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_c_i64), ctypes.POINTER(_c_i64)]),
    "b200ret_csr_build_workspace_bytes": (_c_sz, [_c_i64, _c_i32, _c_i32, _c_int]),
    "b200ret_csr_build": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_i64, _c_i32, _c_i32, _c_int,
                                   _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_sz, _c_ptr]),
    "b200ret_block_table_build": (_c_int, [_c_ptr, _c_ptr, _c_i64, _c_i32, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_ptr]),
    "b200ret_sparse_layout": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_i64, _c_i32, _c_i32, _c_i32, _c_int, _c_ptr, _c_ptr]),
    "b200ret_sparse_layout_f16": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_i64, _c_i32, _c_i32, _c_i32, _c_int, _c_ptr, _c_ptr]),
    "b200ret_sparse_block_docs": (_c_i32, []),
    "b200ret_sparse_search_workspace_bytes": (_c_sz, [_c_i32, _c_i32]),
    "b200ret_sparse_search": (_c_int, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_i32,
                                       _c_ptr, _c_ptr, _c_ptr, _c_i32, _c_i32, _c_f32, _c_i64,
                                       _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_sz, _c_ptr]),
    "b200ret_sparse_scores": (_c_int, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_i32,
                                       _c_ptr, _c_ptr, _c_ptr, _c_i32, _c_ptr, _c_ptr, _c_sz, _c_ptr]),
    "b200ret_sparse_search_f16": (_c_int, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_i32,
                                           _c_ptr, _c_ptr, _c_ptr, _c_i32, _c_i32, _c_f32, _c_i64,
                                           _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_sz, _c_ptr]),
    "b200ret_sparse_scores_f16": (_c_int, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_i32,
                                           _c_ptr, _c_ptr, _c_ptr, _c_i32, _c_ptr, _c_ptr, _c_sz, _c_ptr]),
    "b200ret_exchange_growth": (_c_i32, [_c_i32]),
    "b200ret_sparse_exchange_rounds": (_c_i32, [_c_i32, _c_i32]),
    "b200ret_dense_exchange_rounds": (_c_i32, [_c_i32, _c_i32]),
    "b200ret_sparse_search_sharded": (_c_int, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_i32,
                                               _c_ptr, _c_ptr, _c_ptr, _c_i32, _c_i32, _c_f32, _c_i64,
                                               _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_sz, _c_ptr, ctypes.POINTER(RoundExchange)]),
    "b200ret_dense_search_sharded": (_c_int, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_i32, _c_i32, _c_i64,
                                              _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_sz, _c_ptr, ctypes.POINTER(RoundExchange)]),
    "b200ret_dense_search_workspace_bytes": (_c_sz, [_c_i32, _c_i32, _c_i32, _c_i32]),
    "b200ret_dense_search": (_c_int, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_i32, _c_i32, _c_i64,
                                      _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_sz, _c_ptr]),
    "b200ret_f32_to_bf16": (_c_int, [_c_ptr, _c_ptr, _c_i64, _c_ptr]),
    "b200ret_merge_topk": (_c_int, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_ptr, _c_ptr]),
    "b200ret_pack_keys": (_c_int, [_c_ptr, _c_ptr, _c_i64, _c_ptr, _c_ptr]),
    "b200ret_merge_max_shards": (_c_i32, [_c_i32]),
    "b200ret_merge_keys": (_c_int, [_c_ptr, _c_i32, _c_i32, _c_i32, _c_ptr, _c_ptr]),
    "b200ret_unpack_keys": (_c_int, [_c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_ptr, _c_ptr]),
    "b200ret_term_search_workspace_bytes": (_c_sz, [_c_i32, _c_i32]),
    "b200ret_term_search": (_c_int, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_i32, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_ptr,
                                     _c_ptr, _c_sz, _c_ptr]),
    "b200ret_term_scores": (_c_int, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_i32, _c_i32, _c_ptr, _c_ptr]),
    "b200ret_rank_metrics": (_c_int, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_i32, ctypes.POINTER(_c_i32), _c_i32,
                                      _c_ptr, _c_ptr, _c_ptr]),
    "b200ret_host_copy": (_c_int, [_c_ptr, _c_ptr, _c_sz, _c_int]),
    "b200ret_write_run_json": (_c_int, [ctypes.c_char_p, _c_ptr, _c_ptr, _c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr,
                                        _c_ptr, _c_ptr, _c_ptr, _c_i64, ctypes.POINTER(_c_i64)]),
}

ERROR_NAMES = {-1: "EINVAL", -2: "ECUDA", -3: "EWORKSPACE", -4: "EUNSORTED", -5: "EOVERFLOW"}

_lock = threading.Lock()
_lib = None


class B200RetError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"b200ret {ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


def lib_path():
    return _build.LIB_PATH


def load():
    """Return the loaded CDLL (building it first if it is missing or stale and nvcc is available)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("B200RET_LIB")                      # override: tuning builds of the same sources
        if not path:
            path = _build.LIB_PATH
            if not _build.is_current():                            # missing, or csrc/ was edited after the last build
                try:
                    _build.build()
                except Exception as exc:  # no silent fallback: the CUDA library is the product
                    if not os.path.exists(path):
                        raise RuntimeError(
                            f"libb200ret.so is missing at {path} and could not be built ({exc}). "
                            "Run `python -m scaling_retriever_b200.build`; there is no CPU fallback.") from exc
                    import warnings
                    warnings.warn(f"libb200ret.so is older than csrc/ and could not be rebuilt ({exc}); using the stale library")
        elif not os.path.exists(path):
            raise RuntimeError(f"B200RET_LIB={path} does not exist")
        lib = ctypes.CDLL(path)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.b200ret_version() != 1:
            raise RuntimeError(f"libb200ret.so version {lib.b200ret_version()} does not match the binding (1)")
        _lib = lib
        return _lib


def check(code):
    if code != 0:
        msg = load().b200ret_last_error()
        raise B200RetError(code, msg.decode() if msg else "")
