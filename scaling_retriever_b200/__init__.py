"""scaling_retriever_b200 — B200-native first-stage retrieval engine (sparse inverted index + dense flat IP).

Drop-in for the scoring / top-k path of HansiZeng/scaling-retriever's scaling_retriever/indexer.py: the same Python
class API (indexer.py, inverted_index.py, utils.py here) over hand-written sm_100a CUDA kernels (csrc/) reached
through a C ABI (include/b200ret.h, bound with ctypes in _lib.py).  No CPU fallback.
"""
__version__ = "0.1.0"
