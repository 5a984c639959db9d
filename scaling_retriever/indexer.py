"""Drop-in module: `from scaling_retriever.indexer import ...` resolves to the B200 engine.

Put this repository root before the reference checkout on PYTHONPATH; `scaling_retriever` is a namespace package in
both trees, so `scaling_retriever.modeling`, `.dataset`, `.tasks` still come from the reference while `indexer`,
`utils.utils` and `utils.inverted_index` come from here."""
from scaling_retriever_b200.indexer import (  # noqa: F401
    DEVICE_EMBEDDINGS, DenseFlatIndexer, DenseIndexer, HybridIndexer, HybridRetriever, L0, SparseIndexer, SparseRetrieval,
    TermEncoderRetriever, pack_queries, store_embs)
from scaling_retriever_b200.inverted_index import IndexDictOfArray  # noqa: F401
from scaling_retriever_b200.utils import is_first_worker, obtain_doc_vec_dir_files, supports_bfloat16, to_list  # noqa: F401
