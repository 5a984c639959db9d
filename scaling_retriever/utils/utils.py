"""Drop-in for scaling_retriever/utils/utils.py (the helpers the eval drivers import)."""
from scaling_retriever_b200.utils import (  # noqa: F401
    is_first_worker, obtain_doc_vec_dir_files, rank, supports_bfloat16, to_list, world_size)
