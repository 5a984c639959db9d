"""Boundary helpers the eval drivers import from scaling_retriever.utils.utils (reference utils/utils.py:20-43,69-75)."""
import json
import os

import torch
import torch.distributed


def is_first_worker():
    return not torch.distributed.is_available() or not torch.distributed.is_initialized() or torch.distributed.get_rank() == 0


def to_list(tensor):
    return tensor.detach().cpu().tolist()


def obtain_doc_vec_dir_files(doc_embed_dir):
    """plan.json -> (embs_{rank}_{chunk}.npy, ids_{rank}_{chunk}.npy) lists, rank-major (utils/utils.py:26-43)."""
    with open(os.path.join(doc_embed_dir, "plan.json")) as fin:
        plan = json.load(fin)
    doc_vec_files, doc_id_files = [], []
    for i in range(plan["nranks"]):
        for j in range(plan["num_chunks"]):
            vec_file = os.path.join(doc_embed_dir, f"embs_{i}_{j}.npy")
            doc_id_file = os.path.join(doc_embed_dir, f"ids_{i}_{j}.npy")
            assert os.path.exists(vec_file) and os.path.exists(doc_id_file)
            doc_vec_files.append(vec_file)
            doc_id_files.append(doc_id_file)
    return doc_vec_files, doc_id_files


def supports_bfloat16():
    if torch.cuda.is_available():
        props = torch.cuda.get_device_properties(torch.cuda.current_device())
        return props.major >= 8
    return False


def world_size():
    return torch.distributed.get_world_size() if torch.distributed.is_available() and torch.distributed.is_initialized() else 1


def rank():
    return torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0
