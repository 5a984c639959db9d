"""B200-native drop-in for the retrieval engine of scaling_retriever/indexer.py (reference).

Same public names, constructor arguments, return values and side-effect files as the reference module, so
eval_sparse.py (:104-106, :149-151) and eval_dense.py (:188, :193-199, :223-227) drive it unchanged:

    store_embs · DenseIndexer · DenseFlatIndexer · SparseIndexer · SparseRetrieval

What moved to the GPU (hand-written sm_100a kernels behind include/b200ret.h, bound in ops.py):
  * SparseIndexer.index / IndexDictOfArray.add_batch_document: COO batches stay on the device, one radix-sort CSR build;
  * SparseRetrieval._sparse_retrieve_multithreaded: the numba scorer + argpartition become one batched search call;
  * DenseFlatIndexer.index_data / search_knn: faiss.IndexFlatIP becomes a bf16 corpus in HBM + a fused GEMM/top-k kernel.
There is no CPU fallback: constructing a retriever or indexer without a CUDA device raises.
"""
import json
import logging
import os
import pickle
from collections import defaultdict
from typing import List, Tuple

import numpy as np
import torch
from tqdm import tqdm

from . import ops, shard
from .inverted_index import MAX_SHARD_POSTINGS, IndexDictOfArray
from .results import ExternalIds, IdRows, LazyRun, owned_copy
from .utils import is_first_worker, obtain_doc_vec_dir_files, rank as _rank, supports_bfloat16, to_list, world_size as _world_size

logger = logging.getLogger()


def _unwrap_model(model):
    # transformers.modeling_utils.unwrap_model (used by the reference, indexer.py:10,52) without the import cost
    while hasattr(model, "module"):
        model = model.module
    return model


class L0:
    """non-differentiable (reference modeling/losses/regulariaztion.py:9-14)"""

    def __call__(self, batch_rep):
        return torch.count_nonzero(batch_rep, dim=-1).float().mean()


def _cuda_device(device):
    if not torch.cuda.is_available():
        raise RuntimeError("the B200 retrieval engine needs a CUDA device; there is no CPU fallback")
    if isinstance(device, torch.device):
        return device if device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
    if isinstance(device, int):
        return torch.device("cuda", device)
    if isinstance(device, str) and device.startswith("cuda"):
        return torch.device(device)
    return torch.device("cuda", torch.cuda.current_device())


def _pinned(staging, name, shape, dtype):
    """Reusable pinned host staging buffers (grown on demand) so host<->device copies run at full PCIe rate."""
    need = int(np.prod(shape)) if len(shape) else 1
    buf = staging.get(name)
    if buf is None or buf.dtype != dtype or buf.numel() < need:
        buf = torch.empty(max(need, 1), dtype=dtype, pin_memory=True)
        staging[name] = buf
    return buf[:need].view(shape)


def _stage_in(staging, device, name, host_array, dtype):
    """Host array -> device tensor.  Pageable input goes through a reusable pinned staging buffer; an input that already lives
    in pinned memory (a numpy view of a pinned tensor) is copied straight from where it is."""
    src = torch.from_numpy(np.ascontiguousarray(host_array))
    if src.dtype == dtype and src.numel() and src.is_pinned():
        return src.to(device, non_blocking=True)
    src = src.to(dtype)
    pinned = _pinned(staging, name, tuple(src.shape), dtype)
    pinned.copy_(src)
    return pinned.to(device, non_blocking=True)


def _stage_out(staging, name, dev_tensor):
    pinned = _pinned(staging, name, tuple(dev_tensor.shape), dev_tensor.dtype)
    pinned.copy_(dev_tensor, non_blocking=True)
    return pinned


# ------------------------------------------------------------------------------------------------------------------
# dense: embeddings dump (unchanged format) + flat inner-product index
# ------------------------------------------------------------------------------------------------------------------

# embs_{rank}_{chunk}.npy path -> bf16 CUDA tensor of the same rows, kept when the corpus is encoded with keep_on_device=True
# (SURVEY §8 f1): a DenseFlatIndexer in the same process ingests these instead of re-reading the fp32 files.
DEVICE_EMBEDDINGS = {}


def _write_emb_chunk(index_dir, local_rank, chunk_idx, embeddings, embeddings_ids, force_int64=False, keep_on_device=False,
                     staging=None):
    """embs_{rank}_{chunk}.npy fp32 [n, d] + ids_{rank}_{chunk}.npy (reference indexer.py:58-70; the hybrid indexer always
    casts the ids to int64, :796).  `embeddings` is a list of fp32 CUDA batches: they are concatenated on the device and
    leave it with ONE copy through pinned memory per chunk (the reference syncs with .cpu() once per batch, :56)."""
    reps = torch.cat(embeddings) if len(embeddings) > 1 else embeddings[0]
    if reps.is_cuda:   # bounded pinned staging (256 MB slices), not one pinned buffer of the whole 2 M-doc chunk
        staging = {} if staging is None else staging
        embeddings_np = np.empty(tuple(reps.shape), dtype=np.float32)
        rows = max(1, (256 << 20) // (4 * max(reps.shape[1], 1)))
        with torch.inference_mode():   # the staging buffer may have been created under the encoder loop's inference mode
            for a in range(0, reps.shape[0], rows):
                host = _stage_out(staging, "embs", reps[a:a + rows])
                torch.cuda.current_stream().synchronize()
                embeddings_np[a:a + rows] = host.numpy()
    else:
        embeddings_np = reps.numpy()
    if force_int64 or isinstance(embeddings_ids[0], int):
        embeddings_ids = np.array(embeddings_ids, dtype=np.int64)
    assert len(embeddings_np) == len(embeddings_ids), (len(embeddings_np), len(embeddings_ids))
    emb_path = os.path.join(index_dir, "embs_{}_{}.npy".format(local_rank, chunk_idx))
    np.save(emb_path, embeddings_np)
    np.save(os.path.join(index_dir, "ids_{}_{}.npy".format(local_rank, chunk_idx)), embeddings_ids)
    if keep_on_device and reps.is_cuda:
        DEVICE_EMBEDDINGS[os.path.abspath(emb_path)] = ops.f32_to_bf16(reps.contiguous())
    return embeddings_np.shape


def store_embs(model, collection_loader, local_rank, index_dir, device,
               chunk_size=2_000_000, use_fp16=False, is_query=False, idx_to_id=None, keep_on_device=False):
    """Encode the corpus and write embs_{rank}_{chunk}.npy / ids_{rank}_{chunk}.npy / plan.json (reference
    indexer.py:26-97).  The on-disk format is the dense index INPUT and is kept byte-compatible.  Encoder outputs are
    accumulated on the device (no per-batch host sync); `keep_on_device=True` additionally keeps every chunk in HBM as bf16
    (DEVICE_EMBEDDINGS) so that a search in the same job skips the fp32 .npy round trip."""
    write_freq = chunk_size // collection_loader.batch_size
    if is_first_worker():
        print("write_freq: {}, batch_size: {}, chunk_size: {}".format(write_freq, collection_loader.batch_size, chunk_size))
    dtype = torch.bfloat16 if supports_bfloat16() else torch.float32
    print("Using bfloat16" if dtype == torch.bfloat16 else "Using float32")
    staging = {}

    embeddings, embeddings_ids, chunk_idx = [], [], 0
    for idx, batch in tqdm(enumerate(collection_loader), disable=not is_first_worker(),
                           desc=f"encode # {len(collection_loader)} seqs", total=len(collection_loader)):
        with torch.inference_mode():
            with torch.amp.autocast("cuda", dtype=dtype):
                inputs = {k: v.to(device) for k, v in batch.items() if k != "ids"}
                if is_query:
                    raise NotImplementedError
                reps = _unwrap_model(model).doc_encode(**inputs)
                text_ids = batch["ids"]
        embeddings.append(reps.float())
        assert isinstance(text_ids, list)
        embeddings_ids.extend(text_ids)
        if (idx + 1) % write_freq == 0:
            _write_emb_chunk(index_dir, local_rank, chunk_idx, embeddings, embeddings_ids, keep_on_device=keep_on_device,
                             staging=staging)
            embeddings, embeddings_ids = [], []
            chunk_idx += 1
    if len(embeddings) != 0:
        shape = _write_emb_chunk(index_dir, local_rank, chunk_idx, embeddings, embeddings_ids, keep_on_device=keep_on_device,
                                 staging=staging)
        print("last embedddings shape = {}".format(tuple(shape)))
        chunk_idx += 1

    plan = {"nranks": _world_size(), "num_chunks": chunk_idx, "index_path": os.path.join(index_dir, "model.index")}
    print("plan: ", plan)
    if is_first_worker():
        with open(os.path.join(index_dir, "plan.json"), "w") as fout:
            json.dump(plan, fout)


FAISS_FLAT_IP_FOURCC = b"IxFI"


def write_faiss_flat_ip(path, corpus_bf16, chunk_rows=262144):
    """index.dpr in the on-disk layout of faiss.write_index(faiss.IndexFlatIP) (faiss 1.8 impl/index_write.cpp — restated
    from the published format, faiss itself is not installable here): fourcc "IxFI" | d int32 | ntotal int64 | 2 x int64
    (1 << 20, unused) | is_trained u8 | metric_type int32 (0 = inner product) | n_floats uint64 | fp32 [ntotal, d]."""
    n, d = corpus_bf16.shape
    with open(path, "wb") as f:
        f.write(FAISS_FLAT_IP_FOURCC)
        f.write(np.int32(d).tobytes() + np.int64(n).tobytes() + np.int64(1 << 20).tobytes() * 2 + b"\x01" + np.int32(0).tobytes())
        f.write(np.uint64(n * d).tobytes())
        for i in range(0, n, chunk_rows):
            f.write(corpus_bf16[i:i + chunk_rows].float().cpu().numpy().tobytes())


def read_faiss_flat_ip(path):
    """(fp32 [ntotal, d] memory map, d) of a faiss IndexFlatIP file (see write_faiss_flat_ip)."""
    with open(path, "rb") as f:
        head = f.read(45)
    if head[:4] != FAISS_FLAT_IP_FOURCC:
        raise ValueError(f"{path}: not a faiss IndexFlatIP file (fourcc {head[:4]!r}); only flat inner-product indexes "
                         "(what DenseFlatIndexer.serialize writes, reference indexer.py:145-158) are supported")
    d = int(np.frombuffer(head, dtype=np.int32, count=1, offset=4)[0])
    n = int(np.frombuffer(head, dtype=np.int64, count=1, offset=8)[0])
    metric = int(np.frombuffer(head, dtype=np.int32, count=1, offset=33)[0])
    n_floats = int(np.frombuffer(head, dtype=np.uint64, count=1, offset=37)[0])
    if metric != 0 or n_floats != n * d:
        raise ValueError(f"{path}: unexpected IndexFlatIP header (metric {metric}, {n_floats} floats for {n} x {d})")
    return np.memmap(path, dtype=np.float32, mode="r", offset=45, shape=(n, d)), d


class DenseIndexer(object):
    """Base class with the reference's (de)serialisation surface (indexer.py:127-188).  `self.index` is a bf16 CUDA
    tensor [N, d] instead of a faiss object.  index.dpr is written in faiss's own IndexFlatIP file layout (fp32 rows), so
    the reference can read what this class writes and vice versa; `serialize(file, fmt="bf16")` writes a compact .npy of the
    raw bf16 words instead (half the size, read back by deserialize as well).  index_meta.dpr is the pickled id list
    exactly as in the reference."""

    def __init__(self, buffer_size: int = 50000):
        self.buffer_size = buffer_size
        self.index_id_to_db_id = []
        self.index = None

    def init_index(self, vector_sz: int):
        raise NotImplementedError

    def index_data(self, data: List[Tuple[object, np.array]]):
        raise NotImplementedError

    def get_index_name(self):
        raise NotImplementedError

    def search_knn(self, query_vectors: np.array, top_docs: int):
        raise NotImplementedError

    def serialize(self, file: str, fmt: str = "faiss"):
        logger.info("Serializing index to %s", file)
        if os.path.isdir(file):
            index_file = os.path.join(file, "index.dpr")
            meta_file = os.path.join(file, "index_meta.dpr")
        else:
            index_file = file + ".index.dpr"
            meta_file = file + ".index_meta.dpr"
        if _world_size() > 1:
            raise NotImplementedError("serialize() of a sharded index: every rank holds only its doc-row range")
        if fmt == "faiss":
            write_faiss_flat_ip(index_file, self.index)
        elif fmt == "bf16":
            with open(index_file, "wb") as f:
                np.save(f, self.index.view(torch.int16).cpu().numpy())
        else:
            raise ValueError(f"serialize: unknown format {fmt!r}")
        with open(meta_file, mode="wb") as f:
            pickle.dump(self.index_id_to_db_id, f)

    def get_files(self, path: str):
        if os.path.isdir(path):
            index_file = os.path.join(path, "index.dpr")
            meta_file = os.path.join(path, "index_meta.dpr")
        else:
            index_file = path + ".{}.dpr".format(self.get_index_name())
            meta_file = path + ".{}_meta.dpr".format(self.get_index_name())
        return index_file, meta_file

    def index_exists(self, path: str):
        index_file, meta_file = self.get_files(path)
        return os.path.isfile(index_file) and os.path.isfile(meta_file)

    def deserialize(self, path: str):
        """Works without a prior init_index (eval_dense.py:194-196 calls deserialize on a fresh object).  Under
        torch.distributed every rank keeps only its doc-row range of the file."""
        logger.info("Loading index from %s", path)
        index_file, meta_file = self.get_files(path)
        with open(index_file, "rb") as f:
            magic = f.read(6)
        device = _cuda_device(getattr(self, "device", None))
        if magic == b"\x93NUMPY":
            words = np.load(index_file, mmap_mode="r")
            n, dim = words.shape
            lo, hi = shard.ShardPlan(n, _world_size()).bounds(_rank())
            corpus = torch.from_numpy(np.ascontiguousarray(words[lo:hi])).to(device).view(torch.bfloat16)
        elif magic[:4] == FAISS_FLAT_IP_FOURCC:
            rows, dim = read_faiss_flat_ip(index_file)
            n = rows.shape[0]
            lo, hi = shard.ShardPlan(n, _world_size()).bounds(_rank())
            corpus = torch.empty((hi - lo, dim), dtype=torch.bfloat16, device=device)
            step = max(1, (256 << 20) // (4 * dim))
            for a in range(lo, hi, step):
                b = min(hi, a + step)
                corpus[a - lo:b - lo] = ops.f32_to_bf16(torch.from_numpy(np.ascontiguousarray(rows[a:b])).to(device))
        else:
            raise ValueError(f"{index_file}: neither a faiss IndexFlatIP file nor the bf16 .npy written by "
                             f"serialize(fmt='bf16') (first bytes {magic!r})")
        self.index = corpus
        self.device = device
        self.hidden_dim = int(dim)
        self._row_lo = lo
        self._n_total = n
        self._ext = None
        logger.info("Loaded index of type %s and size %d", type(self.index), n)
        with open(meta_file, "rb") as reader:
            self.index_id_to_db_id = pickle.load(reader)
        assert len(self.index_id_to_db_id) == n, "Deserialized index_id_to_db_id should match index size"

    def _update_id_mapping(self, db_ids: List):
        self.index_id_to_db_id.extend(db_ids)
        return len(self.index_id_to_db_id)


class DenseFlatIndexer(DenseIndexer):
    """Exact inner-product index (reference indexer.py:191-217, faiss.IndexFlatIP): the corpus lives in HBM as bf16
    [N, d]; search is the fused tcgen05 GEMM + top-k kernel (ops.dense_search).  Under torch.distributed with
    world_size > 1 each rank keeps the doc-row range shard.ShardPlan assigns it and results are merged after an
    all-gather."""

    def __init__(self, buffer_size: int = 50000, device=None):
        super().__init__(buffer_size)
        self.device = device
        self.hidden_dim = None
        self._row_lo = 0
        self._n_total = 0
        self._staging = {}
        self._ext = None

    def init_index(self, hidden_dim):
        self.device = _cuda_device(self.device)
        self.hidden_dim = int(hidden_dim)
        self.index = torch.empty((0, self.hidden_dim), dtype=torch.bfloat16, device=self.device)
        self._row_lo = 0
        self._n_total = 0

    def _rows_to_bf16(self, doc_reps, a, b):
        """Rows [a, b) of the batch as a bf16 CUDA tensor: host fp32 -> device -> cast kernel, or straight from CUDA tensors
        (fp32 / bf16) when the encoder output never left the device."""
        if isinstance(doc_reps, torch.Tensor):
            chunk = doc_reps[a:b].to(self.device)
            return chunk.contiguous() if chunk.dtype == torch.bfloat16 else ops.f32_to_bf16(chunk.float().contiguous())
        chunk = torch.from_numpy(np.ascontiguousarray(doc_reps[a:b], dtype=np.float32)).to(self.device)
        return ops.f32_to_bf16(chunk)

    def index_data(self, doc_reps, doc_ids):
        """Additive like faiss's index.add (reference indexer.py:198-208): rows are appended behind what the index already
        holds (a second call, or a call after deserialize, extends the corpus and the id list consistently).
        `doc_reps`: fp32 ndarray [n, d] as in the reference, or a torch tensor (CUDA fp32/bf16: no host round trip).
        Sharded (world_size > 1): each rank keeps its doc-row range of the batch; only the first batch can be sharded."""
        assert len(doc_reps) == len(doc_ids)
        n = len(doc_reps)
        n_before = len(self.index_id_to_db_id)
        if self.hidden_dim is None:
            self.init_index(doc_reps.shape[1])
        world = _world_size()
        if world > 1 and n_before > 0:
            raise NotImplementedError("sharded DenseFlatIndexer: the corpus must arrive in ONE index_data call (each rank keeps "
                                      "one contiguous doc-row range); got a second batch")
        lo, hi = shard.ShardPlan(n, world).bounds(_rank())
        fresh = torch.empty((hi - lo, self.hidden_dim), dtype=torch.bfloat16, device=self.device)
        n_total = n_before
        for i in tqdm(range(0, n, self.buffer_size), total=n // self.buffer_size, desc="indexing", disable=not is_first_worker()):
            a, b = max(i, lo), min(i + self.buffer_size, hi)
            if a < b:   # this slice (partly) belongs to our shard
                fresh[a - lo:b - lo] = self._rows_to_bf16(doc_reps, a, b)
            n_total = self._update_id_mapping(doc_ids[i:i + self.buffer_size])
            logger.info("data indexed %d", n_total)
        assert n_total == n_before + n, (n_total, n_before, n)
        if n_before > 0:
            self.index = torch.cat([self.index, fresh])
        else:
            self.index = fresh
            self._row_lo = lo
        self._n_total = n_total
        self._ext = None
        logger.info("total data indexed %d", n_total)

    def search_arrays(self, query_reps, top_docs, host_ranks="all"):
        """(scores fp32 [Q, k] descending, row labels int64 [Q, k], -1 padded) as host numpy arrays — VIEWS of reusable pinned
        buffers, overwritten by the next call of this object (search_knn returns copies).
        `host_ranks="first"` (sharded search only): the merged result is copied to the host of the first worker alone — the
        rank that writes the run file — and the other ranks return (None, None)."""
        if self.index is None or self.device is None:
            raise RuntimeError("DenseFlatIndexer: init_index / index_data / deserialize first")
        with torch.cuda.device(self.device):
            q = _stage_in(self._staging, self.device, "queries", query_reps, torch.float32)   # pinned -> device
            q16 = ops.f32_to_bf16(q)
            exchange = None
            if _world_size() > 1:     # the shards exchange their bounds between the rounds (shard.TauExchange)
                if self._staging.get("tau_exchange") is None:
                    self._staging["tau_exchange"] = shard.TauExchange("dense", self._n_total, self.device)
                exchange = self._staging["tau_exchange"]
            scores, ids, _ = ops.dense_search(self.index, q16, int(top_docs), doc_id_base=self._row_lo, exchange=exchange)
            if _world_size() > 1:
                if host_ranks == "first" and self._n_total < shard.KEY_ID_LIMIT:
                    key = (scores.shape[0], int(top_docs))
                    if self._staging.get("shared_rows_key") != key:
                        self._staging["shared_rows"] = shard.SharedHostRows(*key)
                        self._staging["shared_rows_key"] = key
                    return shard.merge_shards_to_first_host(scores, ids, int(top_docs), self._staging["shared_rows"])[:2]
                scores, ids, _ = shard.merge_shards(scores, ids, int(top_docs), n_docs_total=self._n_total)
            if host_ranks == "first" and not is_first_worker():
                torch.cuda.current_stream().synchronize()
                return None, None
            out_s, out_i = _stage_out(self._staging, "scores", scores), _stage_out(self._staging, "ids", ids)
            torch.cuda.current_stream().synchronize()
            return out_s.numpy(), out_i.numpy()

    def search_knn(self, query_reps: np.array, top_docs: int):
        """reference indexer.py:210-214.  Sharded (torchrun, world_size > 1): like retrieve() on the sparse side, the merged rows
        are delivered to the FIRST worker only (each GPU copies its query slice into host rows shared with it); the other ranks
        get zero rows."""
        scores, indexes = self.search_arrays(query_reps, top_docs, host_ranks="first")
        if scores is None:
            return IdRows(self.external_ids(), np.zeros((0, int(top_docs)), np.int64)), np.zeros((0, int(top_docs)), np.float32)
        # reference indexer.py:212: [[index_id_to_db_id[idx] ...]] (a -1 label indexes the last id there; kept identical) as a
        # sequence of rows gathered on access (results.IdRows) instead of Q*k eager Python list lookups
        top_doc_ids = IdRows(self.external_ids(), owned_copy(indexes))
        return top_doc_ids, owned_copy(scores)   # search_arrays returns views of reusable pinned staging buffers

    def external_ids(self):
        if self._ext is None or self._ext.size != len(self.index_id_to_db_id):
            self._ext = ExternalIds(self.index_id_to_db_id)
        return self._ext

    def get_index_name(self):
        return "flat_index"


# ------------------------------------------------------------------------------------------------------------------
# sparse: index build + retrieval
# ------------------------------------------------------------------------------------------------------------------

def _batch_ids(batch, id_dict=None):
    if isinstance(batch["ids"], torch.Tensor):
        batch_ids = to_list(batch["ids"])
    else:
        assert isinstance(batch["ids"], list)
        batch_ids = batch["ids"]
    if id_dict:
        batch_ids = [id_dict[x] for x in batch_ids]
    return batch_ids


def _add_sparse_batch(sparse_index, batch_documents, batch_ids, count, world_size, rank, doc_ids, require_all=False):
    """One encoder batch [bz, V] -> COO postings appended to the GPU-resident log of `sparse_index`, and the row -> external
    id entries of `doc_ids` (reference indexer.py:259-284 / :776-790): global row = (row + count) * world_size + rank."""
    row, col = torch.nonzero(batch_documents, as_tuple=True)   # row-major: row asc, col asc
    data = batch_documents[row, col]
    g_row = (row + count) * world_size + rank                  # indexer.py:261-262, kept on the device
    present = torch.zeros(len(batch_ids), dtype=torch.bool, device=row.device)
    present[row] = True
    if bool(present.all()):
        base = count * world_size + rank
        doc_ids.update({base + i * world_size: y for i, y in enumerate(batch_ids)})
    else:   # docs without any posting get no doc_ids entry (indexer.py:273-283); the hybrid indexer forbids them (:784)
        assert not require_all, (int(present.sum()), len(batch_ids))
        for i in torch.nonzero(present).flatten().tolist():
            doc_ids[(count + i) * world_size + rank] = batch_ids[i]
    sparse_index.add_batch_document(g_row, col, data.float(), n_docs=len(batch_ids))


class SparseIndexer:
    def __init__(self, model, index_dir, device, compute_stats=False, dim_voc=None, force_new=True,
                 filename="array_index.h5py", **kwargs):
        self.model = model
        self.model.eval()
        self.index_dir = index_dir
        self.device = device
        self._cuda = _cuda_device(device)
        self.sparse_index = IndexDictOfArray(self.index_dir, dim_voc=dim_voc, force_new=force_new, filename=filename,
                                             device=self._cuda)
        self.compute_stats = compute_stats
        if self.compute_stats:
            self.l0 = L0()
        self.model.to(self._cuda)
        self.local_rank = self.device
        # the reference asserts device == dist.get_rank() (indexer.py:235); without a process group rank is 0 / world 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            assert self.local_rank == torch.distributed.get_rank(), (self.local_rank, torch.distributed.get_rank())
        self.rank = _rank()
        self.world_size = _world_size()
        print("world_size: {}, local_rank: {}".format(self.world_size, self.local_rank))

    def index(self, collection_loader, id_dict=None):
        dtype = torch.bfloat16 if supports_bfloat16() else torch.float32
        print("Using bfloat16" if dtype == torch.bfloat16 else "Using float32")
        doc_ids = {}
        stats = defaultdict(float)
        count = 0
        with torch.inference_mode():
            for t, batch in enumerate(tqdm(collection_loader, disable=not is_first_worker())):
                inputs = {k: v.to(self._cuda) for k, v in batch.items() if k not in {"ids"}}
                with torch.amp.autocast("cuda", dtype=dtype):
                    batch_documents = self.model.encode(**inputs)   # [bz, vocab_size]
                if self.compute_stats:
                    stats["L0_d"] += self.l0(batch_documents).item()
                batch_ids = _batch_ids(batch, id_dict)
                _add_sparse_batch(self.sparse_index, batch_documents, batch_ids, count, self.world_size, self.rank, doc_ids)
                count += len(batch_ids)

        if self.compute_stats:
            stats = {key: value / len(collection_loader) for key, value in stats.items()}
        if self.index_dir is not None:
            self.sparse_index.save()
            pickle.dump(doc_ids, open(os.path.join(self.index_dir, "doc_ids.pkl"), "wb"))
            print("done iterating over the corpus...")
            print("index contains {} posting lists".format(len(self.sparse_index)))
            print("index contains {} documents".format(len(doc_ids)))
            if self.compute_stats:
                with open(os.path.join(self.index_dir, "index_stats.json"), "w") as handler:
                    json.dump(stats, handler)
        else:
            self.sparse_index.finalize()
            out = {"index": self.sparse_index, "ids_mapping": doc_ids}
            if self.compute_stats:
                out["stats"] = stats
            return out


def pack_queries(sparse_query_vecs):
    """list of (col int32[nnz], values fp32[nnz]) (indexer.py:400-401) -> CSR-packed host arrays."""
    n = len(sparse_query_vecs)
    q_offsets = np.zeros(n + 1, dtype=np.int32)
    if n:
        np.cumsum(np.fromiter((len(c) for c, _ in sparse_query_vecs), dtype=np.int64, count=n), out=q_offsets[1:])
    if n and q_offsets[-1]:
        q_terms = np.concatenate([c for c, _ in sparse_query_vecs]).astype(np.int32, copy=False)
        q_weights = np.concatenate([v for _, v in sparse_query_vecs]).astype(np.float32, copy=False)
    else:
        q_terms, q_weights = np.zeros(0, np.int32), np.zeros(0, np.float32)
    return q_offsets, q_terms, q_weights


class SparseRetrieval:
    """retrieval from SparseIndexing (reference indexer.py:311-540), scored on the GPU."""

    @staticmethod
    def select_topk(filtered_indexes, scores, k):
        """Reference semantics (indexer.py:315-322): `scores` are NEGATED; returns the k best rows, scores positive.
        Host-side convenience kept for API parity; the retrieval path selects inside the search kernel."""
        if len(filtered_indexes) > k:
            sorted_ = np.argpartition(scores, k)[:k]
            filtered_indexes, scores = filtered_indexes[sorted_], -scores[sorted_]
        else:
            scores = -scores
        return filtered_indexes, scores

    @staticmethod
    def numba_score_float(inverted_index_ids, inverted_index_floats, indexes_to_retrieve, query_values, threshold, size_collection):
        """The reference's static scorer (indexer.py:324-344) with its signature and return values — `(filtered_indexes
        int64[h], -scores[filtered] fp32[h])` for ONE query over dict-like {term: int32 ids} / {term: fp32 weights} — computed
        by the GPU kernel: the posting lists of the query's terms are uploaded as a small CSR (lists may be in any doc
        order), scored by `b200ret_sparse_scores` with the reference's arithmetic (fp32 multiply, then add, terms in the
        given order) and filtered with the strict `> threshold`.  An API-parity entry point for code that calls the kernel
        directly; the retrieval path scores whole query batches against the HBM-resident index instead."""
        dev = _cuda_device(None)
        terms = [int(t) for t in np.asarray(indexes_to_retrieve).tolist()]
        lists = [(np.asarray(inverted_index_ids[t], dtype=np.int32), np.asarray(inverted_index_floats[t], dtype=np.float32))
                 for t in terms]
        n_local = max(len(terms), 1)
        size_collection = int(size_collection)
        rows = np.concatenate([a for a, _ in lists]) if lists else np.zeros(0, np.int32)
        vals = np.concatenate([v for _, v in lists]) if lists else np.zeros(0, np.float32)
        cols = np.repeat(np.arange(len(lists), dtype=np.int32), [len(a) for a, _ in lists]) if lists else np.zeros(0, np.int32)
        with torch.cuda.device(dev):
            index = ops.SparseDeviceIndex.from_coo(torch.from_numpy(rows).to(dev), torch.from_numpy(cols).to(dev),
                                                   torch.from_numpy(vals).to(dev), n_local, size_collection)
            off = torch.tensor([0, len(terms)], dtype=torch.int32, device=dev)
            q_t = torch.arange(len(terms), dtype=torch.int32, device=dev)     # occurrence j of the query -> local list j
            q_w = torch.from_numpy(np.asarray(query_values, dtype=np.float32)).to(dev)
            scores = ops.sparse_scores(index, off, q_t, q_w)[0]
            filtered = torch.nonzero(scores > threshold).flatten()
            return filtered.cpu().numpy(), (-scores[filtered]).cpu().numpy()

    def __init__(self, model, config, dim_voc, device, dataset_name=None, index_d=None, compute_stats=False, is_beir=False,
                 **kwargs):
        self.model = model
        self.model.eval()
        self.device = device
        self._cuda = _cuda_device(device)
        assert ("index_dir" in config and index_d is None) or ("index_dir" not in config and index_d is not None)
        if "index_dir" in config:
            self.sparse_index = IndexDictOfArray(config["index_dir"], dim_voc=dim_voc, device=self._cuda)
            self.doc_ids = pickle.load(open(os.path.join(config["index_dir"], "doc_ids.pkl"), "rb"))
        else:
            self.sparse_index = index_d["index"]
            self.doc_ids = index_d["ids_mapping"]
            self.sparse_index.fill_missing_terms(dim_voc)   # indexer.py:359-363
        self.dim_voc = dim_voc
        self.size_collection = self.sparse_index.nb_docs()

        # The index moves to HBM once, for the lifetime of the retriever (replaces the numba.typed.Dict copy, :365-370).
        with torch.cuda.device(self._cuda):
            self.sparse_index.device = self._cuda
            self.shard_plan = shard.ShardPlan(self.size_collection, _world_size())
            lo, hi = self.shard_plan.bounds(_rank()) if self.shard_plan.world_size > 1 else (0, self.size_collection)
            # one search-side index normally; several consecutive doc ranges when this rank's share holds >= 2^32 postings
            # kwargs (swallowed like the reference's **kwargs): weight_format="fp16" selects the compressed posting array
            self.device_shards = self.sparse_index.device_shards(lo, hi, kwargs.get("max_shard_postings") or MAX_SHARD_POSTINGS,
                                                                 kwargs.get("weight_format", "fp32"))
            self.device_index, self.doc_id_base = self.device_shards[0]

        self.out_dir = os.path.join(config["out_dir"], dataset_name) if (dataset_name is not None and not is_beir) \
            else config["out_dir"]
        self.doc_stats = index_d["stats"] if (index_d is not None and compute_stats) else None
        self.compute_stats = compute_stats
        if self.compute_stats:
            self.l0 = L0()
        self.model.to(self._cuda)
        self._ext_ids = None
        self._staging = {}

    def _generate_query_vecs(self, q_loader):
        sparse_query_vecs = []
        qids = []
        with torch.inference_mode():
            for t, batch in enumerate(tqdm(q_loader, total=len(q_loader), desc="generate query vecs",
                                           disable=not is_first_worker())):
                inputs = {k: v.to(self._cuda) for k, v in batch.items() if k not in {"ids"}}
                with torch.amp.autocast("cuda", dtype=torch.bfloat16 if supports_bfloat16() else torch.float32):
                    batch_sparse_reps = self.model.encode(**inputs)
                qids.extend(batch["ids"] if isinstance(batch["ids"], list) else to_list(batch["ids"]))
                # one nonzero for the whole batch instead of one per query (indexer.py:394-401): same (col, value) lists
                row, col = torch.nonzero(batch_sparse_reps, as_tuple=True)
                data = batch_sparse_reps[row, col].float().cpu().numpy().astype(np.float32)
                counts = torch.bincount(row, minlength=batch_sparse_reps.shape[0]).cpu().numpy()
                col = col.cpu().numpy().astype(np.int32)
                bounds = np.concatenate([[0], np.cumsum(counts)])
                for i in range(len(counts)):
                    sparse_query_vecs.append((col[bounds[i]:bounds[i + 1]], data[bounds[i]:bounds[i + 1]]))
        return sparse_query_vecs, qids

    # ---- the hot path -----------------------------------------------------------------------------------------
    def search_arrays(self, q_offsets, q_terms, q_weights, topk, threshold=0.0, host_ranks="all"):
        """HOST query arrays -> HOST result arrays (scores fp32 [Q,k], row ids int64 [Q,k], counts int32 [Q]).
        Copies through pinned memory; this is the call bench.py times end to end.  The returned arrays are VIEWS of reusable
        pinned buffers: the next search_arrays call of this object overwrites them (copy what must be kept).  `host_ranks="first"` (sharded search
        only): the merged result goes to the host of the first worker alone — the rank that writes run.json — and the
        other ranks return (None, None, None)."""
        dev = self._cuda
        with torch.cuda.device(dev):
            d_off = self._stage_in("q_off", q_offsets, torch.int32)
            d_terms = self._stage_in("q_terms", q_terms, torch.int32)
            d_w = self._stage_in("q_w", q_weights, torch.float32)
            scores, ids, counts = self._search_local(d_off, d_terms, d_w, int(topk), float(threshold))
            if self.shard_plan.world_size > 1:
                if host_ranks == "first" and self.size_collection < shard.KEY_ID_LIMIT:
                    # every GPU copies its merged query slice into host rows shared with the first worker (own PCIe link each)
                    host = self._shared_rows(scores.shape[0], int(topk))
                    return shard.merge_shards_to_first_host(scores, ids, int(topk), host)[:3]
                scores, ids, counts = shard.merge_shards(scores, ids, int(topk), n_docs_total=self.size_collection)
            if host_ranks == "first" and not is_first_worker():
                torch.cuda.current_stream().synchronize()
                return None, None, None
            out = [self._stage_out(name, t) for name, t in (("scores", scores), ("ids", ids), ("counts", counts))]
            torch.cuda.current_stream().synchronize()
            return tuple(o.numpy() for o in out)

    def _search_local(self, d_off, d_terms, d_w, topk, threshold):
        """This rank's rows: one kernel pass per local doc-range index, merged when there are several.  Sharded over ranks
        (one fp32 index per rank): the shards exchange their bounds between the rounds (shard.TauExchange)."""
        if self.shard_plan.world_size > 1 and self._staging.get("exchange_ok") is None:
            # the exchange is a collective inside the search: every rank must take the same decision (first search only)
            mine = len(self.device_shards) == 1 and self.device_shards[0][0].weight_format == "fp32"
            flag = torch.tensor([int(mine)], dtype=torch.int32, device=self._cuda)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
            self._staging["exchange_ok"] = bool(flag.item())
        if self.shard_plan.world_size > 1 and self._staging["exchange_ok"]:
            if self._staging.get("tau_exchange") is None:
                self._staging["tau_exchange"] = shard.TauExchange("sparse", self.size_collection, self._cuda)
            index, base = self.device_shards[0]
            return ops.sparse_search(index, d_off, d_terms, d_w, topk, threshold, doc_id_base=base,
                                     exchange=self._staging["tau_exchange"])
        parts = [ops.sparse_search(index, d_off, d_terms, d_w, topk, threshold, doc_id_base=base) for index, base in self.device_shards]
        if len(parts) == 1:
            return parts[0]
        return ops.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), topk)

    def _shared_rows(self, n_queries, k):
        key = (n_queries, k)
        if self._staging.get("shared_rows_key") != key:
            self._staging["shared_rows"] = shard.SharedHostRows(n_queries, k)
            self._staging["shared_rows_key"] = key
        return self._staging["shared_rows"]

    def _stage_in(self, name, host_array, dtype):
        return _stage_in(self._staging, self._cuda, name, host_array, dtype)

    def _stage_out(self, name, dev_tensor):
        return _stage_out(self._staging, name, dev_tensor)

    @classmethod
    def from_device_index(cls, device_index, doc_ids=None, doc_id_base=0, size_collection=None, out_dir=None):
        """Wrap an index that is already resident in HBM (bench.py, tests): no encoder, no files."""
        self = cls.__new__(cls)
        self.model = None
        self._cuda = device_index.device
        self.device = self._cuda.index
        self.sparse_index = None
        self.device_index = device_index
        self.doc_id_base = doc_id_base
        self.device_shards = [(device_index, doc_id_base)]
        self.size_collection = device_index.n_docs if size_collection is None else size_collection
        self.shard_plan = shard.ShardPlan(self.size_collection, _world_size())
        self.doc_ids = doc_ids if doc_ids is not None else range(self.size_collection)
        self.dim_voc = device_index.n_terms
        self.out_dir = out_dir
        self.compute_stats = False
        self.doc_stats = None
        self._ext_ids = None
        self._staging = {}
        return self

    def score_float(self, indexes_to_retrieve, query_values, threshold):
        """API analogue of the reference's numba_score_float (indexer.py:324-344) for ONE query, on the GPU:
        returns (filtered_indexes int64[h], -scores[filtered] fp32[h])."""
        dev = self._cuda
        with torch.cuda.device(dev):
            off = torch.tensor([0, len(indexes_to_retrieve)], dtype=torch.int32, device=dev)
            t = torch.as_tensor(np.asarray(indexes_to_retrieve, dtype=np.int32)).to(dev)
            w = torch.as_tensor(np.asarray(query_values, dtype=np.float32)).to(dev)
            scores = ops.sparse_scores(self.device_index, off, t, w)[0]
            filtered = torch.nonzero(scores > threshold).flatten()
            return (filtered + self.doc_id_base).cpu().numpy(), (-scores[filtered]).cpu().numpy()

    def _external_ids(self):
        if self._ext_ids is None:
            self._ext_ids = ExternalIds(self.doc_ids, self.size_collection)
        return self._ext_ids

    def _sparse_retrieve_multithreaded(self, sparse_query_vecs, qids, threshold=0., topk=1000):
        """Same contract as the reference (indexer.py:405-474): returns (res, stats) with
        res[str(qid)][str(doc_ids[row])] = float(score); queries without an eligible doc get no key.  The 4-thread
        numba loop is replaced by one batched GPU search over all queries, and `res` is a read-only Mapping over the
        result arrays (results.LazyRun: inner dicts are built on access, `.to_dict()` gives the eager dict of dicts) —
        the reference spends ~5 s in its per-pair insert loop at 6,980 x 1000 pairs (:429-430)."""
        q_offsets, q_terms, q_weights = pack_queries(sparse_query_vecs)
        # Sharded search (torchrun, world_size > 1): the merged rows are copied to the host of the first worker only — the rank
        # that writes run.json / q_stats.json in retrieve() — and the other ranks return an empty run.
        scores, ids, counts = self.search_arrays(q_offsets, q_terms, q_weights, topk, threshold, host_ranks="first")
        if scores is None:
            res = LazyRun([], np.zeros((0, topk), np.int64), np.zeros((0, topk), np.float32), np.zeros(0, np.int32), self._external_ids())
        else:   # search_arrays hands out views of reusable pinned staging buffers: the run owns copies
            res = LazyRun(qids, owned_copy(ids), owned_copy(scores), np.array(counts), self._external_ids())
        stats = defaultdict(float)
        for n in np.diff(q_offsets).tolist():
            stats["L0_q"] += n / len(qids)
        return res, stats

    def retrieve(self, q_loader, topk, threshold=0.):
        sparse_query_vecs, qids = self._generate_query_vecs(q_loader)
        res, stats = self._sparse_retrieve_multithreaded(sparse_query_vecs, qids, threshold=threshold, topk=topk)
        if is_first_worker():
            if self.compute_stats:
                with open(os.path.join(self.out_dir, "q_stats.json"), "w") as handler:
                    json.dump(stats, handler)
            _write_run(res, os.path.join(self.out_dir, "run.json"))   # native formatter: same bytes as json.dump(res)
        return res


class TermEncoderRetriever:
    """Retrieval over semantic-id codes (reference indexer.py:615-707): every document is a fixed-length list of term ids,
    its score for a query is the sum of the query's predicted term scores at those ids, and the k best documents are kept.
    The gather-sum and the top-k run in one GPU kernel pass per query batch (ops.term_search); the reference's
    pred_scores[:, doc_encodings] intermediate and the [bz, N] score matrix are never built by retrieve()."""

    def __init__(self, model, args):
        self.model = model
        self.model.eval()
        self.args = args

    def _model_device(self):
        if hasattr(self.model, "base_model"):
            return self.model.base_model.device
        if hasattr(self.model, "encoder_decoder"):
            return self.model.encoder_decoder.device
        raise NotImplementedError

    def get_doc_scores(self, pred_scores, doc_encodings):
        """pred_scores [bz, vocab_size], doc_encodings [N, L] -> doc_scores [bz, N] (reference :621-641), full matrix."""
        codes = doc_encodings if doc_encodings.dtype == torch.int32 else doc_encodings.to(torch.int32)
        return ops.term_scores(pred_scores.float().contiguous(), codes.contiguous())

    def retrieve(self, collection_loader, docid_to_smtids, topk, out_dir, use_fp16=False, run_name=None):
        if is_first_worker():
            if not os.path.exists(out_dir):
                os.mkdir(out_dir)
        docids = list(docid_to_smtids.keys())
        code_len = len(next(iter(docid_to_smtids.values()))) if docids else 0
        assert code_len in {16, 32, 64, 128} and all(len(v) == code_len for v in docid_to_smtids.values()), code_len
        print("length of doc_encodings = {}, docids = {}".format(len(docids), len(docids)))
        device = _cuda_device(self._model_device())
        codes = torch.from_numpy(np.asarray(list(docid_to_smtids.values()), dtype=np.int64))
        vocab_checked = None
        codes = codes.to(torch.int32).to(device).contiguous()
        if topk > len(docids):       # torch.topk raises here (reference :688)
            raise RuntimeError(f"selected index k out of range: topk={topk} > {len(docids)} documents")

        ext = ExternalIds(docids)
        all_qids, all_ids, all_scores = [], [], []
        for i, batch in tqdm(enumerate(collection_loader), disable=not is_first_worker(),
                             desc=f"encode # {len(collection_loader)} seqs", total=len(collection_loader)):
            with torch.inference_mode():
                with torch.amp.autocast("cuda", enabled=use_fp16):
                    inputs = {k: v.to(device) for k, v in batch.items() if k != "queries"}
                    batch_preds = self.model.lex_encode(**inputs)     # [bz, vocab_size]
                    if isinstance(batch_preds, tuple):
                        assert len(batch_preds) == 2 and batch_preds[1] is None, batch_preds
                        batch_preds = batch_preds[0]
                    elif not isinstance(batch_preds, torch.Tensor):
                        raise NotImplementedError
                pred = batch_preds.float().contiguous()
                if vocab_checked != pred.shape[1]:                    # the kernel trusts the codes: range-check once
                    assert codes.numel() == 0 or (int(codes.min()) >= 0 and int(codes.max()) < pred.shape[1]), "doc code out of vocabulary"
                    vocab_checked = pred.shape[1]
                top_scores, top_idxes, _ = ops.term_search(pred, codes, int(topk))
            if not isinstance(batch["queries"], list):
                raise ValueError("query_ids with type {} is not valid".format(type(batch["queries"])))
            all_qids.extend(batch["queries"])
            all_ids.append(top_idxes.cpu().numpy())
            all_scores.append(top_scores.cpu().numpy())

        k = int(topk)
        ids = np.concatenate(all_ids) if all_ids else np.zeros((0, k), np.int64)
        scores = np.concatenate(all_scores) if all_scores else np.zeros((0, k), np.float32)
        # {qid: {docid: score}} (:690-696): a repeated qid REPLACES its earlier entry there (`qid_to_rankdata[qid] = {}`) while
        # the key keeps its first-seen position — exactly what this dict comprehension gives
        last = {qid: pos for pos, qid in enumerate(all_qids)}
        if len(last) != len(all_qids):
            keep = list(last.values())
            all_qids, ids, scores = list(last.keys()), ids[keep], scores[keep]
        qid_to_rankdata = LazyRun(all_qids, ids, scores, None, ext)
        _write_run(qid_to_rankdata, os.path.join(out_dir, "run.json" if run_name is None else run_name))
        return qid_to_rankdata


def _write_run(res, path):
    """run.json (reference indexer.py:537-538): formatted from the result arrays when `res` is a LazyRun."""
    if isinstance(res, LazyRun):
        return res.write_json(path)
    with open(path, "w") as handler:
        handler.write(json.dumps(res))
    return "python"


# ------------------------------------------------------------------------------------------------------------------
# hybrid: one encoder pass -> sparse + dense (reference indexer.py:710-1019), on top of the same two engines
# ------------------------------------------------------------------------------------------------------------------

class HybridIndexer:
    """`model.encode` returns (sparse [bz, V], dense [bz, d]); the sparse half goes to the GPU-resident COO log / CSR build of
    SparseIndexer, the dense half to the embs_/ids_/plan.json shard files of store_embs (reference indexer.py:710-856)."""

    def __init__(self, model, sparse_index_dir, dense_index_dir, device, chunk_size=2_000_000, compute_stats=False,
                 dim_voc=None, force_new=True, filename="array_index.h5py", keep_on_device=False, **kwargs):
        self.keep_on_device = keep_on_device   # also keep the dense chunks in HBM as bf16 (DEVICE_EMBEDDINGS)
        self.model = model
        self.model.eval()
        self.sparse_index_dir = sparse_index_dir
        self.dense_index_dir = dense_index_dir
        self.chunk_size = chunk_size
        self.device = device
        self._cuda = _cuda_device(device)
        self.sparse_index = IndexDictOfArray(self.sparse_index_dir, dim_voc=dim_voc, force_new=force_new, filename=filename,
                                             device=self._cuda)
        self.compute_stats = compute_stats
        if self.compute_stats:
            self.l0 = L0()
        self.model.to(self._cuda)
        self.local_rank = self.device
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            assert self.local_rank == torch.distributed.get_rank(), (self.local_rank, torch.distributed.get_rank())
        self.rank = _rank()
        self.world_size = _world_size()
        print("world_size: {}, local_rank: {}".format(self.world_size, self.local_rank))

    def index(self, collection_loader, id_dict=None):
        dtype = torch.bfloat16 if supports_bfloat16() else torch.float32
        print("Using bfloat16" if dtype == torch.bfloat16 else "Using float32")
        doc_ids = {}
        stats = defaultdict(float)
        count = 0
        embeddings, embeddings_ids, chunk_idx = [], [], 0
        staging = {}
        write_freq = self.chunk_size // collection_loader.batch_size
        with torch.inference_mode():
            for idx, batch in enumerate(tqdm(collection_loader, disable=not is_first_worker())):
                inputs = {k: v.to(self._cuda) for k, v in batch.items() if k not in {"ids"}}
                with torch.amp.autocast("cuda", dtype=dtype):
                    batch_sparse_reps, batch_dense_reps = self.model.encode(**inputs)
                batch_ids = _batch_ids(batch, id_dict)
                if self.compute_stats:
                    stats["L0_d"] += self.l0(batch_sparse_reps).item()
                _add_sparse_batch(self.sparse_index, batch_sparse_reps, batch_ids, count, self.world_size, self.rank, doc_ids,
                                  require_all=True)
                count += len(batch_ids)
                embeddings.append(batch_dense_reps.float())
                embeddings_ids.extend(batch_ids)
                if (idx + 1) % write_freq == 0:
                    _write_emb_chunk(self.dense_index_dir, self.local_rank, chunk_idx, embeddings, embeddings_ids, force_int64=True,
                                     keep_on_device=self.keep_on_device, staging=staging)
                    embeddings, embeddings_ids = [], []
                    chunk_idx += 1
        if len(embeddings) != 0:
            shape = _write_emb_chunk(self.dense_index_dir, self.local_rank, chunk_idx, embeddings, embeddings_ids, force_int64=True,
                                     keep_on_device=self.keep_on_device, staging=staging)
            print("last embedddings shape = {}".format(tuple(shape)))
            chunk_idx += 1
        plan = {"nranks": _world_size(), "num_chunks": chunk_idx, "index_path": os.path.join(self.dense_index_dir, "model.index")}
        print("plan: ", plan)
        if is_first_worker():
            with open(os.path.join(self.dense_index_dir, "plan.json"), "w") as fout:
                json.dump(plan, fout)

        if self.compute_stats:
            stats = {key: value / len(collection_loader) for key, value in stats.items()}
        if self.sparse_index_dir is not None:
            self.sparse_index.save()
            pickle.dump(doc_ids, open(os.path.join(self.sparse_index_dir, "doc_ids.pkl"), "wb"))
            print("done iterating over the corpus...")
            print("index contains {} posting lists".format(len(self.sparse_index)))
            print("index contains {} documents".format(len(doc_ids)))
            if self.compute_stats:
                with open(os.path.join(self.sparse_index_dir, "index_stats.json"), "w") as handler:
                    json.dump(stats, handler)
        else:
            self.sparse_index.finalize()
            out = {"index": self.sparse_index, "ids_mapping": doc_ids}
            if self.compute_stats:
                out["stats"] = stats
            return out


class HybridRetriever:
    """One encoder pass per query batch -> sparse run + dense run (reference indexer.py:859-1019).  Both searches run on the
    GPU engines: the sparse index through SparseRetrieval, the dense shards through DenseFlatIndexer."""

    select_topk = staticmethod(SparseRetrieval.select_topk)

    def __init__(self, model, sparse_index_dir, dense_index_dir, out_dir, dim_voc, device, **kwargs):
        self.model = model
        self.model.eval()
        self.sparse_out_dir = os.path.join(out_dir, "sparse")
        self.dense_out_dir = os.path.join(out_dir, "dense")
        if is_first_worker():
            os.makedirs(self.sparse_out_dir, exist_ok=True)
            os.makedirs(self.dense_out_dir, exist_ok=True)
        self.sparse = SparseRetrieval(model, {"index_dir": sparse_index_dir, "out_dir": self.sparse_out_dir}, dim_voc, device)
        self.sparse_index = self.sparse.sparse_index
        self.doc_ids = self.sparse.doc_ids
        self.l0 = L0()
        self.device = device
        self._cuda = _cuda_device(device)
        self.dense_index = DenseFlatIndexer(device=self._cuda)
        self.dense_index.init_index(self.model.hidden_size)
        doc_vector_files, doc_id_files = obtain_doc_vec_dir_files(dense_index_dir)
        self._index_encoded_data(doc_vector_files, doc_id_files)
        self.model.to(self._cuda)

    def _generate_query_vecs(self, q_loader):
        sparse_query_vecs, dense_query_vecs, qids = [], [], []
        with torch.inference_mode():
            for t, batch in enumerate(tqdm(q_loader, total=len(q_loader), desc="generate query vecs",
                                           disable=not is_first_worker())):
                inputs = {k: v.to(self._cuda) for k, v in batch.items() if k not in {"ids"}}
                with torch.amp.autocast("cuda", dtype=torch.bfloat16 if supports_bfloat16() else torch.float32):
                    batch_sparse_reps, batch_dense_reps = self.model.encode(**inputs)
                qids.extend(batch["ids"] if isinstance(batch["ids"], list) else to_list(batch["ids"]))
                dense_query_vecs.append(batch_dense_reps.float().cpu().numpy())
                row, col = torch.nonzero(batch_sparse_reps, as_tuple=True)
                data = batch_sparse_reps[row, col].float().cpu().numpy().astype(np.float32)
                counts = torch.bincount(row, minlength=batch_sparse_reps.shape[0]).cpu().numpy()
                col = col.cpu().numpy().astype(np.int32)
                bounds = np.concatenate([[0], np.cumsum(counts)])
                for i in range(len(counts)):
                    sparse_query_vecs.append((col[bounds[i]:bounds[i + 1]], data[bounds[i]:bounds[i + 1]]))
        dense_query_vecs = np.concatenate(dense_query_vecs, axis=0)
        assert len(sparse_query_vecs) == len(dense_query_vecs) == len(qids), (len(sparse_query_vecs), len(dense_query_vecs), len(qids))
        return sparse_query_vecs, dense_query_vecs, qids

    def _index_encoded_data(self, doc_vec_files, doc_id_files):
        if all(os.path.abspath(f) in DEVICE_EMBEDDINGS for f in doc_vec_files):   # encoded in this job: bf16 chunks are in HBM
            doc_reps = torch.cat([DEVICE_EMBEDDINGS[os.path.abspath(f)] for f in doc_vec_files])
        else:
            doc_reps = np.concatenate([np.load(f) for f in doc_vec_files], axis=0)
        doc_ids = np.concatenate([np.load(f) for f in doc_id_files]).tolist()
        assert len(doc_reps) == len(doc_ids), (len(doc_reps), len(doc_ids))
        print("size of doc reps to index: ", tuple(doc_reps.shape))
        self.dense_index.index_data(doc_reps, doc_ids)
        print("finished indexing")

    def _dense_retrieve(self, query_reps, qids, topk=1000):
        """reference indexer.py:973-982; the run is a LazyRun over the [Q, k] arrays (rows padded with label -1 map to the
        last id like the reference's list indexing, :212)."""
        assert len(qids) == len(query_reps), (len(qids), len(query_reps))
        scores, indexes = self.dense_index.search_arrays(query_reps, topk, host_ranks="first")
        if scores is None:     # sharded search: the run lives on the first worker (the rank that writes dense/run.json)
            return LazyRun([], np.zeros((0, topk), np.int64), np.zeros((0, topk), np.float32), None, self.dense_index.external_ids())
        return LazyRun(qids, owned_copy(indexes), owned_copy(scores), None, self.dense_index.external_ids())

    def _sparse_retrieve(self, sparse_query_vecs, qids, threshold=0., topk=1000):
        return self.sparse._sparse_retrieve_multithreaded(sparse_query_vecs, qids, threshold=threshold, topk=topk)

    def retrieve(self, q_loader, topk, id_dict=False, threshold=0.):
        """Writes sparse/q_stats.json, sparse/run.json and dense/run.json like the reference; additionally returns the two
        result dicts (the reference returns None)."""
        sparse_query_vecs, dense_query_vecs, qids = self._generate_query_vecs(q_loader)
        sparse_res, sparse_stats = self._sparse_retrieve(sparse_query_vecs, qids, threshold=threshold, topk=topk)
        dense_res = self._dense_retrieve(dense_query_vecs, qids, topk=topk)
        if is_first_worker():
            with open(os.path.join(self.sparse_out_dir, "q_stats.json"), "w") as handler:
                json.dump(sparse_stats, handler)
            _write_run(sparse_res, os.path.join(self.sparse_out_dir, "run.json"))
            _write_run(dense_res, os.path.join(self.dense_out_dir, "run.json"))
        return sparse_res, dense_res
