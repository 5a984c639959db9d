"""GPU-backed inverted-index container with the reference's IndexDictOfArray interface.

Mirrors scaling_retriever/utils/inverted_index.py of the reference (class IndexDictOfArray :15-105, merge_indexes
:108-170): same constructor arguments, attributes (`index_doc_id`, `index_doc_value`, `n`), methods
(`add_batch_document`, `__len__`, `nb_docs`, `save`) and on-disk artefacts.  What changes is where the work is done:

* add_batch_document (reference: one CPython list append per posting, :74-76) only logs the COO batch — on the
  GPU if the batch arrives as CUDA tensors, so SparseIndexer.index never synchronises per batch;
* the per-term arrays are produced by ONE stable GPU radix sort of the log (ops.csr_build), bit-exact with what the
  appends would have built; `index_doc_id[t]` / `index_doc_value[t]` are numpy views into that CSR;
* save() writes the reference's HDF5 layout (`dim`, `index_doc_id_{t}`, `index_doc_value_{t}`; through h5py when it is
  installed, else through the built-in writer hdf5_lite.py) and a native CSR bundle (three .npy files) next to it; the
  loader reads either (the newer one).

There is no CPU build path: finalising an index without a CUDA device raises.
"""
import json
import os
import pickle
from collections.abc import Mapping

import numpy as np
import torch

from . import ops

CSR_FILES = ("csr_term_offsets.npy", "csr_doc_ids.npy", "csr_weights.npy")
MAX_SHARD_POSTINGS = (1 << 32) - (1 << 20)     # the kernels address postings with 32-bit positions (csr_build.cu, skip table)


class _TermArrays(Mapping):
    """dict[int -> ndarray] view over CSR arrays (what the reference stores as a real dict of arrays)."""

    def __init__(self, term_offsets, data, keys):
        self._off = term_offsets
        self._data = data
        self._keys = keys            # term ids exposed as dict keys (ascending)
        self._keyset = None

    def __getitem__(self, key):
        key = int(key)
        if key < 0 or key >= len(self._off) - 1 or (self._keyset is not None and key not in self._keyset):
            raise KeyError(key)
        return self._data[self._off[key]:self._off[key + 1]]

    def __contains__(self, key):
        if self._keyset is None:
            self._keyset = set(int(k) for k in self._keys)
        try:
            return int(key) in self._keyset
        except (TypeError, ValueError):
            return False

    def __iter__(self):
        return iter(int(k) for k in self._keys)

    def __len__(self):
        return len(self._keys)


def _as_device_i32(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x).astype(np.int32, copy=False)).to(device, non_blocking=True)


def _as_device_f32(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x).astype(np.float32, copy=False)).to(device, non_blocking=True)


class IndexDictOfArray:
    def __init__(self, index_path=None, force_new=False, filename="array_index.h5py", dim_voc=None, device=None):
        self.dim_voc = dim_voc
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None and torch.cuda.is_available() \
            else (torch.device(device) if device is not None else None)
        self._log = []              # COO batches (rows, cols, vals) as device tensors, in feed order
        self._csr_host = None       # (term_offsets, doc_ids, weights) numpy
        self._csr_dev = None        # same as CUDA tensors
        self._views = None
        self._all_keys = False      # loaded indexes expose every term id in range(dim) like the reference loader
        self.n = 0
        if index_path is not None:
            self.index_path = index_path
            if not os.path.exists(index_path):
                os.makedirs(index_path)
            self.filename = os.path.join(self.index_path, filename)
            if self._exists_on_disk() and not force_new:
                print("index already exists, loading...")
                self._load(dim_voc)
                print("done loading index...")
                doc_ids = pickle.load(open(os.path.join(self.index_path, "doc_ids.pkl"), "rb"))
                if isinstance(doc_ids, list):
                    self.n = len(doc_ids)
                else:   # dict row id -> external id; the reference takes max key + 1 (inverted_index.py:47-55)
                    keys = np.fromiter(doc_ids.keys(), dtype=np.int64, count=len(doc_ids))
                    print("min_val: ", keys.min(), "max_val: ", keys.max())
                    assert keys.min() == 0, keys.min()
                    self.n = int(keys.max()) + 1
            else:
                print("initializing new index...")
        else:
            print("initializing new index...")

    # ---- disk -----------------------------------------------------------------------------------------------
    def _csr_paths(self):
        return [os.path.join(self.index_path, f) for f in CSR_FILES]

    def _exists_on_disk(self):
        return all(os.path.exists(p) for p in self._csr_paths()) or os.path.exists(self.filename)

    def _load(self, dim_voc):
        self._set_csr_host(*read_index_dir(self.index_path, os.path.basename(self.filename), dim_voc))
        self._all_keys = True

    def _set_csr_host(self, off, ids, w):
        self._csr_host = (off, ids, w)
        self._csr_dev = None
        self._views = None
        self.dim_voc = len(off) - 1

    # ---- build ----------------------------------------------------------------------------------------------
    def add_batch_document(self, row, col, data, n_docs=-1):
        """add a batch of documents to the index (reference inverted_index.py:67-76): logged, sorted later."""
        if self._csr_host is not None or self._csr_dev is not None:
            raise RuntimeError("add_batch_document after the index was finalised/loaded is not supported")
        if self.device is None or self.device.type != "cuda":
            raise RuntimeError("IndexDictOfArray needs a CUDA device to build an index (no CPU fallback)")
        r = _as_device_i32(row, self.device)
        c = _as_device_i32(col, self.device)
        v = _as_device_f32(data, self.device)
        if n_docs < 0:
            self.n += int(torch.unique(r).numel())
        else:
            self.n += n_docs
        self._log.append((r, c, v))

    def _n_terms_for_build(self):
        if self.dim_voc is not None:
            return int(self.dim_voc)
        mx = max((int(c.max().item()) for _, c, _ in self._log if c.numel()), default=-1)
        return mx + 1 if mx >= 0 else 1

    def finalize(self):
        """Run the GPU CSR build over the logged batches (idempotent). Returns device (term_offsets, doc_ids, weights)."""
        if self._csr_dev is not None:
            return self._csr_dev
        if self._csr_host is not None:
            dev = self.device if self.device is not None else torch.device("cuda", torch.cuda.current_device())
            self._csr_dev = tuple(torch.as_tensor(a).to(dev) for a in self._csr_host)
            return self._csr_dev
        if self.device is None or self.device.type != "cuda":
            raise RuntimeError("IndexDictOfArray needs a CUDA device to build an index (no CPU fallback)")
        n_terms = self._n_terms_for_build()
        if self._log:
            rows = torch.cat([r for r, _, _ in self._log])
            cols = torch.cat([c for _, c, _ in self._log])
            vals = torch.cat([v for _, _, v in self._log])
        else:
            rows = torch.empty(0, dtype=torch.int32, device=self.device)
            cols = rows.clone()
            vals = torch.empty(0, dtype=torch.float32, device=self.device)
        self._log = []
        if rows.numel():   # the kernels trust term and row ids: one cheap device reduction each before the sort
            c_lo, c_hi, r_lo = int(cols.min().item()), int(cols.max().item()), int(rows.min().item())
            if c_lo < 0 or c_hi >= n_terms or r_lo < 0:
                raise ValueError(f"add_batch_document: term ids must be in [0, {n_terms}) and rows >= 0 "
                                 f"(got cols in [{c_lo}, {c_hi}], min row {r_lo})")
        n_docs = max(self.n, int(rows.max().item()) + 1 if rows.numel() else 0)
        self._csr_dev = ops.csr_build(rows, cols, vals, n_terms, n_docs, sort_docs=False)
        self.dim_voc = n_terms
        return self._csr_dev

    def csr_host(self):
        if self._csr_host is None:
            off, ids, w = self.finalize()
            self._csr_host = (off.cpu().numpy(), ids.cpu().numpy(), w.cpu().numpy())
        return self._csr_host

    def _make_views(self):
        if self._views is None:
            off, ids, w = self.csr_host()
            keys = np.arange(len(off) - 1) if self._all_keys else np.nonzero(np.diff(off))[0]
            self._views = (_TermArrays(off, ids, keys), _TermArrays(off, w, keys))
        return self._views

    @property
    def index_doc_id(self):
        return self._make_views()[0]

    @property
    def index_doc_value(self):
        return self._make_views()[1]

    def fill_missing_terms(self, dim_voc):
        """SparseRetrieval.__init__ fills absent posting lists with empty arrays (indexer.py:359-363)."""
        off, ids, w = self.csr_host()
        if dim_voc != len(off) - 1:
            self._set_csr_host(_resize_offsets(off, dim_voc), ids, w)
        self._all_keys = True
        self._views = None

    def __len__(self):
        return len(self.index_doc_id)

    def nb_docs(self):
        return self.n

    def save(self, dim=None):
        print("converting to numpy")
        off, ids, w = self.csr_host()
        print("save to disk")
        print("filename: ", self.filename)
        keys = np.nonzero(np.diff(off))[0]
        # the reference's HDF5 layout (inverted_index.py:92-100): scalar `dim` + one dataset pair per non-empty posting list
        n_dim = int(dim) if dim else len(keys)
        try:
            import h5py
        except ImportError:
            h5py = None
        if h5py is not None:
            with h5py.File(self.filename, "w") as f:
                f.create_dataset("dim", data=n_dim)
                for key in keys:
                    f.create_dataset("index_doc_id_{}".format(key), data=ids[off[key]:off[key + 1]])
                    f.create_dataset("index_doc_value_{}".format(key), data=w[off[key]:off[key + 1]])
        else:   # no h5py in this environment: the same file structure from the built-in writer (hdf5_lite.py)
            from . import hdf5_lite
            datasets = {"dim": np.int64(n_dim)}
            for key in keys:
                datasets["index_doc_id_{}".format(key)] = ids[off[key]:off[key + 1]]
                datasets["index_doc_value_{}".format(key)] = w[off[key]:off[key + 1]]
            hdf5_lite.write_file(self.filename, datasets)
        for path, arr in zip(self._csr_paths(), (off, ids, w)):     # written last: the loader prefers the newer artefact
            np.save(path, arr)
        print("saving index distribution...")
        index_dist = {int(k): int(off[k + 1] - off[k]) for k in keys}
        json.dump(index_dist, open(os.path.join(self.index_path, "index_dist.json"), "w"))

    # ---- search-side view -----------------------------------------------------------------------------------
    def device_shards(self, doc_lo=0, doc_hi=None, max_postings=MAX_SHARD_POSTINGS, weight_format="fp32"):
        """[(SparseDeviceIndex, first doc row)] covering doc rows [doc_lo, doc_hi): ONE entry normally; a range that holds more
        postings than the kernels' 32-bit positions address (>= 2^32, e.g. 20 M docs x 200 terms on one GPU) is cut into
        consecutive doc ranges that are searched one after the other and merged (SparseRetrieval.search_arrays)."""
        doc_hi = int(self.n) if doc_hi is None else doc_hi
        nnz = len(self._csr_host[1]) if self._csr_host is not None else (self._csr_dev[1].numel() if self._csr_dev is not None else
                                                                          sum(r.numel() for r, _, _ in self._log))
        if nnz <= max_postings:
            return [(self.device_index(doc_lo, doc_hi, weight_format), doc_lo)]
        if self._csr_host is None:
            raise NotImplementedError(f"an index of {nnz} postings built in memory exceeds the 32-bit posting positions of one "
                                      "CSR build; save and load it (the loader cuts it into doc ranges), or index under torchrun")
        off, ids, _ = self._csr_host
        return [(self.device_index(a, b, weight_format), a) for a, b in split_doc_range(off, ids, doc_lo, doc_hi, max_postings)]

    def device_index(self, doc_lo=0, doc_hi=None, weight_format="fp32"):
        """Search-side index on the GPU for doc rows [doc_lo, doc_hi) (default: all; row ids become local to the range):
        doc-sorted CSR + doc-block skip table, slices in bank order.  The canonical CSR of this object is not modified."""
        n_docs = int(self.n)
        sharded = doc_lo != 0 or (doc_hi is not None and doc_hi != n_docs)
        if sharded and self._csr_dev is None and self._csr_host is not None:
            # loaded from disk: cut the shard out on the HOST and upload only that (every rank used to upload the whole CSR,
            # 14 GB at 8.8 M docs, before slicing it on the device)
            dev = self.device if self.device is not None else torch.device("cuda", torch.cuda.current_device())
            doc_hi = n_docs if doc_hi is None else doc_hi
            off, ids, w = (torch.as_tensor(a).to(dev) for a in shard_csr_host(*self._csr_host, doc_lo, doc_hi))
            n_docs = doc_hi - doc_lo
        else:
            off, ids, w = self.finalize()
            if sharded:
                from . import shard
                doc_hi = n_docs if doc_hi is None else doc_hi
                off, ids, w = shard.shard_sparse_csr(off, ids, w, doc_lo, doc_hi)
                n_docs = doc_hi - doc_lo
        try:
            return ops.SparseDeviceIndex.from_csr(off, ids, w, n_docs, weight_format=weight_format)
        except Exception as exc:   # lists not ascending (merged multi-rank index): re-sort by (term, doc) on the GPU
            from ._lib import B200RetError
            if not (isinstance(exc, B200RetError) and exc.code == -4):
                raise
        counts = (off[1:] - off[:-1])
        cols = torch.repeat_interleave(torch.arange(off.numel() - 1, dtype=torch.int32, device=off.device), counts,
                                       output_size=ids.numel())
        return ops.SparseDeviceIndex.from_coo(ids, cols, w, off.numel() - 1, n_docs, weight_format=weight_format)


def split_doc_range(off, ids, lo, hi, max_postings):
    """Cut doc rows [lo, hi) into consecutive ranges holding <= max_postings postings each (host CSR; one bincount pass)."""
    if hi <= lo:
        return [(lo, hi)]
    per_doc = np.zeros(hi - lo, dtype=np.int64)
    step = 1 << 26
    for a in range(0, len(ids), step):
        seg = ids[a:a + step]
        seg = seg[(seg >= lo) & (seg < hi)] - lo
        per_doc += np.bincount(seg, minlength=hi - lo)
    csum = np.cumsum(per_doc)
    bounds, start, base = [], lo, 0
    while start < hi:
        end = lo + int(np.searchsorted(csum, base + max_postings, side="right"))
        end = max(end, start + 1)
        end = min(end, hi)
        bounds.append((start, end))
        base = int(csum[end - lo - 1])
        start = end
    return bounds


def shard_csr_host(off, ids, w, lo, hi, terms_per_chunk=2048):
    """Host CSR restricted to doc rows [lo, hi) with LOCAL row ids, in bounded-memory chunks of terms (numpy only)."""
    n_terms = len(off) - 1
    new_off = np.zeros(n_terms + 1, dtype=np.int64)
    out_ids, out_w = [], []
    for t0 in range(0, n_terms, terms_per_chunk):
        t1 = min(n_terms, t0 + terms_per_chunk)
        a, b = int(off[t0]), int(off[t1])
        seg = ids[a:b]
        keep = (seg >= lo) & (seg < hi)
        csum = np.concatenate([[0], np.cumsum(keep, dtype=np.int64)])
        new_off[t0 + 1:t1 + 1] = new_off[t0] + csum[off[t0 + 1:t1 + 1] - a]
        out_ids.append((seg[keep] - lo).astype(np.int32))
        out_w.append(w[a:b][keep])
    ids_out = np.concatenate(out_ids) if out_ids else np.zeros(0, np.int32)
    w_out = np.concatenate(out_w) if out_w else np.zeros(0, np.float32)
    return new_off, ids_out, w_out.astype(np.float32, copy=False)


def _resize_offsets(off, dim_voc):
    n = len(off) - 1
    if dim_voc > n:
        return np.concatenate([off, np.full(dim_voc - n, off[-1], dtype=off.dtype)])
    if off[dim_voc] != off[-1]:
        raise ValueError(f"dim_voc={dim_voc} would drop non-empty posting lists (index has {n} terms)")
    return off[:dim_voc + 1].copy()


def _load_hdf5(filename, dim_voc):
    """The reference loader (inverted_index.py:24-41): iterate range(dim), missing datasets -> empty lists.  Read with h5py
    when it is installed, else with the built-in reader of the file structure h5py writes by default (hdf5_lite.py)."""
    try:
        import h5py
        opener = lambda: h5py.File(filename, "r")   # noqa: E731
    except ImportError:
        from . import hdf5_lite
        opener = lambda: hdf5_lite.File(filename)   # noqa: E731
    with opener() as f:
        dim = dim_voc if dim_voc is not None else int(f["dim"][()])
        off = np.zeros(dim + 1, dtype=np.int64)
        ids, vals = [], []
        for key in range(dim):
            name = "index_doc_id_{}".format(key)
            if name in f:
                a = np.array(f[name], dtype=np.int32)
                v = np.array(f["index_doc_value_{}".format(key)], dtype=np.float32)
                ids.append(a)
                vals.append(v)
                off[key + 1] = off[key] + len(a)
            else:
                off[key + 1] = off[key]
    ids = np.concatenate(ids) if ids else np.array([], dtype=np.int32)
    vals = np.concatenate(vals) if vals else np.array([], dtype=np.float32)
    return off, ids, vals


def read_index_dir(index_path, filename="array_index.h5py", dim_voc=None):
    """Posting lists of an index directory as host CSR arrays (term_offsets int64[V+1], doc_ids int32, weights fp32): the
    native csr_*.npy bundle when present, else the reference's HDF5 file (needs h5py).  No doc_ids.pkl involved."""
    paths = [os.path.join(index_path, f) for f in CSR_FILES]
    h5_path = os.path.join(index_path, filename)
    have_csr = all(os.path.exists(p) for p in paths)
    if have_csr and os.path.exists(h5_path) and os.path.getmtime(h5_path) > max(os.path.getmtime(p) for p in paths) + 1.0:
        have_csr = False     # the directory was re-indexed by the reference (HDF5 only) after the bundle was written: bundle is stale
    if have_csr:
        off, ids, w = (np.load(p) for p in paths)
        if dim_voc is not None and dim_voc != len(off) - 1:   # the reference trusts dim_voc (inverted_index.py:25-26)
            off = _resize_offsets(off, dim_voc)
    else:
        off, ids, w = _load_hdf5(os.path.join(index_path, filename), dim_voc)
    return off.astype(np.int64), ids.astype(np.int32), w.astype(np.float32)


def convert_hdf5_to_csr(index_path, filename="array_index.h5py", dim_voc=None):
    """One-off converter: read the reference's HDF5 index file of `index_path` and write the native CSR bundle
    next to it, so later loads are three np.load calls instead of 2 x dim_voc dataset reads.  Returns the bundle paths."""
    off, ids, w = _load_hdf5(os.path.join(index_path, filename), dim_voc)
    paths = [os.path.join(index_path, f) for f in CSR_FILES]
    for path, arr in zip(paths, (off.astype(np.int64), ids.astype(np.int32), w.astype(np.float32))):
        np.save(path, arr)
    return paths


def merge_indexes(model_name_or_path, filename="array_index.h5py", index_name="index", index_dir=None):
    """Merge per-rank index dirs `index_0`, `index_1`, ... into `index` (reference inverted_index.py:108-170):
    per term, the posting arrays of the shards are appended in directory order; doc_ids maps are unioned; L0_d is
    averaged; index_dist.json is dict.update()d shard after shard exactly like the reference (:149-150), i.e. a term's
    entry is its count in the LAST shard that holds it, not the sum.  Here the append is one stable GPU sort of the concatenated shard postings by term."""
    with open(os.path.join(model_name_or_path, "config.json")) as fin:
        config = json.load(fin)
    dim_voc = config["vocab_size"]
    print("dim_voc: ", dim_voc)
    root = index_dir if index_dir is not None else model_name_or_path
    # the reference takes os.listdir order as it comes (the order of the shards inside a posting list follows it);
    # sorted here so that the merged index is deterministic: index_0, index_1, ...
    index_dirs = [os.path.join(root, d) for d in sorted(os.listdir(root)) if d.startswith(index_name) and d != index_name]
    assert len(index_dirs) in [1, 2, 4], index_dirs
    if len(index_dirs) == 1:
        print("only one index, no need to merge")
        return
    rows, cols, vals = [], [], []
    doc_ids, index_dist, index_stats = dict(), {}, {"L0_d": 0}
    for idx_dir in index_dirs:
        off, ids, w = read_index_dir(idx_dir, filename, dim_voc)   # like the reference: the posting file only (a shard's
        #                                                             doc_ids.pkl need not start at row 0)
        cols.append(np.repeat(np.arange(dim_voc, dtype=np.int32), np.diff(off)))
        rows.append(ids)
        vals.append(w)
        with open(os.path.join(idx_dir, "doc_ids.pkl"), "rb") as f:
            doc_ids.update(pickle.load(f))
        with open(os.path.join(idx_dir, "index_dist.json"), "r") as f:
            index_dist.update(json.load(f))
        with open(os.path.join(idx_dir, "index_stats.json"), "r") as f:
            index_stats["L0_d"] += json.load(f)["L0_d"] / len(index_dirs)
    out_index_dir = os.path.join(root, index_name)
    merged = IndexDictOfArray(out_index_dir, force_new=True, filename=filename, dim_voc=dim_voc)
    merged.add_batch_document(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals), n_docs=len(doc_ids))
    merged.save(dim=dim_voc)
    with open(os.path.join(out_index_dir, "doc_ids.pkl"), "wb") as f:
        pickle.dump(doc_ids, f)
    with open(os.path.join(out_index_dir, "index_dist.json"), "w") as f:
        json.dump(index_dist, f)
    with open(os.path.join(out_index_dir, "index_stats.json"), "w") as f:
        json.dump(index_stats, f)


if __name__ == "__main__":
    import argparse
    parser = argparse.ArgumentParser()
    parser.add_argument("--model_name_or_path", type=str, default=None)
    parser.add_argument("--index_name", default="index", type=str)
    parser.add_argument("--index_dir", default=None, type=str)
    parser.add_argument("--convert_hdf5", default=None, type=str,
                        help="(extension) index directory whose array_index.h5py is converted to the native CSR bundle; no merge")
    args = parser.parse_args()
    if args.convert_hdf5:
        print(convert_hdf5_to_csr(args.convert_hdf5))
    else:
        merge_indexes(args.model_name_or_path, index_name=args.index_name, index_dir=args.index_dir)
