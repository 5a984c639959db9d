"""The reference's HDF5 index layout (utils/inverted_index.py:24-41 load, :92-100 save) through a FAKE h5py module.

h5py / libhdf5 are not installable in this image, so `IndexDictOfArray.save`'s HDF5 branch and `_load_hdf5` would otherwise
never execute.  The fake below implements exactly the h5py surface those two code paths (and the reference's) use —
File(name, mode) as a context manager, create_dataset(name, data=...), `name in f`, f[name][()], np.array(f[name]) — on top
of an .npz container, and the test drives the product code AND a verbatim restatement of the reference's loader over the
same file: the datasets written are the ones the reference reads (names, dtypes, `dim` semantics)."""
import json
import os
import sys
import types

import numpy as np
import pytest


class _FakeDataset:
    def __init__(self, arr):
        self._a = np.asarray(arr)

    def __getitem__(self, key):
        return self._a[key]

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    def __len__(self):
        return len(self._a)

    @property
    def dtype(self):
        return self._a.dtype


class _FakeFile:
    def __init__(self, name, mode="r"):
        self.name, self.mode, self.data = name, mode, {}
        if mode == "r":
            with np.load(name + ".fake.npz") as z:
                self.data = {k: z[k] for k in z.files}

    def create_dataset(self, name, data=None):
        self.data[name] = np.asarray(data)

    def __contains__(self, name):
        return name in self.data

    def __getitem__(self, name):
        return _FakeDataset(self.data[name])

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if self.mode == "w":
            np.savez(self.name + ".fake.npz", **self.data)
            open(self.name, "wb").write(b"\x89HDF\r\n\x1a\n fake")     # the reference checks os.path.exists(filename)


@pytest.fixture
def fake_h5py(monkeypatch):
    mod = types.ModuleType("h5py")
    mod.File = _FakeFile
    monkeypatch.setitem(sys.modules, "h5py", mod)
    return mod


def reference_loader(filename, dim_voc):
    """utils/inverted_index.py:24-41 restated line for line (h5py calls included) -> dict of arrays."""
    import h5py
    file = h5py.File(filename, "r")
    dim = dim_voc if dim_voc is not None else file["dim"][()]
    index_doc_id, index_doc_value = dict(), dict()
    for key in range(dim):
        try:
            index_doc_id[key] = np.array(file["index_doc_id_{}".format(key)], dtype=np.int32)
            index_doc_value[key] = np.array(file["index_doc_value_{}".format(key)], dtype=np.float32)
        except Exception:
            index_doc_id[key] = np.array([], dtype=np.int32)
            index_doc_value[key] = np.array([], dtype=np.float32)
    file.close()
    return index_doc_id, index_doc_value


def test_hdf5_save_and_load_paths_with_fake_h5py(fake_h5py, golden, tmp_path):
    from scaling_retriever_b200 import inverted_index as inv
    off, ids, vals = golden["C_offsets"], golden["C_ids"], golden["C_vals"]      # case C has 5 empty posting lists
    n_terms = len(off) - 1
    index = inv.IndexDictOfArray(str(tmp_path), force_new=True, dim_voc=n_terms)
    index._set_csr_host(off.astype(np.int64), ids.astype(np.int32), vals.astype(np.float32))   # host CSR: no GPU needed to save
    index.n = int(ids.max()) + 1
    index.save()
    h5 = os.path.join(str(tmp_path), "array_index.h5py")
    assert os.path.exists(h5)

    # what the reference would read back from that file
    ref_ids, ref_vals = reference_loader(h5, n_terms)
    for t in range(n_terms):
        assert np.array_equal(ref_ids[t], ids[off[t]:off[t + 1]]) and ref_ids[t].dtype == np.int32
        assert np.array_equal(ref_vals[t], vals[off[t]:off[t + 1]]) and ref_vals[t].dtype == np.float32
    with fake_h5py.File(h5, "r") as f:
        assert int(f["dim"][()]) == int(np.count_nonzero(np.diff(off)))          # `dim` = number of non-empty lists (:96)
        assert "index_doc_id_{}".format(int(np.flatnonzero(np.diff(off) == 0)[0])) not in f   # empty lists are not stored
    with open(os.path.join(str(tmp_path), "index_dist.json")) as f:
        dist = json.load(f)
    assert dist == {str(t): int(off[t + 1] - off[t]) for t in range(n_terms) if off[t + 1] > off[t]}

    # the product's HDF5 reader (used when a directory holds only the reference's file)
    l_off, l_ids, l_vals = inv._load_hdf5(h5, n_terms)
    assert np.array_equal(l_off, off) and np.array_equal(l_ids, ids) and np.array_equal(l_vals.view(np.uint32), vals.view(np.uint32))
    # dim_voc larger than what was stored: the reference trusts the argument and fills empty lists (:25-26, :36-41)
    l_off2, _, _ = inv._load_hdf5(h5, n_terms + 5)
    assert len(l_off2) == n_terms + 6 and l_off2[-1] == off[-1]

    # a directory re-indexed by the reference (HDF5 newer than the CSR bundle) is served from the HDF5 file, not the stale bundle
    for p in [os.path.join(str(tmp_path), f) for f in inv.CSR_FILES]:
        np.save(p, np.zeros(3, dtype=np.int64))                                   # garbage bundle ...
        os.utime(p, (1, 1))                                                       # ... that is older than the HDF5 file
    r_off, r_ids, r_vals = inv.read_index_dir(str(tmp_path), "array_index.h5py", n_terms)
    assert np.array_equal(r_off, off) and np.array_equal(r_ids, ids)


def test_without_h5py_the_builtin_hdf5_reader_and_writer_are_used(golden, tmp_path, monkeypatch):
    """No h5py (this image): save() writes array_index.h5py with hdf5_lite.write_file, and the reference's loader logic reads it
    back through hdf5_lite.File — same datasets, dtypes and `dim` semantics as with h5py."""
    from scaling_retriever_b200 import hdf5_lite, inverted_index as inv
    monkeypatch.setitem(sys.modules, "h5py", None)                                # import h5py -> ImportError
    off, ids, vals = golden["C_offsets"], golden["C_ids"], golden["C_vals"]
    n_terms = len(off) - 1
    index = inv.IndexDictOfArray(str(tmp_path), force_new=True, dim_voc=n_terms)
    index._set_csr_host(off.astype(np.int64), ids.astype(np.int32), vals.astype(np.float32))
    index.n = int(ids.max()) + 1
    index.save()
    h5 = str(tmp_path / "array_index.h5py")
    with open(h5, "rb") as f:
        assert f.read(8) == b"\x89HDF\r\n\x1a\n"
    fake = types.ModuleType("h5py")
    fake.File = lambda name, mode="r": hdf5_lite.File(name)
    monkeypatch.setitem(sys.modules, "h5py", fake)                                # the reference's loader, over the real file
    ref_ids, ref_vals = reference_loader(h5, n_terms)
    monkeypatch.setitem(sys.modules, "h5py", None)
    for t in range(n_terms):
        assert np.array_equal(ref_ids[t], ids[off[t]:off[t + 1]]) and np.array_equal(ref_vals[t], vals[off[t]:off[t + 1]])
    with hdf5_lite.File(h5) as f:
        assert int(f["dim"][()]) == int(np.count_nonzero(np.diff(off))) and f["dim"].dtype == np.int64
        assert f["index_doc_id_0"].dtype == np.int32 and f["index_doc_value_0"].dtype == np.float32
    for p in [os.path.join(str(tmp_path), f) for f in inv.CSR_FILES]:             # HDF5-only directory (what the reference writes)
        os.remove(p)
    l_off, l_ids, l_vals = inv.read_index_dir(str(tmp_path), "array_index.h5py", n_terms)
    assert np.array_equal(l_off, off) and np.array_equal(l_ids, ids) and np.array_equal(l_vals.view(np.uint32), vals.view(np.uint32))
    assert not any(os.path.exists(os.path.join(str(tmp_path), f)) for f in inv.CSR_FILES)
    inv.convert_hdf5_to_csr(str(tmp_path), dim_voc=n_terms)                       # HDF5 -> CSR bundle converter
    c_off, c_ids, c_vals = (np.load(os.path.join(str(tmp_path), f)) for f in inv.CSR_FILES)
    assert np.array_equal(c_off, off) and np.array_equal(c_ids, ids) and np.array_equal(c_vals.view(np.uint32), vals.view(np.uint32))
