"""hdf5_lite.py: the built-in reader / writer of the HDF5 file structure h5py writes by default for the reference's
array_index.h5py (utils/inverted_index.py:92-100).  PARITY UNPINNED against libhdf5 (none in this image): these tests check
the writer against the reader, the on-disk structures against hand-decoded bytes of the format specification, and the error
behaviour on unsupported files."""
import struct

import numpy as np
import pytest

from scaling_retriever_b200 import hdf5_lite


def test_roundtrip_many_datasets_multi_level_btree(tmp_path):
    rng = np.random.default_rng(0)
    datasets = {"dim": np.int64(5000)}
    for t in range(3000):                      # 6001 links -> 751 SNODs -> 24 level-0 B-tree nodes -> 1 level-1 root
        n = int(rng.integers(0, 40)) if t % 17 else 0
        datasets[f"index_doc_id_{t}"] = rng.integers(0, 1 << 20, size=n).astype(np.int32)
        datasets[f"index_doc_value_{t}"] = rng.random(n).astype(np.float32)
    path = str(tmp_path / "array_index.h5py")
    hdf5_lite.write_file(path, datasets)
    with hdf5_lite.File(path) as f:
        assert set(f.keys()) == set(datasets)
        names = list(f.keys())
        assert names == sorted(names, key=lambda s: s.encode())          # group B-tree order = name order
        assert "index_doc_id_3000" not in f
        for name, want in datasets.items():
            got = f[name]
            assert got.dtype == np.asarray(want).dtype and got.shape == np.asarray(want).shape
            assert np.array_equal(got, want)
    assert hdf5_lite.read_datasets(path)["dim"] == 5000


def test_on_disk_structures_follow_the_format_specification(tmp_path):
    path = str(tmp_path / "t.h5")
    hdf5_lite.write_file(path, {"b": np.arange(5, dtype=np.int32), "a": np.array([1.5, -2.0], dtype=np.float32), "s": np.int64(7)})
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0                       # superblock version 0
    assert raw[13] == 8 and raw[14] == 8                                           # 8-byte offsets and lengths
    leaf_k, internal_k = struct.unpack_from("<HH", raw, 16)
    assert (leaf_k, internal_k) == (4, 16)
    base, free, eof, driver = struct.unpack_from("<QQQQ", raw, 24)
    assert base == 0 and eof == len(raw) and free == driver == 0xFFFFFFFFFFFFFFFF
    name_off, root_hdr, cache_type, _ = struct.unpack_from("<QQII", raw, 56)
    btree, heap = struct.unpack_from("<QQ", raw, 80)
    assert cache_type == 1 and raw[btree:btree + 4] == b"TREE" and raw[heap:heap + 4] == b"HEAP"
    version, _, n_msgs, refs, hsize = struct.unpack_from("<BBHII", raw, root_hdr)
    assert (version, n_msgs, refs) == (1, 1, 1)
    mtype, msize, flags = struct.unpack_from("<HHB", raw, root_hdr + 16)
    assert mtype == 0x0011 and struct.unpack_from("<QQ", raw, root_hdr + 24) == (btree, heap)
    node_type, level, used = struct.unpack_from("<BBH", raw, btree + 4)
    assert (node_type, level, used) == (0, 0, 1)
    key0, child0, key1 = struct.unpack_from("<QQQ", raw, btree + 24)
    assert key0 == 0 and raw[child0:child0 + 4] == b"SNOD"
    heap_size, free_head, heap_data = struct.unpack_from("<QQQ", raw, heap + 8)
    n_sym = struct.unpack_from("<H", raw, child0 + 6)[0]
    names = []
    for i in range(n_sym):
        off, hdr = struct.unpack_from("<QQ", raw, child0 + 8 + 40 * i)
        end = raw.index(b"\x00", heap_data + off)
        names.append(raw[heap_data + off:end].decode())
    assert names == ["a", "b", "s"]
    end = raw.index(b"\x00", heap_data + key1)
    assert raw[heap_data + key1:end] == b"s"                                       # last key = largest name of the child
    # dataset "a": dataspace v1 rank 1 dim 2, IEEE little-endian float32, contiguous layout pointing at the raw floats
    a_hdr = struct.unpack_from("<QQ", raw, child0 + 8)[1]
    p = a_hdr + 16
    msgs = {}
    for _ in range(3):
        mtype, msize, _f = struct.unpack_from("<HHB", raw, p)
        msgs[mtype] = p + 8
        p += 8 + msize
    assert struct.unpack_from("<BB", raw, msgs[1]) == (1, 1) and struct.unpack_from("<Q", raw, msgs[1] + 8)[0] == 2
    cv, b0, b1, _b2, tsize = struct.unpack_from("<BBBBI", raw, msgs[3])
    assert cv == 0x11 and b0 == 0x20 and b1 == 31 and tsize == 4
    assert struct.unpack_from("<HHBBBBI", raw, msgs[3] + 8) == (0, 32, 23, 8, 0, 23, 127)
    lver, lcls, addr, size = struct.unpack_from("<BBQQ", raw, msgs[8])
    assert (lver, lcls, size) == (3, 1, 8) and struct.unpack_from("<ff", raw, addr) == (1.5, -2.0)


def test_unsupported_files_fail_loudly(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"definitely not hdf5" * 10)
    with pytest.raises(hdf5_lite.HDF5FormatError):
        hdf5_lite.File(str(p))
    p.write_bytes(b"\x89HDF\r\n\x1a\n" + bytes([2]) + b"\x00" * 200)                # superblock version 2 (libver="latest")
    with pytest.raises(NotImplementedError, match="superblock version 2"):
        hdf5_lite.File(str(p))


def test_empty_file_and_empty_datasets(tmp_path):
    path = str(tmp_path / "e.h5")
    hdf5_lite.write_file(path, {})
    with hdf5_lite.File(path) as f:
        assert list(f.keys()) == []
    hdf5_lite.write_file(path, {"z": np.zeros(0, dtype=np.int32)})
    with hdf5_lite.File(path) as f:
        assert f["z"].shape == (0,) and f["z"].dtype == np.int32
