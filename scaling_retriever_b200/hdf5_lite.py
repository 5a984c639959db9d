"""Minimal pure-Python HDF5 reader / writer for the reference's sparse index file `array_index.h5py`.

The reference stores its inverted index with h5py (utils/inverted_index.py:92-100): a scalar int dataset `dim` and, per non-empty
posting list t, two 1-D datasets `index_doc_id_{t}` (int32) and `index_doc_value_{t}` (float32), all in the root group, created
with `create_dataset(name, data=array)` — i.e. contiguous, unfiltered storage.  h5py / libhdf5 are not installable in this
image, so this module restates exactly the part of the HDF5 file format (HDF5 File Format Specification 2.0/3.0) that such a
file uses with libhdf5's default settings ("earliest" library version bounds):

    superblock version 0 or 1 -> root symbol-table entry -> local heap + version-1 B-tree of symbol-table nodes (SNOD)
    -> version-1 object headers -> dataspace (v1/v2), datatype (fixed-point / IEEE float), layout v3 (contiguous or compact;
       v1/v2 contiguous) messages, header continuation blocks.

`read_datasets(path)` yields the datasets as numpy arrays (memory-mapped); `write_file(path, datasets)` writes a file of the same
structure.  PARITY UNPINNED: no HDF5 implementation or HDF5 file exists in this image to check against — the writer is validated
by this reader only, and both follow the published format specification.  Unsupported structures (new-style groups of
`libver="latest"`, chunked / filtered datasets, superblock >= 2) raise NotImplementedError naming the feature.
"""
import mmap
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5FormatError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------------------ reader

class _Reader:
    def __init__(self, buf):
        self.buf = buf
        self.base = 0

    def u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self.buf, off)

    def superblock(self):
        off = 0
        while off < len(self.buf) and self.buf[off:off + 8] != SIGNATURE:   # the superblock may sit at 0, 512, 1024, ...
            off = 512 if off == 0 else off * 2
        if off >= len(self.buf):
            raise HDF5FormatError("not an HDF5 file (signature not found)")
        version = self.buf[off + 8]
        if version not in (0, 1):
            raise NotImplementedError(f"HDF5 superblock version {version} (files written with libver='latest'); only the default "
                                      "version-0/1 layout of h5py is supported")
        size_offsets, size_lengths = self.buf[off + 13], self.buf[off + 14]
        if size_offsets != 8 or size_lengths != 8:
            raise NotImplementedError(f"HDF5 offsets/lengths of {size_offsets}/{size_lengths} bytes (expected 8/8)")
        p = off + 24 + (4 if version == 1 else 0)
        base, _free, _eof, _driver = self.u("QQQQ", p)
        self.base = base
        p += 32
        _name_off, header_addr, cache_type, _res = self.u("QQII", p)
        btree = heap = None
        if cache_type == 1:
            btree, heap = self.u("QQ", p + 24)
        return header_addr, btree, heap

    # -- object headers ---------------------------------------------------------------------------------------------
    def messages(self, addr):
        """[(type, flags, data offset, size)] of a version-1 object header at `addr`, following continuation blocks."""
        a = self.base + addr
        if self.buf[a:a + 4] == b"OHDR":
            raise NotImplementedError("version-2 object headers (libver='latest')")
        version, _res, n_msgs, _refs, hdr_size = self.u("BBHII", a)
        if version != 1:
            raise HDF5FormatError(f"object header version {version} at {addr}")
        blocks = [(a + 16, hdr_size)]
        out = []
        while blocks and len(out) < n_msgs:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < n_msgs:
                mtype, msize, flags = self.u("HHB", p)
                data = p + 8
                out.append((mtype, flags, data, msize))
                if mtype == 0x0010:                                   # continuation: (offset, length)
                    c_off, c_len = self.u("QQ", data)
                    blocks.append((self.base + c_off, c_len))
                p = data + msize
        return out

    def dataset(self, addr):
        shape = dtype = None
        layout = None
        for mtype, flags, p, size in self.messages(addr):
            if flags & 0x02:
                raise NotImplementedError("shared object-header messages")
            if mtype == 0x0001:                                       # dataspace
                version, rank, dflags = self.u("BBB", p)
                if version == 1:
                    dims_at = p + 8
                elif version == 2:
                    dims_at = p + 4
                else:
                    raise HDF5FormatError(f"dataspace message version {version}")
                shape = tuple(self.u("Q" * rank, dims_at)) if rank else ()
            elif mtype == 0x0003:                                     # datatype
                cv, b0, _b1, _b2, tsize = self.u("BBBBI", p)
                cls = cv & 0x0F
                order = ">" if (b0 & 1) else "<"
                if cls == 0:
                    dtype = np.dtype(f"{order}{'i' if (b0 & 0x08) else 'u'}{tsize}")
                elif cls == 1:
                    dtype = np.dtype(f"{order}f{tsize}")
                else:
                    raise NotImplementedError(f"HDF5 datatype class {cls} (only fixed-point and floating-point datasets are used)")
            elif mtype == 0x0008:                                     # data layout
                version = self.buf[p]
                if version == 3:
                    cls = self.buf[p + 1]
                    if cls == 1:
                        d_addr, d_size = self.u("QQ", p + 2)
                        layout = ("contiguous", d_addr, d_size)
                    elif cls == 0:
                        c_size = self.u("H", p + 2)[0]
                        layout = ("compact", p + 4, c_size)
                    else:
                        raise NotImplementedError("chunked / filtered HDF5 datasets (the reference writes contiguous ones)")
                elif version in (1, 2):
                    rank, cls = self.buf[p + 1], self.buf[p + 2]
                    if cls != 1:
                        raise NotImplementedError("non-contiguous HDF5 dataset (layout message version 1/2)")
                    d_addr = self.u("Q", p + 8)[0]
                    layout = ("contiguous", d_addr, None)
                else:
                    raise HDF5FormatError(f"data layout message version {version}")
            elif mtype == 0x000B:
                raise NotImplementedError("filtered (compressed) HDF5 datasets")
        if shape is None or dtype is None or layout is None:
            raise HDF5FormatError(f"object at {addr} is not a simple dataset")
        count = int(np.prod(shape)) if shape else 1
        kind, a, _size = layout
        if kind == "compact":
            arr = np.frombuffer(self.buf, dtype=dtype, count=count, offset=a)
        elif a == UNDEF or count == 0:
            arr = np.zeros(count, dtype=dtype)
        else:
            arr = np.frombuffer(self.buf, dtype=dtype, count=count, offset=self.base + a)
        return arr.reshape(shape)

    # -- old-style groups -------------------------------------------------------------------------------------------
    def heap_data(self, heap_addr):
        a = self.base + heap_addr
        if self.buf[a:a + 4] != b"HEAP":
            raise HDF5FormatError("local heap signature missing")
        _size, _free, data_addr = self.u("QQQ", a + 8)
        return self.base + data_addr

    def name_at(self, heap_data, off):
        start = heap_data + off
        end = self.buf.find(b"\x00", start)
        return self.buf[start:end].decode("utf-8")

    def group_entries(self, btree_addr, heap_addr):
        """(name, object header address) of every link in an old-style group, in B-tree (name) order."""
        heap = self.heap_data(heap_addr)
        stack = [btree_addr]
        while stack:
            addr = stack.pop()
            a = self.base + addr
            sig = self.buf[a:a + 4]
            if sig == b"TREE":
                node_type, level, used = self.u("BBH", a + 4)
                if node_type != 0:
                    raise HDF5FormatError("B-tree node is not a group node")
                children = [self.u("Q", a + 24 + 8 + 16 * i)[0] for i in range(used)]
                stack.extend(reversed(children))
            elif sig == b"SNOD":
                n_sym = self.u("H", a + 6)[0]
                for i in range(n_sym):
                    name_off, header = self.u("QQ", a + 8 + 40 * i)
                    yield self.name_at(heap, name_off), header
            else:
                raise HDF5FormatError(f"unexpected block {sig!r} in a group B-tree")

    def root_links(self):
        header_addr, btree, heap = self.superblock()
        if btree is None:                                             # symbol-table message of the root object header
            for mtype, _flags, p, _size in self.messages(header_addr):
                if mtype == 0x0011:
                    btree, heap = self.u("QQ", p)
                elif mtype in (0x0002, 0x0006):
                    raise NotImplementedError("new-style (link-message) HDF5 groups (libver='latest')")
        if btree is None:
            raise HDF5FormatError("root group has no symbol table")
        return self.group_entries(btree, heap)


class File:
    """Read-only view of the datasets in the root group: `name in f`, `f[name]` (numpy array), `f.keys()`."""

    def __init__(self, path):
        self._fh = open(path, "rb")
        self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        self._r = _Reader(self._mm)
        self._links = dict(self._r.root_links())

    def keys(self):
        return self._links.keys()

    def __contains__(self, name):
        return name in self._links

    def __getitem__(self, name):
        return self._r.dataset(self._links[name])

    def close(self):
        self._r = None
        try:
            self._mm.close()
        except BufferError:       # numpy views of the mapping are still alive; the mapping goes away with them
            pass
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def read_datasets(path):
    """{name: ndarray copy} of every dataset in the root group."""
    with File(path) as f:
        return {name: np.array(f[name]) for name in f.keys()}


# ------------------------------------------------------------------------------------------------------------ writer
# Layout produced (what libhdf5 writes for such a file, in a simpler block order):
#   [superblock v0 | root object header | local heap header | heap data (names) | dataset object headers | SNODs | B-tree
#    nodes | raw data]

LEAF_K, INTERNAL_K = 4, 16          # libhdf5 defaults: <= 2*4 symbols per SNOD, <= 2*16 children per B-tree node


def _pad8(n):
    return (n + 7) & ~7


def _dataspace_msg(shape):
    rank = len(shape)
    body = struct.pack("<BBBB4x", 1, rank, 0, 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)
    return 0x0001, body


def _datatype_msg(dtype):
    dtype = np.dtype(dtype)
    if dtype.byteorder == ">":
        raise NotImplementedError("big-endian arrays")
    size = dtype.itemsize
    if dtype.kind in "iu":
        bits0 = 0x08 if dtype.kind == "i" else 0x00
        body = struct.pack("<BBBBI", 0x10 | 0, bits0, 0, 0, size) + struct.pack("<HH", 0, size * 8)
    elif dtype.kind == "f" and size in (4, 8):
        # IEEE: byte order LE, padding 0, mantissa normalisation 2 (implied msb), sign bit position in byte 1
        sign_pos = size * 8 - 1
        if size == 4:
            exp_loc, exp_size, man_loc, man_size, bias = 23, 8, 0, 23, 127
        else:
            exp_loc, exp_size, man_loc, man_size, bias = 52, 11, 0, 52, 1023
        body = struct.pack("<BBBBI", 0x10 | 1, 0x20, sign_pos, 0, size) + struct.pack("<HHBBBBI", 0, size * 8, exp_loc, exp_size, man_loc,
                                                                                         man_size, bias)
    else:
        raise NotImplementedError(f"dtype {dtype}")
    return 0x0003, body


def _object_header(messages):
    """Version-1 object header: 16-byte prefix (12 used + 4 pad) then 8-byte aligned messages."""
    body = b""
    for mtype, data in messages:
        data = data + b"\x00" * (_pad8(len(data)) - len(data))
        body += struct.pack("<HHB3x", mtype, len(data), 0) + data
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def write_file(path, datasets):
    """Write `datasets` ({name: array}, scalars allowed) as contiguous datasets of the root group of a new HDF5 file."""
    names = sorted(datasets, key=lambda s: s.encode("utf-8"))       # group B-trees are ordered by name (strcmp)
    arrays = {n: np.asarray(datasets[n], order="C") for n in names}   # (ascontiguousarray would turn scalars into 1-D)

    # local heap data: offset 0 holds the empty string (key 0 of the B-tree), then the names, 8-byte aligned each
    heap = bytearray(b"\x00" * 8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap)
        raw = n.encode("utf-8") + b"\x00"
        heap += raw + b"\x00" * (_pad8(len(raw)) - len(raw))
    free_off = len(heap)
    heap += struct.pack("<QQ", 1, 16)                               # one free block at the end: (next = 1 (none), size 16)
    heap_size = len(heap)

    # group the names into SNODs and B-tree levels
    per_leaf = 2 * LEAF_K
    leaves = [names[i:i + per_leaf] for i in range(0, len(names), per_leaf)] or [[]]
    levels = [leaves]                                               # levels[0] = SNODs; above: lists of child index ranges
    width = len(leaves)
    while width > 2 * INTERNAL_K or len(levels) == 1:
        below = len(levels[-1])
        groups = [list(range(i, min(i + 2 * INTERNAL_K, below))) for i in range(0, below, 2 * INTERNAL_K)]
        levels.append(groups)
        width = len(groups)
        if width == 1:
            break
    if len(levels[-1]) != 1:
        levels.append([list(range(len(levels[-1])))])

    # ---- assign addresses -----------------------------------------------------------------------------------------
    pos = 96                                                        # superblock v0 (56 bytes) + root symbol-table entry (40)
    root_header_addr = pos
    root_header = _object_header([(0x0011, struct.pack("<QQ", 0, 0))])          # patched below
    pos += len(root_header)
    heap_addr = pos
    pos += 32
    heap_data_addr = pos
    pos += heap_size
    header_addr, header_len = {}, {}
    data_addr = {}
    for n in names:
        a = arrays[n]
        probe = _object_header([_dataspace_msg(a.shape), _datatype_msg(a.dtype), (0x0008, struct.pack("<BBQQ", 3, 1, 0, a.nbytes))])
        header_addr[n], header_len[n] = pos, len(probe)
        pos += len(probe)
    snod_size = 8 + 40 * per_leaf
    snod_addr = []
    for _ in leaves:
        snod_addr.append(pos)
        pos += snod_size
    node_size = 24 + 8 + 16 * (2 * INTERNAL_K)
    node_addr = [snod_addr]
    for lvl in levels[1:]:
        addrs = []
        for _ in lvl:
            addrs.append(pos)
            pos += node_size
        node_addr.append(addrs)
    pos = _pad8(pos)
    for n in names:
        a = arrays[n]
        data_addr[n] = pos if a.nbytes else UNDEF
        pos += _pad8(a.nbytes)
    eof = pos

    # largest name (as heap offset) under every node, for the B-tree keys
    max_name = [[name_off[leaf[-1]] if leaf else 0 for leaf in leaves]]
    for li, lvl in enumerate(levels[1:], 1):
        max_name.append([max_name[li - 1][grp[-1]] for grp in lvl])
    root_btree = node_addr[-1][0]

    with open(path, "wb") as f:
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root_header_addr, 1, 0) + struct.pack("<QQ", root_btree, heap_addr)
        assert len(sb) == 96
        f.write(sb)
        f.write(_object_header([(0x0011, struct.pack("<QQ", root_btree, heap_addr))]))
        f.write(b"HEAP" + struct.pack("<B3xQQQ", 0, heap_size, free_off, heap_data_addr))
        f.write(bytes(heap))
        for n in names:
            a = arrays[n]
            hdr = _object_header([_dataspace_msg(a.shape), _datatype_msg(a.dtype),
                                  (0x0008, struct.pack("<BBQQ", 3, 1, data_addr[n], a.nbytes))])
            assert len(hdr) == header_len[n] and f.tell() == header_addr[n]
            f.write(hdr)
        for leaf in leaves:
            blk = b"SNOD" + struct.pack("<BBH", 1, 0, len(leaf))
            for n in leaf:
                blk += struct.pack("<QQII16x", name_off[n], header_addr[n], 0, 0)
            f.write(blk + b"\x00" * (snod_size - len(blk)))
        for li, lvl in enumerate(levels[1:], 1):
            for gi, grp in enumerate(lvl):
                left = node_addr[li][gi - 1] if gi > 0 else UNDEF
                right = node_addr[li][gi + 1] if gi + 1 < len(lvl) else UNDEF
                blk = b"TREE" + struct.pack("<BBHQQ", 0, li - 1, len(grp), left, right)
                first_key = max_name[li - 1][grp[0] - 1] if grp[0] > 0 else 0     # key i bounds child i from below
                blk += struct.pack("<Q", first_key)
                for c in grp:
                    blk += struct.pack("<QQ", node_addr[li - 1][c], max_name[li - 1][c])
                f.write(blk + b"\x00" * (node_size - len(blk)))
        f.write(b"\x00" * (_pad8(f.tell()) - f.tell()))
        for n in names:
            a = arrays[n]
            if a.nbytes:
                assert f.tell() == data_addr[n]
                f.write(a.tobytes())
                f.write(b"\x00" * (_pad8(a.nbytes) - a.nbytes))
        assert f.tell() == eof
