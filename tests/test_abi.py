"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/b200ret.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

from scaling_retriever_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "b200ret.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(b200ret_\w+)\s*\(", text)))


def test_library_builds_and_is_current():
    path = build.build()
    assert os.path.exists(path) and build.is_current()


def test_header_symbols_are_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/b200ret.h but not exported"
        assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) == set(names)


def test_binary_targets_sm_100a():
    out = subprocess.run(["cuobjdump", "--list-elf", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_non_compute_entry_points_work_without_gpu():
    lib = _lib.load()
    assert lib.b200ret_version() == 1
    assert lib.b200ret_sparse_block_docs() > 0 and lib.b200ret_sparse_block_docs() % 128 == 0
    small = lib.b200ret_sparse_search_workspace_bytes(1, 10)
    big = lib.b200ret_sparse_search_workspace_bytes(6980, 1000)
    assert 0 < small < big < 2 << 30
    assert lib.b200ret_csr_build_workspace_bytes(1000, 128256, 100, 0) >= 1000 * 24


def test_argument_validation_reports_errors_without_touching_the_gpu():
    lib = _lib.load()
    rc = lib.b200ret_sparse_search(None, None, 10, 10, 1, None, None, None, 1, 10, 0.0, 0, None, None, None, None, 0, None)
    assert rc == -1 and b"block_docs" in lib.b200ret_last_error()
    rc = lib.b200ret_sparse_search(None, None, 10, 10, lib.b200ret_sparse_block_docs(), None, None, None, 1, 100000, 0.0, 0,
                                   None, None, None, None, 0, None)
    assert rc == -1 and b"k=" in lib.b200ret_last_error()
    with pytest.raises(_lib.B200RetError):
        _lib.check(rc)
