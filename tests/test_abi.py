"""The C-ABI library loads without a GPU and exports every symbol include/kge_b200.h declares;
host-only entry points behave; argument errors come back as codes, never as crashes."""
import ctypes
import os
import re

import pytest

from torchkge_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "kge_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kge_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), "missing export " + n
        assert n in _lib.SIGNATURES, "ctypes binding lacks " + n


def test_abi_version_and_planes():
    lib = _lib.load()
    assert lib.kge_abi_version() == _lib.ABI_VERSION
    assert [lib.kge_cand_planes(m) for m in range(6)] == [1, 1, 1, 1, 2, 2]
    assert lib.kge_query_planes(_lib.TRANSE_L2, _lib.SIDE_TAIL) == 1
    assert lib.kge_query_planes(_lib.TRANSE_L2, _lib.SIDE_HEAD) == 2
    assert lib.kge_query_planes(_lib.COMPLEX, _lib.SIDE_TAIL) == 2
    assert lib.kge_cand_planes(42) == 0


def test_size_queries():
    lib = _lib.load()
    # 1000 rows -> 8 candidate tiles of 128; dim 200; 1 plane
    assert lib.kge_packed_table_floats(_lib.DISTMULT, 1000, 200) == 8 * 200 * 128
    assert lib.kge_packed_table_floats(_lib.COMPLEX, 1000, 200) == 8 * 200 * 2 * 128
    assert lib.kge_packed_table_floats(_lib.DISTMULT, 0, 200) == 0
    small = lib.kge_rank_workspace_bytes(_lib.TRANSE_L2, _lib.SIDE_TAIL, 200, 64, 0, 0)
    big = lib.kge_rank_workspace_bytes(_lib.TRANSE_L2, _lib.SIDE_HEAD, 200, 64, 0, 0)
    assert 0 < small < big  # head side carries two query planes
    with_tc = lib.kge_rank_workspace_bytes(_lib.TRANSE_L2, _lib.SIDE_TAIL, 200, 64, 100000,
                                           _lib.FLAG_TENSOR_CORE)
    assert with_tc > small + (64 * 100000 // 128) * 8  # near-tie list (1/128 of the pairs) + operand image
    # tensor-core operand image per 256-row tile: k-blocks x (hi, lo) x 256 rows x row bytes, + norms
    # + 256 bytes of per-image facts (scale, maxima)
    try:
        lib.kge_tc_configure(32, -1, -1, -1, -1)   # 64-byte swizzle: 7 k-blocks of 32 for k = 200
        assert lib.kge_tc_packed_bytes(_lib.DISTMULT, 1000, 200) == 4 * (7 * 2 * 256 * 64) + 2 * 1024 * 4 + 256
        lib.kge_tc_configure(64, -1, -1, -1, -1)   # 128-byte swizzle: 4 k-blocks of 64
        assert lib.kge_tc_packed_bytes(_lib.DISTMULT, 1000, 200) == 4 * (4 * 2 * 256 * 128) + 2 * 1024 * 4 + 256
    finally:
        lib.kge_tc_configure(32, -1, -1, -1, -1)
    assert lib.kge_tc_packed_bytes(_lib.TRANSE_L1, 1000, 200) == 0   # no tensor-core path
    assert lib.kge_tc_packed_bytes(_lib.ROTATE, 1000, 200) == 0


def test_bad_arguments_return_error_codes():
    lib = _lib.load()
    assert lib.kge_rank_side(None) == 1
    assert b"null" in lib.kge_last_error()
    args = _lib.RankArgs()
    args.model, args.side, args.dim, args.n = 77, 0, 16, 4
    assert lib.kge_rank_side(ctypes.byref(args)) == 1
    assert lib.kge_score_all(None) == 1
    assert lib.kge_pack_table(_lib.DISTMULT, None, None, 10, 16, None, None) == 1
    assert lib.kge_finalize_ranks(None, None, 0, None, None, None) == 0  # n == 0 is a no-op


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.KgeLibraryError, match="no CPU fallback"):
        _lib.load()
