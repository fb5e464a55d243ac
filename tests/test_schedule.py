"""The reduction schedules built by libkge_b200 (host code, no GPU) replayed in numpy must
reproduce ATen's CPU reductions bit for bit -- the arithmetic contract of include/kge_b200.h."""
import numpy as np
import pytest
import torch

from oracle import replay
from torchkge_b200 import _lib

DIMS = [1, 3, 7, 8, 9, 16, 31, 33, 50, 64, 100, 127, 200, 256, 400, 511, 512, 513, 1000, 1003, 2049]


@pytest.mark.parametrize("d", DIMS)
def test_schedule_is_a_permutation(d):
    for model in (_lib.TRANSE_L1, _lib.TRANSE_L2, _lib.DISTMULT, _lib.COMPLEX, _lib.ROTATE):
        perm, code = _lib.build_schedule(model, d)
        assert sorted(perm.tolist()) == list(range(d))
        assert code.shape == (d,)


@pytest.mark.parametrize("d", DIMS)
def test_l1_schedule_matches_aten_norm1(d):
    torch.manual_seed(d)
    x = torch.randn(300, d)
    perm, code = _lib.build_schedule(_lib.TRANSE_L1, d)
    got = replay.replay(perm, code, np.abs(x.numpy()))
    want = x.norm(p=1, dim=-1).numpy()
    assert np.array_equal(replay.bits(got), replay.bits(want))


@pytest.mark.parametrize("d", DIMS)
def test_l2_schedule_matches_aten_norm2(d):
    torch.manual_seed(d)
    x = torch.randn(3, 100, d)
    perm, code = _lib.build_schedule(_lib.TRANSE_L2, d)
    t = replay.replay(perm, code, x.numpy().reshape(-1, d), l2=True)
    got = np.sqrt(t).astype(np.float32)
    want = x.norm(p=2, dim=-1).numpy().reshape(-1)
    assert np.array_equal(replay.bits(got), replay.bits(want))
    # and the full dissimilarity: norm ** 2 is an exact fp32 square of the rounded norm
    want2 = (x.norm(p=2, dim=-1) ** 2).numpy().reshape(-1)
    assert np.array_equal(replay.bits((got * got).astype(np.float32)), replay.bits(want2))


@pytest.mark.parametrize("d", DIMS)
def test_sum_schedule_matches_aten_sum(d):
    torch.manual_seed(d)
    x = torch.randn(3, 100, d)
    perm, code = _lib.build_schedule(_lib.DISTMULT, d)
    got = replay.replay(perm, code, x.numpy().reshape(-1, d))
    want = x.sum(dim=2).numpy().reshape(-1)
    assert np.array_equal(replay.bits(got), replay.bits(want))


def test_unsupported_dim_is_an_error_not_a_crash():
    with pytest.raises(_lib.KgeLibraryError):
        _lib.build_schedule(_lib.DISTMULT, 8192)
    with pytest.raises(_lib.KgeLibraryError):
        _lib.build_schedule(99, 16)


def test_schedule_depth_is_the_depth_of_the_reduction_tree():
    """kge_schedule_depth: rounded additions on the longest path of the ATen-order reduction."""
    lib = _lib.load()
    # norm(p=1): one sequential chain over d terms -> d - 1 additions (the first is exact)
    assert [lib.kge_schedule_depth(_lib.TRANSE_L1, d) for d in (1, 2, 50)] == [0, 1, 49]
    # norm(p=2): 8 lanes of d/8 terms, lanes folded one after the other, short tail
    for d in (8, 64, 200, 203, 1000):
        depth = lib.kge_schedule_depth(_lib.TRANSE_L2, d)
        assert d // 8 - 1 <= depth <= d // 8 + 15
    # cascade sum: 32 chains of d/32 terms (+ spill levels), rows and lanes folded
    for d in (7, 8, 100, 200, 400, 800, 2049):
        depth = lib.kge_schedule_depth(_lib.DISTMULT, d)
        assert 0 <= depth <= d // 32 + 21
        assert depth == lib.kge_schedule_depth(_lib.COMPLEX, d)
    assert lib.kge_schedule_depth(_lib.DISTMULT, 10 ** 6) == -1
