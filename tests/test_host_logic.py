"""Host-side logic that needs no GPU: column staging (back-to-back host columns move as one copy), the c5 id space,
the layer keyword parsing of the sharding modes."""
import numpy as np
import torch


def test_copy_columns_groups_back_to_back_columns():
    from etr_b200.CustomLayers import Trainer
    B, n = 64, 6
    block = torch.arange(n * B, dtype=torch.int64).reshape(n, B)            # columns of one contiguous block
    cols = [block[i] for i in range(n)]
    dst = torch.zeros((n, B), dtype=torch.int64)
    calls = []
    orig = torch.Tensor.copy_

    def spy(self, src, non_blocking=False):
        calls.append(tuple(src.shape))
        return orig(self, src, non_blocking=non_blocking)

    torch.Tensor.copy_ = spy
    try:
        Trainer._copy_columns(dst, cols, torch.int64)
    finally:
        torch.Tensor.copy_ = orig
    assert torch.equal(dst, block) and calls == [(n, B)]                    # ONE copy for the whole block

    # separate allocations, a [B,1] column, a wrong dtype and a gap in the middle: still correct, column by column
    cols2 = [block[0], block[1].clone(), block[2].reshape(-1, 1), block[3].to(torch.int32), block[5], block[5]]
    dst2 = torch.zeros((n, B), dtype=torch.int64)
    Trainer._copy_columns(dst2, cols2, torch.int64)
    want = torch.stack([block[0], block[1], block[2], block[3], block[5], block[5]])
    assert torch.equal(dst2, want)

    # numpy columns of a Fortran-ordered matrix are back to back as well
    X = np.asfortranarray(np.arange(B * 3, dtype=np.float32).reshape(B, 3))
    dst3 = torch.zeros((3, B))
    Trainer._copy_columns(dst3, [X[:, j] for j in range(3)], torch.float32)
    assert np.array_equal(dst3.numpy(), X.T)


def test_c5_id_space():
    import bench
    cards = bench.c5_cards()
    assert len(cards) == 26 and sum(cards) == 100_000_000 and min(cards) >= 1
    base = np.asarray(bench.CRITEO_CARDS, dtype=np.float64)
    ratio = np.asarray(cards) / (base * 1e8 / base.sum())
    assert np.all(ratio[base > 1000] > 0.99) and np.all(ratio[base > 1000] < 1.01)      # proportional to Criteo
    (X, Xc, y), = bench.make_batches(1, 512, "zipf", cards=cards)
    offs = np.concatenate([[0], np.cumsum(cards)[:-1]])
    assert X.shape == (512, 26) and np.all(X >= offs[None, :]) and np.all(X < (offs + np.asarray(cards))[None, :])
    assert Xc.shape == (512, 13) and set(np.unique(y)) <= {0.0, 1.0}


def test_sharding_modes_are_validated_without_a_gpu():
    import pytest
    from etr_b200 import CustomLayers as L
    if torch.cuda.is_available():
        pytest.skip("constructor would succeed on a GPU box")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.FMRankingLayer(["a"], 10, 16, shard=("peer", 2, 0))
