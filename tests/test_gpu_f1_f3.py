"""SURVEY 8 f1 (masked / weighted pooling of a padded behaviour series behind the Embedding drop-in; FiBiNet++'s
weighted lookup) and f3 (L2 on the rows a batch used) against the oracle restatement of the reference code."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import reference_layers as R                                          # noqa: E402
from tests.util import assert_close, cpu, dense_table_grad_to_slices, oracle_deepfm, zipf_ids   # noqa: E402


@pytest.fixture(scope="module")
def L():
    from etr_b200 import CustomLayers
    return CustomLayers


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
@pytest.mark.parametrize("B,Ls,C,k", [(64, 50, 3, 16), (33, 7, 1, 8), (130, 200, 2, 32)])
def test_sequence_pool_forward_backward(L, dtype, B, Ls, C, k):
    rng = np.random.default_rng(B + Ls)
    V, pad = 500, 0
    emb = L.Embedding(V, k, table_dtype=dtype, seed=4)
    ids = rng.integers(1, V, size=(B, Ls, C))
    lens = rng.integers(0, Ls + 1, size=B)                                   # ragged histories, some empty
    ids[np.arange(Ls)[None, :] >= lens[:, None]] = pad
    scores = rng.normal(size=(B, Ls)).astype(np.float32)
    extra_mask = rng.random((B, Ls)) < 0.8
    table = cpu(emb.embeddings, torch.float64).requires_grad_(True)
    for reduce in (True, False):
        for w, m in ((scores, None), (None, None), (scores, extra_mask)):
            out = emb.sequence_pool(torch.tensor(ids), None if w is None else torch.tensor(w),
                                    None if m is None else torch.tensor(m), padding_index=pad, reduce=reduce, training=True)
            table.grad = None
            wr = None if w is None else torch.tensor(w, dtype=torch.float64, requires_grad=True)
            ref = R.sequence_pool(table, torch.tensor(ids), wr, None if m is None else torch.tensor(m), pad, reduce)
            assert out.shape == ref.shape
            # a pooled entry is a signed sum of up to L weighted rows: error scales with the terms, as for gradients
            assert_close(cpu(out).numpy(), ref.detach().numpy(), 1e-5, f"sequence_pool reduce={reduce}", grad=reduce)
            g = rng.normal(size=tuple(ref.shape)).astype(np.float32)
            sg, dw = emb.sequence_pool_backward(torch.tensor(g).cuda())
            ref.backward(torch.tensor(g, dtype=torch.float64))
            uid, rows = sg.indexed_slices()
            rows = cpu(rows).numpy()[:, :k]
            keep = np.abs(rows).sum(1) > 0
            ref_ids, ref_rows = dense_table_grad_to_slices(table.grad)
            assert np.array_equal(uid.cpu().numpy()[keep], ref_ids)
            assert_close(rows[keep], ref_rows, 1e-5, "sequence_pool table grad", grad=True)
            if wr is not None:
                assert_close(cpu(dw).numpy(), wr.grad.numpy(), 1e-5, "d scores", grad=True)


def test_sequence_pool_is_the_reference_eager_chain(L):
    """bit-level statement of what is fused: gather [B, L*C] -> reshape [B, L, C*k] -> mask -> weight -> sum over L."""
    rng = np.random.default_rng(2)
    emb = L.Embedding(100, 16, seed=1)
    ids = rng.integers(0, 100, size=(8, 5, 2))
    rows = emb(torch.tensor(ids.reshape(8, 10)))                             # the no-reduce Embedding.call, bit-exact rows
    assert rows.shape == (8, 10, 16)
    plain = emb.sequence_pool(torch.tensor(ids), reduce=False)
    assert torch.equal(plain, rows.reshape(8, 5, 32))                        # weight 1: the very rows, [B, L, C*k]
    keys = rng.integers(0, 100, size=(8, 6))
    vals = rng.normal(size=(8, 6)).astype(np.float32)
    wl = emb.weighted_lookup(torch.tensor(keys), torch.tensor(vals))
    assert torch.equal(wl, emb(torch.tensor(keys)) * torch.tensor(vals).cuda().unsqueeze(-1))     # FiBiNet++ :124-126
    with pytest.raises(IndexError):
        emb.sequence_pool(torch.tensor([[[100]]]))                           # out-of-range id raises like TF-CPU


def test_used_rows_l2_matches_reference_loop(L):
    """Trainer(used_rows_l2 = factor): loss term and gradients of 5.DIN/ModelManager.py:175-190 (unique ids of the batch,
    factor * tf.nn.l2_loss of their rows) -- two steps vs the oracle with the same term under autograd."""
    rng = np.random.default_rng(3)
    B, F, k, V, factor = 300, 5, 16, 200, 1e-3
    names = [f"f{i}" for i in range(F)]
    lay = L.DeepFMRankingLayer(names, V, k, seed=9)
    orc = oracle_deepfm(lay, torch.float64)
    tr = L.Trainer(lay, lr=1e-2)
    tr.used_rows_l2 = factor
    opt = R.KerasAdam(lr=1e-2, mode="rowwise")
    for step in range(2):
        X = zipf_ids(rng, [V // F] * F, B)
        y = (rng.random(B) < 0.3).astype(np.float32)
        loss = tr.train_step(torch.tensor(X), torch.tensor(y))
        for v in orc.variables():
            v.grad = None
        bce = R.keras_bce(torch.tensor(y, dtype=torch.float64).reshape(-1, 1), torch.sigmoid(orc.logit(torch.tensor(X))))
        l2 = R.used_rows_l2(torch.cat([orc.embed, orc.w], 1), torch.tensor(X), factor)
        (bce + l2).backward()
        lr_t = opt.step_begin()
        for v in orc.variables():
            if v is orc.embed or v is orc.w:
                nz, rows = R.dedup_dense_grad(v.grad)
                opt.apply_sparse(v, nz, rows, lr_t)
            else:
                opt.apply_dense(v, v.grad, lr_t)
        assert abs(float(loss.item()) - float(bce)) <= 1e-5 * abs(float(bce))
        assert abs(float(tr.last_l2.item()) - float(l2)) <= 1e-5 * abs(float(l2))
    assert np.abs(cpu(lay.embed, torch.float64).numpy() - orc.embed.detach().numpy()).max() <= 2e-3 * 1e-2 * 2
    assert np.abs(cpu(lay.w, torch.float64).numpy() - orc.w.detach().numpy()).max() <= 2e-3 * 1e-2 * 2
