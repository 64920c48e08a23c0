"""SURVEY 8 f4 on the device: TFRecord files -> pinned column blocks -> Trainer (graph mode) trains like the oracle;
serving rank() == the oracle's forward on the batch the reference would have assembled (2.FM/OnlineServer.py:77-101)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import reference_layers as R                          # noqa: E402
from tests.util import cpu, oracle_deepfm                         # noqa: E402

FEATS = ['user_tag0', 'user_tag1', 'item_tag1', 'item_tag2', 'item_tag3']


def test_tfrecord_to_trainer_matches_oracle(tmp_path):
    from etr_b200 import CustomLayers as L
    from etr_b200 import tfrecord as T
    rng = np.random.default_rng(3)
    N, V, B = 1024, 600, 256
    ints = {f: rng.integers(0, V, size=N) for f in FEATS}
    y = (rng.random(N) < 0.3).astype(np.float32)
    path = str(tmp_path / "data_train_0.tfrecord")
    T.write_examples(path, ints, {"label": y})
    lay = L.DeepFMRankingLayer(FEATS, V, 16, seed=5)
    orc = oracle_deepfm(lay, torch.float64)
    tr = L.Trainer(lay, lr=1e-2, graph=True)
    opt = R.KerasAdam(lr=1e-2, mode="rowwise")
    step = 0
    for batch in T.TFRecordDataset(path, FEATS, ["label"], batch=B):
        assert batch[FEATS[0]].is_pinned()
        target = batch.pop("label")                                # the reference's inputs.pop(label_name)
        loss = float(tr.train_step(batch, target).item())
        X = np.stack([ints[f][step * B:(step + 1) * B] for f in FEATS], 1)
        for v in orc.variables():
            v.grad = None
        l_ref = R.keras_bce(torch.tensor(y[step * B:(step + 1) * B], dtype=torch.float64).reshape(-1, 1),
                            torch.sigmoid(orc.logit(torch.tensor(X))))
        l_ref.backward()
        lr_t = opt.step_begin()
        for v in orc.variables():
            if v is orc.embed or v is orc.w:
                nz, rows = R.dedup_dense_grad(v.grad)
                opt.apply_sparse(v, nz, rows, lr_t)
            else:
                opt.apply_dense(v, v.grad, lr_t)
        assert abs(loss - float(l_ref)) <= 2e-5 * abs(float(l_ref)), step
        step += 1
    assert step == 4
    assert np.abs(cpu(lay.embed, torch.float64).numpy() - orc.embed.detach().numpy()).max() <= 2e-3 * 1e-2 * 4


def test_rank_matches_reference_batch_assembly():
    from etr_b200 import CustomLayers as L
    from etr_b200.serving import Ranker
    rng = np.random.default_rng(8)
    V = 500
    lay = L.DeepFMRankingLayer(FEATS, V, 16, seed=6)
    users = {str(10 ** 18 + i): rng.integers(0, V, size=2).tolist() for i in range(20)}
    items = {str(2 * 10 ** 18 + i): rng.integers(0, V, size=3).tolist() for i in range(300)}
    rk = Ranker(lay, users, items, max_items=256)
    uid = list(users)[3]
    cand = [list(items)[i] for i in rng.permutation(300)[:100]]
    got = rk.rank(uid, cand)
    # the reference's assembly: user features repeated, item features appended per candidate
    X = {f: [] for f in FEATS}
    for name, val in zip(FEATS[:2], users[uid]):
        X[name] = [val] * len(cand)
    for it in cand:
        for name, val in zip(FEATS[2:], items[it]):
            X[name].append(val)
    ref = oracle_deepfm(lay, torch.float64).call({k: torch.tensor(v) for k, v in X.items()})["output"].detach().numpy()
    assert list(got) == cand
    np.testing.assert_allclose(np.array(list(got.values())), ref[:, 0], rtol=1e-5)
    assert rk.rank(uid, []) == {}
    with pytest.raises(KeyError):
        rk.rank("nobody", cand)
    with pytest.raises(KeyError):
        rk.rank(uid, ["no such item"])
