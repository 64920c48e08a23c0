"""GPU parity: embedding gather, FM / DeepFM forward + backward, sparse plan,
segment reduction, Adam -- CUDA path (through the C ABI) vs the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import reference_layers as R                     # noqa: E402
from tests.util import (assert_close, cpu, dense_table_grad_to_slices, oracle_deepfm, oracle_fm,  # noqa: E402
                        zipf_ids)

FP32_RTOL = 1e-5      # north_star: fp32 outputs and gradients within 1e-5 relative
BF16_RTOL = 1e-2


@pytest.fixture(scope="module")
def L():
    from etr_b200 import CustomLayers
    return CustomLayers


def _names(F):
    return [f"f{i}" for i in range(F)]


def _dict_inputs(X, names, rank2_every=2):
    d = {}
    for i, n in enumerate(names):
        col = torch.tensor(X[:, i])
        d[n] = col.reshape(-1, 1) if i % rank2_every == 0 else col       # mix [B,1] and [B] like Keras Input / tf.constant
    return d


# ------------------------------------------------------------------ gather
@pytest.mark.parametrize("V,k,dtype", [(20, 16, "float32"), (1000, 8, "float32"), (777, 10, "float32"),
                                        (5000, 64, "float32"), (300, 16, "bfloat16"), (300, 64, "bfloat16")])
def test_embedding_gather_bit_exact(L, V, k, dtype):
    emb = L.Embedding(V, k, table_dtype=dtype, seed=1)
    rng = np.random.default_rng(0)
    ids = rng.integers(0, V, size=(37, 5))
    out = emb(torch.tensor(ids))
    ref = emb.embeddings.float().cpu()[torch.tensor(ids)]
    assert out.shape == (37, 5, k)
    assert torch.equal(out.cpu(), ref)                         # gathered rows bit-exact


def test_out_of_range_id_raises(L):
    from etr_b200 import EtrIdRangeError
    lay = L.FMRankingLayer(_names(3), feature_dims=20, embedding_dims=16)
    with pytest.raises(EtrIdRangeError):
        lay(torch.tensor([[0, 1, 20]]))
    with pytest.raises(IndexError):
        lay(torch.tensor([[0, -1, 2]]))
    out = lay(torch.tensor([[0, 1, 19]]))["output"]           # error word was cleared
    assert out.shape == (1, 1)


# -------------------------------------------------------------- FM forward
@pytest.mark.parametrize("B,F,k,V", [(3, 3, 16, 20), (4096, 26, 16, 160000), (257, 5, 8, 5600), (100, 7, 10, 999),
                                      (64, 26, 64, 4000), (33, 4, 128, 500), (1, 1, 4, 3)])
def test_fm_forward_matches_oracle(L, B, F, k, V):
    rng = np.random.default_rng(B + F)
    lay = L.FMRankingLayer(_names(F), feature_dims=V, embedding_dims=k, seed=3)
    X = zipf_ids(rng, [V // F] * F, B) if V >= F else rng.integers(0, V, size=(B, F))
    out = lay(_dict_inputs(X, lay.feature_names))["output"]
    assert out.shape == (B, 1) and out.dtype == torch.float32
    ref32 = oracle_fm(lay, torch.float32).call(torch.tensor(X))["output"].detach().numpy()
    ref64 = oracle_fm(lay, torch.float64).call(torch.tensor(X))["output"].detach().numpy()
    assert_close(out.cpu().numpy(), ref32, FP32_RTOL, "vs fp32 oracle")
    assert_close(out.cpu().numpy(), ref64, FP32_RTOL, "vs fp64 oracle")
    # [B,F] matrix input is the same function
    out2 = lay(torch.tensor(X))["output"]
    assert torch.equal(out, out2)


def test_fm_reference_docstring_inputs(L):
    """2.FM/CustomLayers.py:88-90 docstring: ids 0..8 over 3 features, V=20."""
    lay = L.FMRankingLayer(feature_names=['item_tag1', 'item_tag2', 'item_tag3'])
    inp = {'item_tag1': torch.tensor([0, 1, 2]), 'item_tag2': torch.tensor([3, 4, 5]),
           'item_tag3': torch.tensor([6, 7, 8])}
    out = lay(inp)["output"]
    X = torch.tensor([[0, 3, 6], [1, 4, 7], [2, 5, 8]])
    assert_close(out.cpu().numpy(), oracle_fm(lay).call(X)["output"].detach().numpy(), FP32_RTOL)
    assert [tuple(v.shape) for v in lay.variables] == [(1,), (20, 16), (20, 1)]      # bias, embed, w


def test_fm_numpy_and_dlpack_inputs(L):
    lay = L.FMRankingLayer(_names(4), feature_dims=50, embedding_dims=16)
    X = np.random.default_rng(5).integers(0, 50, size=(9, 4))
    a = lay({n: X[:, i] for i, n in enumerate(lay.feature_names)})["output"]

    class Producer:                       # any DLPack producer (tf.Tensor implements the same protocol)
        def __init__(self, t):
            self.t = t

        def __dlpack__(self, **kw):
            return self.t.__dlpack__(**kw)

        def __dlpack_device__(self):
            return self.t.__dlpack_device__()

        shape = property(lambda s: s.t.shape)

    b = lay({n: Producer(torch.tensor(X[:, i]).cuda()) for i, n in enumerate(lay.feature_names)})["output"]
    assert torch.equal(a, b)
    assert torch.equal(torch.from_dlpack(a), a)                # output exports DLPack


def test_device_resident_dict_uses_assemble_kernel(L):
    """dict of CUDA int64 columns -> one etr_assemble_ids launch into the field-major [F,B] buffer."""
    lay = L.FMRankingLayer(_names(70), feature_dims=500, embedding_dims=8)      # 70 fields: two launches of <= 64
    X = np.random.default_rng(2).integers(0, 500, size=(33, 70))
    host = lay({n: torch.tensor(X[:, i]) for i, n in enumerate(lay.feature_names)})["output"]
    l0 = lay.rt.launches
    dev = lay({n: torch.tensor(X[:, i]).cuda().reshape(-1, 1) for i, n in enumerate(lay.feature_names)})["output"]
    assert lay.rt.launches - l0 >= 3                      # 2 assemble launches + the gather kernel
    assert torch.equal(host, dev)


def test_fm_bf16_table(L):
    rng = np.random.default_rng(11)
    lay = L.FMRankingLayer(_names(26), feature_dims=10000, embedding_dims=16, table_dtype="bfloat16")
    X = rng.integers(0, 10000, size=(512, 26))
    out = lay(torch.tensor(X))["output"]
    ref = oracle_fm(lay, torch.float64).call(torch.tensor(X))["output"].detach().numpy()   # same (rounded) weights
    assert_close(out.cpu().numpy(), ref, FP32_RTOL, "bf16 rows, fp32 accumulate")


# ----------------------------------------------- streaming single-hot kernel, every template shape
@pytest.mark.parametrize("flat", [None, "float32", "bfloat16"])
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
@pytest.mark.parametrize("B,F,k,has_w", [(1, 1, 8, True), (37, 26, 16, True), (1000, 39, 8, True), (513, 13, 32, True),
                                         (130, 26, 64, False), (64, 3, 16, True), (2049, 9, 128, False)])
def test_streaming_gather_every_shape(L, B, F, k, has_w, dtype, flat):
    """gather_fm_fwd_stream_kernel<Elem, LPR, U, FLAT>: LPR in {2,4,8,16}, U in {8,13}, ragged tiles,
    partial field batches, fp32 / bf16 rows, no / fp32 / bf16 flattened operand with front columns."""
    from etr_b200.runtime import EmbeddingTable, IdsBatch, Runtime, gather_fm_forward
    rt = Runtime.get()
    tdt = torch.float32 if dtype == "float32" else torch.bfloat16
    V = 4099
    tab = EmbeddingTable(rt, V, k + (1 if has_w else 0), tdt)
    g = torch.Generator(device=rt.device).manual_seed(5)
    tab.init_uniform(-0.05, 0.05, g)
    rng = np.random.default_rng(B * 131 + F)
    X = torch.tensor(rng.integers(0, V, size=(B, F))).cuda()
    ids = IdsBatch.from_matrix(rt, X)
    C_ = 5
    cont = torch.randn(B, C_, device=rt.device)
    logit, prob, sumv = rt.empty((B,)), rt.empty((B, 1)), rt.empty((B, k))
    fl = None
    col0 = 8
    if flat:
        fl = torch.full((B, col0 + F * k), 7.0, device=rt.device, dtype=getattr(torch, flat))
    bias = torch.tensor([0.125], device=rt.device)
    gather_fm_forward(tab, k, has_w, ids, bias=bias, logit=logit, prob=prob, sumv=sumv, flat=fl, flat_col0=col0,
                      cont=cont if flat else None)
    rt.poll_error()
    rows = tab.data[X].double()                              # [B,F,stride]
    v = rows[..., :k]
    S = v.sum(1)
    z = 0.125 + 0.5 * ((S * S).sum(1) - (v * v).sum((1, 2)))
    if has_w:
        z = z + rows[..., k].sum(1)
    assert_close(logit.cpu().numpy(), z.cpu().numpy(), FP32_RTOL, "logit")
    assert_close(prob.cpu().numpy().ravel(), torch.sigmoid(z).cpu().numpy(), FP32_RTOL, "prob")
    assert_close(sumv.cpu().numpy(), S.cpu().numpy(), FP32_RTOL, "sum_v")
    if flat:
        want = torch.cat([torch.zeros(B, col0 - C_, device=rt.device), cont, tab.data[X][..., :k].float().reshape(B, F * k)], 1)
        assert torch.equal(fl, want.to(fl.dtype)), "flattened operand (front padding | dense | rows) must be exact"


# --------------------------------------------------------------------- bags
@pytest.mark.parametrize("pooling", ["sum", "mean"])
@pytest.mark.parametrize("Lmax", [1, 5, 50])
def test_fm_bags_padded_and_csr(L, pooling, Lmax):
    rng = np.random.default_rng(Lmax)
    B, F, k, V = 65, 6, 16, 400
    lay = L.FMRankingLayer(_names(F), feature_dims=V, embedding_dims=k, pad_id=0, pooling=pooling)
    lens = rng.integers(0 if Lmax > 1 else 1, Lmax + 1, size=(B, F))      # empty bags included
    X = np.zeros((B, F, Lmax), dtype=np.int64)
    for b in range(B):
        for f in range(F):
            X[b, f, : lens[b, f]] = rng.integers(1, V, size=lens[b, f])
    out = lay(torch.tensor(X))["output"]
    ref = oracle_fm(lay, torch.float64).call(torch.tensor(X))["output"].detach().numpy()
    assert_close(out.cpu().numpy(), ref, FP32_RTOL, "padded bags")
    # CSR form of the same bags
    from etr_b200 import IdsBatch
    vals = np.concatenate([X[b, f, : lens[b, f]] for b in range(B) for f in range(F)] + [np.zeros(0, np.int64)])
    offs = np.concatenate([[0], np.cumsum(lens.ravel())]).astype(np.int32)
    ids = IdsBatch.from_csr(lay.rt, vals, offs, B, F, pooling=pooling)
    out_csr = lay(ids)["output"]
    assert_close(out_csr.cpu().numpy(), ref, FP32_RTOL, "CSR bags")


def test_bag_L1_is_single_hot(L):
    lay = L.FMRankingLayer(_names(5), feature_dims=100, embedding_dims=16, pad_id=None)
    X = np.random.default_rng(1).integers(0, 100, size=(40, 5))
    assert torch.equal(lay(torch.tensor(X))["output"], lay(torch.tensor(X[:, :, None]))["output"])


# ------------------------------------------------------------- FM backward
def _fm_grad_check(L, lay, orc, X, rtol):
    B = X.shape[0]
    rng = np.random.default_rng(7)
    dz = rng.normal(size=(B,)).astype(np.float32)
    lay(torch.tensor(X), training=True)
    grads = lay.backward(torch.tensor(dz).cuda())
    ids, rows = grads[0].indexed_slices()
    z = orc.logit(torch.tensor(X))
    (z.squeeze(1) * torch.tensor(dz, dtype=z.dtype)).sum().backward()
    k = lay.embedding_dims
    full = torch.cat([orc.embed.grad, orc.w.grad], dim=1)
    ref_ids, ref_rows = dense_table_grad_to_slices(full)
    got_ids = ids.cpu().numpy()
    keep = np.abs(rows.cpu().numpy()).sum(1) > 0
    assert np.array_equal(got_ids[keep], ref_ids)               # ID routing bit-exact
    assert_close(rows.cpu().numpy()[keep], ref_rows, rtol, "table grads (IndexedSlices, dedup'd)", grad=True)
    assert_close(lay.params.g("bias").cpu().numpy(), orc.bias.grad.numpy(), rtol, "bias grad", grad=True)


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("B,F,k,V", [(16, 3, 16, 20), (4096, 26, 16, 160000), (300, 5, 8, 64), (128, 26, 64, 3000),
                                      (6000, 3, 16, 30)])
def test_fm_backward_matches_autograd(L, B, F, k, V, fused):
    """fused=True: backward + segment reduction in one pass over the sorted runs (no materialised row
    gradients; (6000,3,16,30) has runs of thousands of occurrences -> the chunked long-run path);
    fused=False: the generic per-bag gradient rows + segment reduction."""
    rng = np.random.default_rng(B)
    lay = L.FMRankingLayer(_names(F), feature_dims=V, embedding_dims=k, seed=2, fused_apply=fused)
    X = zipf_ids(rng, [V // F] * F, B)
    _fm_grad_check(L, lay, oracle_fm(lay, torch.float64), X, FP32_RTOL)
    from etr_b200.runtime import FusedFMGrad
    lay(torch.tensor(X), training=True)
    g = lay.backward(torch.zeros(B, device="cuda"))[0]
    assert isinstance(g, FusedFMGrad) == fused


def test_fm_backward_bags_mean(L):
    rng = np.random.default_rng(4)
    B, F, Lm, V = 50, 4, 7, 60
    lay = L.FMRankingLayer(_names(F), feature_dims=V, embedding_dims=16, pad_id=0, pooling="mean")
    X = rng.integers(0, V, size=(B, F, Lm))                    # zeros are pads
    _fm_grad_check(L, lay, oracle_fm(lay, torch.float64), X, FP32_RTOL)


# ------------------------------------------------- segment reduce: long runs
def test_segment_reduce_long_runs(L):
    """ids drawn from 3 values -> runs far longer than the short-run and chunk
    thresholds (64 / 1024); result must equal an fp64 scatter-add."""
    from etr_b200 import EmbeddingTable, IdsBatch, Runtime, SparseGrad
    rt = Runtime.get()
    rng = np.random.default_rng(0)
    B, F = 20000, 2
    X = np.stack([rng.integers(0, 3, size=B), 3 + (rng.random(B) ** 3 * 5000).astype(np.int64)], axis=1)
    table = EmbeddingTable(rt, 6000, 17)
    ids = IdsBatch.from_matrix(rt, X)
    g = torch.randn(B * F, table.grad_ld, device=rt.device)
    sg = SparseGrad(table, ids, g).reduce()
    uid, rows = sg.indexed_slices()
    ref = np.zeros((6000, 17))
    np.add.at(ref, X.ravel(), g.cpu().numpy().astype(np.float64)[:, :17])
    ref_ids = np.unique(X)
    assert np.array_equal(uid.cpu().numpy(), ref_ids)
    got = rows.cpu().numpy()
    # fp32 summation of up to ~7000 N(0,1) terms: error ~ 1e-7 * sqrt(n) * |terms|
    assert np.max(np.abs(got - ref[ref_ids])) < 2e-3
    sg2 = SparseGrad(table, ids, g).reduce()
    assert torch.equal(sg2.indexed_slices()[1], rows)           # deterministic


# ------------------------------------------------------------------ DeepFM
@pytest.mark.parametrize("B,F,k,V,C", [(4, 5, 8, 20, 0), (513, 26, 16, 50000, 13), (200, 5, 16, 300, 3)])
def test_deepfm_forward_backward(L, B, F, k, V, C):
    rng = np.random.default_rng(B)
    cont = [f"c{i}" for i in range(C)]
    lay = L.DeepFMRankingLayer(_names(F), feature_dims=V, embedding_dims=k, continuous_features=cont, seed=5)
    X = zipf_ids(rng, [V // F] * F, B)
    Xc = rng.normal(size=(B, C)).astype(np.float32)
    inputs = _dict_inputs(X, lay.feature_names)
    inputs.update({n: torch.tensor(Xc[:, i]) for i, n in enumerate(cont)})
    out = lay(inputs, training=True)["output"]
    orc = oracle_deepfm(lay, torch.float64)
    xc64 = torch.tensor(Xc, dtype=torch.float64) if C else None
    z = orc.logit(torch.tensor(X), xc64)
    assert_close(out.cpu().numpy(), torch.sigmoid(z).detach().numpy(), FP32_RTOL, "DeepFM output")
    orc32 = oracle_deepfm(lay, torch.float32)
    z32 = orc32.logit(torch.tensor(X), torch.tensor(Xc) if C else None)
    assert_close(out.cpu().numpy(), torch.sigmoid(z32).detach().numpy(), FP32_RTOL, "DeepFM output vs fp32")
    dz = rng.normal(size=(B,)).astype(np.float32)
    grads = lay.backward(torch.tensor(dz).cuda())
    (z.squeeze(1) * torch.tensor(dz, dtype=torch.float64)).sum().backward()
    ids, rows = grads[0].indexed_slices()
    full = torch.cat([orc.embed.grad, orc.w.grad], dim=1)
    ref_ids, ref_rows = dense_table_grad_to_slices(full)
    assert np.array_equal(ids.cpu().numpy(), ref_ids)
    assert_close(rows.cpu().numpy(), ref_rows, FP32_RTOL, "DeepFM table grads", grad=True)
    for dev_mlp, o_mlp in ((lay.MLP_layer1, orc.MLP_layer1), (lay.MLP_layer2, orc.MLP_layer2)):
        for i in range(len(dev_mlp.units)):
            assert_close(lay.params.g(f"{dev_mlp.name}/kernel_{i}").cpu().numpy(), o_mlp.kernels[i].grad.numpy(),
                         FP32_RTOL, f"{dev_mlp.name} kernel_{i} grad", grad=True)
            assert_close(lay.params.g(f"{dev_mlp.name}/bias_{i}").cpu().numpy(), o_mlp.biases[i].grad.numpy(),
                         FP32_RTOL, f"{dev_mlp.name} bias_{i} grad", grad=True)


def test_deepfm_reference_docstring(L):
    """2.FM/CustomLayers.py:242-254 docstring inputs; embedding_dims=8."""
    lay = L.DeepFMRankingLayer(embedding_dims=8)
    inp = {'user_tag0': torch.tensor([12, 13, 14, 15]), 'user_tag1': torch.tensor([16, 17, 18, 19]),
           'item_tag1': torch.tensor([0, 1, 2, 3]), 'item_tag2': torch.tensor([4, 5, 6, 7]),
           'item_tag3': torch.tensor([8, 9, 10, 11])}
    out = lay(inp)["output"]
    X = torch.stack([inp[n] for n in lay.feature_names], dim=1)
    ref = torch.sigmoid(oracle_deepfm(lay, torch.float64).logit(X)).detach().numpy()
    assert_close(out.cpu().numpy(), ref, FP32_RTOL)


# --------------------------------------------------------------- train step
@pytest.mark.parametrize("mode", ["rowwise", "keras_dense"])
@pytest.mark.parametrize("model", ["fm", "deepfm"])
def test_train_steps_match_oracle_adam(L, mode, model):
    """3 steps of the reference train loop (BCE -> tape.gradient -> Adam,
    2.FM/ModelManager.py:171-181): weights after each step vs the CPU oracle."""
    rng = np.random.default_rng(9)
    B, F, k, V = 256, 6, 16, 120
    if model == "fm":
        lay = L.FMRankingLayer(_names(F), feature_dims=V, embedding_dims=k, seed=8)
        orc = oracle_fm(lay, torch.float64)
    else:
        lay = L.DeepFMRankingLayer(_names(F), feature_dims=V, embedding_dims=k, seed=8)
        orc = oracle_deepfm(lay, torch.float64)
    tr = L.Trainer(lay, lr=1e-2, apply_mode=mode)
    opt = R.KerasAdam(lr=1e-2, mode=mode)
    for step in range(3):
        X = zipf_ids(rng, [V // F] * F, B)
        y = (rng.random(B) < 0.25).astype(np.float32)
        loss = tr.train_step(torch.tensor(X), torch.tensor(y))
        # oracle step
        for v in orc.variables():
            v.grad = None
        p = torch.sigmoid(orc.logit(torch.tensor(X)))
        l_ref = R.keras_bce(torch.tensor(y, dtype=torch.float64).reshape(-1, 1), p)
        l_ref.backward()
        lr_t = opt.step_begin()
        sparse = {id(orc.embed), id(orc.w)}
        for v in orc.variables():
            if id(v) in sparse:
                nz, rows = R.dedup_dense_grad(v.grad)
                opt.apply_sparse(v, nz, rows, lr_t)
            else:
                opt.apply_dense(v, v.grad, lr_t)
        assert abs(float(loss.item()) - float(l_ref)) <= 1e-5 * abs(float(l_ref)), step
        # Adam's update lr*m/(sqrt(v)+eps) is ill-conditioned where a row gradient cancels to ~eps, so
        # weights are compared in units of the step size: |dw| <= 2e-3 * lr per step taken
        tol = 2e-3 * 1e-2 * (step + 1)
        for name, got, ref in (("embed", lay.embed, orc.embed), ("w", lay.w, orc.w), ("bias", lay.bias, orc.bias)):
            err = np.abs(cpu(got, torch.float64).numpy() - ref.detach().numpy()).max()
            assert err <= tol, (name, step, err)


# --------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K,ta,tb", [(513, 32, 429, 0, 0), (1000, 8, 32, 0, 0), (777, 1, 8, 0, 0),
                                           (429, 32, 20000, 1, 0), (300, 429, 32, 0, 1), (130, 70, 50, 1, 1),
                                           (64, 1677, 1677, 0, 1), (64, 8, 65536, 1, 0), (32, 8, 8192, 1, 0),
                                           (56, 32, 5000, 1, 0)])
def test_gemm_f32(L, M, N, K, ta, tb):
    from etr_b200 import Runtime
    from etr_b200.runtime import gemm_f32
    rt = Runtime.get()
    g = torch.Generator(device=rt.device)
    g.manual_seed(M + N)
    A = torch.randn((K, M) if ta else (M, K), device=rt.device, generator=g)
    Bm = torch.randn((N, K) if tb else (K, N), device=rt.device, generator=g)
    bias = torch.randn(N, device=rt.device, generator=g)
    Cm = rt.empty((M, N))
    gemm_f32(rt, A, Bm, Cm, M, N, K, A.stride(0), Bm.stride(0), N, trans_a=bool(ta), trans_b=bool(tb), bias=bias,
             act="relu")
    ref = torch.relu((A.double().T if ta else A.double()) @ (Bm.double().T if tb else Bm.double()) + bias.double())
    err = (Cm.double() - ref).abs().max().item()
    assert err <= 1e-5 * ref.abs().max().item() * max(1.0, (K / 1000) ** 0.5), err


def test_fused_and_generic_apply_agree(L):
    """Two DeepFM replicas, one with the fused backward+reduce+Adam, one with materialised per-bag
    gradient rows: same weights after training steps (up to fp32 summation order)."""
    rng = np.random.default_rng(77)
    B, F, k, V, C = 2048, 8, 16, 900, 3
    names, cont = _names(F), [f"c{i}" for i in range(C)]
    lays = [L.DeepFMRankingLayer(names, feature_dims=V, embedding_dims=k, continuous_features=cont, seed=4,
                                 fused_apply=f) for f in (True, False)]
    trs = [L.Trainer(l, lr=1e-2) for l in lays]
    for step in range(3):
        X = zipf_ids(rng, [V // F] * F, B)
        d = {n: torch.tensor(X[:, i]) for i, n in enumerate(names)}
        d.update({n: torch.tensor(rng.normal(size=B).astype(np.float32)) for n in cont})
        y = torch.tensor((rng.random(B) < 0.3).astype(np.float32))
        la, lb = trs[0].train_step(d, y), trs[1].train_step(d, y)
        assert abs(float(la.item()) - float(lb.item())) < 1e-6
    assert (lays[0].table.data - lays[1].table.data).abs().max().item() < 2e-5      # < 0.2% of one Adam step (lr 1e-2)
    assert (lays[0].params.value - lays[1].params.value).abs().max().item() < 2e-5


def test_graph_trainer_matches_eager(L):
    """The captured-CUDA-graph train step (static staged inputs, device-resident
    optimizer clock) must produce exactly the eager step's weights."""
    rng = np.random.default_rng(21)
    B, F, k, V, C = 512, 6, 16, 3000, 3
    names, cont = _names(F), [f"c{i}" for i in range(C)]
    lays = [L.DeepFMRankingLayer(names, feature_dims=V, embedding_dims=k, continuous_features=cont, seed=4)
            for _ in range(2)]
    trs = [L.Trainer(lays[0], lr=1e-2, graph=False), L.Trainer(lays[1], lr=1e-2, graph=True)]
    for step in range(9):
        X = zipf_ids(rng, [V // F] * F, B)
        Xc = rng.normal(size=(B, C)).astype(np.float32)
        y = (rng.random(B) < 0.3).astype(np.float32)
        d = {n: torch.tensor(X[:, i]) for i, n in enumerate(names)}
        d.update({n: torch.tensor(Xc[:, i]) for i, n in enumerate(cont)})
        la = trs[0].train_step(d, torch.tensor(y))
        lb = trs[1].train_step(d, torch.tensor(y))
        assert float(la.item()) == float(lb.item()), step
    assert torch.equal(lays[0].table.data, lays[1].table.data)
    assert torch.equal(lays[0].params.value, lays[1].params.value)
    assert trs[0].iterations == trs[1].iterations == 9        # per buffer set: 2 eager steps, capture, replays
    # asynchronous stepping reads every loss, one step late
    hs = [trs[1].train_step_async(d, torch.tensor(y)) for _ in range(3)]
    vals = [h.result() for h in hs]
    assert all(np.isfinite(v) for v in vals) and trs[1].iterations == 12


# ------------------------------------------------------------ checkpoint / resume (tf.train.Checkpoint analogue)
@pytest.mark.parametrize("graph", [False, True])
def test_trainer_checkpoint_resume_is_bit_exact(L, graph, tmp_path):
    F, k, V, C_, B = 26, 16, 20011, 13, 512
    names, cont = _names(F), [f"c{i}" for i in range(C_)]
    rng = np.random.default_rng(7)

    def batch():
        d = {n: torch.tensor(rng.integers(0, V, size=B)) for n in names}
        d.update({n: torch.tensor(rng.normal(size=B).astype(np.float32)) for n in cont})
        return d, torch.tensor((rng.random(B) < 0.3).astype(np.float32))

    data = [batch() for _ in range(8)]
    mk = lambda: L.DeepFMRankingLayer(names, V, k, continuous_features=cont, seed=5, mlp_precision="bf16")   # noqa: E731
    lay = mk()
    tr = L.Trainer(lay, lr=1e-2, graph=graph)
    for d, y in data[:4]:
        tr.train_step(d, y)
    path = str(tmp_path / "ckpt.pt")
    tr.save(path)
    ref = [float(tr.train_step(d, y).item()) for d, y in data[4:]]
    lay2 = mk()
    tr2 = L.Trainer(lay2, lr=1e-2, graph=graph)
    tr2.restore(path)
    assert tr2.iterations == 4
    got = [float(tr2.train_step(d, y).item()) for d, y in data[4:]]
    assert got == ref                                                        # same losses, bit for bit
    torch.cuda.synchronize()
    assert torch.equal(lay2.table.data, lay.table.data) and torch.equal(lay2.params.value, lay.params.value)
    assert torch.equal(lay2.table.m, lay.table.m) and torch.equal(lay2.params.v, lay.params.v)


def test_plan_ahead_unsharded_is_bit_exact(L):
    """Trainer(plan_ahead=True): the sorted-id plan of batch i+1 is built on the side stream during step i
    (train_step(batch, next_batch=staged), CUDA graph, 3 buffer sets).  Same kernels on the same plans: weights, Adam slots
    and losses equal those of the plain graph trainer bit for bit, over enough steps to replay every captured variant."""
    F, k, V, C_, B = 26, 16, 20011, 13, 1024
    names, cont = _names(F), [f"c{i}" for i in range(C_)]
    rng = np.random.default_rng(11)

    def batch():
        d = {n: torch.tensor(rng.integers(0, V, size=B)) for n in names}
        d.update({n: torch.tensor(rng.normal(size=B).astype(np.float32)) for n in cont})
        return d, torch.tensor((rng.random(B) < 0.3).astype(np.float32))

    data = [batch() for _ in range(16)]
    mk = lambda: L.DeepFMRankingLayer(names, V, k, continuous_features=cont, seed=5, mlp_precision="bf16")   # noqa: E731
    lay_a, lay_b = mk(), mk()
    tr_a = L.Trainer(lay_a, lr=1e-2, graph=True)
    tr_b = L.Trainer(lay_b, lr=1e-2, graph=True, plan_ahead=True)
    assert tr_b.plan_ahead and tr_b.depth == 3 and not tr_a.plan_ahead
    ref = [float(tr_a.train_step(d, y).item()) for d, y in data]
    got = []
    cur = tr_b.stage(*data[0])
    for i in range(len(data)):
        nxt = tr_b.stage(*data[i + 1]) if i + 1 < len(data) else None
        got.append(float(tr_b.train_step(cur, None, nxt).item()))
        cur = nxt
    torch.cuda.synchronize()
    assert got == ref
    assert torch.equal(lay_b.table.data, lay_a.table.data) and torch.equal(lay_b.params.value, lay_a.params.value)
    assert torch.equal(lay_b.table.m, lay_a.table.m) and torch.equal(lay_b.table.v, lay_a.table.v)
