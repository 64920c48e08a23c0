"""GPU parity: field-aware pair interaction (FFM / FwFM), PNN products, DCN
cross layers and the DCN model -- CUDA path vs the CPU oracle, plus the
known-answer vectors of SURVEY 8c run through the kernels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import kats                                     # noqa: E402
from oracle import reference_layers as R                    # noqa: E402
from tests.util import (assert_close, cpu, dense_table_grad_to_slices, leaf, oracle_dcn, oracle_ffm,  # noqa: E402
                        oracle_ffm_ranking, oracle_mlp, oracle_pnn, table_slices, zipf_ids)

RTOL = 1e-5


@pytest.fixture(scope="module")
def L():
    from etr_b200 import CustomLayers
    return CustomLayers


def _names(F):
    return [f"f{i}" for i in range(F)]


# ------------------------------------------------------------- KATs on GPU
def test_kat1_inner_product_gpu(L):
    x, ipn, _ = kats.kat1_inner_product()
    out = L.InnerProductNetwork()(torch.tensor(x, dtype=torch.float32))
    assert np.array_equal(out.cpu().numpy(), ipn)
    assert np.array_equal(L.IpnLayer()(torch.tensor(x, dtype=torch.float32)).cpu().numpy(), ipn)


def test_kat2_field_aware_gpu(L):
    T, X, pv, term = kats.kat2_field_aware()
    lay = L.FieldAwareInteractionLayer(3, feature_dims=6, embedding_dims=2)
    lay.embedding_lookup_table.copy_(torch.tensor(T, dtype=torch.float32))
    out = lay(torch.tensor(X))
    assert out.shape == (2, 3, 2)
    assert np.array_equal(out.cpu().numpy(), pv)
    # the fused FFM head on the same table: sigma(bias + 0 + term) with bias = w = 0
    ffm = L.FFMLayer(["a", "b", "c"], feature_dims=6, embedding_dims=2)
    ffm.fa_interaction_layer.embedding_lookup_table.copy_(torch.tensor(T, dtype=torch.float32) * 1e-3)
    ffm.w.zero_()
    ffm.params.set("bias", [0.0])
    p = ffm(torch.tensor(X))["output"].cpu().numpy().ravel()
    assert_close(p, 1 / (1 + np.exp(-term * 1e-6)), RTOL)


def test_kat3_cross_vector_gpu(L):
    x0, w, b, out = kats.kat3_cross_vector()
    lay = L.CrossLayer(2)
    lay.build(2)
    for i in range(2):
        lay.cross_weight[i].copy_(torch.tensor(w[i], dtype=torch.float32))
        lay.cross_bias[i].copy_(torch.tensor(b[i], dtype=torch.float32))
    got = lay(torch.tensor(x0, dtype=torch.float32))
    assert_close(got.cpu().numpy(), out, 1e-6)


def test_kat4_cross_matrix_gpu(L):
    x0, W, b, out = kats.kat4_cross_matrix()
    lay = L.MatrixCrossLayer(2)
    lay.build(2)
    for i in range(2):
        lay.cross_weight[i].copy_(torch.tensor(W[i], dtype=torch.float32))
        lay.cross_bias[i].copy_(torch.tensor(b[i], dtype=torch.float32))
    got = lay(torch.tensor(x0, dtype=torch.float32))
    assert_close(got.cpu().numpy(), out, 1e-6)               # y = W x, not x W


def test_kat5_outer_product_mat_gpu(L):
    x, K, out = kats.kat5_outer_product_mat()
    lay = L.OuterProductNetwork(3, 2, 'mat')
    lay.kernel.copy_(torch.tensor(K, dtype=torch.float32))
    got = lay(torch.tensor(x, dtype=torch.float32))
    assert np.array_equal(got.cpu().numpy(), out)


# ----------------------------------------------------------------- FFM/FwFM
@pytest.mark.parametrize("cls", ["FFMLayer", "FwFMLayer"])
@pytest.mark.parametrize("B,F,k,V", [(8, 5, 16, 20), (300, 39, 8, 5000), (64, 26, 16, 2000)])
def test_ffm_fwfm_forward_backward(L, cls, B, F, k, V):
    rng = np.random.default_rng(B + F)
    lay = getattr(L, cls)(_names(F), feature_dims=V, embedding_dims=k, seed=2)
    X = zipf_ids(rng, [V // F] * F, B)
    out = lay({n: torch.tensor(X[:, i]) for i, n in enumerate(lay.feature_names)}, training=True)["output"]
    orc = oracle_ffm(lay)
    z = orc.logit(torch.tensor(X))
    assert_close(out.cpu().numpy(), torch.sigmoid(z).detach().numpy(), RTOL, cls)
    dz = rng.normal(size=(B,)).astype(np.float32)
    grads = lay.backward(torch.tensor(dz).cuda())
    (z.squeeze(1) * torch.tensor(dz, dtype=torch.float64)).sum().backward()
    Tg = orc.fa_interaction_layer.embedding_lookup_table.grad.reshape(V, F * k)
    full = torch.cat([Tg, orc.w.grad], dim=1)
    ref_ids, ref_rows = dense_table_grad_to_slices(full)
    ids, rows = table_slices(grads)
    assert np.array_equal(ids, ref_ids)
    assert_close(rows, ref_rows, RTOL, "pair-table grads", grad=True)
    assert_close(lay.params.g("bias").cpu().numpy(), orc.bias.grad.numpy(), RTOL, "bias", grad=True)
    if cls == "FwFMLayer":
        assert_close(lay.params.g("interaction_weights/kernel").cpu().numpy(), orc.r.grad.numpy(), RTOL, "r",
                     grad=True)
        assert_close(lay.params.g("interaction_weights/bias").cpu().numpy(), orc.r0.grad.numpy(), RTOL, "r0",
                     grad=True)


def test_ffm_ranking_equals_field_aware_form(L):
    """FFMRankingLayer (F tables, Python double loop in the reference) is the
    same function as the vectorised FFMLayer when T_i[v] = T[v,i]."""
    rng = np.random.default_rng(3)
    B, F, k, V = 100, 6, 16, 300
    lay = L.FFMRankingLayer(_names(F), feature_dims=V, embedding_dims=k, seed=1)
    X = rng.integers(0, V, size=(B, F))
    out = lay(torch.tensor(X), training=True)["output"]
    orc = oracle_ffm_ranking(lay)
    z = orc.logit(torch.tensor(X))
    assert_close(out.cpu().numpy(), torch.sigmoid(z).detach().numpy(), RTOL)
    assert [tuple(v.shape) for v in lay.variables] == [(1,), (V, 1)] + [(V, k)] * F
    dz = rng.normal(size=(B,)).astype(np.float32)
    grads = lay.backward(torch.tensor(dz).cuda())
    (z.squeeze(1) * torch.tensor(dz, dtype=torch.float64)).sum().backward()
    full = torch.cat([t.grad for t in orc.embedding_list] + [orc.w.grad], dim=1)
    ref_ids, ref_rows = dense_table_grad_to_slices(full)
    ids, rows = table_slices(grads)
    assert np.array_equal(ids, ref_ids)
    assert_close(rows, ref_rows, RTOL, "FFM tables", grad=True)


@pytest.mark.parametrize("pooling", ["sum", "mean"])
def test_ffm_multi_hot_bags(L, pooling):
    """c4 shape in small: 39 fields, k=8, bags of up to 50 ids (pad id 0)."""
    rng = np.random.default_rng(8)
    B, F, k, V, Lm = 24, 39, 8, 3000, 50
    lay = L.FwFMLayer(_names(F), feature_dims=V, embedding_dims=k, pad_id=0, pooling=pooling, seed=6)
    lens = rng.integers(1, Lm + 1, size=(B, F))
    X = np.zeros((B, F, Lm), dtype=np.int64)
    for b in range(B):
        for f in range(F):
            X[b, f, : lens[b, f]] = rng.integers(1, V, size=lens[b, f])
    out = lay(torch.tensor(X), training=True)["output"]
    orc = oracle_ffm(lay)
    z = orc.logit(torch.tensor(X))
    assert_close(out.cpu().numpy(), torch.sigmoid(z).detach().numpy(), RTOL, "FwFM bags")
    dz = rng.normal(size=(B,)).astype(np.float32)
    grads = lay.backward(torch.tensor(dz).cuda())
    (z.squeeze(1) * torch.tensor(dz, dtype=torch.float64)).sum().backward()
    Tg = orc.fa_interaction_layer.embedding_lookup_table.grad.reshape(V, F * k)
    full = torch.cat([Tg, orc.w.grad], dim=1)
    ref_ids, ref_rows = dense_table_grad_to_slices(full)
    ids, rows = table_slices(grads)
    keep = np.abs(rows).sum(1) > 0
    assert np.array_equal(ids[keep], ref_ids)
    assert_close(rows[keep], ref_rows, RTOL, "bag pair-table grads", grad=True)


def test_field_aware_layer_backward_pairvec(L):
    rng = np.random.default_rng(5)
    B, F, k, V = 40, 5, 16, 50
    lay = L.FieldAwareInteractionLayer(F, feature_dims=V, embedding_dims=k, seed=9)
    X = rng.integers(0, V, size=(B, F))
    out = lay(torch.tensor(X), training=True)
    T = leaf(lay.embedding_lookup_table, torch.float64)
    ref = R.field_aware_interaction(T, torch.tensor(X))
    assert_close(out.cpu().numpy(), ref.detach().numpy(), RTOL)
    g = rng.normal(size=ref.shape).astype(np.float32)
    grads = lay.backward(torch.tensor(g).cuda())
    (ref * torch.tensor(g, dtype=torch.float64)).sum().backward()
    ref_ids, ref_rows = dense_table_grad_to_slices(T.grad.reshape(V, F * k))
    ids, rows = table_slices(grads)
    keep = np.abs(rows).sum(1) > 0
    assert np.array_equal(ids[keep], ref_ids)
    assert_close(rows[keep], ref_rows, RTOL, "v grads", grad=True)


# ---------------------------------------------------------------------- PNN
@pytest.mark.parametrize("method,kt", [("inner", None), ("outer", "mat"), ("outer", "vec"), ("outer", "num")])
def test_pnn_forward_backward(L, method, kt):
    rng = np.random.default_rng(12)
    B, F, k, V = 130, 7, 8, 400
    lay = L.PNNRankingLayer(_names(F), feature_dims=V, embedding_dims=k, method=method, kernel_type=kt, seed=3)
    X = zipf_ids(rng, [V // F] * F, B)
    out = lay(torch.tensor(X), training=True)["output"]
    orc = oracle_pnn(lay)
    ref = orc.call(torch.tensor(X))["output"]
    assert_close(out.cpu().numpy(), ref.detach().numpy(), RTOL, "PNN output")
    # upstream gradient w.r.t. the pre-sigmoid logit
    dz = rng.normal(size=(B,)).astype(np.float32)
    grads = lay.backward(torch.tensor(dz).cuda())
    p = ref.squeeze(1)
    z = torch.log(p) - torch.log1p(-p)
    (z * torch.tensor(dz, dtype=torch.float64)).sum().backward()
    ref_ids, ref_rows = dense_table_grad_to_slices(orc.embed.grad)
    ids, rows = table_slices(grads)
    assert np.array_equal(ids, ref_ids)
    assert_close(rows, ref_rows, RTOL, "PNN embed grads", grad=True)
    if method == "outer":
        assert_close(lay.params.g("pn/kernel").cpu().numpy(), orc.kernel.grad.numpy(), RTOL, "pn kernel", grad=True)
    assert_close(lay.params.g("MLP_layer1/kernel_0").cpu().numpy(), orc.MLP_layer1.kernels[0].grad.numpy(), RTOL,
                 "mlp1 k0", grad=True)


def test_pnn_layers_criteo_shape(L):
    """F=26, k=16 (P=325): inner products; FM second order == sum_p IPN."""
    rng = np.random.default_rng(1)
    x = rng.normal(size=(50, 26, 16)).astype(np.float32)
    out = L.InnerProductNetwork()(torch.tensor(x))
    ref = R.inner_product_network(torch.tensor(x, dtype=torch.float64))
    assert_close(out.cpu().numpy(), ref.numpy(), RTOL, grad=True)
    _, second = R.fm_terms_from_rows(torch.tensor(x, dtype=torch.float64), torch.zeros(50, 26, 1, dtype=torch.float64))
    assert_close(out.sum(1, keepdim=True).cpu().numpy(), second.numpy(), 1e-4, grad=True)


# -------------------------------------------------------------------- cross
@pytest.mark.parametrize("cls,D,B", [("CrossLayer", 163, 257), ("CrossLayer", 1677, 64), ("MatrixCrossLayer", 163, 130),
                                      ("CrossLayer", 7, 5), ("MatrixCrossLayer", 20, 33)])
def test_cross_layers_forward_backward(L, cls, D, B):
    rng = np.random.default_rng(D)
    lay = getattr(L, cls)(3, seed=4)
    lay.build(D)
    for b_ in lay.cross_bias:                           # zeros in the reference; make them count
        b_.copy_(torch.tensor(rng.normal(size=(D, 1)) * 0.1, dtype=torch.float32))
    x = rng.normal(size=(B, D)).astype(np.float32)
    out = lay(torch.tensor(x), training=True)
    ws = [leaf(w, torch.float64) for w in lay.cross_weight]
    bs = [leaf(b_, torch.float64) for b_ in lay.cross_bias]
    x64 = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    fn = R.matrix_cross_layer if lay.matrix else R.cross_layer
    ref = fn(x64, ws, bs)
    assert_close(out.cpu().numpy(), ref.detach().numpy(), RTOL, cls, grad=True)
    g = rng.normal(size=(B, D)).astype(np.float32)
    dx = lay.backward(torch.tensor(g).cuda())
    (ref * torch.tensor(g, dtype=torch.float64)).sum().backward()
    assert_close(dx.cpu().numpy(), x64.grad.numpy(), RTOL, "dx0", grad=True)
    key = "cross/W" if lay.matrix else "cross/w"
    for i in range(3):
        gw = lay.params.g(key)[i].cpu().numpy()
        assert_close(gw.reshape(ws[i].shape), ws[i].grad.numpy(), RTOL, f"dW{i}", grad=True)
        assert_close(lay.params.g("cross/b")[i].cpu().numpy().reshape(-1, 1), bs[i].grad.numpy(), RTOL, f"db{i}",
                     grad=True)


# ---------------------------------------------------------------------- DCN
@pytest.mark.parametrize("type_", ["vec", "matrix"])
def test_dcn_forward_backward(L, type_):
    """Reference default shape: 10 categorical + 3 continuous, k=16 -> D=163."""
    rng = np.random.default_rng(2)
    B, V = 200, 5000
    lay = L.DeepCrossNetworkLayer(feature_dims=V, type=type_, seed=7)
    F, C = len(lay.categorical_features), len(lay.continuous_features)
    X = zipf_ids(rng, [V // F] * F, B)
    Xc = rng.normal(size=(B, C)).astype(np.float32)
    inputs = {n: torch.tensor(X[:, i]) for i, n in enumerate(lay.categorical_features)}
    inputs.update({n: torch.tensor(Xc[:, i]) for i, n in enumerate(lay.continuous_features)})
    out = lay(inputs, training=True)["output"]
    orc = oracle_dcn(lay)
    o_in = {n: torch.tensor(X[:, i]) for i, n in enumerate(lay.categorical_features)}
    o_in.update({n: torch.tensor(Xc[:, i], dtype=torch.float64) for i, n in enumerate(lay.continuous_features)})
    ref = orc.call(o_in)["output"]
    assert out.shape == (B, 1)
    assert_close(out.cpu().numpy(), ref.detach().numpy(), RTOL, f"DCN-{type_} output")
    dz = rng.normal(size=(B,)).astype(np.float32)
    grads = lay.backward(torch.tensor(dz).cuda())
    p = ref.squeeze(1)
    z = torch.log(p) - torch.log1p(-p)
    (z * torch.tensor(dz, dtype=torch.float64)).sum().backward()
    ref_ids, ref_rows = dense_table_grad_to_slices(orc.embedding.grad)
    ids, rows = table_slices(grads)
    assert np.array_equal(ids, ref_ids)
    assert_close(rows, ref_rows, RTOL, "DCN embedding grads", grad=True)
    for i in range(lay.cross_layer.layer_num):
        key = "cross/W" if type_ == "matrix" else "cross/w"
        p0 = lay.front_pad
        gw = lay.params.g(key)[i]
        gw = gw[p0:, p0:] if type_ == "matrix" else gw[p0:].unsqueeze(1)
        assert_close(gw.cpu().numpy(), orc.cross_layer.cross_weight[i].grad.numpy(), RTOL, f"cross w{i}", grad=True)
    assert_close(lay.params.g("dense_layer/kernel_0").cpu().numpy(), orc.dense_layer.kernels[0].grad.numpy(), RTOL,
                 "dense k0", grad=True)
    assert_close(lay.params.g("output_layer/kernel_0").cpu().numpy(), orc.output_layer.kernels[0].grad.numpy(), RTOL,
                 "output k0", grad=True)


def test_dcn_train_steps_decrease_loss(L):
    rng = np.random.default_rng(0)
    B, V = 512, 2000
    lay = L.DeepCrossNetworkLayer(feature_dims=V, type="matrix", seed=1)
    F, C = len(lay.categorical_features), len(lay.continuous_features)
    X = rng.integers(0, V, size=(B, F))
    inputs = {n: torch.tensor(X[:, i]) for i, n in enumerate(lay.categorical_features)}
    inputs.update({n: torch.randn(B) for n in lay.continuous_features})
    y = torch.tensor((rng.random(B) < 0.3).astype(np.float32))
    tr = L.Trainer(lay, lr=1e-2)
    losses = [float(tr.train_step(inputs, y).item()) for _ in range(8)]
    assert losses[-1] < losses[0]
