"""torch.autograd binding (etr_b200.autograd): ``loss.backward()`` through the layer shims == the oracle's autograd,
and the torch-style train loop (EtrModule + EtrAdam) == the oracle's Keras-Adam loop."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import reference_layers as R                                          # noqa: E402
from tests.util import (assert_close, cpu, dense_table_grad_to_slices, oracle_dcn, oracle_deepfm, oracle_ffm,    # noqa: E402
                        oracle_fm, oracle_pnn, table_slices, zipf_ids)


@pytest.fixture(scope="module")
def L():
    from etr_b200 import CustomLayers
    return CustomLayers


def _inputs(names, X, cont=(), Xc=None):
    d = {n: torch.tensor(X[:, i]) for i, n in enumerate(names)}
    for i, n in enumerate(cont):
        d[n] = torch.tensor(Xc[:, i])
    return d


@pytest.mark.parametrize("model", ["fm", "deepfm", "ffm", "fwfm", "pnn", "dcn"])
def test_loss_backward_matches_oracle_autograd(L, model):
    from etr_b200.autograd import EtrModule
    rng = np.random.default_rng(11)
    B, F, k, V, C = 300, 6, 16, 240, 3
    names, cont = [f"f{i}" for i in range(F)], [f"c{i}" for i in range(C)]
    X = zipf_ids(rng, [V // F] * F, B)
    Xc = rng.normal(size=(B, C)).astype(np.float32)
    y = (rng.random(B) < 0.3).astype(np.float32)
    if model == "fm":
        lay = L.FMRankingLayer(names, V, k, seed=2); orc = oracle_fm(lay, torch.float64); ins = _inputs(names, X)
        ref_p = lambda: torch.sigmoid(orc.logit(torch.tensor(X)))
        table = lambda: torch.cat([orc.embed.grad, orc.w.grad], 1)
    elif model == "deepfm":
        lay = L.DeepFMRankingLayer(names, V, k, continuous_features=cont, seed=2); orc = oracle_deepfm(lay, torch.float64)
        ins = _inputs(names, X, cont, Xc)
        ref_p = lambda: torch.sigmoid(orc.logit(torch.tensor(X), torch.tensor(Xc, dtype=torch.float64)))
        table = lambda: torch.cat([orc.embed.grad, orc.w.grad], 1)
    elif model in ("ffm", "fwfm"):
        k = 4
        lay = (L.FFMLayer if model == "ffm" else L.FwFMLayer)(names, V, k, seed=2); orc = oracle_ffm(lay); ins = _inputs(names, X)
        ref_p = lambda: orc.call({n: torch.tensor(X[:, i]) for i, n in enumerate(names)})["output"]
        table = lambda: torch.cat([orc.fa_interaction_layer.embedding_lookup_table.grad.reshape(V, -1), orc.w.grad], 1)
    elif model == "pnn":
        lay = L.PNNRankingLayer(names, V, k, seed=2); orc = oracle_pnn(lay); ins = _inputs(names, X)
        ref_p = lambda: orc.call({n: torch.tensor(X[:, i]) for i, n in enumerate(names)})["output"]
        table = lambda: orc.embed.grad
    else:
        lay = L.DeepCrossNetworkLayer(names, cont, feature_dims=V, embedding_dims=k, type="vec", seed=2); orc = oracle_dcn(lay)
        ins = _inputs(names, X, cont, Xc)
        o_in = {n: torch.tensor(X[:, i]) for i, n in enumerate(names)}
        o_in.update({n: torch.tensor(Xc[:, i], dtype=torch.float64) for i, n in enumerate(cont)})
        ref_p = lambda: orc.call(o_in)["output"]
        table = lambda: orc.embedding.grad
    mod = EtrModule(lay)
    out = mod(ins)["output"]
    assert out.requires_grad and out.shape == (B, 1)
    loss = torch.nn.functional.binary_cross_entropy(out.reshape(-1), torch.tensor(y).cuda())
    loss.backward()
    p = ref_p()
    l_ref = torch.nn.functional.binary_cross_entropy(p.reshape(-1), torch.tensor(y, dtype=torch.float64))
    l_ref.backward()
    assert abs(float(loss.item()) - float(l_ref)) <= 1e-5 * abs(float(l_ref))
    ids, rows = table_slices(mod.sparse_grads)
    ref_ids, ref_rows = dense_table_grad_to_slices(table())
    assert np.array_equal(ids, ref_ids)
    assert_close(rows[:, :ref_rows.shape[1]], ref_rows, 1e-5, f"{model} table grad", grad=True)
    # every dense variable's gradient, by its reference name
    if model in ("fm", "deepfm", "ffm", "fwfm"):
        assert_close(cpu(mod.grad_of("bias")).numpy(), orc.bias.grad.numpy(), 1e-5, "bias grad", grad=True)
    if model == "deepfm":
        for i in range(2):
            assert_close(cpu(mod.grad_of(f"MLP_layer1/kernel_{i}")).numpy(), orc.MLP_layer1.kernels[i].grad.numpy(), 1e-5,
                         f"kernel_{i}", grad=True)
    if model == "fwfm":
        assert_close(cpu(mod.grad_of("interaction_weights/kernel")).numpy(), orc.r.grad.numpy(), 1e-5, "r grad", grad=True)
    if model == "dcn":
        p0 = lay.front_pad
        for i in range(lay.cross_layer.layer_num):
            assert_close(cpu(mod.grad_of("cross/w"))[i][p0:].unsqueeze(1).numpy(), orc.cross_layer.cross_weight[i].grad.numpy(),
                         1e-5, f"cross w{i}", grad=True)


def test_interaction_ops_backward(L):
    from etr_b200.autograd import EtrModule
    rng = np.random.default_rng(5)
    x = rng.normal(size=(64, 5, 8)).astype(np.float32)
    g = rng.normal(size=(64, 10)).astype(np.float32)
    xt = torch.tensor(x, device="cuda", requires_grad=True)
    out = EtrModule(L.InnerProductNetwork())(xt)
    out.backward(torch.tensor(g).cuda())
    xr = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    R.inner_product_network(xr).backward(torch.tensor(g, dtype=torch.float64))
    assert_close(cpu(xt.grad).numpy(), xr.grad.numpy(), 1e-5, "IPN dx", grad=True)
    # CrossLayer: dense parameters + input gradient
    xc = rng.normal(size=(128, 20)).astype(np.float32)
    gc = rng.normal(size=(128, 20)).astype(np.float32)
    cl = L.CrossLayer(3)
    cl.build(20)
    mod = EtrModule(cl)
    xt = torch.tensor(xc, device="cuda", requires_grad=True)
    mod(xt).backward(torch.tensor(gc).cuda())
    ws = [cpu(w, torch.float64).requires_grad_(True) for w in cl.cross_weight]
    bs = [cpu(b, torch.float64).requires_grad_(True) for b in cl.cross_bias]
    xr = torch.tensor(xc, dtype=torch.float64, requires_grad=True)
    R.cross_layer(xr, ws, bs).backward(torch.tensor(gc, dtype=torch.float64))
    assert_close(cpu(xt.grad).numpy(), xr.grad.numpy(), 1e-5, "cross dx", grad=True)
    for i in range(3):
        assert_close(cpu(mod.grad_of("cross/w"))[i].unsqueeze(1).numpy(), ws[i].grad.numpy(), 1e-5, f"cross w{i}", grad=True)
    # FieldAwareInteractionLayer: sparse table gradient from a pair-vector upstream gradient
    fa = L.FieldAwareInteractionLayer(4, feature_dims=50, embedding_dims=4)
    X = rng.integers(0, 50, size=(32, 4))
    gp = rng.normal(size=(32, 6, 4)).astype(np.float32)
    mod = EtrModule(fa)
    out = mod(torch.tensor(X))
    out.backward(torch.tensor(gp).cuda())
    T = cpu(fa.embedding_lookup_table, torch.float64).requires_grad_(True)
    R.field_aware_interaction(T, torch.tensor(X)).backward(torch.tensor(gp, dtype=torch.float64))
    ids, rows = table_slices(mod.sparse_grads)
    ref_ids, ref_rows = dense_table_grad_to_slices(T.grad.reshape(50, -1))
    assert np.array_equal(ids, ref_ids)
    assert_close(rows[:, :16], ref_rows, 1e-5, "field-aware table grad", grad=True)


@pytest.mark.parametrize("mode", ["rowwise", "keras_dense"])
def test_torch_train_loop_matches_oracle_loop(L, mode):
    """EtrModule + EtrAdam driven like the reference loop == the oracle's BCE -> autograd -> KerasAdam loop."""
    from etr_b200.autograd import EtrAdam, EtrModule
    rng = np.random.default_rng(9)
    B, F, k, V = 256, 6, 16, 120
    names = [f"f{i}" for i in range(F)]
    lay = L.DeepFMRankingLayer(names, V, k, seed=8)
    orc = oracle_deepfm(lay, torch.float64)
    mod = EtrModule(lay)
    opt = EtrAdam(mod, learning_rate=1e-2, apply_mode=mode)
    ref = R.KerasAdam(lr=1e-2, mode=mode)
    for step in range(3):
        X = zipf_ids(rng, [V // F] * F, B)
        y = (rng.random(B) < 0.25).astype(np.float32)
        opt.zero_grad()
        out = mod(_inputs(names, X))["output"]
        loss = torch.nn.functional.binary_cross_entropy(out.reshape(-1), torch.tensor(y).cuda())
        loss.backward()
        opt.step()
        for v in orc.variables():
            v.grad = None
        l_ref = R.keras_bce(torch.tensor(y, dtype=torch.float64).reshape(-1, 1), torch.sigmoid(orc.logit(torch.tensor(X))))
        l_ref.backward()
        lr_t = ref.step_begin()
        sparse = {id(orc.embed), id(orc.w)}
        for v in orc.variables():
            if id(v) in sparse:
                nz, rows = R.dedup_dense_grad(v.grad)
                ref.apply_sparse(v, nz, rows, lr_t)
            else:
                ref.apply_dense(v, v.grad, lr_t)
        assert abs(float(loss.item()) - float(l_ref)) <= 2e-5 * abs(float(l_ref)), step
    tol = 2e-3 * 1e-2 * 3
    for name, got, want in (("embed", lay.embed, orc.embed), ("w", lay.w, orc.w), ("bias", lay.bias, orc.bias),
                            ("kernel_0", lay.MLP_layer1.kernels[0], orc.MLP_layer1.kernels[0])):
        assert np.abs(cpu(got, torch.float64).numpy() - want.detach().numpy()).max() <= tol, name
    assert opt.iterations == 3
