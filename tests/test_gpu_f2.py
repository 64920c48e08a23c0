"""AFM pair vectors, FiBiNet bilinear interaction, NFM bi-interaction and the ONN interaction against the
oracle restatement of the reference loops (forward and backward, fp32 1e-5), and the deterministic PNN kernel gradient."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import reference_layers as R                                       # noqa: E402
from tests.util import assert_close, cpu, dense_table_grad_to_slices, zipf_ids  # noqa: E402


@pytest.fixture(scope="module")
def L():
    from etr_b200 import CustomLayers
    return CustomLayers


@pytest.mark.parametrize("B,F,k", [(64, 5, 8), (257, 26, 16), (33, 3, 4)])
def test_afm_pair_vectors_and_bi_interaction(L, B, F, k):
    rng = np.random.default_rng(B + F)
    x = rng.normal(size=(B, F, k)).astype(np.float32)
    P = F * (F - 1) // 2
    for cls, fn, gshape in ((L.InteractionLayer, R.interaction_layer, (B, P, k)), (L.BiInteractionPooling, R.bi_interaction, (B, k))):
        lay = cls()
        out = lay(torch.tensor(x), training=True)
        xr = torch.tensor(x, dtype=torch.float64, requires_grad=True)
        ref = fn(xr)
        assert out.shape == gshape
        assert_close(cpu(out).numpy(), ref.detach().numpy(), 1e-5, cls.__name__)
        g = rng.normal(size=gshape).astype(np.float32)
        dx = lay.backward(torch.tensor(g).cuda())
        ref.backward(torch.tensor(g, dtype=torch.float64))
        assert_close(cpu(dx).numpy(), xr.grad.numpy(), 1e-5, cls.__name__ + " dx", grad=True)


def test_afm_pair_order_is_the_reference_loop_order(L):
    x = np.arange(24, dtype=np.float32).reshape(2, 3, 4)            # the InnerProductNetwork docstring tensor (KAT-1)
    out = cpu(L.InteractionLayer()(torch.tensor(x))).numpy()
    assert np.array_equal(out[:, 0], x[:, 0] * x[:, 1]) and np.array_equal(out[:, 1], x[:, 0] * x[:, 2])
    assert np.array_equal(out[:, 2], x[:, 1] * x[:, 2])
    assert np.array_equal(out.sum(2), np.array([[38.0, 62.0, 214.0], [950.0, 1166.0, 1510.0]]))     # == KAT-1
    assert np.array_equal(cpu(L.BiInteractionPooling()(torch.tensor(x))).numpy().sum(1), np.array([314.0, 3626.0]))


@pytest.mark.parametrize("btype", ["all", "each", "interaction"])
@pytest.mark.parametrize("B,F,k", [(130, 6, 8), (64, 26, 16)])
def test_bilinear_interaction(L, btype, B, F, k):
    rng = np.random.default_rng(F + k)
    x = rng.normal(size=(B, F, k)).astype(np.float32)
    lay = L.BilinearInteractionLayer(btype, seed=3)
    out = lay(torch.tensor(x), training=True)
    n_w = {"all": 1, "each": F - 1, "interaction": F * (F - 1) // 2}[btype]
    assert lay.W.shape == (n_w, k, k) and len(lay.W_list) == n_w
    assert abs(float(lay.W.std()) - (1.0 / k) ** 0.5) < 0.25 * (1.0 / k) ** 0.5          # glorot_normal scale
    W = cpu(lay.W, torch.float64).requires_grad_(True)
    xr = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    ref = R.bilinear_interaction(xr, W[0] if btype == "all" else W, btype)
    assert_close(cpu(out).numpy(), ref.detach().numpy(), 1e-5, f"bilinear {btype}")
    g = rng.normal(size=out.shape).astype(np.float32)
    dx = lay.backward(torch.tensor(g).cuda())
    ref.backward(torch.tensor(g, dtype=torch.float64))
    assert_close(cpu(dx).numpy(), xr.grad.numpy(), 1e-5, f"bilinear {btype} dx", grad=True)
    assert_close(cpu(lay.W_grad).numpy(), W.grad.numpy(), 1e-5, f"bilinear {btype} dW", grad=True)
    dx2 = lay.backward(torch.tensor(g).cuda())           # deterministic: same bits again
    assert torch.equal(dx, dx2)
    w1 = lay.W_grad.clone()
    lay(torch.tensor(x), training=True)
    lay.backward(torch.tensor(g).cuda())
    assert torch.equal(w1, lay.W_grad)


@pytest.mark.parametrize("reduce", [False, True])
def test_onn_interaction(L, reduce):
    rng = np.random.default_rng(4)
    B, F, k, V = 200, 4, 8, 80
    names = ['item_tag1', 'item_tag2', 'item_tag3', 'user_tag0']
    lay = L.ParralledOnnLayer(names, V, k, reduce=reduce, seed=2)
    X = zipf_ids(rng, [V // F] * F, B)                               # disjoint id ranges per field (DataGenerator id space)
    comb = lay.interaction({n: torch.tensor(X[:, i]) for i, n in enumerate(names)}, training=True)
    P = F * (F - 1) // 2
    assert comb.shape == (B, F * k + (P if reduce else P * k)) == (B, lay.out_dim)
    E = cpu(lay.embedding_single, torch.float64).requires_grad_(True)
    T = cpu(lay.fa_interaction_layer.embedding_lookup_table, torch.float64).requires_grad_(True)
    ref = R.onn_combined(E, T, torch.tensor(X), reduce)
    assert_close(cpu(comb).numpy(), ref.detach().numpy(), 1e-5, "ONN combined")
    # the loop form (one table pair per field pair) is the same function on this id space
    tables = {(i, j): (T[:, j, :], T[:, i, :]) for i in range(F) for j in range(i + 1, F)}
    loop = R.onn_loop_combined(E, tables, torch.tensor(X), reduce)
    assert torch.equal(loop, ref)
    g = rng.normal(size=comb.shape).astype(np.float32)
    grads = lay.interaction_backward(torch.tensor(g).cuda())
    ref.backward(torch.tensor(g, dtype=torch.float64))
    for sg, want, w in ((grads[0], E.grad, k), (grads[1], T.grad.reshape(V, -1), F * k)):
        ids, rows = sg.indexed_slices()
        rows = cpu(rows).numpy()[:, :w]
        keep = np.abs(rows).sum(1) > 0
        ref_ids, ref_rows = dense_table_grad_to_slices(want)
        assert np.array_equal(ids.cpu().numpy()[keep], ref_ids)
        assert_close(rows[keep], ref_rows, 1e-5, "ONN table grad", grad=True)
    with pytest.raises(NotImplementedError):
        lay({n: torch.tensor(X[:, i]) for i, n in enumerate(names)})
    lay.mlp_layer = lambda c: torch.sigmoid(c.sum(1, keepdim=True))
    assert lay({n: torch.tensor(X[:, i]) for i, n in enumerate(names)})["output"].shape == (B, 1)


@pytest.mark.parametrize("ktype", ["mat", "vec", "num"])
def test_pnn_outer_kernel_gradient_is_deterministic(L, ktype):
    rng = np.random.default_rng(6)
    B, F, k = 700, 6, 8
    x = torch.tensor(rng.normal(size=(B, F, k)).astype(np.float32))
    g = torch.tensor(rng.normal(size=(B, F * (F - 1) // 2)).astype(np.float32)).cuda()
    lay = L.OuterProductNetwork(F, k, ktype, seed=1)
    lay(x, training=True)
    lay.backward(g)
    first = lay.kernel_grad.clone()
    K = cpu(lay.kernel, torch.float64).requires_grad_(True)
    xr = x.double()
    R.outer_product_network(xr, K, ktype).backward(g.double().cpu())
    assert_close(cpu(first).numpy(), K.grad.numpy(), 1e-5, f"OPN {ktype} dK", grad=True)
    for _ in range(3):
        lay(x, training=True)
        lay.backward(g)
        assert torch.equal(first, lay.kernel_grad)                     # bit-identical: no atomics
