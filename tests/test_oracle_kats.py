"""The oracle against the known-answer vectors and its own
internal consistency properties.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import kats
from oracle import reference_layers as R

T = lambda a: torch.tensor(np.asarray(a))


def test_kat1_inner_product_and_fm_second_order():
    x, ipn, fm2 = kats.kat1_inner_product()
    for dt in (torch.float32, torch.float64):
        xt = T(x).to(dt)
        assert np.array_equal(R.inner_product_network(xt).numpy(), ipn)
        assert np.array_equal(R.ipn_layer(xt).numpy(), ipn)               # vectorised twin
        _, second = R.fm_terms_from_rows(xt, torch.zeros(2, 3, 1, dtype=dt))
        assert np.array_equal(second.numpy(), fm2)
        assert np.array_equal(ipn.sum(1, keepdims=True), fm2)             # FM2 == sum_p IPN


def test_kat2_field_aware():
    Tt, X, pv, term = kats.kat2_field_aware()
    out = R.field_aware_interaction(T(Tt), T(X))
    assert np.array_equal(out.numpy(), pv)
    assert np.array_equal(out.sum((1, 2)).numpy(), term)
    # FFMRankingLayer (F tables) == FieldAwareInteractionLayer when T_i[v] = T[v,i]
    lay = R.FFMRankingLayer(["a", "b", "c"], feature_dims=6, embedding_dims=2)
    lay.bias = torch.zeros(1, dtype=torch.float64)
    lay.w = torch.zeros(6, 1, dtype=torch.float64)
    lay.embedding_list = [T(Tt[:, i, :]) for i in range(3)]
    assert np.array_equal(lay.logit(T(X)).numpy().ravel(), term)


def test_kat3_cross_vector():
    x0, w, b, out = kats.kat3_cross_vector()
    got = R.cross_layer(T(x0), [T(a) for a in w], [T(a) for a in b])
    np.testing.assert_allclose(got.numpy(), out, rtol=1e-12)


def test_kat4_cross_matrix_is_Wx_not_xW():
    x0, W, b, out = kats.kat4_cross_matrix()
    got = R.matrix_cross_layer(T(x0), [T(a) for a in W], [T(a) for a in b])
    np.testing.assert_allclose(got.numpy(), out, rtol=1e-12)
    wrong = R.matrix_cross_layer(T(x0), [T(a.T.copy()) for a in W], [T(a) for a in b])
    assert not np.allclose(wrong.numpy(), out)


def test_kat5_outer_product_mat():
    x, K, out = kats.kat5_outer_product_mat()
    assert np.array_equal(R.outer_product_network(T(x), T(K), "mat").numpy(), out)
    assert np.array_equal(np.einsum("bpc,apc,bpa->bp", x[:, [0, 0, 1]], K, x[:, [1, 2, 2]]), out)


def test_outer_vec_num_reduce_to_inner():
    rng = np.random.default_rng(1)
    x = T(rng.normal(size=(5, 4, 3)))
    P = 6
    ipn = R.inner_product_network(x)
    assert torch.allclose(R.outer_product_network(x, torch.ones(P, 3, dtype=x.dtype), "vec"), ipn)
    assert torch.allclose(R.outer_product_network(x, torch.ones(P, 1, dtype=x.dtype), "num"), ipn)


def test_bag_L1_equals_single_hot_and_mean():
    rng = np.random.default_rng(2)
    table = T(rng.normal(size=(11, 4)))
    X = T(rng.integers(1, 11, size=(6, 3)))
    a = R.embedding_lookup(table, X)
    b = R.pooled_lookup(table, X.unsqueeze(-1), pad_id=0)
    assert torch.equal(a, b)
    Xb = T(np.array([[[1, 2, 0], [3, 0, 0]]]))
    s = R.pooled_lookup(table, Xb, pad_id=0, mode="sum")
    m = R.pooled_lookup(table, Xb, pad_id=0, mode="mean")
    assert torch.allclose(s[0, 0], table[1] + table[2]) and torch.allclose(m[0, 0], (table[1] + table[2]) / 2)
    assert torch.allclose(m[0, 1], table[3])


def test_fm_layer_matches_closed_form_and_autograd_formula():
    rng = np.random.default_rng(3)
    lay = R.FMRankingLayer(["a", "b", "c", "d"], feature_dims=30, embedding_dims=5).init_weights(rng, torch.float64)
    X = T(rng.integers(0, 30, size=(7, 4)))
    z = lay.logit(X)
    v = lay.embed[X]
    pair = sum((v[:, i] * v[:, j]).sum(1) for i in range(4) for j in range(i + 1, 4))
    ref = lay.bias + lay.w[X].sum(1).squeeze(1) + pair
    assert torch.allclose(z.squeeze(1), ref)
    # SURVEY a': dL/dv_f = g (S - v_f)
    g = T(rng.normal(size=(7, 1)))
    (z * g).sum().backward()
    S = v.sum(1, keepdim=True)
    occ = (g.unsqueeze(-1) * (S - v)).detach().reshape(-1, 5).numpy()
    ids, rows = R.indexed_slices_dedup(X.numpy().ravel(), occ)
    np.testing.assert_allclose(lay.embed.grad.numpy()[ids], rows, rtol=1e-10, atol=1e-12)


def test_indexed_slices_dedup_first_occurrence_order():
    ids, rows = R.indexed_slices_dedup(np.array([5, 2, 5, 9, 2]), np.arange(10.0).reshape(5, 2))
    assert ids.tolist() == [5, 2, 9]
    assert rows.tolist() == [[4.0, 6.0], [10.0, 12.0], [6.0, 7.0]]


def test_keras_adam_modes_agree_on_touched_rows_first_step():
    var_a = torch.ones(6, 2, dtype=torch.float64)
    var_b = var_a.clone()
    ids = torch.tensor([1, 4])
    g = torch.tensor([[0.5, -1.0], [2.0, 0.25]], dtype=torch.float64)
    a, b = R.KerasAdam(mode="keras_dense"), R.KerasAdam(mode="rowwise")
    a.apply_sparse(var_a, ids, g, a.step_begin())
    b.apply_sparse(var_b, ids, g, b.step_begin())
    assert torch.allclose(var_a, var_b)
    # second step with other rows: the dense mode keeps moving rows 1 and 4, the lazy one does not
    ids2 = torch.tensor([0])
    before = var_b[1].clone()
    a.apply_sparse(var_a, ids2, g[:1], a.step_begin())
    b.apply_sparse(var_b, ids2, g[:1], b.step_begin())
    assert not torch.allclose(var_a[1], var_b[1]) and torch.equal(var_b[1], before)


def test_keras_bce_matches_definition():
    p = torch.tensor([[0.9], [0.2], [1.0], [0.0]])
    y = torch.tensor([[1.0], [0.0], [1.0], [1.0]])
    got = float(R.keras_bce(y, p))
    e = 1e-7
    pc = np.clip(p.numpy().astype(np.float64), e, 1 - e)
    exp = np.mean(-(y.numpy() * np.log(pc + e) + (1 - y.numpy()) * np.log(1 - pc + e)))
    assert abs(got - exp) < 1e-5 * abs(exp)


# ---------------------------------------------------------------- round 2: gradient / Adam / BCE KATs
def test_kat6_fm_gradient():
    import torch
    from oracle import kats, reference_layers as R
    embed, w, bias, X, dz, logit, g_embed, g_w, g_bias = kats.kat6_fm_gradient()
    fm = R.FMRankingLayer(["a", "b"], 4, 2)
    fm.bias = torch.tensor(bias, requires_grad=True)
    fm.embed = torch.tensor(embed, requires_grad=True)
    fm.w = torch.tensor(w, requires_grad=True)
    z = fm.logit(torch.tensor(X))
    assert np.array_equal(z.detach().numpy(), logit)
    (z.squeeze(1) * torch.tensor(dz)).sum().backward()
    assert np.array_equal(fm.embed.grad.numpy(), g_embed)
    assert np.array_equal(fm.w.grad.numpy(), g_w)
    assert np.array_equal(fm.bias.grad.numpy(), g_bias)


def test_kat7_ffm_pair_gradient():
    import torch
    from oracle import kats, reference_layers as R
    T, X, g = kats.kat7_ffm_pair_gradient()
    Tt = torch.tensor(T, requires_grad=True)
    R.field_aware_interaction(Tt, torch.tensor(X)).sum().backward()
    assert np.array_equal(Tt.grad.numpy(), g)


@pytest.mark.parametrize("mode", ["keras_dense", "rowwise"])
def test_kat8_keras_adam(mode):
    import torch
    from oracle import kats, reference_layers as R
    lr, steps, expect = kats.kat8_keras_adam()
    var = torch.tensor([[1.0], [2.0]], dtype=torch.float64)
    opt = R.KerasAdam(lr=lr, mode=mode)
    for (idx, vals), (e_var, e_m, e_v) in zip(steps, expect[mode]):
        uniq, summed = R.indexed_slices_dedup(idx, vals)
        opt.apply_sparse(var, torch.tensor(uniq), torch.tensor(summed), opt.step_begin())
        m, v = opt._slots(var)
        np.testing.assert_allclose(var.numpy().ravel(), e_var, rtol=1e-13)
        np.testing.assert_allclose(m.numpy().ravel(), e_m, rtol=1e-13, atol=1e-300)
        np.testing.assert_allclose(v.numpy().ravel(), e_v, rtol=1e-13, atol=1e-300)


def test_kat9_keras_bce():
    import torch
    from oracle import kats, reference_layers as R
    p, y, expect = kats.kat9_keras_bce()
    got = float(R.keras_bce(torch.tensor(y).reshape(-1, 1), torch.tensor(p).reshape(-1, 1)))
    assert abs(got - expect) < 1e-13
