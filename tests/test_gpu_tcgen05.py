"""GPU numerics of the tcgen05 / TMA / TMEM bf16 GEMM and the fused DCN-matrix
cross epilogue.  Reference = the same bf16-rounded operands multiplied in fp64
(a floating-point kernel: tolerance is about accumulation order only), plus the
CPU oracle for the cross layer within the bf16 tolerance (1e-2)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import reference_layers as R        # noqa: E402


@pytest.fixture(scope="module")
def rt():
    from etr_b200 import Runtime
    return Runtime.get()


def _rand_bf16(rt, shape, ld, seed, scale=1.0):
    g = torch.Generator(device=rt.device)
    g.manual_seed(seed)
    buf = torch.zeros((shape[0], ld), dtype=torch.bfloat16, device=rt.device)
    buf[:, : shape[1]] = (torch.randn(shape, device=rt.device, generator=g) * scale).to(torch.bfloat16)
    return buf


@pytest.mark.parametrize("M,N,K", [(128, 32, 64), (128, 128, 128), (300, 32, 432), (1000, 240, 1680),
                                    (257, 1677, 1677), (129, 64, 40), (4096, 500, 1000)])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gemm_bf16_tn(rt, M, N, K, out_dtype):
    from etr_b200.runtime import gemm_bf16_tn
    ldk = (K + 7) // 8 * 8
    A = _rand_bf16(rt, (M, K), ldk, 1)
    B = _rand_bf16(rt, (N, K), ldk, 2)
    bias = torch.randn(N, device=rt.device)
    ldc = (N + 7) // 8 * 8
    C = torch.full((M, ldc), 7.0, dtype=out_dtype, device=rt.device)
    gemm_bf16_tn(rt, A, B, C, M, N, K, bias=bias, act="relu")
    torch.cuda.synchronize()
    ref = torch.relu(A[:, :K].double() @ B[:, :K].double().T + bias.double())
    got = C[:, :N].double()
    tol = 2e-3 if out_dtype == torch.float32 else 1e-2
    err = (got - ref).abs().max().item()
    assert err <= tol * max(1.0, ref.abs().max().item()), (err, ref.abs().max().item())
    if ldc > N:
        assert torch.all(C[:, N:] == 7.0)               # padding columns untouched


def test_gemm_bf16_split_k_wgrad_shape(rt):
    """wgrad of MLP layer 1: [432, B] x [B, 32] with K = batch -> split-K, fixed-order reduce."""
    from etr_b200.runtime import gemm_bf16_tn
    M, N, K = 432, 32, 65536
    A = _rand_bf16(rt, (M, K), K, 3, scale=0.1)
    B = _rand_bf16(rt, (N, K), K, 4, scale=0.1)
    C = rt.empty((M, N))
    gemm_bf16_tn(rt, A, B, C, M, N, K)
    C2 = rt.empty((M, N))
    gemm_bf16_tn(rt, A, B, C2, M, N, K)
    ref = A.double() @ B.double().T
    err = (C.double() - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item(), err
    assert torch.equal(C, C2)                            # deterministic


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (200, 100, 1000), (1680, 1680, 4096), (432, 8, 65536), (1677 + 3, 240, 777),
                                    (64, 1, 300)])
def test_gemm_bf16_wgrad_mn_major(rt, M, N, K):
    """C = A^T B with A [K,M], B [K,N] as stored (MN-major UMMA operands: no transposes); ragged M / N / K tails."""
    from etr_b200.runtime import gemm_bf16_wgrad
    ldm, ldn = (M + 7) // 8 * 8, (N + 7) // 8 * 8
    A = _rand_bf16(rt, (K, M), ldm, 5, scale=0.1)
    B = _rand_bf16(rt, (K, N), ldn, 6, scale=0.1)
    C = torch.full((M, ldn), 7.0, device=rt.device)
    gemm_bf16_wgrad(rt, A, B, C, M, N, K)
    C2 = torch.full((M, ldn), 7.0, device=rt.device)
    gemm_bf16_wgrad(rt, A, B, C2, M, N, K)
    torch.cuda.synchronize()
    ref = A[:, :M].double().T @ B[:, :N].double()
    err = (C[:, :N].double() - ref).abs().max().item()
    assert err <= 2e-3 * max(ref.abs().max().item(), 1e-6), err
    assert torch.equal(C, C2)                            # deterministic
    if ldn > N:
        assert torch.all(C[:, N:] == 7.0)               # nothing written past the N columns


@pytest.mark.parametrize("M,N,K", [(1000, 1680, 64), (257, 240, 40), (128, 33, 8)])
def test_gemm_bf16_tn_accumulate(rt, M, N, K):
    from etr_b200.runtime import gemm_bf16_tn_accumulate
    ldk = (K + 7) // 8 * 8
    A = _rand_bf16(rt, (M, K), ldk, 7)
    B = _rand_bf16(rt, (N, K), ldk, 8)
    ldc = (N + 3) // 4 * 4 + 4
    C0 = torch.randn((M, ldc), device=rt.device)
    C = C0.clone()
    gemm_bf16_tn_accumulate(rt, A, B, C, M, N, K)
    torch.cuda.synchronize()
    ref = C0[:, :N].double() + A[:, :K].double() @ B[:, :K].double().T
    err = (C[:, :N].double() - ref).abs().max().item()
    assert err <= 1e-4 * max(ref.abs().max().item(), 1.0), err
    assert torch.equal(C[:, N:], C0[:, N:])


def test_cast_and_transpose_bf16(rt):
    from etr_b200.runtime import cast_bf16, transpose_bf16
    x = torch.randn(37, 13, device=rt.device)
    a = cast_bf16(rt, x)
    assert a.shape == (37, 16) and torch.equal(a[:, :13], x.to(torch.bfloat16)) and torch.all(a[:, 13:] == 0)
    b = cast_bf16(rt, x, transpose=True)
    assert b.shape == (13, 40) and torch.equal(b[:, :37], x.T.to(torch.bfloat16)) and torch.all(b[:, 37:] == 0)
    c = transpose_bf16(rt, a, 37, 13)
    assert torch.equal(c[:, :37], a[:, :13].T) and torch.all(c[:, 37:] == 0)


@pytest.mark.parametrize("B,D", [(256, 163), (1000, 1677), (130, 64)])
def test_cross_mat_layer_bf16(rt, B, D):
    import ctypes as C
    from etr_b200._lib import check
    ld = (D + 7) // 8 * 8
    x0 = _rand_bf16(rt, (B, D), ld, 5)
    xl = _rand_bf16(rt, (B, D), ld, 6)
    W = _rand_bf16(rt, (D, D), ld, 7, scale=0.05)
    b = torch.randn(ld, device=rt.device) * 0.1
    b[D:] = 0
    out = torch.zeros((B, ld), dtype=torch.bfloat16, device=rt.device)
    u = torch.zeros((B, ld), dtype=torch.bfloat16, device=rt.device)
    check(rt.lib.etr_cross_mat_layer_bf16(rt.ctx, x0.data_ptr(), xl.data_ptr(), ld, B, D, W.data_ptr(), ld,
                                          b.data_ptr(), out.data_ptr(), ld, u.data_ptr(), ld, rt.stream))
    torch.cuda.synchronize()
    x0d, xld, Wd, bd = x0[:, :D].double().cpu(), xl[:, :D].double().cpu(), W[:, :D].double().cpu(), b[:D].double().cpu()
    # one layer of the oracle's matrix cross with xl as the running state: x0*(W xl + b) + xl
    ref_u = xld @ Wd.T + bd
    ref = x0d * ref_u + xld
    assert (u[:, :D].double().cpu() - ref_u).abs().max().item() <= 1e-2 * max(1.0, ref_u.abs().max().item())
    assert (out[:, :D].double().cpu() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())
    # and it is the reference formula (KAT-4 orientation): matches R.matrix_cross_layer for a single layer from x0 == xl
    ref1 = R.matrix_cross_layer(x0d, [Wd], [bd.reshape(-1, 1)])
    out1 = torch.zeros_like(out)
    check(rt.lib.etr_cross_mat_layer_bf16(rt.ctx, x0.data_ptr(), x0.data_ptr(), ld, B, D, W.data_ptr(), ld,
                                          b.data_ptr(), out1.data_ptr(), ld, None, 0, rt.stream))
    assert (out1[:, :D].double().cpu() - ref1).abs().max().item() <= 1e-2 * max(1.0, ref1.abs().max().item())


def test_deepfm_bf16_mlp_matches_oracle(rt):
    """DeepFM with the first MLP layer on the tensor cores (bf16 operands, fp32
    accumulate): output and gradients vs the fp64 oracle within the bf16 tolerance."""
    from etr_b200 import CustomLayers as L
    from tests.util import assert_close, dense_table_grad_to_slices, oracle_deepfm, table_slices, zipf_ids
    rng = np.random.default_rng(3)
    B, F, k, V, C = 1000, 26, 16, 50000, 13
    names, cont = [f"f{i}" for i in range(F)], [f"c{i}" for i in range(C)]
    lay = L.DeepFMRankingLayer(names, feature_dims=V, embedding_dims=k, continuous_features=cont, seed=5,
                               mlp_precision="bf16")
    # Make every tensor-core operand bf16-representable (table rows, dense inputs, kernel_0), so the
    # fp64 oracle sees the same layer-1 pre-activations and hence the same ReLU mask; otherwise a
    # ~0.3% fraction of flipped masks alone moves a random-sign weight gradient by ~5% (not a kernel
    # property).  What is left is the bf16 rounding of the backward operand (delta_1) and fp32 accumulation.
    lay.table.data.copy_(lay.table.data.to(torch.bfloat16).float())
    lay.MLP_layer1.kernels[0].copy_(lay.MLP_layer1.kernels[0].to(torch.bfloat16).float())
    X = zipf_ids(rng, [V // F] * F, B)
    Xc = torch.tensor(rng.normal(size=(B, C)).astype(np.float32)).to(torch.bfloat16).float().numpy()
    inputs = {n: torch.tensor(X[:, i]) for i, n in enumerate(names)}
    inputs.update({n: torch.tensor(Xc[:, i]) for i, n in enumerate(cont)})
    out = lay(inputs, training=True)["output"]
    orc = oracle_deepfm(lay, torch.float64)
    z = orc.logit(torch.tensor(X), torch.tensor(Xc, dtype=torch.float64))
    assert_close(out.cpu().numpy(), torch.sigmoid(z).detach().numpy(), 1e-2, "DeepFM bf16-MLP output")
    dz = rng.normal(size=(B,)).astype(np.float32)
    grads = lay.backward(torch.tensor(dz).cuda())
    (z.squeeze(1) * torch.tensor(dz, dtype=torch.float64)).sum().backward()
    ids, rows = table_slices(grads)
    ref_ids, ref_rows = dense_table_grad_to_slices(torch.cat([orc.embed.grad, orc.w.grad], dim=1))
    assert np.array_equal(ids, ref_ids)
    assert_close(rows, ref_rows, 1e-2, "table grads (bf16 MLP)", grad=True)
    assert_close(lay.params.g("MLP_layer1/kernel_0").cpu().numpy(), orc.MLP_layer1.kernels[0].grad.numpy(), 1e-2,
                 "kernel_0 grad (tcgen05 split-K wgrad)", grad=True)
    assert_close(lay.params.g("MLP_layer1/bias_0").cpu().numpy(), orc.MLP_layer1.biases[0].grad.numpy(), 1e-2,
                 "bias_0 grad", grad=True)
    # and a few train steps run (graph capture included)
    tr = L.Trainer(lay, lr=1e-2, graph=True)
    y = torch.tensor((rng.random(B) < 0.3).astype(np.float32))
    losses = [float(tr.train_step(inputs, y).item()) for _ in range(6)]
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("B", [256, 1000])
def test_dcn_matrix_bf16_forward_backward(rt, B):
    """DeepCrossNetworkLayer(type='matrix', precision='bf16'): cross layers + wide dense layers on the
    tensor cores; output and gradients vs the fp64 oracle within the bf16 tolerance (operands made
    bf16-representable so ReLU masks agree, see test_deepfm_bf16_mlp_matches_oracle)."""
    from etr_b200 import CustomLayers as L
    from tests.util import assert_close, dense_table_grad_to_slices, oracle_dcn, table_slices, zipf_ids
    rng = np.random.default_rng(B)
    V = 5000
    lay = L.DeepCrossNetworkLayer(feature_dims=V, type="matrix", precision="bf16", seed=7)
    lay.table.data.copy_(lay.table.data.to(torch.bfloat16).float())
    lay.params.value.copy_(lay.params.value.to(torch.bfloat16).float())
    F, C = len(lay.categorical_features), len(lay.continuous_features)
    X = zipf_ids(rng, [V // F] * F, B)
    Xc = torch.tensor(rng.normal(size=(B, C)).astype(np.float32)).to(torch.bfloat16).float().numpy()
    inputs = {n: torch.tensor(X[:, i]) for i, n in enumerate(lay.categorical_features)}
    inputs.update({n: torch.tensor(Xc[:, i]) for i, n in enumerate(lay.continuous_features)})
    out = lay(inputs, training=True)["output"]
    orc = oracle_dcn(lay)
    o_in = {n: torch.tensor(X[:, i]) for i, n in enumerate(lay.categorical_features)}
    o_in.update({n: torch.tensor(Xc[:, i], dtype=torch.float64) for i, n in enumerate(lay.continuous_features)})
    ref = orc.call(o_in)["output"]
    assert_close(out.cpu().numpy(), ref.detach().numpy(), 1e-2, "DCN-matrix bf16 output")
    dz = rng.normal(size=(B,)).astype(np.float32)
    grads = lay.backward(torch.tensor(dz).cuda())
    p = ref.squeeze(1)
    z = torch.log(p) - torch.log1p(-p)
    (z * torch.tensor(dz, dtype=torch.float64)).sum().backward()
    ref_ids, ref_rows = dense_table_grad_to_slices(orc.embedding.grad)
    ids, rows = table_slices(grads)
    assert np.array_equal(ids, ref_ids)
    assert_close(rows, ref_rows, 2e-2, "DCN bf16 embedding grads", grad=True)
    p0 = lay.front_pad
    for i in range(lay.cross_layer.layer_num):
        gw = lay.params.g("cross/W")[i][p0:, p0:]
        assert_close(gw.cpu().numpy(), orc.cross_layer.cross_weight[i].grad.numpy(), 2e-2, f"cross W{i}", grad=True)
        gb = lay.params.g("cross/b")[i][p0:].unsqueeze(1)
        assert_close(gb.cpu().numpy(), orc.cross_layer.cross_bias[i].grad.numpy(), 2e-2, f"cross b{i}", grad=True)
    assert_close(lay.params.g("dense_layer/kernel_0").cpu().numpy(), orc.dense_layer.kernels[0].grad.numpy(), 2e-2,
                 "dense k0", grad=True)
    assert_close(lay.params.g("output_layer/kernel_0").cpu().numpy(), orc.output_layer.kernels[0].grad.numpy(), 2e-2,
                 "output k0", grad=True)


@pytest.mark.parametrize("B,n_in", [(1000, 432), (64, 16), (4097, 208), (130, 512)])
def test_mlp_skinny_backward_one_pass(rt, B, n_in):
    """K7b (etr_mlp_skinny_backward): dX, dK, db of a [n_in -> 32] layer in one pass == the same products
    of the bf16-rounded operands in fp64 (dX is bf16: 2^-8 relative)."""
    import ctypes as C
    from etr_b200._lib import check
    g = torch.Generator(device=rt.device)
    g.manual_seed(B + n_in)
    X = _rand_bf16(rt, (B, n_in), n_in, 3)
    dy = torch.randn((B, 32), device=rt.device, generator=g) * 0.1
    K = torch.randn((n_in, 32), device=rt.device, generator=g) * 0.2
    dX = torch.full((B, n_in), 7.0, dtype=torch.bfloat16, device=rt.device)
    dK = torch.full((n_in, 32), 7.0, device=rt.device)
    db = torch.full((32,), 7.0, device=rt.device)
    check(rt.lib.etr_mlp_skinny_backward(rt.ctx, X.data_ptr(), n_in, dy.data_ptr(), K.data_ptr(), B, n_in, 32,
                                         dX.data_ptr(), n_in, dK.data_ptr(), db.data_ptr(), rt.stream))
    torch.cuda.synchronize()
    dyb, Kb = dy.to(torch.bfloat16).double(), K.to(torch.bfloat16).double()
    ref_dX = dyb @ Kb.T
    ref_dK = X.double().T @ dyb
    ref_db = dy.double().sum(0)
    assert (dX.double() - ref_dX).abs().max().item() <= 2.0 ** -7 * ref_dX.abs().max().item()
    assert (dK.double() - ref_dK).abs().max().item() <= 1e-4 * max(ref_dK.abs().max().item(), 1.0)
    assert (db.double() - ref_db).abs().max().item() <= 1e-5 * max(ref_db.abs().max().item(), 1.0)


def test_mlp_backward_fused_skinny_matches_gemm_path(rt):
    """MLPLayer(bf16) backward: the one-pass kernel and the transpose + tcgen05 GEMM path agree."""
    from etr_b200.dense import DenseParams, MLPLayer
    outs = []
    for fused in (True, False):
        params = DenseParams(rt)
        gen = torch.Generator(device=rt.device)
        gen.manual_seed(5)
        mlp = MLPLayer(units=[32, 8], activation="relu", precision="bf16", fused_skinny=fused)
        mlp.build(429, params, gen, front_pad=3)
        params.finalize()
        g = torch.Generator(device=rt.device)
        g.manual_seed(9)
        x = torch.zeros((777, 432), dtype=torch.bfloat16, device=rt.device)
        x[:, 3:] = (torch.randn((777, 429), device=rt.device, generator=g) * 0.3).to(torch.bfloat16)
        y = mlp(x, training=True)
        dy = torch.randn(y.shape, device=rt.device, generator=g).contiguous()
        dx = mlp.backward(dy)
        torch.cuda.synchronize()
        outs.append((dx.float().clone(), params.grad.clone()))
    (dx1, g1), (dx0, g0) = outs
    assert dx1.dtype == dx0.dtype and (dx1 - dx0).abs().max().item() <= 2.0 ** -6 * dx0.abs().max().item()
    assert (g1 - g0).abs().max().item() <= 2e-3 * g0.abs().max().item()


def test_deepfm_fused_train_step_matches_layerwise(rt):
    """Trainer on DeepFM(bf16 MLP): the fused step (K7c tail + loss, K7b layer-1 backward) and the layer-by-layer
    step give the same losses and the same weights after 3 Adam steps (same bf16 products, fp32 sums reordered)."""
    from etr_b200 import CustomLayers as L
    F, k, V, C_, B = 26, 16, 5000, 13, 1000
    names, cont = [f"f{i}" for i in range(F)], [f"c{i}" for i in range(C_)]
    rng = np.random.default_rng(3)
    batches = []
    for _ in range(3):
        d = {n: torch.tensor(rng.integers(0, V, size=B)) for n in names}
        d.update({n: torch.tensor(rng.normal(size=B).astype(np.float32)) for n in cont})
        batches.append((d, torch.tensor((rng.random(B) < 0.3).astype(np.float32))))
    res = []
    for fused in (True, False):
        lay = L.DeepFMRankingLayer(names, V, k, continuous_features=cont, seed=4, mlp_precision="bf16", fused_tail=fused)
        assert lay.fused_train_ok() == fused
        tr = L.Trainer(lay, lr=1e-2)
        losses = [float(tr.train_step(d, y).item()) for d, y in batches]
        torch.cuda.synchronize()
        res.append((losses, lay.params.value.clone(), lay.table.data.clone()))
    (l1, p1, t1), (l0, p0, t0) = res
    assert max(abs(a - b) for a, b in zip(l1, l0)) < 2e-5
    assert (p1 - p0).abs().max().item() < 2e-3          # 3 Adam steps of lr 1e-2 (sign-like updates of tiny grads)
    assert (t1 - t0).abs().max().item() < 2e-3


# ---- persistent forms of the big K-major products (TMEM double-buffered; ETR_GEMM_PERSIST = 1: one CTA per tile row
# block, 2: CTA pairs with tcgen05 cta_group::2).  Same operands through the per-tile kernel (mode 0) must agree: the
# k order of the fp32 accumulation is the same, so the results are compared bit for bit as well as with fp64.
_PERSIST_MODES = [int(x) for x in __import__("os").environ.get("ETR_TEST_PERSIST_MODES", "1,2").split(",")]


@pytest.fixture
def persist_env():
    import os
    old = os.environ.get("ETR_GEMM_PERSIST")
    yield lambda m: os.environ.__setitem__("ETR_GEMM_PERSIST", str(m))
    if old is None:
        os.environ.pop("ETR_GEMM_PERSIST", None)
    else:
        os.environ["ETR_GEMM_PERSIST"] = old


@pytest.mark.parametrize("mode", _PERSIST_MODES)
@pytest.mark.parametrize("B,D", [(40000, 1677), (38400 + 130, 512), (50000, 384), (65536, 256)])
def test_cross_mat_layer_bf16_persistent(rt, persist_env, mode, B, D):
    from etr_b200._lib import check
    ld = (D + 7) // 8 * 8
    x0 = _rand_bf16(rt, (B, D), ld, 15)
    xl = _rand_bf16(rt, (B, D), ld, 16)
    W = _rand_bf16(rt, (D, D), ld, 17, scale=0.05)
    b = torch.randn(ld, device=rt.device) * 0.1
    b[D:] = 0
    res = {}
    for m in (0, mode):
        persist_env(m)
        out = torch.full((B, ld), 3.0, dtype=torch.bfloat16, device=rt.device)
        u = torch.full((B, ld), 3.0, dtype=torch.bfloat16, device=rt.device)
        check(rt.lib.etr_cross_mat_layer_bf16(rt.ctx, x0.data_ptr(), xl.data_ptr(), ld, B, D, W.data_ptr(), ld,
                                              b.data_ptr(), out.data_ptr(), ld, u.data_ptr(), ld, rt.stream))
        torch.cuda.synchronize()
        res[m] = (out, u)
    ref_u = xl[:, :D].double() @ W[:, :D].double().T + b[:D].double()
    ref = x0[:, :D].double() * ref_u + xl[:, :D].double()
    out, u = res[mode]
    eu = (u[:, :D].double() - ref_u).abs().max().item()
    eo = (out[:, :D].double() - ref).abs().max().item()
    assert eu <= 1e-2 * max(1.0, ref_u.abs().max().item()), eu
    assert eo <= 1e-2 * max(1.0, ref.abs().max().item()), eo
    if ld > D:
        assert torch.all(out[:, D:] == 3.0) and torch.all(u[:, D:] == 3.0)       # padding columns untouched
    assert torch.equal(out, res[0][0]) and torch.equal(u, res[0][1])


@pytest.mark.parametrize("mode", _PERSIST_MODES)
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gemm_bf16_tn_persistent(rt, persist_env, mode, out_dtype):
    from etr_b200.runtime import gemm_bf16_tn, gemm_bf16_tn_accumulate
    M, N, K = 40000 + 77, 1677, 200
    ldk = (K + 7) // 8 * 8
    A = _rand_bf16(rt, (M, K), ldk, 21)
    B = _rand_bf16(rt, (N, K), ldk, 22)
    bias = torch.randn(N, device=rt.device)
    ldc = (N + 7) // 8 * 8
    outs = {}
    for m in (0, mode):
        persist_env(m)
        C = torch.full((M, ldc), 7.0, dtype=out_dtype, device=rt.device)
        gemm_bf16_tn(rt, A, B, C, M, N, K, bias=bias, act="relu")
        torch.cuda.synchronize()
        outs[m] = C
    ref = torch.relu(A[:, :K].double() @ B[:, :K].double().T + bias.double())
    tol = 2e-3 if out_dtype == torch.float32 else 1e-2
    err = (outs[mode][:, :N].double() - ref).abs().max().item()
    assert err <= tol * max(1.0, ref.abs().max().item()), err
    assert torch.all(outs[mode][:, N:] == 7.0)
    assert torch.equal(outs[mode], outs[0])
    if out_dtype == torch.float32:                       # C += A B^T (the shared-input gradient of the cross backward)
        C0 = torch.randn((M, ldc), device=rt.device)
        acc = {}
        for m in (0, mode):
            persist_env(m)
            C = C0.clone()
            gemm_bf16_tn_accumulate(rt, A, B, C, M, N, K)
            torch.cuda.synchronize()
            acc[m] = C
        ref2 = C0[:, :N].double() + A[:, :K].double() @ B[:, :K].double().T
        err2 = (acc[mode][:, :N].double() - ref2).abs().max().item()
        assert err2 <= 1e-4 * max(ref2.abs().max().item(), 1.0), err2
        assert torch.equal(acc[mode], acc[0])


@pytest.mark.parametrize("B,D,slice_off", [(1000, 1680, 0), (513, 64, 8), (2048 + 3, 168, 0)])
def test_cross_backward_one_pass_kernels(rt, B, D, slice_off):
    """du = G (.) x0 with db = colsum(du) in one pass, and dx0 = sum_l G_{l+1} (.) u_l + G_0 written once, against the
    per-layer read-modify-write kernels they replace (same arithmetic order: dx0 bit-identical, du bit-identical, db to
    summation-order rounding) and against fp64."""
    import ctypes as C
    from etr_b200._lib import check
    L_ = 3
    wide = _rand_bf16(rt, (B, D + slice_off + 8), D + slice_off + 8, 31)
    Gs = [_rand_bf16(rt, (B, D), D, 40 + l) for l in range(L_)] + [wide[:, slice_off:slice_off + D]]   # G_L = a column slice
    us = [_rand_bf16(rt, (B, D), D, 50 + l) for l in range(L_)]
    x0 = _rand_bf16(rt, (B, D), D, 60)
    # (a) du + db
    G = Gs[L_]
    du = torch.empty((B, D), dtype=torch.bfloat16, device=rt.device)
    db = torch.empty((D,), device=rt.device)
    check(rt.lib.etr_cross_mat_bwd_du_colsum_bf16(rt.ctx, G.data_ptr(), G.stride(0), x0.data_ptr(), B, D, du.data_ptr(),
                                                  db.data_ptr(), rt.stream))
    du_ref = torch.empty_like(du)
    dx0_old = torch.empty((B, D), device=rt.device)
    check(rt.lib.etr_cross_mat_bwd_elementwise_bf16(rt.ctx, G.data_ptr(), G.stride(0), x0.data_ptr(), us[L_ - 1].data_ptr(), B, D,
                                                    du_ref.data_ptr(), dx0_old.data_ptr(), 1, rt.stream))
    torch.cuda.synchronize()
    assert torch.equal(du, du_ref)
    ref_db = du.double().sum(0)
    assert (db.double() - ref_db).abs().max().item() <= 1e-5 * max(1.0, ref_db.abs().max().item())
    # (b) dx0 in one pass == the layer-by-layer accumulation
    for l in reversed(range(L_ - 1)):
        tmp = torch.empty_like(du)
        check(rt.lib.etr_cross_mat_bwd_elementwise_bf16(rt.ctx, Gs[l + 1].data_ptr(), Gs[l + 1].stride(0), x0.data_ptr(),
                                                        us[l].data_ptr(), B, D, tmp.data_ptr(), dx0_old.data_ptr(), 0, rt.stream))
    check(rt.lib.etr_add_bf16_into_f32(rt.ctx, Gs[0].data_ptr(), B * D, dx0_old.data_ptr(), rt.stream))
    dx0 = torch.empty((B, D), device=rt.device)
    gp = (C.c_void_p * (L_ + 1))(*[g.data_ptr() for g in Gs])
    gl = (C.c_int64 * (L_ + 1))(*[g.stride(0) for g in Gs])
    up = (C.c_void_p * L_)(*[u.data_ptr() for u in us])
    check(rt.lib.etr_cross_mat_bwd_dx0_bf16(rt.ctx, L_, gp, gl, up, None, 0, B, D, dx0.data_ptr(), rt.stream))
    torch.cuda.synchronize()
    ref = sum(Gs[l + 1].double() * us[l].double() for l in range(L_)) + Gs[0].double()
    assert (dx0.double() - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())
    assert torch.equal(dx0, dx0_old)
    ex = _rand_bf16(rt, (B, D), D, 70)
    dx1 = torch.empty_like(dx0)
    check(rt.lib.etr_cross_mat_bwd_dx0_bf16(rt.ctx, L_, gp, gl, up, ex.data_ptr(), D, B, D, dx1.data_ptr(), rt.stream))
    torch.cuda.synchronize()
    assert torch.equal(dx1, dx0 + ex.float())


@pytest.mark.parametrize("B,n,ld", [(1000, 1688, 1688), (257, 13, 16), (64, 40, 40)])
def test_outer_bf16(rt, B, n, ld):
    """dL/dx of a Dense(1) layer as a rank-1 pass: out[b, j] = bf16(dz[b] * k[j]); padding columns untouched."""
    from etr_b200._lib import check
    g = torch.Generator(device=rt.device)
    g.manual_seed(B + n)
    dz = torch.randn(B, device=rt.device, generator=g)
    k = torch.randn(n, device=rt.device, generator=g)
    out = torch.full((B, ld), 7.0, dtype=torch.bfloat16, device=rt.device)
    check(rt.lib.etr_outer_bf16(rt.ctx, dz.data_ptr(), k.data_ptr(), B, n, out.data_ptr(), ld, rt.stream))
    torch.cuda.synchronize()
    ref = (dz[:, None] * k[None, :]).to(torch.bfloat16)
    assert torch.equal(out[:, :n], ref)
    if ld > n:
        assert torch.all(out[:, n:].float() == 0.0) or torch.all(out[:, n:] == 7.0)
