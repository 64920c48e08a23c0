"""Committed golden fixtures (tests/golden/layers_v1.npz, made by
tests/golden/make_golden.py from the fp64 oracle on seeded inputs).
CPU: the oracle still reproduces them (regression pin).
GPU: the CUDA path matches them through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import reference_layers as R

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layers_v1.npz"))
T = lambda a: torch.tensor(np.asarray(a))
NAMES6, CONT3 = [f"f{i}" for i in range(6)], [f"c{i}" for i in range(3)]


# ------------------------------------------------------------------ CPU: oracle vs fixtures
def test_oracle_reproduces_fm_deepfm_golden():
    fm = R.FMRankingLayer(NAMES6, 60, 8)
    fm.bias, fm.embed, fm.w = T(G["fm_bias"]), T(G["fm_embed"]), T(G["fm_w"])
    np.testing.assert_allclose(fm.call(T(G["fm_X"]))["output"].numpy(), G["fm_out"], rtol=1e-13)
    fmb = R.FMRankingLayer(NAMES6, 60, 8, pad_id=0, pooling="mean")
    fmb.bias, fmb.embed, fmb.w = fm.bias, fm.embed, fm.w
    np.testing.assert_allclose(fmb.call(T(G["bag_X"]))["output"].numpy(), G["bag_out"], rtol=1e-13)


def test_oracle_reproduces_interaction_golden():
    x = T(G["pnn_x"])
    np.testing.assert_allclose(R.inner_product_network(x).numpy(), G["pnn_inner"], rtol=1e-13)
    np.testing.assert_allclose(R.outer_product_network(x, T(G["pnn_Kmat"]), "mat").numpy(), G["pnn_outer_mat"], rtol=1e-13)
    xc = T(G["cross_x"])
    bs = [T(b) for b in G["cross_b"]]
    np.testing.assert_allclose(R.cross_layer(xc, [T(w) for w in G["cross_w"]], bs).numpy(), G["cross_vec_out"], rtol=1e-13)
    np.testing.assert_allclose(R.matrix_cross_layer(xc, [T(w) for w in G["cross_W"]], bs).numpy(), G["cross_mat_out"],
                               rtol=1e-13)
    fw = R.FwFMLayer([f"g{i}" for i in range(5)], 40, 4)
    fw.bias, fw.w, fw.r, fw.r0 = T(G["fwfm_bias"]), T(G["fwfm_w"]), T(G["fwfm_r"]), T(G["fwfm_r0"])
    fw.fa_interaction_layer.embedding_lookup_table = T(G["fwfm_T"])
    np.testing.assert_allclose(fw.call(T(G["fwfm_X"]))["output"].numpy(), G["fwfm_out"], rtol=1e-13)


# ------------------------------------------------------------------ GPU: CUDA path vs fixtures
def _close(got, ref, rtol, grad=False):
    from tests.util import assert_close
    assert_close(got, ref, rtol, grad=grad)


@pytest.mark.gpu
def test_cuda_fm_deepfm_match_golden():
    from etr_b200 import CustomLayers as L
    from tests.util import table_slices
    lay = L.FMRankingLayer(NAMES6, 60, 8)
    lay.set_weights(G["fm_bias"], G["fm_embed"].astype(np.float32), G["fm_w"].astype(np.float32))
    # fp32 weights are the rounded fp64 fixtures: compare at the fp32 tolerance
    out = lay(torch.tensor(G["fm_X"]), training=True)["output"]
    _close(out.cpu().numpy(), G["fm_out"], 1e-5)
    grads = lay.backward(torch.tensor(G["fm_dz"], dtype=torch.float32).cuda())
    ids, rows = table_slices(grads)
    full = np.concatenate([G["fm_gembed"], G["fm_gw"]], axis=1)
    _close(rows, full[ids], 1e-5, grad=True)
    assert set(np.nonzero(np.abs(full).sum(1))[0]) == set(ids)
    bag = L.FMRankingLayer(NAMES6, 60, 8, pad_id=0, pooling="mean")
    bag.set_weights(G["fm_bias"], G["fm_embed"].astype(np.float32), G["fm_w"].astype(np.float32))
    _close(bag(torch.tensor(G["bag_X"]))["output"].cpu().numpy(), G["bag_out"], 1e-5)
    d = L.DeepFMRankingLayer(NAMES6, 60, 8, mlp_dims=[16, 4], continuous_features=CONT3)
    d.set_weights(G["dfm_bias"], G["dfm_embed"].astype(np.float32), G["dfm_w"].astype(np.float32))
    for li, mlp in enumerate((d.MLP_layer1, d.MLP_layer2)):
        for i in range(len(mlp.units)):
            mlp.kernels[i].copy_(torch.tensor(G[f"dfm_mlp{li + 1}_k{i}"], dtype=torch.float32))
            mlp.biases[i].copy_(torch.tensor(G[f"dfm_mlp{li + 1}_b{i}"], dtype=torch.float32))
    inputs = {n: torch.tensor(G["fm_X"][:, i]) for i, n in enumerate(NAMES6)}
    inputs.update({n: torch.tensor(G["dfm_Xc"][:, i], dtype=torch.float32) for i, n in enumerate(CONT3)})
    out = d(inputs, training=True)["output"]
    _close(out.cpu().numpy(), G["dfm_out"], 1e-5)
    grads = d.backward(torch.tensor(G["fm_dz"], dtype=torch.float32).cuda())
    ids, rows = table_slices(grads)
    full = np.concatenate([G["dfm_gembed"], G["dfm_gw"]], axis=1)
    _close(rows, full[ids], 1e-5, grad=True)
    _close(d.params.g("MLP_layer1/kernel_0").cpu().numpy(), G["dfm_mlp1_gk0"], 1e-5, grad=True)


@pytest.mark.gpu
def test_cuda_interactions_match_golden():
    from etr_b200 import CustomLayers as L
    x = torch.tensor(G["pnn_x"], dtype=torch.float32)
    _close(L.InnerProductNetwork()(x).cpu().numpy(), G["pnn_inner"], 1e-5, grad=True)
    opn = L.OuterProductNetwork(5, 4, "mat")
    opn.kernel.copy_(torch.tensor(G["pnn_Kmat"], dtype=torch.float32))
    _close(opn(x).cpu().numpy(), G["pnn_outer_mat"], 1e-5, grad=True)
    for cls, wkey, okey in ((L.CrossLayer, "cross_w", "cross_vec_out"), (L.MatrixCrossLayer, "cross_W", "cross_mat_out")):
        lay = cls(3)
        lay.build(10)
        for i in range(3):
            lay.cross_weight[i].copy_(torch.tensor(G[wkey][i], dtype=torch.float32))
            lay.cross_bias[i].copy_(torch.tensor(G["cross_b"][i], dtype=torch.float32))
        _close(lay(torch.tensor(G["cross_x"], dtype=torch.float32)).cpu().numpy(), G[okey], 1e-5, grad=True)
    fw = L.FwFMLayer([f"g{i}" for i in range(5)], 40, 4)
    fw.params.set("bias", G["fwfm_bias"])
    fw.w.copy_(torch.tensor(G["fwfm_w"], dtype=torch.float32))
    fw.fa_interaction_layer.embedding_lookup_table.copy_(torch.tensor(G["fwfm_T"], dtype=torch.float32))
    fw.params.set("interaction_weights/kernel", G["fwfm_r"])
    fw.params.set("interaction_weights/bias", G["fwfm_r0"])
    _close(fw(torch.tensor(G["fwfm_X"]))["output"].cpu().numpy(), G["fwfm_out"], 1e-5)
