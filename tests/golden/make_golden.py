#!/usr/bin/env python
"""Generates tests/golden/layers_v1.npz: seeded inputs, weights and fp64 outputs /
gradients of every hot-path layer, computed by the CPU oracle
(oracle/reference_layers.py).  The reference itself cannot produce fixtures
(TensorFlow is not installable here), so these pin the
ORACLE (regression) and give the CUDA path a second, file-based target.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_layers as R  # noqa: E402

SEED = 20260


def cases():
    out = {}
    rng = np.random.default_rng(SEED)
    T = lambda a: torch.tensor(np.asarray(a))
    dt = torch.float64

    # ---- FM / DeepFM: B=32, F=6, k=8, V=60, C=3
    names, cont = [f"f{i}" for i in range(6)], [f"c{i}" for i in range(3)]
    X = rng.integers(0, 60, size=(32, 6))
    Xc = rng.normal(size=(32, 3))
    dz = rng.normal(size=(32,))
    fm = R.FMRankingLayer(names, 60, 8).init_weights(rng, dt)
    z = fm.logit(T(X))
    (z.squeeze(1) * T(dz)).sum().backward()
    out.update(fm_X=X, fm_dz=dz, fm_bias=fm.bias.detach().numpy(), fm_embed=fm.embed.detach().numpy(),
               fm_w=fm.w.detach().numpy(), fm_out=torch.sigmoid(z).detach().numpy(),
               fm_gembed=fm.embed.grad.numpy(), fm_gw=fm.w.grad.numpy(), fm_gbias=fm.bias.grad.numpy())
    dfm = R.DeepFMRankingLayer(names, 60, 8, [16, 4], continuous_features=cont).init_weights(rng, dt)
    z = dfm.logit(T(X), T(Xc))
    (z.squeeze(1) * T(dz)).sum().backward()
    out.update(dfm_Xc=Xc, dfm_bias=dfm.bias.detach().numpy(), dfm_embed=dfm.embed.detach().numpy(),
               dfm_w=dfm.w.detach().numpy(), dfm_out=torch.sigmoid(z).detach().numpy(),
               dfm_gembed=dfm.embed.grad.numpy(), dfm_gw=dfm.w.grad.numpy())
    for li, mlp in enumerate((dfm.MLP_layer1, dfm.MLP_layer2)):
        for i, (k, b) in enumerate(zip(mlp.kernels, mlp.biases)):
            out[f"dfm_mlp{li + 1}_k{i}"] = k.detach().numpy()
            out[f"dfm_mlp{li + 1}_b{i}"] = b.detach().numpy()
            out[f"dfm_mlp{li + 1}_gk{i}"] = k.grad.numpy()
            out[f"dfm_mlp{li + 1}_gb{i}"] = b.grad.numpy()

    # ---- bags: [B,F,L] with pad id 0, mean pooling
    Xb = rng.integers(0, 60, size=(16, 6, 5))
    fmb = R.FMRankingLayer(names, 60, 8, pad_id=0, pooling="mean")
    fmb.bias, fmb.embed, fmb.w = fm.bias.detach().clone(), fm.embed.detach().clone(), fm.w.detach().clone()
    out.update(bag_X=Xb, bag_out=fmb.call(T(Xb))["output"].numpy())

    # ---- FwFM: B=16, F=5, k=4, V=40
    n5 = [f"g{i}" for i in range(5)]
    Xf = rng.integers(0, 40, size=(16, 5))
    fw = R.FwFMLayer(n5, 40, 4).init_weights(rng, dt)
    z = fw.logit(T(Xf))
    z.sum().backward()
    out.update(fwfm_X=Xf, fwfm_bias=fw.bias.detach().numpy(), fwfm_w=fw.w.detach().numpy(),
               fwfm_T=fw.fa_interaction_layer.embedding_lookup_table.detach().numpy(), fwfm_r=fw.r.detach().numpy(),
               fwfm_r0=fw.r0.detach().numpy(), fwfm_out=torch.sigmoid(z).detach().numpy(),
               fwfm_gT=fw.fa_interaction_layer.embedding_lookup_table.grad.numpy(), fwfm_gr=fw.r.grad.numpy())

    # ---- PNN products on a fixed tensor
    x = rng.normal(size=(8, 5, 4))
    Km = rng.normal(size=(4, 10, 4))
    out.update(pnn_x=x, pnn_inner=R.inner_product_network(T(x)).numpy(), pnn_Kmat=Km,
               pnn_outer_mat=R.outer_product_network(T(x), T(Km), "mat").numpy())

    # ---- cross layers: B=12, D=10, 3 layers
    xc = rng.normal(size=(12, 10))
    ws = [rng.normal(0, 0.3, size=(10, 1)) for _ in range(3)]
    Ws = [rng.normal(0, 0.3, size=(10, 10)) for _ in range(3)]
    bs = [rng.normal(0, 0.1, size=(10, 1)) for _ in range(3)]
    out.update(cross_x=xc, cross_w=np.stack(ws), cross_W=np.stack(Ws), cross_b=np.stack(bs),
               cross_vec_out=R.cross_layer(T(xc), [T(w) for w in ws], [T(b) for b in bs]).numpy(),
               cross_mat_out=R.matrix_cross_layer(T(xc), [T(w) for w in Ws], [T(b) for b in bs]).numpy())
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "layers_v1.npz")
    np.savez_compressed(path, **cases())
    print(path, os.path.getsize(path), "bytes")
