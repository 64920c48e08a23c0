"""Shared helpers for the GPU parity tests: build the CPU oracle with the very
weights of a device layer, and compare."""
import numpy as np
import torch

from oracle import reference_layers as R


def cpu(t, dtype=torch.float32):
    return t.detach().to("cpu").to(dtype).clone()


def leaf(t, dtype):
    return cpu(t, dtype).requires_grad_(True)


def oracle_mlp(dev_mlp, dtype):
    o = R.MLPLayer(list(dev_mlp.units), dev_mlp.activation, dev_mlp.use_bias)
    o.kernels = [leaf(k, dtype) for k in dev_mlp.kernels]
    o.biases = [leaf(b, dtype) for b in dev_mlp.biases] if dev_mlp.use_bias else []
    return o


def oracle_fm(layer, dtype=torch.float32):
    o = R.FMRankingLayer(layer.feature_names, layer.feature_dims, layer.embedding_dims,
                         pad_id=layer.pad_id, pooling=layer.pooling)
    o.bias, o.embed, o.w = leaf(layer.bias, dtype), leaf(layer.embed, dtype), leaf(layer.w, dtype)
    return o


def oracle_deepfm(layer, dtype=torch.float32):
    o = R.DeepFMRankingLayer(layer.feature_names, layer.feature_dims, layer.embedding_dims, layer.mlp_dims,
                             continuous_features=layer.continuous_features, pad_id=layer.pad_id,
                             pooling=layer.pooling)
    o.bias, o.embed, o.w = leaf(layer.bias, dtype), leaf(layer.embed, dtype), leaf(layer.w, dtype)
    o.MLP_layer1, o.MLP_layer2 = oracle_mlp(layer.MLP_layer1, dtype), oracle_mlp(layer.MLP_layer2, dtype)
    return o


def assert_close(got, ref, rtol, what=""):
    """|got-ref| <= rtol * max(|ref|, max|ref| * 1e-2): relative, with a floor for
    entries that are cancellation residue next to the tensor's scale."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = np.maximum(np.abs(ref), np.abs(ref).max() * 1e-2 if ref.size else 0.0)
    err = np.abs(got - ref)
    bad = err > rtol * scale + 1e-30
    assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} out of tol; max rel err {np.max(err / (scale + 1e-300)):.3e}"


def zipf_ids(rng, cards, B, alpha=3.0):
    """SURVEY 8d id space: field f owns [offset_f, offset_f+card_f); rank = floor(card*u^alpha)."""
    offs = np.concatenate([[0], np.cumsum(cards)[:-1]])
    u = rng.random((B, len(cards)))
    return (offs[None, :] + np.floor(np.asarray(cards)[None, :] * u ** alpha)).astype(np.int64)


def dense_table_grad_to_slices(grad: torch.Tensor):
    nz = torch.nonzero(grad.abs().sum(dim=1) > 0).flatten()
    return nz.numpy(), grad[nz].numpy()
