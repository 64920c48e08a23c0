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


def assert_close(got, ref, rtol, what="", grad=False):
    """Outputs: |got-ref| <= rtol * max(|ref|, 1e-2 max|ref|) -- relative, with a
    floor for entries that are cancellation residue next to the tensor's scale.
    ``grad=True`` (gradients / pre-sigmoid logits): atol = rtol * max|ref|
    per tensor -- a gradient entry is a sum of signed terms, its rounding error
    scales with the terms, not with the (possibly cancelled) result."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    floor = 1.0 if grad else 1e-2
    scale = np.maximum(np.abs(ref), np.abs(ref).max() * floor if ref.size else 0.0)
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


def oracle_ffm(layer, dtype=torch.float64):
    """FFMLayer / FwFMLayer -> oracle twin carrying the same weights."""
    cls = R.FwFMLayer if layer.fwfm else R.FFMLayer
    o = cls(layer.feature_names, layer.feature_dims, layer.embedding_dims, pad_id=layer.pad_id,
            pooling=layer.pooling)
    o.bias, o.w = leaf(layer.bias, dtype), leaf(layer.w, dtype)
    o.fa_interaction_layer.embedding_lookup_table = leaf(layer.fa_interaction_layer.embedding_lookup_table, dtype)
    if layer.fwfm:
        o.r = leaf(layer.params["interaction_weights/kernel"], dtype)
        o.r0 = leaf(layer.params["interaction_weights/bias"], dtype)
    return o


def oracle_ffm_ranking(layer, dtype=torch.float64):
    o = R.FFMRankingLayer(layer.feature_names, layer.feature_dims, layer.embedding_dims)
    o.bias, o.w = leaf(layer.bias, dtype), leaf(layer.w, dtype)
    o.embedding_list = [leaf(t, dtype) for t in layer.embedding_list]
    return o


def oracle_pnn(layer, dtype=torch.float64):
    o = R.PNNRankingLayer(layer.feature_names, layer.feature_dims, layer.embedding_dims, layer.mlp_dims,
                          method=layer.method, kernel_type=None if layer.method == "inner" else layer.kernel_type)
    o.embed = leaf(layer.embed, dtype)
    if layer.method == "outer":
        o.kernel = leaf(layer.pn_kernel, dtype)
    o.MLP_layer1, o.MLP_layer2 = oracle_mlp(layer.MLP_layer1, dtype), oracle_mlp(layer.MLP_layer2, dtype)
    return o


def oracle_dcn(layer, dtype=torch.float64):
    o = R.DeepCrossNetworkLayer(layer.categorical_features, layer.continuous_features, layer.feature_dims,
                                layer.embedding_dims, layer.units, layer.dense_layer.activation,
                                layer.cross_layer.layer_num, type=layer.type)
    o.embedding = leaf(layer.embeddings, dtype)
    o.cross_layer.cross_weight = [leaf(w, dtype) for w in layer.cross_layer.cross_weight]
    o.cross_layer.cross_bias = [leaf(b, dtype) for b in layer.cross_layer.cross_bias]
    o.dense_layer, o.output_layer = oracle_mlp(layer.dense_layer, dtype), oracle_mlp(layer.output_layer, dtype)
    return o


def table_slices(grads, width=None):
    """(ids, rows) of the deduplicated table gradient; rows that are exactly zero
    are dropped (the oracle's dense autograd gradient cannot tell them from
    untouched rows)."""
    ids, rows = grads[0].indexed_slices()
    rows = rows.cpu().numpy()
    if width is not None:
        rows = rows[:, :width]
    keep = np.abs(rows).sum(1) > 0
    return ids.cpu().numpy()[keep], rows[keep]
