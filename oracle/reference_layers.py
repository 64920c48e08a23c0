"""Op-for-op CPU restatement of the reference Keras layers (torch-CPU tensors).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED by the
reference itself; pinned here by ``oracle/kats.py``.

Every ``call`` below follows the reference ``call()`` body op by op, in the
same order and with the same intermediate dtype, so that fp32 rounding matches
the TF eager path up to the reduction order inside a single reduce op.
``torch.autograd`` plays the role of ``tf.GradientTape``
(2.FM/ModelManager.py:172-176).  All citations are relative to the reference
repository root.

Weights are plain ``torch.Tensor`` attributes that tests inject; ``dtype``
selects fp32 (the reference's type) or fp64 (the "truth" both the oracle and
the CUDA path are compared with).
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------
# a1: input assembly idiom                         2.FM/CustomLayers.py:138-144
# --------------------------------------------------------------------------
def assemble_ids(inputs: Dict[str, Tensor], names: Sequence[str]) -> Tensor:
    """dict name -> [B] or [B,1] (or [B,L]) int64  ==>  X [B,F] int64.

    Rank-1 tensors get ``expand_dims(axis=1)``; then ``concat(axis=1)``
    (2.FM/CustomLayers.py:139-144; same idiom 3.DCN/CustomLayers.py:240-254).
    """
    cols = []
    for name in names:
        t = torch.as_tensor(inputs[name])
        if t.dim() == 1:
            t = t.unsqueeze(1)
        cols.append(t)
    return torch.cat(cols, dim=1)


# --------------------------------------------------------------------------
# a2: tf.keras.layers.Embedding / tf.nn.embedding_lookup  == plain row gather
# --------------------------------------------------------------------------
def embedding_lookup(table: Tensor, ids: Tensor) -> Tensor:
    """``out[..., :] = table[ids[...], :]`` (Keras ``Embedding.call`` ->
    ``ResourceGather``; call sites 2.FM/CustomLayers.py:146-147, 438, 490).
    TF-CPU raises InvalidArgumentError on an out-of-range id; so do we."""
    if ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= table.shape[0]):
        raise IndexError("embedding id out of range")
    return table[ids]


# --------------------------------------------------------------------------
# a17: multi-hot bag pooling -- DEFINED BY THE BUILD (no reference type).
# Nearest reference idioms: pad-id mask 7.SIM/CustomLayers.py:116, masked sum
# 7.SIM/CustomLayers.py:88-95, plain mean 5.DIN/CustomLayers.py:662.
# --------------------------------------------------------------------------
def pooled_lookup(table: Tensor, ids: Tensor, pad_id: Optional[int], mode: str = "sum") -> Tensor:
    """ids [B,F,L] (padded with ``pad_id``) -> [B,F,width].

    e_f = sum over valid l of table[id_{f,l}]   (mode 'sum'), divided by
    max(count,1) for mode 'mean'.  L=1 without padding is exactly
    :func:`embedding_lookup`.
    """
    assert mode in ("sum", "mean")
    if pad_id is None:
        valid = torch.ones_like(ids, dtype=torch.bool)
    else:
        valid = ids != pad_id
    safe = torch.where(valid, ids, torch.zeros_like(ids))
    rows = embedding_lookup(table, safe)                       # [B,F,L,w]
    rows = rows * valid.unsqueeze(-1).to(rows.dtype)
    out = rows[:, :, 0, :].clone()
    for l in range(1, ids.shape[2]):                           # sequential sum over l
        out = out + rows[:, :, l, :]
    if mode == "mean":
        cnt = valid.sum(dim=2).clamp(min=1).to(rows.dtype)
        out = out / cnt.unsqueeze(-1)
    return out


def csr_to_padded(values: np.ndarray, offsets: np.ndarray, B: int, F: int, pad_id: int):
    """CSR (values[nnz], offsets[B*F+1]) -> padded [B,F,Lmax] with pad_id."""
    lens = np.diff(offsets)
    L = max(int(lens.max()) if lens.size else 1, 1)
    out = np.full((B * F, L), pad_id, dtype=np.int64)
    for i in range(B * F):
        out[i, : lens[i]] = values[offsets[i]: offsets[i + 1]]
    return out.reshape(B, F, L)


# --------------------------------------------------------------------------
# a15: MLPLayer                                     2.FM/CustomLayers.py:15-84
#      DenseLayer                                3.DCN/CustomLayers.py:153-167
# --------------------------------------------------------------------------
_ACT = {
    None: lambda x: x,
    "linear": lambda x: x,
    "relu": torch.relu,
    "sigmoid": torch.sigmoid,
    "tanh": torch.tanh,
}


class MLPLayer:
    """per layer: act(x @ K_i + b_i)  (MatMul, BiasAdd, activation;
    2.FM/CustomLayers.py:72-84).  Batch-norm / dropout branches are never
    active on the hot path (``is_train`` is never passed, :72,82)."""

    def __init__(self, units, activation=None, use_bias=True):
        self.units = [units] if not isinstance(units, list) else list(units)
        self.activation = activation
        self.use_bias = use_bias
        self.kernels: List[Tensor] = []
        self.biases: List[Tensor] = []

    def init_weights(self, in_dim: int, rng: np.random.Generator, dtype=torch.float32):
        dims = [in_dim] + self.units
        self.kernels, self.biases = [], []
        for i in range(len(dims) - 1):
            lim = math.sqrt(6.0 / (dims[i] + dims[i + 1]))            # glorot_uniform
            k = rng.uniform(-lim, lim, size=(dims[i], dims[i + 1]))
            self.kernels.append(torch.tensor(k, dtype=dtype, requires_grad=True))
            self.biases.append(torch.zeros(dims[i + 1], dtype=dtype, requires_grad=True))
        return self

    def variables(self):
        out = []
        for i in range(len(self.kernels)):
            out.append(self.kernels[i])
            if self.use_bias:
                out.append(self.biases[i])
        return out

    def __call__(self, x: Tensor) -> Tensor:
        act = _ACT[self.activation]
        for i in range(len(self.units)):
            x = x @ self.kernels[i]
            if self.use_bias:
                x = x + self.biases[i]
            x = act(x)
        return x


DenseLayer = MLPLayer  # Dense(x, activation) chain == MLPLayer(units, activation)


# --------------------------------------------------------------------------
# a3: FMRankingLayer                              2.FM/CustomLayers.py:87-157
# --------------------------------------------------------------------------
def fm_terms(embed: Tensor, w: Tensor, X: Tensor):
    """first_order [B,1], second_order [B,1]   (2.FM/CustomLayers.py:146-153).

    ``embed`` [V,k], ``w`` [V,1]; X [B,F]  (or already pooled rows, see
    :func:`fm_terms_from_rows`)."""
    return fm_terms_from_rows(embedding_lookup(embed, X), embedding_lookup(w, X))


def fm_terms_from_rows(emb_output: Tensor, w_output: Tensor):
    first_order = torch.sum(w_output, dim=1)                               # :149
    sum_of_square = torch.sum(torch.square(emb_output), dim=1)             # :151
    square_of_sum = torch.square(torch.sum(emb_output, dim=1))             # :152
    second_order = 0.5 * torch.sum(square_of_sum - sum_of_square, dim=1, keepdim=True)  # :153
    return first_order, second_order


class FMRankingLayer:
    """sigma(bias + sum_f w[x_f] + 0.5 sum_k((sum_f v)^2 - sum_f v^2));
    variable order bias, embed, w (2.FM/CustomLayers.py:123-135)."""

    def __init__(self, feature_names, feature_dims=20, embedding_dims=16,
                 pad_id: Optional[int] = None, pooling: str = "sum"):
        self.feature_names = list(feature_names)
        self.feature_dims = feature_dims
        self.embedding_dims = embedding_dims
        self.pad_id, self.pooling = pad_id, pooling
        self.bias = self.embed = self.w = None

    def init_weights(self, rng, dtype=torch.float32):
        self.bias = torch.tensor(rng.uniform(-1.0, 1.0, size=(1,)), dtype=dtype, requires_grad=True)
        self.embed = torch.tensor(rng.uniform(-0.05, 0.05, size=(self.feature_dims, self.embedding_dims)),
                                  dtype=dtype, requires_grad=True)
        self.w = torch.tensor(rng.uniform(-0.05, 0.05, size=(self.feature_dims, 1)),
                              dtype=dtype, requires_grad=True)
        return self

    def variables(self):
        return [self.bias, self.embed, self.w]

    def _rows(self, X):
        if X.dim() == 3:                                  # build-defined bag mode (a17)
            return (pooled_lookup(self.embed, X, self.pad_id, self.pooling),
                    pooled_lookup(self.w, X, self.pad_id, self.pooling))
        return embedding_lookup(self.embed, X), embedding_lookup(self.w, X)

    def logit(self, X: Tensor) -> Tensor:
        emb, wv = self._rows(X)
        first, second = fm_terms_from_rows(emb, wv)
        return self.bias + first + second                                   # :155

    def call(self, inputs) -> Dict[str, Tensor]:
        X = inputs if isinstance(inputs, Tensor) else assemble_ids(inputs, self.feature_names)
        return {"output": torch.sigmoid(self.logit(X))}

    __call__ = call


# --------------------------------------------------------------------------
# a4: DeepFMRankingLayer                         2.FM/CustomLayers.py:241-308
# --------------------------------------------------------------------------
class DeepFMRankingLayer(FMRankingLayer):
    """FM logit (bias folded into first order, :293) + MLP([32,8],relu) ->
    MLP([1]) over Flatten(emb) (:300-301); sigma(fm+dnn) (:305).

    Build extension: optional ``continuous`` [B,C] input that is
    concatenated *in front of* the flattened embedding for the MLP only, the
    way 3.DCN/CustomLayers.py:259 does; the FM terms are untouched."""

    def __init__(self, feature_names, feature_dims=20, embedding_dims=16, mlp_dims=(32, 8),
                 continuous_features: Sequence[str] = (), **kw):
        super().__init__(feature_names, feature_dims, embedding_dims, **kw)
        self.mlp_dims = list(mlp_dims)
        self.continuous_features = list(continuous_features)
        self.MLP_layer1 = MLPLayer(units=self.mlp_dims, activation="relu")
        self.MLP_layer2 = MLPLayer(units=[1])

    def init_weights(self, rng, dtype=torch.float32):
        super().init_weights(rng, dtype)
        in_dim = len(self.continuous_features) + len(self.feature_names) * self.embedding_dims
        self.MLP_layer1.init_weights(in_dim, rng, dtype)
        self.MLP_layer2.init_weights(self.mlp_dims[-1], rng, dtype)
        return self

    def variables(self):
        return [self.bias, self.embed, self.w] + self.MLP_layer1.variables() + self.MLP_layer2.variables()

    def logit(self, X: Tensor, X_cont: Optional[Tensor] = None) -> Tensor:
        emb, wv = self._rows(X)
        first_order = torch.sum(wv, dim=1) + self.bias                       # :293
        sum_of_square = torch.sum(torch.square(emb), dim=1)
        square_of_sum = torch.square(torch.sum(emb, dim=1))
        second_order = 0.5 * torch.sum(square_of_sum - sum_of_square, dim=1, keepdim=True)
        fm_part = first_order + second_order                                 # :297
        dense_embedding = emb.reshape(emb.shape[0], -1)                      # Flatten :300
        if X_cont is not None:
            dense_embedding = torch.cat([X_cont, dense_embedding], dim=1)
        dnn_part = self.MLP_layer2(self.MLP_layer1(dense_embedding))         # :301
        return fm_part + dnn_part

    def call(self, inputs) -> Dict[str, Tensor]:
        X = assemble_ids(inputs, self.feature_names)
        X_cont = None
        if self.continuous_features:
            X_cont = assemble_ids(inputs, self.continuous_features).to(self.embed.dtype)
        return {"output": torch.sigmoid(self.logit(X, X_cont))}

    __call__ = call


# --------------------------------------------------------------------------
# a6: FieldAwareInteractionLayer                 2.FM/CustomLayers.py:428-462
# --------------------------------------------------------------------------
def field_aware_interaction(T: Tensor, X: Tensor) -> Tensor:
    """T [V,F,k]; X [B,F] -> [B,P,k] with I[b,(a,c),:] = T[x_a,c,:]*T[x_c,a,:],
    a<c in row-major (a,c) order (band-part mask + boolean_mask, :446-461)."""
    E = embedding_lookup(T, X)                       # [B,F,F,k]   :438
    ET = E.permute(0, 2, 1, 3)                       # :439
    inter = E * ET                                   # :440
    F = X.shape[1]
    iu = torch.triu_indices(F, F, offset=1)          # row-major strict upper triangle
    return inter[:, iu[0], iu[1], :]


def field_aware_interaction_from_rows(E: Tensor) -> Tensor:
    """Same, from already gathered/pooled rows E [B,F,F,k] (bag mode)."""
    F = E.shape[1]
    inter = E * E.permute(0, 2, 1, 3)
    iu = torch.triu_indices(F, F, offset=1)
    return inter[:, iu[0], iu[1], :]


class FieldAwareInteractionLayer:
    def __init__(self, fields_cnt, feature_dims=20, embedding_dims=16):
        self.fields_cnt, self.feature_dims, self.embedding_dims = fields_cnt, feature_dims, embedding_dims
        self.embedding_lookup_table = None

    def init_weights(self, rng, dtype=torch.float32):
        # Keras default for add_weight: glorot_uniform; fans for a rank-3 shape (V,F,k):
        # receptive = V, fan_in = F*V, fan_out = k*V
        V, F, k = self.feature_dims, self.fields_cnt, self.embedding_dims
        lim = math.sqrt(6.0 / (F * V + k * V))
        self.embedding_lookup_table = torch.tensor(rng.uniform(-lim, lim, size=(V, F, k)),
                                                   dtype=dtype, requires_grad=True)
        return self

    def __call__(self, X):
        return field_aware_interaction(self.embedding_lookup_table, X)


# --------------------------------------------------------------------------
# a5: FFMRankingLayer (F separate tables)         2.FM/CustomLayers.py:370-425
# --------------------------------------------------------------------------
class FFMRankingLayer:
    def __init__(self, feature_names, feature_dims=20, embedding_dims=16):
        self.feature_names = list(feature_names)
        self.feature_dims, self.embedding_dims = feature_dims, embedding_dims
        self.fields_cnt = len(self.feature_names)
        self.bias = self.w = None
        self.embedding_list: List[Tensor] = []

    def init_weights(self, rng, dtype=torch.float32):
        self.bias = torch.tensor(rng.uniform(-1.0, 1.0, size=(1,)), dtype=dtype, requires_grad=True)
        self.w = torch.tensor(rng.uniform(-0.05, 0.05, size=(self.feature_dims, 1)), dtype=dtype, requires_grad=True)
        self.embedding_list = [
            torch.tensor(rng.uniform(-0.05, 0.05, size=(self.feature_dims, self.embedding_dims)),
                         dtype=dtype, requires_grad=True) for _ in range(self.fields_cnt)]
        return self

    def variables(self):
        return [self.bias, self.w] + self.embedding_list

    def logit(self, X):
        first_order = torch.sum(embedding_lookup(self.w, X), dim=1)           # :408-410
        ebd_out = [embedding_lookup(self.embedding_list[i], X) for i in range(self.fields_cnt)]  # :412
        interactions = []
        for i in range(self.fields_cnt):                                       # :414-416
            for j in range(i + 1, self.fields_cnt):
                interactions.append(ebd_out[i][:, j] * ebd_out[j][:, i])
        interactions = torch.stack(interactions, dim=1)                        # :420
        second_order = torch.sum(torch.sum(interactions, dim=1), dim=1, keepdim=True)  # :421
        return self.bias + first_order + second_order

    def call(self, inputs):
        X = inputs if isinstance(inputs, Tensor) else assemble_ids(inputs, self.feature_names)
        return {"output": torch.sigmoid(self.logit(X))}

    __call__ = call


# --------------------------------------------------------------------------
# a7/a8: FFMLayer, FwFMLayer                     2.FM/CustomLayers.py:465-533
# --------------------------------------------------------------------------
class FFMLayer:
    """sigma(bias + sum_f w[x_f] + sum_{p,k} I).  NOTE the reference builds
    ``FieldAwareInteractionLayer(self.fields_cnt)`` without forwarding
    feature_dims/embedding_dims (:477), so its pair table is always (20,F,16);
    the oracle (and the build) forward them -- recorded in DESIGN.md."""

    head = "ffm"

    def __init__(self, feature_names, feature_dims=20, embedding_dims=16,
                 pad_id: Optional[int] = None, pooling: str = "sum"):
        self.feature_names = list(feature_names)
        self.feature_dims, self.embedding_dims = feature_dims, embedding_dims
        self.fields_cnt = len(self.feature_names)
        self.pad_id, self.pooling = pad_id, pooling
        self.fa_interaction_layer = FieldAwareInteractionLayer(self.fields_cnt, feature_dims, embedding_dims)
        self.bias = self.w = None
        self.r = self.r0 = None        # FwFM Dense(1): kernel [P,1], bias [1]

    def init_weights(self, rng, dtype=torch.float32):
        self.bias = torch.tensor(rng.uniform(-1.0, 1.0, size=(1,)), dtype=dtype, requires_grad=True)
        lim = math.sqrt(6.0 / (self.feature_dims + 1))
        self.w = torch.tensor(rng.uniform(-lim, lim, size=(self.feature_dims, 1)), dtype=dtype, requires_grad=True)
        self.fa_interaction_layer.init_weights(rng, dtype)
        if self.head == "fwfm":
            P = self.fields_cnt * (self.fields_cnt - 1) // 2
            lim = math.sqrt(6.0 / (P + 1))
            self.r = torch.tensor(rng.uniform(-lim, lim, size=(P, 1)), dtype=dtype, requires_grad=True)
            self.r0 = torch.zeros(1, dtype=dtype, requires_grad=True)
        return self

    def variables(self):
        v = [self.bias, self.w, self.fa_interaction_layer.embedding_lookup_table]
        if self.head == "fwfm":
            v += [self.r, self.r0]
        return v

    def pair_vectors(self, X):
        T = self.fa_interaction_layer.embedding_lookup_table
        if X.dim() == 3:          # bag mode (a17): pool rows of width F*k first
            V, F, k = T.shape
            E = pooled_lookup(T.reshape(V, F * k), X, self.pad_id, self.pooling).reshape(X.shape[0], F, F, k)
            return field_aware_interaction_from_rows(E)
        return field_aware_interaction(T, X)

    def logit(self, X):
        if X.dim() == 3:
            linear_term = torch.sum(pooled_lookup(self.w, X, self.pad_id, self.pooling), dim=1)
        else:
            linear_term = torch.sum(embedding_lookup(self.w, X), dim=1)          # :490 / :526
        iv = self.pair_vectors(X)                                               # :492 / :528
        if self.head == "ffm":
            interaction_term = torch.sum(torch.sum(iv, dim=1), dim=1, keepdim=True)   # :493
        else:
            interaction_term = torch.sum(iv, dim=-1) @ self.r + self.r0         # Dense(1)  :529
        return self.bias + linear_term + interaction_term

    def call(self, inputs):
        X = inputs if isinstance(inputs, Tensor) else assemble_ids(inputs, self.feature_names)
        return {"output": torch.sigmoid(self.logit(X))}

    __call__ = call


class FwFMLayer(FFMLayer):
    head = "fwfm"


# --------------------------------------------------------------------------
# a9: InnerProductNetwork / IpnLayer   2.FM/CustomLayers.py:601-624, 755-792
# --------------------------------------------------------------------------
def pair_index(F: int):
    """(i,j), i<j in itertools.combinations order == row-major upper triangle."""
    return list(itertools.combinations(range(F), 2))


def inner_product_network(x: Tensor) -> Tensor:
    """x [B,F,k] -> [B,P], out[b,p] = <x_i, x_j>  (loop form :618-623)."""
    F = x.shape[1]
    products = [x[:, i, :] * x[:, j, :] for i, j in pair_index(F)]
    return torch.sum(torch.stack(products, dim=1), dim=2)


def shared_fields_interaction(x: Tensor) -> Tensor:
    """Vectorised form (2.FM/CustomLayers.py:755-772): [B,F,k] -> [B,P,k]."""
    F = x.shape[1]
    inter = x.unsqueeze(1) * x.unsqueeze(2)
    iu = torch.triu_indices(F, F, offset=1)
    return inter[:, iu[0], iu[1], :]


def ipn_layer(x: Tensor) -> Tensor:
    return torch.sum(shared_fields_interaction(x), dim=2)                      # :790-791


# --------------------------------------------------------------------------
# a10: OuterProductNetwork / OpnLayer  2.FM/CustomLayers.py:627-682, 795-851
# --------------------------------------------------------------------------
def outer_product_network(x: Tensor, kernel: Tensor, kernel_type: str = "mat") -> Tensor:
    """kernel 'mat' [k,P,k], 'vec' [P,k], 'num' [P,1]; x [B,F,k] -> [B,P]."""
    F = x.shape[1]
    pairs = pair_index(F)
    if kernel_type != "mat":
        stacked = torch.stack([x[:, i, :] * x[:, j, :] for i, j in pairs], dim=1)   # :663-667
        return torch.sum(stacked * kernel.unsqueeze(0), dim=2)                      # :668-669
    p = torch.stack([x[:, i, :] for i, _ in pairs], dim=1)                          # :673-677
    q = torch.stack([x[:, j, :] for _, j in pairs], dim=1)
    kp = p.unsqueeze(1) * kernel                                                    # [B,k,P,k] :678
    kp = torch.sum(kp, dim=-1).permute(0, 2, 1)                                     # [B,P,k]   :679
    return torch.sum(kp * q, dim=-1)                                                # :680


# --------------------------------------------------------------------------
# a11: PNNRankingLayer / PNNLayer       2.FM/CustomLayers.py:536-599, 685-753
# --------------------------------------------------------------------------
class PNNRankingLayer:
    def __init__(self, feature_names, feature_dims=20, embedding_dims=16, mlp_dims=(32, 8),
                 dropout=0, method="inner", kernel_type=None):
        assert method in ("inner", "outer")
        self.feature_names = list(feature_names)
        self.feature_dims, self.embedding_dims = feature_dims, embedding_dims
        self.fields_cnt = len(self.feature_names)
        self.mlp_dims = list(mlp_dims)
        self.method = method
        self.kernel_type = kernel_type or "mat"
        self.MLP_layer1 = MLPLayer(units=self.mlp_dims, activation="relu")
        self.MLP_layer2 = MLPLayer(units=[1], activation="sigmoid")
        self.embed = self.kernel = None

    def init_weights(self, rng, dtype=torch.float32):
        F, k = self.fields_cnt, self.embedding_dims
        P = F * (F - 1) // 2
        self.embed = torch.tensor(rng.uniform(-0.05, 0.05, size=(self.feature_dims, k)), dtype=dtype, requires_grad=True)
        if self.method == "outer":
            shape = {"mat": (k, P, k), "vec": (P, k), "num": (P, 1)}[self.kernel_type]
            self.kernel = torch.tensor(rng.normal(0.0, 0.05, size=shape), dtype=dtype, requires_grad=True)
        self.MLP_layer1.init_weights(F * k + P, rng, dtype)
        self.MLP_layer2.init_weights(self.mlp_dims[-1], rng, dtype)
        return self

    def variables(self):
        v = [self.embed]
        if self.kernel is not None:
            v.append(self.kernel)
        return v + self.MLP_layer1.variables() + self.MLP_layer2.variables()

    def product(self, emb):
        if self.method == "inner":
            return inner_product_network(emb)
        return outer_product_network(emb, self.kernel, self.kernel_type)

    def call(self, inputs):
        X = inputs if isinstance(inputs, Tensor) else assemble_ids(inputs, self.feature_names)
        emb_output = embedding_lookup(self.embed, X)                               # :586
        product_output = self.product(emb_output)                                  # :588
        dense_embedding = emb_output.reshape(emb_output.shape[0], -1)              # :589
        combined = torch.cat([dense_embedding, product_output], dim=1)             # :591
        return {"output": self.MLP_layer2(self.MLP_layer1(combined))}              # :595-596

    __call__ = call


PNNLayer = PNNRankingLayer     # vectorised twin (:685-753) computes the same function


# --------------------------------------------------------------------------
# a12: CrossLayer                               3.DCN/CustomLayers.py:170-203
# --------------------------------------------------------------------------
def cross_layer(x: Tensor, ws: Sequence[Tensor], bs: Sequence[Tensor]) -> Tensor:
    """x_{l+1} = x0 * (x_l^T w_l) + b_l + x_l ; w_l, b_l [D,1] (:195-203)."""
    x0 = x.unsqueeze(2)                                    # [B,D,1]
    xl = x0
    for w, b in zip(ws, bs):
        xl_w = torch.matmul(xl.transpose(1, 2), w)         # [B,1,1]   :199
        xl = torch.matmul(x0, xl_w) + b + xl               # :200
    return xl.squeeze(2)


# --------------------------------------------------------------------------
# a13: MatrixCrossLayer                         3.DCN/CustomLayers.py:272-305
# --------------------------------------------------------------------------
def matrix_cross_layer(x: Tensor, Ws: Sequence[Tensor], bs: Sequence[Tensor]) -> Tensor:
    """x_{l+1} = x0 (.) (W_l x_l + b_l) + x_l ; W_l [D,D], b_l [D,1]; the
    reference's ``tf.matmul(W, x[B,D,1])`` is y = W x, i.e. X W^T in batch
    form (:300-303; pinned by KAT-4)."""
    x0 = x.unsqueeze(2)
    xl = x0
    for W, b in zip(Ws, bs):
        xl_w = torch.matmul(W, xl)                         # [B,D,1]   :301
        xl = x0 * (xl_w + b) + xl                          # :302
    return xl.squeeze(2)


class CrossLayer:
    kind = "vec"

    def __init__(self, layer_num, reg_w=1e-4, reg_b=1e-4):
        self.layer_num = layer_num
        self.cross_weight: List[Tensor] = []
        self.cross_bias: List[Tensor] = []

    def init_weights(self, D, rng, dtype=torch.float32):
        shape = (D, 1) if self.kind == "vec" else (D, D)
        self.cross_weight = [torch.tensor(rng.normal(0.0, 0.05, size=shape), dtype=dtype, requires_grad=True)
                             for _ in range(self.layer_num)]
        self.cross_bias = [torch.zeros((D, 1), dtype=dtype, requires_grad=True) for _ in range(self.layer_num)]
        return self

    def variables(self):
        return self.cross_weight + self.cross_bias

    def __call__(self, x):
        fn = cross_layer if self.kind == "vec" else matrix_cross_layer
        return fn(x, self.cross_weight, self.cross_bias)


class MatrixCrossLayer(CrossLayer):
    kind = "matrix"


# --------------------------------------------------------------------------
# a14: DeepCrossNetworkLayer                    3.DCN/CustomLayers.py:206-269
# --------------------------------------------------------------------------
class DeepCrossNetworkLayer:
    def __init__(self, categorical_features, continuous_features, feature_dims=160000,
                 embedding_dims=16, units=(64, 8), activation="relu", layer_num=3,
                 reg_w=1e-4, reg_b=1e-4, type="vec"):
        self.categorical_features = list(categorical_features)
        self.continuous_features = list(continuous_features)
        self.feature_dims, self.embedding_dims = feature_dims, embedding_dims
        self.cross_layer = CrossLayer(layer_num) if type == "vec" else MatrixCrossLayer(layer_num)  # :226-229
        self.dense_layer = MLPLayer(list(units), activation)
        self.output_layer = MLPLayer([1], "sigmoid")                                 # Dense(1, sigmoid) :234
        self.embedding = None

    @property
    def D(self):
        return len(self.continuous_features) + len(self.categorical_features) * self.embedding_dims

    def init_weights(self, rng, dtype=torch.float32):
        self.embedding = torch.tensor(rng.uniform(-0.05, 0.05, size=(self.feature_dims, self.embedding_dims)),
                                      dtype=dtype, requires_grad=True)
        self.cross_layer.init_weights(self.D, rng, dtype)
        self.dense_layer.init_weights(self.D, rng, dtype)
        self.output_layer.init_weights(self.D + self.dense_layer.units[-1], rng, dtype)
        return self

    def variables(self):
        return [self.embedding] + self.cross_layer.variables() + self.dense_layer.variables() + \
            self.output_layer.variables()

    def assemble(self, inputs):
        X = assemble_ids(inputs, self.categorical_features)                          # :240-246
        X_cont = assemble_ids(inputs, self.continuous_features).to(self.embedding.dtype)  # :248-254
        X_emb = embedding_lookup(self.embedding, X)                                  # :256
        X_flatten = X_emb.reshape(X_emb.shape[0], -1)                                # :257
        return torch.cat([X_cont, X_flatten], dim=1)                                 # continuous first :259

    def call(self, inputs):
        _input = self.assemble(inputs)
        cross_output = self.cross_layer(_input)                                      # :261
        dnn_output = self.dense_layer(_input)                                        # :263
        combine = torch.cat([cross_output, dnn_output], dim=1)                       # :265
        return {"output": self.output_layer(combine)}                                # :267

    __call__ = call


# --------------------------------------------------------------------------
# f2: the other pairwise consumers of the same rows
# --------------------------------------------------------------------------
def interaction_layer(x: Tensor) -> Tensor:
    """AFM ``InteractionLayer.call`` (3.DCN/CustomLayers.py:825-838): the double loop i<j of
    ``inputs[:,i,:] * inputs[:,j,:]``, stacked and transposed to [B,P,k]."""
    F = x.shape[1]
    result = []
    for i in range(F - 1):
        for j in range(i + 1, F):
            result.append(x[:, i, :] * x[:, j, :])
    return torch.stack(result, dim=0).permute(1, 0, 2)


def bilinear_interaction(x: Tensor, W, bilinear_type: str = "interaction") -> Tensor:
    """FiBiNet ``BilinearInteractionLayer.call`` (3.DCN/CustomLayers.py:996-1009): ``tensordot(field_i, W, axes=(-1,0))
    * field_j`` over itertools.combinations; 'all': one W [k,k]; 'each': W_list[i] for the LEFT field i;
    'interaction': W_list[p] per pair.  -> [B,P,k]."""
    import itertools
    F = x.shape[1]
    pairs = list(itertools.combinations(range(F), 2))
    out = []
    for p, (i, j) in enumerate(pairs):
        Wm = W if bilinear_type == "all" else (W[i] if bilinear_type == "each" else W[p])
        out.append((x[:, i, :] @ Wm) * x[:, j, :])
    return torch.stack(out, dim=1)


def bi_interaction(x: Tensor) -> Tensor:
    """NFM second-order pooling (3.DCN/CustomLayers.py:499-501): 0.5 * (square(sum_f x) - sum_f square(x)) -> [B,k]."""
    return 0.5 * (torch.square(torch.sum(x, dim=1)) - torch.sum(torch.square(x), dim=1))


def onn_combined(embed_single: Tensor, T: Tensor, X: Tensor, reduce: bool = False) -> Tensor:
    """``ParralledOnnLayer.call`` up to the tower (2.FM/CustomLayers.py:989-1001): [Flatten(embedding_single(X)) |
    Flatten(FieldAwareInteractionLayer(X))] with the pair vectors summed over k when ``reduce``."""
    X_single = embedding_lookup(embed_single, X).reshape(X.shape[0], -1)
    X_pair = field_aware_interaction(T, X)
    if reduce:
        X_pair = torch.sum(X_pair, dim=2)
    return torch.cat([X_single, X_pair.reshape(X.shape[0], -1)], dim=1)


def onn_loop_combined(embed_single: Tensor, pair_tables, X: Tensor, reduce: bool = False) -> Tensor:
    """``ONNLayer.call`` up to the tower (2.FM/CustomLayers.py:936-953): per pair (i,j) its own two tables,
    ``E1_ij[x_i] * E2_ij[x_j]``; ``pair_tables[(i,j)] = (E1, E2)``."""
    F = X.shape[1]
    X_single = embedding_lookup(embed_single, X).reshape(X.shape[0], -1)
    prods = []
    for i in range(F):
        for j in range(i + 1, F):
            e1, e2 = pair_tables[(i, j)]
            prods.append(embedding_lookup(e1, X[:, i]) * embedding_lookup(e2, X[:, j]))
    X_pair = torch.stack(prods, dim=1)
    if reduce:
        X_pair = torch.sum(X_pair, dim=2)
    return torch.cat([X_single, X_pair.reshape(X.shape[0], -1)], dim=1)


# --------------------------------------------------------------------------
# f1: masked / weighted pooling of a padded behaviour series; f3: used-rows L2
# --------------------------------------------------------------------------
def sequence_pool(table: Tensor, ids: Tensor, weights: Optional[Tensor] = None, mask: Optional[Tensor] = None,
                  padding_index: Optional[int] = None, reduce: bool = True) -> Tensor:
    """7.SIM/CustomLayers.py:107-118 + :88-95: X_series = embed(reshape(ids, [B, L*C])) reshaped to [B, L, C*k];
    valid_mask = ids[:, :, 0] != padding_index; pooled = einsum('bl,ble->be', scores * mask, X_series).
    ``reduce=False`` returns the weighted [B, L, C*k] rows (FiBiNet++'s embed(keys) * values[..., None] is L = F, C = 1)."""
    B, L, C = ids.shape
    x = embedding_lookup(table, ids.reshape(B, L * C)).reshape(B, L, C * table.shape[1])
    w = torch.ones((B, L), dtype=table.dtype)
    if weights is not None:
        w = w * weights
    if mask is not None:
        w = w * mask.to(table.dtype)
    if padding_index is not None:
        w = w * (ids[:, :, 0] != padding_index).to(table.dtype)
    if reduce:
        return torch.einsum("bl,ble->be", w, x)
    return x * w.unsqueeze(-1)


def used_rows_l2(table: Tensor, all_ids: Tensor, factor: float) -> Tensor:
    """5.DIN/ModelManager.py:185-190: tf.unique over every id of the batch, gather, tf.nn.l2_loss (= sum(x^2)/2) * factor."""
    uniq = torch.unique(all_ids.reshape(-1))
    return factor * 0.5 * torch.sum(torch.square(table[uniq]))


# --------------------------------------------------------------------------
# a16: loss, backward, IndexedSlices dedup, Adam     2.FM/ModelManager.py:99-104,171-181
# --------------------------------------------------------------------------
KERAS_EPS = 1e-7


def keras_bce(target: Tensor, output: Tensor) -> Tensor:
    """tf.keras.losses.BinaryCrossentropy() on probabilities (upstream Keras 2.8
    ``backend.binary_crossentropy``): clip to [eps,1-eps]; -(t log(p+eps) +
    (1-t) log(1-p+eps)); mean over the last axis then over the batch.
    ``reduce_sum`` of that scalar (2.FM/ModelManager.py:175) is the same scalar."""
    p = torch.clamp(output, KERAS_EPS, 1.0 - KERAS_EPS)
    bce = target * torch.log(p + KERAS_EPS) + (1.0 - target) * torch.log(1.0 - p + KERAS_EPS)
    return torch.mean(torch.mean(-bce, dim=-1))


def indexed_slices_dedup(indices: np.ndarray, values: np.ndarray):
    """Keras ``_deduplicate_indexed_slices``: ``tf.unique`` (first-occurrence
    order) + ``unsorted_segment_sum``.  Returns (unique_ids, summed rows)."""
    uniq, first_pos, inverse = np.unique(indices, return_index=True, return_inverse=True)
    order = np.argsort(first_pos, kind="stable")          # first-occurrence order
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    seg = rank[inverse]
    summed = np.zeros((uniq.size,) + values.shape[1:], dtype=values.dtype)
    np.add.at(summed, seg, values)
    return uniq[order], summed


class KerasAdam:
    """tf.keras.optimizers.Adam(lr) with Keras defaults (beta1=.9, beta2=.999,
    eps=1e-7), call site 2.FM/ModelManager.py:103-104,178.

    ``mode='keras_dense'`` restates upstream Keras-2.8 ``_resource_apply_sparse``
    for IndexedSlices: m and v are decayed for ALL rows, the (deduplicated)
    gradient is scattered into the touched rows, and var is updated for ALL rows.
    ``mode='rowwise'`` is the lazy variant the build benchmarks: only the unique
    touched rows are read and written.  For dense variables both are plain Adam.
    (Recalled upstream semantics; TF is not present to confirm -- SURVEY a16.)"""

    def __init__(self, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, mode="rowwise"):
        self.lr, self.b1, self.b2, self.eps, self.mode = lr, beta_1, beta_2, epsilon, mode
        self.t = 0
        self.state: Dict[int, tuple] = {}

    def _slots(self, var):
        if id(var) not in self.state:
            self.state[id(var)] = (torch.zeros_like(var), torch.zeros_like(var))
        return self.state[id(var)]

    def step_begin(self):
        self.t += 1
        return self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)

    @torch.no_grad()
    def apply_dense(self, var: Tensor, grad: Tensor, lr_t: float):
        m, v = self._slots(var)
        m.mul_(self.b1).add_(grad, alpha=1.0 - self.b1)
        v.mul_(self.b2).addcmul_(grad, grad, value=1.0 - self.b2)
        var.sub_(lr_t * m / (torch.sqrt(v) + self.eps))

    @torch.no_grad()
    def apply_sparse(self, var: Tensor, uniq_ids: Tensor, grad_rows: Tensor, lr_t: float):
        """grad_rows already summed over duplicate ids."""
        m, v = self._slots(var)
        if self.mode == "keras_dense":
            m.mul_(self.b1)
            m[uniq_ids] += (1.0 - self.b1) * grad_rows
            v.mul_(self.b2)
            v[uniq_ids] += (1.0 - self.b2) * grad_rows * grad_rows
            var.sub_(lr_t * m / (torch.sqrt(v) + self.eps))
        else:
            mr = m[uniq_ids] * self.b1 + (1.0 - self.b1) * grad_rows
            vr = v[uniq_ids] * self.b2 + (1.0 - self.b2) * grad_rows * grad_rows
            m[uniq_ids] = mr
            v[uniq_ids] = vr
            var[uniq_ids] -= lr_t * mr / (torch.sqrt(vr) + self.eps)


def dedup_dense_grad(var_grad: Tensor):
    """torch autograd gives a dense [V,w] gradient for a gathered table; recover
    the IndexedSlices view (unique touched rows, summed) for the sparse apply."""
    nz = torch.nonzero(var_grad.abs().sum(dim=tuple(range(1, var_grad.dim()))) > 0).flatten()
    return nz, var_grad[nz]
