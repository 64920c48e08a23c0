"""Drop-in layer classes for the reference's ``CustomLayers.py`` hot path.

Same class names, constructor kwargs, ``call(inputs: dict) -> {'output': [B,1]}``
protocol and variable order/naming as 2.FM/CustomLayers.py and
3.DCN/CustomLayers.py, but every ``call`` is one or a few sm_100a kernels behind
the C ABI (include/etr.h) instead of a chain of eager TF ops.  TensorFlow is not
required: inputs may be torch tensors (any device), numpy arrays or any DLPack
producer (``tf.Tensor`` included); outputs are torch CUDA tensors, which export
``__dlpack__`` for ``tf.experimental.dlpack.from_dlpack``.

Extra, build-defined kwargs (all optional, defaults reproduce the reference):
``device``, ``seed``, ``table_dtype`` ('float32' | 'bfloat16'), ``pad_id`` and
``pooling`` ('sum' | 'mean') for multi-hot bags, ``check_ids``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import check
from .dense import DenseLayer, DenseParams, MLPLayer, glorot_uniform, random_normal
from .runtime import (EmbeddingTable, IdsBatch, Runtime, SparseGrad, SparsePlan, bce_forward_backward,
                      embedding_gather, gather_fm_backward, gather_fm_forward, lr_t, _p)

_DT = {"float32": torch.float32, "bfloat16": torch.bfloat16, torch.float32: torch.float32,
       torch.bfloat16: torch.bfloat16}


class DeviceBatch:
    """Inputs of one step already staged on the device: an IdsBatch, the
    continuous matrix [B,C] (any strides) and labels [B].  Layers accept it in
    place of the reference's dict so a captured CUDA graph can read static buffers."""

    def __init__(self, ids: IdsBatch, cont: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None):
        self.ids, self.cont, self.labels = ids, cont, labels


class _Layer:
    """Minimal stand-in for tf.keras.layers.Layer: ``layer(inputs)`` -> ``call``."""

    def _setup(self, kwargs):
        self.rt = Runtime.get(kwargs.pop("device", None))
        self.seed = int(kwargs.pop("seed", 0))
        self.table_dtype = _DT[kwargs.pop("table_dtype", "float32")]
        self.pad_id = kwargs.pop("pad_id", None)
        self.pooling = kwargs.pop("pooling", "sum")
        self.check_ids = bool(kwargs.pop("check_ids", True))
        self.name = kwargs.pop("name", type(self).__name__)
        kwargs.pop("trainable", None)
        kwargs.pop("dtype", None)
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        self.gen = torch.Generator(device=self.rt.device)
        self.gen.manual_seed(self.seed)
        self.params = DenseParams(self.rt)
        self._ctx: Dict[str, object] = {}

    def __call__(self, inputs, training: bool = False, **kw):
        return self.call(inputs, training=training, **kw)

    def _finish(self, training: bool):
        if self.check_ids and not training:
            self.rt.poll_error()

    def _ids(self, inputs, names) -> IdsBatch:
        if isinstance(inputs, DeviceBatch):
            return inputs.ids
        return IdsBatch.make(self.rt, inputs, names, self.pad_id, self.pooling)

    def _cont(self, inputs, names) -> torch.Tensor:
        if isinstance(inputs, DeviceBatch):
            return inputs.cont
        return _cont_matrix(self.rt, inputs, names)

    def sparse_tables(self) -> List[EmbeddingTable]:
        return []


def _cont_matrix(rt: Runtime, inputs, names: Sequence[str]) -> torch.Tensor:
    """continuous dict -> X_cont [B,C] fp32 (3.DCN/CustomLayers.py:248-254)."""
    cols = [rt.to_device(inputs[n], torch.float32).reshape(-1) for n in names]
    return torch.stack(cols, dim=1)


# ---------------------------------------------------------------------------
class Embedding(_Layer):
    """tf.keras.layers.Embedding(input_dim, output_dim) drop-in: ``__call__(ids[...])
    -> [..., output_dim]``; ``.embeddings`` / ``.weights[0]`` as used by
    5.DIN/ModelManager.py:189.  Keras default init uniform(-0.05, 0.05)."""

    def __init__(self, input_dim, output_dim, embeddings_regularizer=None, **kwargs):
        self._setup(kwargs)
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        self.table = EmbeddingTable(self.rt, self.input_dim, self.output_dim, self.table_dtype)
        self.table.init_uniform(-0.05, 0.05, self.gen)

    @property
    def embeddings(self) -> torch.Tensor:
        return self.table.cols(0, self.output_dim)

    @property
    def weights(self):
        return [self.embeddings]

    def call(self, ids, training: bool = False):
        out = embedding_gather(self.table, ids)
        self._finish(training)
        return out

    def sparse_tables(self):
        return [self.table]


# ---------------------------------------------------------------------------
class FMRankingLayer(_Layer):
    """sigma(bias + sum_f w[x_f] + 0.5 sum_k((sum_f v)^2 - sum_f v^2))
    (2.FM/CustomLayers.py:87-157).  Variable order: bias, embed, w (:123-135).
    ``embed`` [V,k] and ``w`` [V,1] live fused in one [V,k+1] HBM table."""

    def __init__(self, feature_names=['item_tag1', 'item_tag2', 'item_tag3'], feature_dims=20, embedding_dims=16,
                 **kwargs):
        self._setup(kwargs)
        self.feature_names = list(feature_names)
        self.feature_dims = int(feature_dims)
        self.embedding_dims = int(embedding_dims)
        self._build_fm()
        self._build_extra()
        self.params.finalize()

    def _build_fm(self):
        k = self.embedding_dims
        self.params.add("bias", glorot_uniform((1,), self.gen, self.rt.device))
        self.table = EmbeddingTable(self.rt, self.feature_dims, k + 1, self.table_dtype)
        self.table.init_uniform(-0.05, 0.05, self.gen)      # Keras Embedding default for embed and w

    def _build_extra(self):
        pass

    # -- variables, reference names -----------------------------------------
    @property
    def bias(self) -> torch.Tensor:
        return self.params["bias"]

    @property
    def embed(self) -> torch.Tensor:            # embed/embeddings
        return self.table.cols(0, self.embedding_dims)

    @property
    def w(self) -> torch.Tensor:                # w/embeddings
        return self.table.cols(self.embedding_dims, self.embedding_dims + 1)

    @property
    def variables(self):
        return [self.bias, self.embed, self.w]

    def set_weights(self, bias=None, embed=None, w=None):
        if bias is not None:
            self.params.set("bias", bias)
        if embed is not None:
            self.embed.copy_(torch.as_tensor(embed).to(self.rt.device))
        if w is not None:
            self.w.copy_(torch.as_tensor(w).reshape(-1, 1).to(self.rt.device))

    def sparse_tables(self):
        return [self.table]

    # -- forward / backward ---------------------------------------------------
    def call(self, inputs, training: bool = False):
        ids = self._ids(inputs, self.feature_names)
        prob = self.rt.empty((ids.B, 1))
        gather_fm_forward(self.table, self.embedding_dims, True, ids, bias=self.bias, prob=prob)
        if training:
            self._ctx = {"ids": ids}
        self._finish(training)
        return {"output": prob}

    def backward(self, dlogit: torch.Tensor) -> List[SparseGrad]:
        """dlogit = dL/dz [B] (z the pre-sigmoid logit).  Dense grads land in
        ``self.params.grad``; the table gradient is returned as a SparseGrad."""
        ids = self._ctx["ids"]
        dl = dlogit.reshape(-1)
        bag = gather_fm_backward(self.table, self.embedding_dims, True, ids, dlogit=dl)
        check(self.rt.lib.etr_colsum_f32(self.rt.ctx, dl.data_ptr(), ids.B, 1, 1, self.params.g("bias").data_ptr(),
                                         self.rt.stream))
        return [SparseGrad(self.table, ids, bag)]


# ---------------------------------------------------------------------------
class DeepFMRankingLayer(FMRankingLayer):
    """FM logit + MLP([32,8],relu) -> MLP([1]) over Flatten(emb); sigma(fm+dnn)
    (2.FM/CustomLayers.py:241-308).  Build extension (c2):
    ``continuous_features`` are concatenated in front of the flattened embedding
    for the MLP only, as 3.DCN/CustomLayers.py:259 does."""

    def __init__(self, feature_names=['user_tag0', 'user_tag1', 'item_tag1', 'item_tag2', 'item_tag3'],
                 feature_dims=20, embedding_dims=16, mlp_dims=[32, 8], continuous_features=(), **kwargs):
        self.mlp_dims = list(mlp_dims)
        self.continuous_features = list(continuous_features)
        super().__init__(feature_names, feature_dims, embedding_dims, **kwargs)

    def _build_extra(self):
        in_dim = len(self.continuous_features) + len(self.feature_names) * self.embedding_dims
        self.MLP_layer1 = MLPLayer(units=self.mlp_dims, activation="relu", name="MLP_layer1")
        self.MLP_layer2 = MLPLayer(units=[1], name="MLP_layer2")
        self.MLP_layer1.build(in_dim, self.params, self.gen)
        self.MLP_layer2.build(self.mlp_dims[-1], self.params, self.gen)

    @property
    def variables(self):
        v = [self.bias, self.embed, self.w]
        for mlp in (self.MLP_layer1, self.MLP_layer2):
            for kk, bb in zip(mlp.kernels, mlp.biases):
                v += [kk, bb]
        return v

    def call(self, inputs, training: bool = False):
        rt = self.rt
        ids = self._ids(inputs, self.feature_names)
        C_ = len(self.continuous_features)
        k, F = self.embedding_dims, len(self.feature_names)
        x = rt.empty((ids.B, C_ + F * k))
        if C_:
            x[:, :C_] = self._cont(inputs, self.continuous_features)
        fm_logit = rt.empty((ids.B,))
        gather_fm_forward(self.table, k, True, ids, bias=self.bias, logit=fm_logit, flat=x, flat_col0=C_)
        dnn = self.MLP_layer2(self.MLP_layer1(x, training=training), training=training)      # [B,1]
        prob = rt.empty((ids.B, 1))
        check(rt.lib.etr_add_sigmoid(rt.ctx, fm_logit.data_ptr(), dnn.data_ptr(), ids.B, None, prob.data_ptr(),
                                     rt.stream))
        if training:
            self._ctx = {"ids": ids}
        self._finish(training)
        return {"output": prob}

    def backward(self, dlogit: torch.Tensor) -> List[SparseGrad]:
        rt = self.rt
        ids = self._ctx["ids"]
        dl = dlogit.reshape(-1)
        C_ = len(self.continuous_features)
        d_dnn = dl.clone().reshape(-1, 1)                    # d(fm+dnn)/d dnn = 1
        dh = self.MLP_layer2.backward(d_dnn)
        dx = self.MLP_layer1.backward(dh)                    # [B, C + F*k]
        bag = gather_fm_backward(self.table, self.embedding_dims, True, ids, dlogit=dl, dflat=dx, flat_col0=C_)
        check(rt.lib.etr_colsum_f32(rt.ctx, dl.data_ptr(), ids.B, 1, 1, self.params.g("bias").data_ptr(), rt.stream))
        return [SparseGrad(self.table, ids, bag)]


# ---------------------------------------------------------------------------
class Trainer:
    """The reference train step (2.FM/ModelManager.py:171-181): BCE on the
    layer's 'output', backward, Adam apply.  ``apply_mode='rowwise'`` touches
    only the unique rows of the batch; ``'keras_dense'`` restates Keras 2.8
    ``Adam._resource_apply_sparse`` (every row of the table is decayed/updated).

    ``graph=True`` stages the inputs of every step into static device buffers
    (one async H2D copy per host column) and replays ONE captured CUDA graph of
    the whole step; the optimizer clock lives on the device for that reason."""

    def __init__(self, layer, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, apply_mode="rowwise", graph=False):
        assert apply_mode in ("rowwise", "keras_dense")
        self.layer, self.lr, self.b1, self.b2, self.eps = layer, lr, beta_1, beta_2, epsilon
        self.mode = _lib.ADAM_ROWWISE if apply_mode == "rowwise" else _lib.ADAM_KERAS_DENSE
        self.rt = layer.rt
        self.state = self.rt.zeros((2,))            # [iterations, lr_t] on the device
        self.last_grads: List[SparseGrad] = []
        self.use_graph = graph
        self._graphs: Dict[tuple, tuple] = {}

    @property
    def iterations(self) -> int:                     # synchronises
        return int(self.state[0].item())

    # ------------------------------------------------------------ eager step
    def train_step(self, inputs, labels=None) -> torch.Tensor:
        """One step; returns the (device-resident, un-synchronised) scalar loss."""
        if self.use_graph:
            return self._graph_step(inputs, labels)
        return self._eager_step(inputs, labels)

    def _eager_step(self, inputs, labels=None) -> torch.Tensor:
        rt = self.rt
        if labels is None and isinstance(inputs, DeviceBatch):
            labels = inputs.labels
        out = self.layer(inputs, training=True)["output"]
        y = rt.to_device(labels, torch.float32).reshape(-1)
        loss, dlogit = bce_forward_backward(rt, out.reshape(-1), y)
        grads = self.layer.backward(dlogit)
        self.apply_gradients(grads)
        return loss

    def apply_gradients(self, grads: List[SparseGrad]) -> None:
        rt = self.rt
        check(rt.lib.etr_adam_step_begin(rt.ctx, self.state.data_ptr(), self.lr, self.b1, self.b2, rt.stream))
        d_lr = self.state[1:]
        self.layer.params.adam_step(0.0, d_lr, self.b1, self.b2, self.eps)
        plans: Dict[tuple, SparsePlan] = {}
        for g in grads:
            key = (id(g.ids), g.table.rows)
            g.reduce(plans.get(key))
            plans[key] = g.plan
            t = g.table.desc()
            check(rt.lib.etr_sparse_adam_apply(rt.ctx, C.byref(t), g.table.m.data_ptr(), g.table.v.data_ptr(),
                                               g.plan.unique_ids.data_ptr(), g.plan.counts.data_ptr(),
                                               g.plan.n_slots, g.unique_grad.data_ptr(), g.unique_grad.shape[1],
                                               0.0, d_lr.data_ptr(), self.b1, self.b2, self.eps, self.mode,
                                               rt.stream))
        self.last_grads = grads

    # ----------------------------------------------------------- graph step
    def stage(self, inputs, labels) -> DeviceBatch:
        """Copy one step's host (or device) inputs into the static buffers."""
        lay, rt = self.layer, self.rt
        names = getattr(lay, "feature_names", None) or lay.categorical_features
        cont_names = list(getattr(lay, "continuous_features", []))
        B = int(inputs[names[0]].shape[0])
        key = (B,)
        if key not in self._graphs:
            ids_buf = rt.empty((len(names), B), torch.int64)
            cont_buf = rt.empty((max(len(cont_names), 1), B), torch.float32)
            lab_buf = rt.empty((B,), torch.float32)
            ids = IdsBatch(rt, ids_buf, B, len(names), 1, 1, B, 1, lay.pad_id, lay.pooling)
            batch = DeviceBatch(ids, cont_buf[: len(cont_names)].t() if cont_names else None, lab_buf)
            self._graphs[key] = [batch, ids_buf, cont_buf, lab_buf, None, None]
        batch, ids_buf, cont_buf, lab_buf, _, _ = self._graphs[key]
        for f, n in enumerate(names):
            ids_buf[f].copy_(torch.as_tensor(inputs[n]).reshape(-1), non_blocking=True)
        for c, n in enumerate(cont_names):
            cont_buf[c].copy_(torch.as_tensor(inputs[n]).reshape(-1), non_blocking=True)
        lab_buf.copy_(torch.as_tensor(labels).reshape(-1), non_blocking=True)
        return batch

    def _graph_step(self, inputs, labels) -> torch.Tensor:
        batch = inputs if isinstance(inputs, DeviceBatch) else self.stage(inputs, labels)
        key = (batch.ids.B,)
        slot = self._graphs.setdefault(key, [batch, None, None, None, None, None])
        if slot[4] is None:
            assert slot[0] is batch, "graph mode needs the static DeviceBatch returned by stage()"
            # warm-up eagerly on a side stream (sizes the workspace), then capture
            s = torch.cuda.Stream(device=self.rt.device)
            s.wait_stream(torch.cuda.current_stream(self.rt.device))
            with torch.cuda.stream(s):
                for _ in range(2):
                    self._eager_step(batch)
            torch.cuda.current_stream(self.rt.device).wait_stream(s)
            torch.cuda.synchronize(self.rt.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss = self._eager_step(batch)
            slot[4], slot[5] = g, loss          # capture does not execute: replay runs the step
        return self._replay(slot)

    @staticmethod
    def _replay(slot) -> torch.Tensor:
        slot[4].replay()
        return slot[5]
