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
import os
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import check
from .dense import DenseLayer, DenseParams, MLPLayer, glorot_uniform, random_normal  # noqa: F401
from .runtime import (EmbeddingTable, FusedFMGrad, IdsBatch, Runtime, SparseGrad, SparsePlan, bce_forward_backward,
                      cast_bf16, embedding_gather, gather_fm_backward, FlatSparseGrad, gather_fm_forward, gemm_bf16_tn, lr_t, _p)

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
        self.fused_apply = bool(kwargs.pop("fused_apply", True))     # FM family: fused backward+reduce+Adam
        self.record_layout = bool(kwargs.pop("record_layout", True))  # FM family, k=16: [var|m|v] in one 256-B record
        self.name = kwargs.pop("name", type(self).__name__)
        # row-sharded tables: shard=True | "a2a" (NCCL all-to-all exchange) | "peer" (CUDA-IPC peer memory over
        # NVLink: de-duplicated request/serve exchange in the fused train step, rows pulled by the gather kernel
        # itself elsewhere) | "peer-pull" (always the fused pull); world/rank from torch.distributed, or
        # shard=(world, rank) / shard=("peer", world, rank)
        shard = kwargs.pop("shard", None)
        self.shard_spec, self.shard_mode = None, None
        if shard:
            mode = "a2a"
            if isinstance(shard, str):
                mode, shard = shard, True
            elif isinstance(shard, (tuple, list)) and isinstance(shard[0], str):
                mode, shard = shard[0], tuple(shard[1:])
            assert mode in ("a2a", "peer", "peer-pull"), mode
            if shard is True:
                import torch.distributed as dist
                shard = (dist.get_world_size(), dist.get_rank())
            self.shard_spec, self.shard_mode = (int(shard[0]), int(shard[1])), mode
        self.shard = None
        self.peer = None
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

    # -- row-sharded tables, all-to-all form: shared by every model layer ------------------
    def _make_table(self, rows: int, width: int, record: bool = False) -> EmbeddingTable:
        """the layer's embedding table: plain, or -- ``shard=`` -- this rank's rows r, r+G, r+2G, ... of it"""
        if self.shard_spec and self.shard_mode == "a2a":
            from .sharded import ShardedTable
            self.shard = ShardedTable(self.rt, rows, width, self.shard_spec[0], self.shard_spec[1], self.table_dtype)
            return self.shard.local
        assert not self.shard_spec, f"{type(self).__name__}: row-sharded tables use the all-to-all form (shard='a2a')"
        return EmbeddingTable(self.rt, rows, width, self.table_dtype, record=record)

    def _lookup(self, ids: IdsBatch):
        """(table, ids, route): the local table, or -- row-sharded -- the rows fetched over
        the all-to-all presented as a virtual table indexed by the inverse permutation."""
        if self.shard is None:
            return self.table, ids, None
        return self.shard.lookup(ids)

    def _table_grad(self, bag: torch.Tensor) -> SparseGrad:
        if self.shard is None:
            sg = SparseGrad(self.table, self._ctx["ids"], bag)
            sg.plan = self._ctx.get("plan")          # an overlapped plan started at forward time, if any
            return sg
        return self.shard.sparse_grad(self._ctx["route"], bag)


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

    # -- masked / weighted pooling of looked-up rows ----------------------------------
    def sequence_pool(self, ids, weights=None, mask=None, padding_index=None, reduce: bool = True, training: bool = False):
        """``ids`` [B, L, C] (a padded behaviour series of C feature columns, as ``tf.stack(series, axis=2)`` in
        7.SIM/CustomLayers.py:107) -> ``sum_l w[b,l] * [embed(ids[b,l,0]) | ... | embed(ids[b,l,C-1])]`` [B, C*k]
        (``reduce=False``: the weighted rows themselves, [B, L, C*k]); ``w = weights * mask * (ids[b,l,0] != padding_index)``,
        each factor optional (7.SIM/CustomLayers.py:88-95,116; 5.DIN/CustomLayers.py:258-283).  One fused kernel: the
        [B, L, C*k] tensor of the reference is never written in reduce mode."""
        rt = self.rt
        t = rt.to_device(ids, torch.int64).contiguous()
        assert t.dim() == 3, "ids must be [B, L, C]"
        B, L, C_ = t.shape
        k = self.output_dim
        w = rt.to_device(weights, torch.float32).contiguous() if weights is not None else None
        m = rt.to_device(mask, torch.uint8).contiguous() if mask is not None else None
        out = rt.empty((B, C_ * k) if reduce else (B, L, C_ * k))
        d = self.table.desc()
        check(rt.lib.etr_sequence_pool_forward(rt.ctx, C.byref(d), k, t.data_ptr(), B, L, C_, _p(w), _p(m),
                                               0 if padding_index is None else int(padding_index),
                                               0 if padding_index is None else 1, int(reduce), out.data_ptr(), rt.stream))
        if training:
            self._ctx = {"ids": t, "w": w, "m": m, "pad": padding_index, "reduce": reduce}
        self._finish(training)
        return out

    def sequence_pool_backward(self, dout: torch.Tensor, need_weight_grad: bool = True):
        """-> (SparseGrad of the table, d weights [B, L] or None)"""
        rt = self.rt
        c = self._ctx
        t = c["ids"]
        B, L, C_ = t.shape
        k = self.output_dim
        dout = dout.to(torch.float32).contiguous()
        occ = rt.empty((B * L * C_, self.table.grad_ld))
        dw = rt.empty((B, L)) if need_weight_grad else None
        d = self.table.desc()
        check(rt.lib.etr_sequence_pool_backward(rt.ctx, C.byref(d), k, t.data_ptr(), B, L, C_, _p(c["w"]), _p(c["m"]),
                                                0 if c["pad"] is None else int(c["pad"]), 0 if c["pad"] is None else 1,
                                                int(c["reduce"]), dout.data_ptr(), occ.data_ptr(), self.table.grad_ld, _p(dw),
                                                rt.stream))
        flat = IdsBatch(rt, t.reshape(-1, 1), B * L * C_, 1, 1, 1, 1, 1, pad_id=c["pad"])
        return SparseGrad(self.table, flat, occ), dw

    def weighted_lookup(self, keys, values, training: bool = False):
        """FiBiNet++ continuous-feature embedding: ``embed(keys)[B,F,k] * values[B,F,None]``
        (11.FiBiNet++/CustomLayers.py:124-126), one kernel."""
        keys = self.rt.to_device(keys, torch.int64)
        B, F_ = keys.shape
        out = self.sequence_pool(keys.reshape(B, F_, 1), weights=values, reduce=False, training=training)
        return out.reshape(B, F_, self.output_dim)


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
        if self.shard_spec:
            from .sharded import PeerShardedTable, ShardedTable
            if self.shard_mode in ("peer", "peer-pull"):
                assert self.table_dtype == torch.float32, "peer-sharded tables are fp32 in this round"
                self.peer = PeerShardedTable(self.rt, self.feature_dims, k + 1, self.shard_spec[0], self.shard_spec[1])
                self.table = self.peer                      # kernels see the GLOBAL table; views are the local rows
            else:
                self.shard = ShardedTable(self.rt, self.feature_dims, k + 1, self.shard_spec[0], self.shard_spec[1],
                                          self.table_dtype)
                self.table = self.shard.local               # this rank's rows r, r+G, r+2G, ...
            shard_gen = torch.Generator(device=self.rt.device)
            shard_gen.manual_seed(self.seed * 1000003 + self.shard_spec[1])
            self.table.init_uniform(-0.05, 0.05, shard_gen)
        else:
            # k = 16, fp32: the row and its Adam slots share one 256-byte record (two full lines per row in the
            # fused apply instead of twelve scattered chunks); ``record_layout=False`` keeps three plain arrays
            rec = self.record_layout and self.table_dtype == torch.float32 and k == 16
            self.table = EmbeddingTable(self.rt, self.feature_dims, k + 1, self.table_dtype, record=rec)
            self.table.init_uniform(-0.05, 0.05, self.gen)  # Keras Embedding default for embed and w

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
        tab, vids, route = self._lookup(ids)
        sumv = self._sumv_buffer(training, tab, vids)
        plan = SparsePlan(self.rt, vids, tab.rows, overlap=True) if sumv is not None else None   # sort || fwd/bwd
        self._prepare_plan(plan, tab)
        gather_fm_forward(tab, self.embedding_dims, True, vids, bias=self.bias, prob=prob, sumv=sumv)
        if training:
            self._ctx = {"ids": vids, "table": tab, "route": route, "sumv": sumv, "plan": plan}
        self._finish(training)
        return {"output": prob}

    def _prepare_plan(self, plan, tab) -> None:
        """Row descriptors / long-run items of the tiled fused apply: built behind the sort, off the critical path."""
        if plan is not None and getattr(tab, "record", False) and self.embedding_dims == 16 and FusedFMGrad.apply_kernel == "tile":
            plan.prepare_fm()

    def _sumv_buffer(self, training: bool, tab, vids) -> Optional[torch.Tensor]:
        """S = sum_f v_f [B,k], saved for the fused backward+apply (single-hot, fp32, unsharded)."""
        if training and self.shard is None and self.fused_apply and FusedFMGrad.eligible(tab, vids, self.embedding_dims):
            return self.rt.empty((vids.B, self.embedding_dims))
        return None

    def _fused_grad(self, fused: FusedFMGrad):
        if self.peer is None:
            return fused
        from .sharded import PeerFMGrad
        return PeerFMGrad(self.peer, fused)

    def backward(self, dlogit: torch.Tensor) -> List[SparseGrad]:
        """dlogit = dL/dz [B] (z the pre-sigmoid logit).  Dense grads land in
        ``self.params.grad``; the table gradient is returned as a SparseGrad."""
        ids = self._ctx["ids"]
        dl = dlogit.reshape(-1)
        check(self.rt.lib.etr_colsum_f32(self.rt.ctx, dl.data_ptr(), ids.B, 1, 1, self.params.g("bias").data_ptr(),
                                         self.rt.stream))
        if self._ctx.get("sumv") is not None:           # fused backward + segment reduction + Adam
            return [self._fused_grad(FusedFMGrad(self.table, ids, self.embedding_dims, dl, self._ctx["sumv"],
                                                 plan=self._ctx["plan"]))]
        assert self.peer is None, "peer-sharded tables need the fused FM gradient (single-hot ids, fp32, k % 4 == 0)"
        bag = gather_fm_backward(self._ctx["table"], self.embedding_dims, True, ids, dlogit=dl)
        return [self._table_grad(bag)]


# ---------------------------------------------------------------------------
class DeepFMRankingLayer(FMRankingLayer):
    """FM logit + MLP([32,8],relu) -> MLP([1]) over Flatten(emb); sigma(fm+dnn)
    (2.FM/CustomLayers.py:241-308).  Build extension (c2):
    ``continuous_features`` are concatenated in front of the flattened embedding
    for the MLP only, as 3.DCN/CustomLayers.py:259 does."""

    def __init__(self, feature_names=['user_tag0', 'user_tag1', 'item_tag1', 'item_tag2', 'item_tag3'],
                 feature_dims=20, embedding_dims=16, mlp_dims=[32, 8], continuous_features=(),
                 mlp_precision="fp32", **kwargs):
        self.mlp_dims = list(mlp_dims)
        self.continuous_features = list(continuous_features)
        # "bf16": the wide first MLP layer runs on the tensor cores (bf16 operands, fp32 accumulate;
        # parity 1e-2); "fp32": exact-parity SIMT path (1e-5)
        self.mlp_precision = mlp_precision
        self.fused_tail = bool(kwargs.pop("fused_tail", True))   # K7c: tower tail + loss fwd/bwd in one launch
        super().__init__(feature_names, feature_dims, embedding_dims, **kwargs)

    def _build_extra(self):
        C_ = len(self.continuous_features)
        in_dim = C_ + len(self.feature_names) * self.embedding_dims
        # front padding so that Flatten(emb) starts on a 16-byte boundary behind [pad | X_cont]
        self.front_pad = (-C_) % (8 if self.mlp_precision == "bf16" else 4)
        self.MLP_layer1 = MLPLayer(units=self.mlp_dims, activation="relu", name="MLP_layer1",
                                   precision=self.mlp_precision)
        self.MLP_layer2 = MLPLayer(units=[1], name="MLP_layer2")
        self.MLP_layer1.build(in_dim, self.params, self.gen, front_pad=self.front_pad)
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
        col0 = self.front_pad + C_
        x = rt.empty((ids.B, col0 + F * k),                  # [pad | X_cont | Flatten(emb)]
                     torch.bfloat16 if self.mlp_precision == "bf16" else torch.float32)
        cont = self._cont(inputs, self.continuous_features) if C_ else None
        fm_logit = rt.empty((ids.B,))
        tab, vids, route = self._lookup(ids)
        sumv = self._sumv_buffer(training, tab, vids)
        plan = SparsePlan(rt, vids, tab.rows, overlap=True) if sumv is not None else None        # sort || fwd/bwd
        self._prepare_plan(plan, tab)
        gather_fm_forward(tab, k, True, vids, bias=self.bias, logit=fm_logit, sumv=sumv, flat=x, flat_col0=col0,
                          cont=cont)
        dnn = self.MLP_layer2(self.MLP_layer1(x, training=training), training=training)      # [B,1]
        prob = rt.empty((ids.B, 1))
        check(rt.lib.etr_add_sigmoid(rt.ctx, fm_logit.data_ptr(), dnn.data_ptr(), ids.B, None, prob.data_ptr(),
                                     rt.stream))
        if training:
            self._ctx = {"ids": vids, "table": tab, "route": route, "sumv": sumv, "plan": plan}
        self._finish(training)
        return {"output": prob}

    # -- fused train step (K1 -> layer 1 on tcgen05 -> K7c tail+loss fwd/bwd -> K7b layer-1 backward) ------
    def fused_train_ok(self) -> bool:
        m1 = self.MLP_layer1
        return (self.fused_tail and self.shard is None and self.mlp_precision == "bf16" and self.mlp_dims == [32, 8]
                and m1.activation == "relu"
                and m1.use_bias and self.MLP_layer2.use_bias and self.MLP_layer2.activation is None
                and (self.front_pad + len(self.continuous_features) + len(self.feature_names) * self.embedding_dims) % 16 == 0)

    def train_forward_backward(self, inputs, labels: torch.Tensor, grad_scale: float = 1.0, plan: Optional[SparsePlan] = None):
        """The whole reference train step up to apply_gradients (2.FM/ModelManager.py:172-176) for the default
        DeepFM tower, in 6 launches: returns (loss [1], prob [B,1], table gradients) or None when this batch
        is not eligible (the caller then takes the layer-by-layer path).  Dense grads land in params.grad."""
        rt = self.rt
        ids = self._ids(inputs, self.feature_names)
        C_ = len(self.continuous_features)
        k, F = self.embedding_dims, len(self.feature_names)
        col0 = self.front_pad + C_
        tab, vids, route = self._lookup(ids)
        if route is not None or not FusedFMGrad.eligible(tab, vids, k) or not self.fused_apply or col0 % 4:
            return None
        B, n_in = ids.B, col0 + F * k
        x = rt.empty((B, n_in), torch.bfloat16)                # [pad | X_cont | Flatten(emb)]
        cont = self._cont(inputs, self.continuous_features) if C_ else None
        fm_logit = rt.empty((B,))
        sumv = rt.empty((B, k))
        slot_of_u = None
        if self.peer is not None and self.shard_mode == "peer":
            # de-duplicated exchange: the sorted plan comes first, the owners write the unique rows into this rank's
            # response buffer, and the gather below runs on that buffer (local HBM, mostly L2-resident)
            # ``plan``: this batch's plan, sorted ahead of time during the previous step (Trainer next_batch=)
            plan = plan.join() if plan is not None else SparsePlan(rt, vids, tab.rows)
            gtab, gids, slot_of_u = self.peer.exchange_forward(plan, B, F)
        else:
            if plan is not None:
                plan = plan.join()                                 # sorted (and prepared) during the previous step
            else:
                plan = SparsePlan(rt, vids, tab.rows, overlap=True)    # sort || everything below
                self._prepare_plan(plan, tab)
            gtab, gids = tab, vids
        gather_fm_forward(gtab, k, True, gids, bias=self.bias, logit=fm_logit, sumv=sumv, flat=x, flat_col0=col0,
                          cont=cont)
        m1, m2, P = self.MLP_layer1, self.MLP_layer2, self.params
        k0 = P.full(f"{m1.name}/kernel_0")
        y1 = rt.empty((B, 32))
        gemm_bf16_tn(rt, x, cast_bf16(rt, k0, transpose=True), y1, B, 32, n_in, bias=P[f"{m1.name}/bias_0"], act="relu")
        prob, dlogit, d1, loss = rt.empty((B, 1)), rt.empty((B,)), rt.empty((B, 32)), rt.empty((1,))
        check(rt.lib.etr_deepfm_tail_train(
            rt.ctx, y1.data_ptr(), fm_logit.data_ptr(), labels.data_ptr(), B, P[f"{m1.name}/kernel_1"].data_ptr(),
            P[f"{m1.name}/bias_1"].data_ptr(), P[f"{m2.name}/kernel_0"].data_ptr(), P[f"{m2.name}/bias_0"].data_ptr(),
            float(grad_scale), prob.data_ptr(), dlogit.data_ptr(), d1.data_ptr(), loss.data_ptr(),
            P.g("bias").data_ptr(), P.g(f"{m1.name}/kernel_1").data_ptr(), P.g(f"{m1.name}/bias_1").data_ptr(),
            P.g(f"{m2.name}/kernel_0").data_ptr(), P.g(f"{m2.name}/bias_0").data_ptr(), rt.stream))
        dx = rt.empty((B, n_in), torch.bfloat16)
        check(rt.lib.etr_mlp_skinny_backward(rt.ctx, x.data_ptr(), n_in, d1.data_ptr(), k0.data_ptr(), B, n_in, 32,
                                             dx.data_ptr(), n_in, P.gfull(f"{m1.name}/kernel_0").data_ptr(),
                                             P.g(f"{m1.name}/bias_0").data_ptr(), rt.stream))
        fused = FusedFMGrad(self.table, vids, k, dlogit, sumv, dx, col0, plan=plan)
        if slot_of_u is not None:
            from .sharded import PeerSlotGrad
            return loss, prob, [PeerSlotGrad(self.peer, fused, slot_of_u)]
        return loss, prob, [self._fused_grad(fused)]

    def backward(self, dlogit: torch.Tensor) -> List[SparseGrad]:
        rt = self.rt
        ids = self._ctx["ids"]
        dl = dlogit.reshape(-1)
        col0 = self.front_pad + len(self.continuous_features)
        d_dnn = dl.clone().reshape(-1, 1)                    # d(fm+dnn)/d dnn = 1
        dh = self.MLP_layer2.backward(d_dnn)
        dx = self.MLP_layer1.backward(dh)                    # [B, pad + C + F*k]
        check(rt.lib.etr_colsum_f32(rt.ctx, dl.data_ptr(), ids.B, 1, 1, self.params.g("bias").data_ptr(), rt.stream))
        if self._ctx.get("sumv") is not None and col0 % 4 == 0 and dx.stride(0) % 4 == 0:
            return [self._fused_grad(FusedFMGrad(self.table, ids, self.embedding_dims, dl, self._ctx["sumv"], dx, col0,
                                                 plan=self._ctx["plan"]))]
        assert self.peer is None, "peer-sharded tables need the fused FM gradient (single-hot ids, fp32, k % 4 == 0)"
        bag = gather_fm_backward(self._ctx["table"], self.embedding_dims, True, ids, dlogit=dl, dflat=dx,
                                 flat_col0=col0)
        return [self._table_grad(bag)]


# ---------------------------------------------------------------------------
class FieldAwareInteractionLayer(_Layer):
    """``call(X[B,F]) -> [B,P,k]``: I[b,(a,c),:] = T[x_a,c,:] * T[x_c,a,:], a<c in
    row-major order (2.FM/CustomLayers.py:428-462); weight ``v`` [V,F,k] with the
    reference's ctor defaults feature_dims=20, embedding_dims=16 (:430)."""

    def __init__(self, fields_cnt, feature_dims=20, embedding_dims=16, _table: Optional[EmbeddingTable] = None,
                 **kwargs):
        self._setup(kwargs)
        self.fields_cnt, self.feature_dims, self.embedding_dims = int(fields_cnt), int(feature_dims), int(embedding_dims)
        Fk = self.fields_cnt * self.embedding_dims
        if _table is None:
            self.table = EmbeddingTable(self.rt, self.feature_dims, Fk)
            init = glorot_uniform((self.feature_dims, self.fields_cnt, self.embedding_dims), self.gen, self.rt.device)
            self.table.data[:, :Fk] = init.reshape(self.feature_dims, Fk)
            self.has_w = False
        else:                                   # shared with an FFM/FwFM head (w fused at column F*k)
            self.table, self.has_w = _table, True
        self.params.finalize()

    @property
    def embedding_lookup_table(self) -> torch.Tensor:          # 'v' [V,F,k]
        Fk = self.fields_cnt * self.embedding_dims
        return self.table.data[:, :Fk].unflatten(1, (self.fields_cnt, self.embedding_dims))

    def sparse_tables(self):
        return [self.table]

    def call(self, X, training: bool = False):
        rt = self.rt
        ids = IdsBatch.make(rt, X, None, self.pad_id, self.pooling)
        F, k = self.fields_cnt, self.embedding_dims
        assert ids.F == F
        P = F * (F - 1) // 2
        out = rt.empty((ids.B, P, k))
        pooled = rt.empty((ids.B, F, F * k)) if training else None
        t, d = self.table.desc(), ids.desc()
        check(rt.lib.etr_field_pair_forward(rt.ctx, C.byref(t), k, int(self.has_w), C.byref(d), None, None, None,
                                            out.data_ptr(), None, None, None, _p(pooled), rt.stream))
        if training:
            self._ctx = {"ids": ids, "pooled": pooled}
        self._finish(training)
        return out

    def backward(self, dpairvec: torch.Tensor) -> List[SparseGrad]:
        rt = self.rt
        ids, pooled = self._ctx["ids"], self._ctx["pooled"]
        bag = rt.empty((ids.B * ids.F, self.table.grad_ld))
        d = ids.desc()
        dpv = dpairvec.contiguous()
        check(rt.lib.etr_field_pair_backward(rt.ctx, self.embedding_dims, int(self.has_w), C.byref(d),
                                             pooled.data_ptr(), None, None, dpv.data_ptr(), bag.data_ptr(),
                                             self.table.grad_ld, rt.stream))
        return [SparseGrad(self.table, ids, bag)]


class _FieldPairHead(_Layer):
    """Shared body of FFMLayer / FwFMLayer / FFMRankingLayer: one fused HBM table
    [V, F*k + 1] (pair rows + the linear weight w at column F*k), one kernel per
    forward, one per backward."""

    fwfm = False

    def _init_head(self, feature_names, feature_dims, embedding_dims, kwargs, table_init):
        self._setup(kwargs)
        self.feature_names = list(feature_names)
        self.feature_dims, self.embedding_dims = int(feature_dims), int(embedding_dims)
        self.fields_cnt = len(self.feature_names)
        F, k, V = self.fields_cnt, self.embedding_dims, self.feature_dims
        self.P = F * (F - 1) // 2
        self.params.add("bias", glorot_uniform((1,), self.gen, self.rt.device))
        if self.fwfm:                                           # tf.keras.layers.Dense(1): glorot kernel, zero bias
            self.params.add("interaction_weights/kernel", glorot_uniform((self.P, 1), self.gen, self.rt.device))
            self.params.add("interaction_weights/bias", torch.zeros(1, device=self.rt.device))
        self.table = EmbeddingTable(self.rt, V, F * k + 1)
        table_init(self.table)
        self.params.finalize()

    @property
    def bias(self):
        return self.params["bias"]

    @property
    def w(self) -> torch.Tensor:
        Fk = self.fields_cnt * self.embedding_dims
        return self.table.cols(Fk, Fk + 1)

    def sparse_tables(self):
        return [self.table]

    def call(self, inputs, training: bool = False):
        rt = self.rt
        ids = self._ids(inputs, self.feature_names)
        F, k = self.fields_cnt, self.embedding_dims
        prob = rt.empty((ids.B, 1))
        pooled = rt.empty((ids.B, F, F * k)) if training else None
        pairdot = rt.empty((ids.B, self.P)) if (training and self.fwfm) else None
        r = self.params["interaction_weights/kernel"] if self.fwfm else None
        r0 = self.params["interaction_weights/bias"] if self.fwfm else None
        t, d = self.table.desc(), ids.desc()
        check(rt.lib.etr_field_pair_forward(rt.ctx, C.byref(t), k, 1, C.byref(d), self.bias.data_ptr(), _p(r), _p(r0),
                                            None, _p(pairdot), None, prob.data_ptr(), _p(pooled), rt.stream))
        if training:
            self._ctx = {"ids": ids, "pooled": pooled, "pairdot": pairdot}
        self._finish(training)
        return {"output": prob}

    def backward(self, dlogit: torch.Tensor) -> List[SparseGrad]:
        rt = self.rt
        ids, pooled, pairdot = self._ctx["ids"], self._ctx["pooled"], self._ctx["pairdot"]
        dl = dlogit.reshape(-1).contiguous()
        bag = rt.empty((ids.B * ids.F, self.table.grad_ld))
        r = self.params["interaction_weights/kernel"] if self.fwfm else None
        d = ids.desc()
        check(rt.lib.etr_field_pair_backward(rt.ctx, self.embedding_dims, 1, C.byref(d), pooled.data_ptr(),
                                             dl.data_ptr(), _p(r), None, bag.data_ptr(), self.table.grad_ld,
                                             rt.stream))
        check(rt.lib.etr_colsum_f32(rt.ctx, dl.data_ptr(), ids.B, 1, 1, self.params.g("bias").data_ptr(), rt.stream))
        if self.fwfm:
            from .runtime import gemm_f32
            # dr = pairdot^T dlogit  ([P,B] x [B,1]); dr0 = sum_b dlogit
            gemm_f32(rt, pairdot, dl, self.params.g("interaction_weights/kernel"), self.P, 1, ids.B, self.P, 1, 1,
                     trans_a=True)
            check(rt.lib.etr_colsum_f32(rt.ctx, dl.data_ptr(), ids.B, 1, 1,
                                        self.params.g("interaction_weights/bias").data_ptr(), rt.stream))
        return [SparseGrad(self.table, ids, bag)]


class FFMLayer(_FieldPairHead):
    """sigma(bias + sum_f w[x_f] + sum_{p,k} I)  (2.FM/CustomLayers.py:465-497).

    Decision on a reference quirk: the reference builds
    ``FieldAwareInteractionLayer(self.fields_cnt)`` WITHOUT forwarding
    feature_dims / embedding_dims (:477), so its pair table is always (20,F,16)
    and any id >= 20 is out of range; here they are forwarded."""

    def __init__(self, feature_names=['item_tag1', 'item_tag2', 'item_tag3', 'user_tag0', 'user_tag1'],
                 feature_dims=20, embedding_dims=16, **kwargs):
        def init(table):
            F, k, V = len(feature_names), embedding_dims, feature_dims
            table.data[:, :F * k] = glorot_uniform((V, F, k), self.gen, self.rt.device).reshape(V, F * k)
            table.data[:, F * k:F * k + 1] = glorot_uniform((V, 1), self.gen, self.rt.device)   # add_weight default
        self._init_head(feature_names, feature_dims, embedding_dims, kwargs, init)
        self.fa_interaction_layer = FieldAwareInteractionLayer(self.fields_cnt, self.feature_dims,
                                                               self.embedding_dims, _table=self.table,
                                                               device=self.rt.device, pad_id=self.pad_id,
                                                               pooling=self.pooling)

    @property
    def variables(self):
        v = [self.bias, self.w, self.fa_interaction_layer.embedding_lookup_table]
        if self.fwfm:
            v += [self.params["interaction_weights/kernel"], self.params["interaction_weights/bias"]]
        return v


class FwFMLayer(FFMLayer):
    """sigma(bias + sum w + Dense(1)(sum_k I))  (2.FM/CustomLayers.py:500-533)."""
    fwfm = True


class FFMRankingLayer(_FieldPairHead):
    """F separate Embedding tables T_i [V,k] (2.FM/CustomLayers.py:370-425):
    sum_{i<j} <T_i[x_j], T_j[x_i]> + sum w + bias -> sigma.  Stored as one fused
    [V, F*k+1] table with T_i at columns i*k..(i+1)*k (T_i[v] == T[v,i], KAT-2)."""

    def __init__(self, feature_names=['item_tag1', 'item_tag2', 'item_tag3'], feature_dims=20, embedding_dims=16,
                 **kwargs):
        self._init_head(feature_names, feature_dims, embedding_dims, kwargs,
                        lambda table: table.init_uniform(-0.05, 0.05, self.gen))

    @property
    def embedding_list(self) -> List[torch.Tensor]:
        k = self.embedding_dims
        return [self.table.cols(i * k, (i + 1) * k) for i in range(self.fields_cnt)]

    @property
    def variables(self):
        return [self.bias, self.w] + self.embedding_list


# ---------------------------------------------------------------------------
_PNN_TYPE = {"inner": 0, "mat": 1, "vec": 2, "num": 3}


class InnerProductNetwork(_Layer):
    """``call(x[B,F,k]) -> [B,P]``, out[b,p] = <x_i,x_j>, pairs i<j in combinations
    order (2.FM/CustomLayers.py:601-624; vectorised twin IpnLayer :775-792)."""

    kernel_type = "inner"

    def __init__(self, **kwargs):
        self._setup(kwargs)
        self.kernel = None
        self.params.finalize()

    def call(self, x, training: bool = False):
        rt = self.rt
        x = rt.to_device(x, torch.float32).contiguous()
        B, F, k = x.shape
        P = F * (F - 1) // 2
        out = rt.empty((B, P))
        check(rt.lib.etr_pnn_forward(rt.ctx, x.data_ptr(), F * k, B, F, k, _PNN_TYPE[self.kernel_type],
                                     _p(self.kernel), out.data_ptr(), P, rt.stream))
        if training:
            self._ctx = {"x": x}
        return out

    def backward(self, g: torch.Tensor):
        """returns dx [B,F,k]; the kernel gradient lands in ``self.kernel_grad``."""
        rt = self.rt
        x = self._ctx["x"]
        B, F, k = x.shape
        g = g.contiguous()
        dx = rt.zeros((B, F, k))
        self.kernel_grad = torch.empty_like(self.kernel) if self.kernel is not None else None
        check(rt.lib.etr_pnn_backward(rt.ctx, x.data_ptr(), F * k, B, F, k, _PNN_TYPE[self.kernel_type],
                                      _p(self.kernel), g.data_ptr(), g.shape[1], dx.data_ptr(), F * k,
                                      _p(self.kernel_grad), rt.stream))
        return dx


IpnLayer = InnerProductNetwork


class OuterProductNetwork(InnerProductNetwork):
    """kernel 'mat' [k,P,k]: out[b,p] = sum_{a,c} x_i[c] K[a,p,c] x_j[a]; 'vec' [P,k];
    'num' [P,1]; init random_normal (2.FM/CustomLayers.py:627-682; OpnLayer :795-851)."""

    def __init__(self, fields_cnt, embedding_dims, kernel_type=None, **kwargs):
        self._setup(kwargs)
        kernel_type = kernel_type or "mat"
        assert kernel_type in ("mat", "vec", "num")
        self.kernel_type = kernel_type
        P = fields_cnt * (fields_cnt - 1) // 2
        self.kernel_shape = {"mat": (embedding_dims, P, embedding_dims), "vec": (P, embedding_dims),
                             "num": (P, 1)}[kernel_type]
        self.kernel = random_normal(self.kernel_shape, self.gen, self.rt.device)
        self.params.finalize()


OpnLayer = OuterProductNetwork


class PNNRankingLayer(_Layer):
    """MLP2(sigmoid)(MLP1(relu)(concat[Flatten(emb), product]))
    (2.FM/CustomLayers.py:536-599; vectorised twin PNNLayer :685-753)."""

    def __init__(self, feature_names=['user_tag0', 'user_tag1', 'item_tag1', 'item_tag2', 'item_tag3'],
                 feature_dims=20, embedding_dims=16, mlp_dims=[32, 8], dropout=0, method='inner', kernel_type=None,
                 **kwargs):
        assert method in ('inner', 'outer')
        self._setup(kwargs)
        self.feature_names = list(feature_names)
        self.feature_dims, self.embedding_dims = int(feature_dims), int(embedding_dims)
        self.fields_cnt = len(self.feature_names)
        self.mlp_dims, self.method, self.dropout = list(mlp_dims), method, dropout
        self.kernel_type = "inner" if method == "inner" else (kernel_type or "mat")
        F, k = self.fields_cnt, self.embedding_dims
        self.P = F * (F - 1) // 2
        self.table = self._make_table(self.feature_dims, k)
        self.table.init_uniform(-0.05, 0.05, self.gen)
        if method == "outer":
            shape = {"mat": (k, self.P, k), "vec": (self.P, k), "num": (self.P, 1)}[self.kernel_type]
            self.params.add("pn/kernel", random_normal(shape, self.gen, self.rt.device))
        self.MLP_layer1 = MLPLayer(units=self.mlp_dims, activation='relu', is_dropput=dropout, name="MLP_layer1")
        self.MLP_layer2 = MLPLayer(units=[1], activation='sigmoid', name="MLP_layer2")
        self.MLP_layer1.build(F * k + self.P, self.params, self.gen)
        self.MLP_layer2.build(self.mlp_dims[-1], self.params, self.gen)
        self.params.finalize()

    @property
    def embed(self):
        return self.table.cols(0, self.embedding_dims)

    @property
    def pn_kernel(self):
        return self.params["pn/kernel"] if self.method == "outer" else None

    def sparse_tables(self):
        return [self.table]

    def call(self, inputs, training: bool = False):
        rt = self.rt
        ids = self._ids(inputs, self.feature_names)
        F, k, P = self.fields_cnt, self.embedding_dims, self.P
        comb = rt.empty((ids.B, F * k + P))                        # concat fused: [Flatten(emb) | product]
        tab, vids, route = self._lookup(ids)
        gather_fm_forward(tab, k, False, vids, flat=comb, flat_col0=0)
        check(rt.lib.etr_pnn_forward(rt.ctx, comb.data_ptr(), F * k + P, ids.B, F, k, _PNN_TYPE[self.kernel_type],
                                     _p(self.pn_kernel), comb[:, F * k:].data_ptr(), F * k + P, rt.stream))
        out = self.MLP_layer2(self.MLP_layer1(comb, training=training), training=training)
        if training:
            self._ctx = {"ids": vids, "comb": comb, "table": tab, "route": route}
        self._finish(training)
        return {"output": out}

    def backward(self, dlogit: torch.Tensor) -> List[SparseGrad]:
        """dlogit = dL/dz, z the pre-sigmoid input of MLP_layer2's activation."""
        rt = self.rt
        ids, comb = self._ctx["ids"], self._ctx["comb"]
        F, k, P = self.fields_cnt, self.embedding_dims, self.P
        dz = dlogit.reshape(-1, 1).clone()
        dh = self.MLP_layer2.backward(dz, dy_is_preact=True)
        dcomb = self.MLP_layer1.backward(dh)                       # [B, F*k + P]
        dk = None
        if self.method == "outer":
            dk = self.params.g("pn/kernel")            # written by the deterministic batch reduction
        check(rt.lib.etr_pnn_backward(rt.ctx, comb.data_ptr(), F * k + P, ids.B, F, k, _PNN_TYPE[self.kernel_type],
                                      _p(self.pn_kernel), dcomb[:, F * k:].data_ptr(), F * k + P, dcomb.data_ptr(),
                                      F * k + P, _p(dk), rt.stream))
        bag = gather_fm_backward(self._ctx["table"], k, False, ids, dflat=dcomb, flat_col0=0)
        return [self._table_grad(bag)]


PNNLayer = PNNRankingLayer


# ---------------------------------------------------------------------------
# the other pairwise consumers of the same rows (dense [B,F,k] in, one kernel each way)
class _PairDense(_Layer):
    mode = 0

    def _w(self):
        return None, 0

    def _fwd(self, x, training):
        rt = self.rt
        x = rt.to_device(x, torch.float32).contiguous()
        B, F, k = x.shape
        P = F * (F - 1) // 2
        out = rt.empty((B, k)) if self.mode == 2 else rt.empty((B, P, k))
        W, kind = self._w()
        check(rt.lib.etr_pair_dense_forward(rt.ctx, x.data_ptr(), F * k, B, F, k, self.mode, kind, _p(W), out.data_ptr(),
                                            k if self.mode == 2 else P * k, rt.stream))
        if training:
            self._ctx = {"x": x}
        return out

    def call(self, x, training: bool = False):
        return self._fwd(x, training)

    def backward(self, g: torch.Tensor) -> torch.Tensor:
        """upstream gradient in the output layout -> dx [B,F,k]; a weight gradient lands in ``self.W_grad``."""
        rt = self.rt
        x = self._ctx["x"]
        B, F, k = x.shape
        g = g.to(torch.float32).contiguous()
        dx = rt.zeros((B, F, k))
        W, kind = self._w()
        self.W_grad = torch.empty_like(W) if W is not None else None
        check(rt.lib.etr_pair_dense_backward(rt.ctx, x.data_ptr(), F * k, B, F, k, self.mode, kind, _p(W), g.data_ptr(),
                                             g.numel() // B, dx.data_ptr(), F * k, _p(self.W_grad), rt.stream))
        return dx


class InteractionLayer(_PairDense):
    """AFM pair vectors: ``call(x[B,F,k]) -> [B,P,k]``, out[b,(i,j),:] = x_i * x_j, i<j in loop order
    (3.DCN/CustomLayers.py:825-838)."""
    mode = 0

    def __init__(self, **kwargs):
        self._setup(kwargs)
        self.params.finalize()


class BiInteractionPooling(_PairDense):
    """NFM second-order pooling: ``call(x[B,F,k]) -> [B,k]`` = 0.5 ((sum_f x)^2 - sum_f x^2)
    (inline in NFMLayer.call, 3.DCN/CustomLayers.py:499-501)."""
    mode = 2

    def __init__(self, **kwargs):
        self._setup(kwargs)
        self.params.finalize()


class BilinearInteractionLayer(_PairDense):
    """FiBiNet: ``call(x[B,F,k]) -> [B,P,k]``, out[b,(i,j),:] = (x_i W) * x_j; bilinear_type 'all' (one W [k,k]),
    'each' (W_i per LEFT field, F-1 of them), 'interaction' (W_p per pair); glorot_normal init
    (3.DCN/CustomLayers.py:977-1009).  Weights: ``self.W`` [n_w,k,k] (``W_list[i] == W[i]``)."""
    mode = 1
    _KIND = {"all": 0, "each": 1, "interaction": 2}

    def __init__(self, bilinear_type='interaction', **kwargs):
        if bilinear_type not in self._KIND:
            raise NotImplementedError
        self._setup(kwargs)
        self.bilinear_type = bilinear_type
        self.W = None
        self.params.finalize()

    def build(self, F: int, k: int):
        n_w = {"all": 1, "each": F - 1, "interaction": F * (F - 1) // 2}[self.bilinear_type]
        std = (2.0 / (k + k)) ** 0.5                     # glorot_normal: truncated normal, stddev sqrt(2/(fan_in+fan_out))
        W = torch.empty((n_w, k, k), device=self.rt.device)
        torch.nn.init.trunc_normal_(W, mean=0.0, std=std / 0.87962566103423978, a=-2 * std / 0.87962566103423978,
                                    b=2 * std / 0.87962566103423978, generator=self.gen)
        self.W = W
        return self

    @property
    def W_list(self):
        return [self.W[i] for i in range(self.W.shape[0])]

    def _w(self):
        return self.W, self._KIND[self.bilinear_type]

    def call(self, x, training: bool = False):
        if self.W is None:
            self.build(x.shape[1], x.shape[2])
        return self._fwd(x, training)


class ParralledOnnLayer(_Layer):
    """ONN (2.FM/CustomLayers.py:957-1006): ``interaction(inputs)`` = [Flatten(embedding_single(X)) | Flatten(pair
    vectors of FieldAwareInteractionLayer(X))] (pair vectors summed over k when ``reduce``) -- one gather kernel + one
    field-pair kernel.  The reference then applies ``make_mlp_layer(mlp_units, sigmoid_units=True)`` (Dense +
    LayerNormalization + PReLU stacks, :868-888), which is not on the hot path: pass it as ``mlp_layer`` (any callable on
    a [B, D] CUDA tensor, e.g. a torch module) or use ``interaction`` / ``interaction_backward`` directly.

    Reference quirk fixed as for FFMLayer: the reference builds ``FieldAwareInteractionLayer(self.fields_cnt)`` without
    forwarding feature_dims / embedding_dims (:982); here they are forwarded.  ``ONNLayer`` (the loop form, one table
    pair per field pair, :891-955) computes the same function whenever the fields' id ranges are disjoint (the
    DataGenerator id space): E1_ij[v] = T[v, j, :], E2_ij[v] = T[v, i, :]."""

    def __init__(self, feature_names=['item_tag1', 'item_tag2', 'item_tag3', 'user_tag0'], feature_dims=20, embedding_dims=16,
                 mlp_units=[40, 20], reduce=False, mlp_layer=None, **kwargs):
        self._setup(kwargs)
        self.feature_names = list(feature_names)
        self.feature_dims, self.embedding_dims = int(feature_dims), int(embedding_dims)
        self.fields_cnt = len(self.feature_names)
        self.mlp_units, self.reduce, self.mlp_layer = list(mlp_units), bool(reduce), mlp_layer
        self.table = EmbeddingTable(self.rt, self.feature_dims, self.embedding_dims, self.table_dtype)
        self.table.init_uniform(-0.05, 0.05, self.gen)
        self.fa_interaction_layer = FieldAwareInteractionLayer(self.fields_cnt, self.feature_dims, self.embedding_dims,
                                                               device=self.rt.device, seed=self.seed + 1)
        self.params.finalize()

    @property
    def embedding_single(self) -> torch.Tensor:
        return self.table.cols(0, self.embedding_dims)

    def sparse_tables(self):
        return [self.table, self.fa_interaction_layer.table]

    @property
    def out_dim(self) -> int:
        F, k = self.fields_cnt, self.embedding_dims
        P = F * (F - 1) // 2
        return F * k + (P if self.reduce else P * k)

    def interaction(self, inputs, training: bool = False) -> torch.Tensor:
        rt = self.rt
        ids = self._ids(inputs, self.feature_names)
        F, k = self.fields_cnt, self.embedding_dims
        P = F * (F - 1) // 2
        comb = rt.empty((ids.B, self.out_dim))
        gather_fm_forward(self.table, k, False, ids, flat=comb, flat_col0=0)
        fa = self.fa_interaction_layer
        pooled = rt.empty((ids.B, F, F * k)) if training else None
        t, d = fa.table.desc(), ids.desc()
        if self.reduce:
            pairdot = rt.empty((ids.B, P))
            check(rt.lib.etr_field_pair_forward(rt.ctx, C.byref(t), k, 0, C.byref(d), None, None, None, None,
                                                pairdot.data_ptr(), None, None, _p(pooled), rt.stream))
            comb[:, F * k:] = pairdot
        else:
            pv = rt.empty((ids.B, P, k))
            check(rt.lib.etr_field_pair_forward(rt.ctx, C.byref(t), k, 0, C.byref(d), None, None, None, pv.data_ptr(),
                                                None, None, None, _p(pooled), rt.stream))
            comb[:, F * k:] = pv.reshape(ids.B, P * k)
        if training:
            self._ctx = {"ids": ids, "pooled": pooled}
        self._finish(training)
        return comb

    def interaction_backward(self, dcomb: torch.Tensor) -> List[SparseGrad]:
        """d X_combined [B, out_dim] -> the two table gradients (embedding_single, pair table)."""
        rt = self.rt
        ids, pooled = self._ctx["ids"], self._ctx["pooled"]
        F, k = self.fields_cnt, self.embedding_dims
        P = F * (F - 1) // 2
        dcomb = dcomb.to(torch.float32).contiguous()
        bag1 = gather_fm_backward(self.table, k, False, ids, dflat=dcomb, flat_col0=0)
        dpv = dcomb[:, F * k:]
        if self.reduce:
            dpv = dpv.unsqueeze(2).expand(ids.B, P, k)
        dpv = dpv.reshape(ids.B, P, k).contiguous()
        fa = self.fa_interaction_layer
        bag2 = rt.empty((ids.B * ids.F, fa.table.grad_ld))
        d = ids.desc()
        check(rt.lib.etr_field_pair_backward(rt.ctx, k, 0, C.byref(d), pooled.data_ptr(), None, None, dpv.data_ptr(),
                                             bag2.data_ptr(), fa.table.grad_ld, rt.stream))
        return [SparseGrad(self.table, ids, bag1), SparseGrad(fa.table, ids, bag2)]

    def call(self, inputs, training: bool = False):
        if self.mlp_layer is None:
            raise NotImplementedError("ParralledOnnLayer.call needs mlp_layer= (the reference's Dense + LayerNormalization + "
                                      "PReLU tower is outside the hot path); interaction() is the kernel part")
        return {"output": self.mlp_layer(self.interaction(inputs, training=training))}


ONNLayer = ParralledOnnLayer


# ---------------------------------------------------------------------------
class CrossLayer(_Layer):
    """x_{l+1} = x0*(x_l^T w_l) + b_l + x_l  (3.DCN/CustomLayers.py:170-203);
    ``cross_weight`` / ``cross_bias``: layer_num variables of shape [D,1]."""

    matrix = False

    def __init__(self, layer_num, reg_w=1e-4, reg_b=1e-4, precision="fp32", **kwargs):
        self._setup(kwargs)
        assert precision in ("fp32", "bf16")
        self.precision = precision
        self.layer_num, self.reg_w, self.reg_b = int(layer_num), reg_w, reg_b
        self.D = None
        self.front_pad = 0
        self._external_params = False

    def build(self, D: int, params: Optional[DenseParams] = None, gen=None, front_pad: int = 0):
        self.D, self.front_pad = int(D), int(front_pad)
        Di = self.D + self.front_pad
        gen = gen or self.gen
        if params is not None:
            self.params, self._external_params = params, True
        L = self.layer_num
        if self.matrix:
            W = torch.zeros((L, Di, Di), device=self.rt.device)
            W[:, front_pad:, front_pad:] = random_normal((L, self.D, self.D), gen, self.rt.device)
            self.params.add("cross/W", W)
        else:
            w = torch.zeros((L, Di), device=self.rt.device)
            w[:, front_pad:] = random_normal((L, self.D), gen, self.rt.device)
            self.params.add("cross/w", w)
        self.params.add("cross/b", torch.zeros((L, Di), device=self.rt.device))
        if not self._external_params:
            self.params.finalize()
        return self

    @property
    def cross_weight(self) -> List[torch.Tensor]:
        p0 = self.front_pad
        if self.matrix:
            return [self.params["cross/W"][i, p0:, p0:] for i in range(self.layer_num)]
        return [self.params["cross/w"][i, p0:].unsqueeze(1) for i in range(self.layer_num)]

    @property
    def cross_bias(self) -> List[torch.Tensor]:
        return [self.params["cross/b"][i, self.front_pad:].unsqueeze(1) for i in range(self.layer_num)]

    def call(self, inputs, training: bool = False, out: Optional[torch.Tensor] = None):
        """inputs [B, front_pad + D] fp32 (unit column stride); ``out`` may be a
        strided [B, front_pad + D] view (e.g. the left part of the DCN concat)."""
        rt = self.rt
        x = rt.to_device(inputs, torch.float32)
        if self.D is None:
            self.build(x.shape[1])
        Di = self.D + self.front_pad
        assert x.dim() == 2 and x.shape[1] == Di and x.stride(1) == 1
        B = x.shape[0]
        if out is None:
            out = rt.empty((B, Di))
        check(rt.lib.etr_cross_vec_forward(rt.ctx, x.data_ptr(), x.stride(0), B, Di, self.layer_num,
                                           self.params["cross/w"].data_ptr(), self.params["cross/b"].data_ptr(),
                                           out.data_ptr(), out.stride(0), rt.stream))
        if training:
            self._ctx = {"x0": x}
        return out

    def backward(self, gout: torch.Tensor) -> torch.Tensor:
        """gout [B, Di] (unit column stride, any row stride) -> dx0 [B, Di]."""
        from .runtime import gemm_f32
        rt = self.rt
        x0 = self._ctx["x0"]
        B, Di, L = x0.shape[0], self.D + self.front_pad, self.layer_num
        dx0 = rt.empty((B, Di))
        scal = rt.empty((B, 2 * L))
        w, b = self.params["cross/w"], self.params["cross/b"]
        check(rt.lib.etr_cross_vec_backward(rt.ctx, x0.data_ptr(), x0.stride(0), B, Di, L, w.data_ptr(), b.data_ptr(),
                                            gout.data_ptr(), gout.stride(0), dx0.data_ptr(), Di, scal.data_ptr(),
                                            rt.stream))
        XtC = rt.empty((Di, 2 * L))
        gemm_f32(rt, x0, scal, XtC, Di, 2 * L, B, x0.stride(0), 2 * L, 2 * L, trans_a=True)
        sums, gsum = rt.empty((2 * L,)), rt.empty((Di,))
        check(rt.lib.etr_colsum_f32(rt.ctx, scal.data_ptr(), B, 2 * L, 2 * L, sums.data_ptr(), rt.stream))
        check(rt.lib.etr_colsum_f32(rt.ctx, gout.data_ptr(), B, Di, gout.stride(0), gsum.data_ptr(), rt.stream))
        check(rt.lib.etr_cross_vec_finish(rt.ctx, XtC.data_ptr(), sums.data_ptr(), gsum.data_ptr(), w.data_ptr(),
                                          b.data_ptr(), Di, L, self.params.g("cross/w").data_ptr(),
                                          self.params.g("cross/b").data_ptr(), rt.stream))
        return dx0


class MatrixCrossLayer(CrossLayer):
    """x_{l+1} = x0 (.) (W_l x_l + b_l) + x_l  (3.DCN/CustomLayers.py:272-305);
    ``tf.matmul(W, x[B,D,1])`` is y = W x, i.e. X_l W_l^T in batch form (KAT-4).
    This is the fp32 exact-parity path; the bf16 tcgen05 path is
    ``precision='bf16'`` (tensor cores, fp32 accumulate)."""

    matrix = True

    # ---- bf16 tensor-core path (tcgen05 + TMA, fp32 accumulate; parity 1e-2) ------------------
    def _call_bf16(self, x: torch.Tensor, training: bool, out: Optional[torch.Tensor]):
        from .runtime import cast_bf16
        rt = self.rt
        Di = self.D + self.front_pad
        assert x.dtype == torch.bfloat16 and x.shape[1] == Di and x.stride(1) == 1 and x.stride(0) % 8 == 0
        B = x.shape[0]
        W, b = self.params["cross/W"], self.params["cross/b"]
        xs, us = [x], []
        xl = x
        for l in range(self.layer_num):
            last = l == self.layer_num - 1
            nxt = out if (last and out is not None) else rt.empty((B, Di), torch.bfloat16)
            u = rt.empty((B, Di), torch.bfloat16) if training else None
            Wb = cast_bf16(rt, W[l])                      # [Di, Di] bf16: W itself is the [N,K] operand (y = W x)
            assert xl.stride(0) == x.stride(0)
            check(rt.lib.etr_cross_mat_layer_bf16(rt.ctx, x.data_ptr(), xl.data_ptr(), x.stride(0), B, Di,
                                                  Wb.data_ptr(), Wb.stride(0), b[l].data_ptr(), nxt.data_ptr(),
                                                  nxt.stride(0), _p(u), Di, rt.stream))
            xs.append(nxt)
            us.append(u)
            xl = nxt
        if training:
            self._ctx = {"xs": xs, "us": us, "bf16": True}
        return xl

    def _backward_bf16(self, gout: torch.Tensor, extra: Optional[torch.Tensor] = None) -> torch.Tensor:
        from .runtime import cast_bf16, gemm_bf16_wgrad
        rt = self.rt
        xs, us = self._ctx["xs"], self._ctx["us"]
        x0 = xs[0]
        B, Di = x0.shape[0], self.D + self.front_pad
        W = self.params["cross/W"]
        # G starts as the caller's gradient (possibly a column slice of a wider bf16 matrix: read through its row pitch)
        G = gout if (gout.dtype == torch.bfloat16 and gout.stride(1) == 1 and gout.stride(0) % 2 == 0) else \
            gout.to(torch.bfloat16).contiguous()
        dx0 = rt.empty((B, Di))
        assert x0.stride(0) == Di
        # one-pass form (round 2): du + db in one kernel per layer, dx0 = sum_l G_{l+1} (.) u_l + G_0 written once at the end
        # (needs 16-byte aligned rows everywhere; otherwise the per-layer read-modify-write form below)
        one_pass = (Di % 8 == 0 and G.stride(0) % 8 == 0 and G.data_ptr() % 16 == 0 and self.layer_num <= 8
                    and os.environ.get("ETR_CROSS_BWD_ONEPASS", "1") != "0")
        Gs = [None] * (self.layer_num + 1)
        Gs[self.layer_num] = G
        for l in reversed(range(self.layer_num)):
            du = rt.empty((B, Di), torch.bfloat16)
            if one_pass:
                check(rt.lib.etr_cross_mat_bwd_du_colsum_bf16(rt.ctx, G.data_ptr(), G.stride(0), x0.data_ptr(), B, Di, du.data_ptr(),
                                                              self.params.g("cross/b")[l].data_ptr(), rt.stream))
            else:
                check(rt.lib.etr_cross_mat_bwd_elementwise_bf16(rt.ctx, G.data_ptr(), G.stride(0), x0.data_ptr(), us[l].data_ptr(),
                                                                B, Di, du.data_ptr(), dx0.data_ptr(),
                                                                int(l == self.layer_num - 1), rt.stream))
            # dW_l = dU^T X_l : both operands as stored ([B, Di], MN-major UMMA operands), split-K over the batch; fp32 result
            gemm_bf16_wgrad(rt, du, xs[l], self.params.g("cross/W")[l], Di, Di, B)
            if not one_pass:
                check(rt.lib.etr_colsum_bf16(rt.ctx, du.data_ptr(), B, Di, Di, self.params.g("cross/b")[l].data_ptr(),
                                             rt.stream))
            # G_l = G_{l+1} + dU W_l : B operand [N=j, K=i] = W[i,j]  ->  W^T
            Wt = cast_bf16(rt, W[l], transpose=True)
            Gn = rt.empty((B, Di), torch.bfloat16)
            check(rt.lib.etr_gemm_bf16_tn_residual(rt.ctx, B, Di, Di, du.data_ptr(), Di, Wt.data_ptr(), Wt.stride(0),
                                                   G.data_ptr(), G.stride(0), Gn.data_ptr(), Di, rt.stream))
            G = Gn
            Gs[l] = G
        if one_pass:
            L_ = self.layer_num
            gp = (C.c_void_p * (L_ + 1))(*[g.data_ptr() for g in Gs])
            gl = (C.c_int64 * (L_ + 1))(*[g.stride(0) for g in Gs])
            up = (C.c_void_p * L_)(*[u.data_ptr() for u in us])
            ex_ok = (extra is not None and extra.dtype == torch.bfloat16 and extra.stride(1) == 1 and extra.stride(0) % 8 == 0
                     and extra.data_ptr() % 16 == 0)
            check(rt.lib.etr_cross_mat_bwd_dx0_bf16(rt.ctx, L_, gp, gl, up, extra.data_ptr() if ex_ok else None,
                                                    extra.stride(0) if ex_ok else 0, B, Di, dx0.data_ptr(), rt.stream))
            if ex_ok:
                extra = None
        else:
            check(rt.lib.etr_add_bf16_into_f32(rt.ctx, G.data_ptr(), B * Di, dx0.data_ptr(), rt.stream))
        if extra is not None:                         # the other branch's gradient for the same input, not folded above
            if extra.dtype == torch.bfloat16 and extra.is_contiguous():
                check(rt.lib.etr_add_bf16_into_f32(rt.ctx, extra.data_ptr(), B * Di, dx0.data_ptr(), rt.stream))
            else:
                dx0 += extra.float()
        return dx0

    def call(self, inputs, training: bool = False, out: Optional[torch.Tensor] = None):
        rt = self.rt
        if self.precision == "bf16":
            if self.D is None:
                assert inputs.shape[1] % 8 == 0, "bf16 cross layer: width must be a multiple of 8"
                self.build(inputs.shape[1])
            x = inputs if (isinstance(inputs, torch.Tensor) and inputs.dtype == torch.bfloat16 and
                           inputs.device == rt.device) else rt.to_device(inputs, torch.float32).to(torch.bfloat16)
            return self._call_bf16(x.contiguous() if x.stride(1) != 1 else x, training, out)
        x = rt.to_device(inputs, torch.float32)
        if self.D is None:
            self.build(x.shape[1])
        Di = self.D + self.front_pad
        assert x.dim() == 2 and x.shape[1] == Di and x.stride(1) == 1
        B = x.shape[0]
        W, b = self.params["cross/W"], self.params["cross/b"]
        xs, us = [x], []
        xl = x
        for l in range(self.layer_num):
            last = l == self.layer_num - 1
            nxt = out if (last and out is not None) else rt.empty((B, Di))
            u = rt.empty((B, Di))
            assert xl.stride(0) == x.stride(0)
            check(rt.lib.etr_cross_mat_layer_f32(rt.ctx, x.data_ptr(), xl.data_ptr(), xl.stride(0), B, Di,
                                                 W[l].data_ptr(), b[l].data_ptr(), nxt.data_ptr(), nxt.stride(0),
                                                 u.data_ptr(), Di, rt.stream))
            xs.append(nxt)
            us.append(u)
            xl = nxt
        if training:
            self._ctx = {"xs": xs, "us": us}
        return xl

    def backward(self, gout: torch.Tensor, extra: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``extra`` (bf16 tensor-core path only): a bf16 [B, Di] gradient another branch computed for the same input; it is
        added in the single pass that writes dx0."""
        from .runtime import gemm_f32
        if self._ctx.get("bf16"):
            return self._backward_bf16(gout, extra)
        assert extra is None
        rt = self.rt
        xs, us = self._ctx["xs"], self._ctx["us"]
        x0 = xs[0]
        B, Di = x0.shape[0], self.D + self.front_pad
        W = self.params["cross/W"]
        G = gout.contiguous().clone() if gout.stride(0) != Di else gout.clone()
        x0c = x0 if x0.stride(0) == Di else x0.contiguous()
        dx0 = rt.zeros((B, Di))
        du = rt.empty((B, Di))
        for l in reversed(range(self.layer_num)):
            # du = G (.) x0 ; dx0 += G (.) U_l
            check(rt.lib.etr_cross_mat_bwd_elementwise(rt.ctx, G.data_ptr(), x0c.data_ptr(), us[l].data_ptr(), B * Di,
                                                       du.data_ptr(), dx0.data_ptr(), rt.stream))
            xl = xs[l]
            # dW_l = du^T X_l   ([Di,B] x [B,Di])
            gemm_f32(rt, du, xl, self.params.g("cross/W")[l], Di, Di, B, Di, xl.stride(0), Di, trans_a=True)
            check(rt.lib.etr_colsum_f32(rt.ctx, du.data_ptr(), B, Di, Di, self.params.g("cross/b")[l].data_ptr(),
                                        rt.stream))
            # G_l = G_{l+1} + du W_l
            gemm_f32(rt, du, W[l], G, B, Di, Di, Di, Di, Di, beta=1.0)
        check(rt.lib.etr_add_sigmoid(rt.ctx, dx0.data_ptr(), G.data_ptr(), B * Di, dx0.data_ptr(), None, rt.stream))
        return dx0


# ---------------------------------------------------------------------------
class DeepCrossNetworkLayer(_Layer):
    """_input = [X_cont || Flatten(Embedding(X))] (continuous FIRST, :259) -> cross
    (vector or matrix by ``type``) || DenseLayer(units, activation) -> concat ->
    Dense(1, sigmoid)  (3.DCN/CustomLayers.py:206-269)."""

    def __init__(self, categorical_features=['uid', 'iid', 'utag1', 'utag2', 'utag3', 'utag4', 'itag1', 'itag2',
                                             'itag3', 'itag4'],
                 continuous_features=['itag4_origin', 'itag4_square', 'itag4_cube'], feature_dims=160000,
                 embedding_dims=16, units=[64, 8], activation='relu', layer_num=3, reg_w=1e-4, reg_b=1e-4,
                 type='vec', precision="fp32", **kwargs):
        self._setup(kwargs)
        # "bf16" (matrix type only): cross layers and the wide dense layers on the tensor cores
        assert precision in ("fp32", "bf16") and not (precision == "bf16" and type == 'vec')
        self.precision = precision
        self.categorical_features = list(categorical_features)
        self.continuous_features = list(continuous_features)
        self.feature_names = self.categorical_features
        self.feature_dims, self.embedding_dims = int(feature_dims), int(embedding_dims)
        self.units, self.type = list(units), type
        C_, F, k = len(self.continuous_features), len(self.categorical_features), self.embedding_dims
        self.D = C_ + F * k
        self.front_pad = (-C_) % (8 if precision == "bf16" else 4)
        if precision == "bf16":
            assert (F * k) % 8 == 0 and self.units[-1] % 8 == 0, "bf16 DCN: F*k and units[-1] must be multiples of 8"
        self.embedding_layer = self.table = self._make_table(self.feature_dims, k)
        self.table.init_uniform(-0.05, 0.05, self.gen)
        cls = CrossLayer if type == 'vec' else MatrixCrossLayer
        self.cross_layer = cls(layer_num, reg_w, reg_b, precision=precision, device=self.rt.device)
        self.cross_layer.build(self.D, self.params, self.gen, front_pad=self.front_pad)
        self.dense_layer = DenseLayer(self.units, activation, name="dense_layer", precision=precision)
        self.dense_layer.build(self.D, self.params, self.gen, front_pad=self.front_pad)
        self.output_layer = MLPLayer([1], 'sigmoid', name="output_layer", precision=precision)
        self.output_layer.build(self.D + self.units[-1], self.params, self.gen, front_pad=self.front_pad)
        self.params.finalize()

    @property
    def embeddings(self):
        return self.table.cols(0, self.embedding_dims)

    def sparse_tables(self):
        return [self.table]

    def call(self, inputs, training: bool = False):
        rt = self.rt
        ids = self._ids(inputs, self.categorical_features)
        C_, k = len(self.continuous_features), self.embedding_dims
        Di = self.front_pad + self.D
        dt = torch.bfloat16 if self.precision == "bf16" else torch.float32
        x = rt.empty((ids.B, Di), dt)                           # [pad | X_cont | Flatten(emb)]
        cont = self._cont(inputs, self.continuous_features) if C_ else None
        tab, vids, route = self._lookup(ids)
        gather_fm_forward(tab, k, False, vids, flat=x, flat_col0=self.front_pad + C_, cont=cont)
        comb = rt.empty((ids.B, Di + self.units[-1]), dt)       # concat fused: [cross_output | dnn_output]
        self.cross_layer.call(x, training=training, out=comb[:, :Di])
        dnn = self.dense_layer(x, training=training)
        comb[:, Di:] = dnn
        out = self.output_layer(comb, training=training)
        if training:
            self._ctx = {"ids": vids, "table": tab, "route": route}
        self._finish(training)
        return {"output": out}

    def backward(self, dlogit: torch.Tensor) -> List[SparseGrad]:
        rt = self.rt
        ids = self._ctx["ids"]
        Di = self.front_pad + self.D
        dz = dlogit.reshape(-1, 1).clone()
        dcomb = self.output_layer.backward(dz, dy_is_preact=True)           # [B, Di + units[-1]]
        ddnn = dcomb[:, Di:].float().contiguous()
        if self.cross_layer._ctx.get("bf16") and getattr(self.cross_layer, "matrix", False):
            # tensor-core path: the Dense branch's input gradient (bf16) first, folded into the ONE pass that writes dx0
            # (instead of a GEMM epilogue that reads and rewrites the fp32 dx0: 0.67 -> 0.15 ms at c3)
            dxd = self.dense_layer.backward(ddnn)
            dx = self.cross_layer.backward(dcomb[:, :Di], extra=dxd)       # [B, Di] fp32
        else:
            dx = self.cross_layer.backward(dcomb[:, :Di])                   # [B, Di] fp32
            self.dense_layer.backward(ddnn, accumulate_into=dx)
        col0 = self.front_pad + len(self.continuous_features)
        if self.shard is None and FlatSparseGrad.eligible(self._ctx["table"], ids, dx, col0, self.embedding_dims):
            # the rows' gradients ARE the embedding block of dx: the segment reduction reads them in place
            sg = FlatSparseGrad(self._ctx["table"], ids, dx, col0, self.embedding_dims)
            sg.plan = self._ctx.get("plan")
            return [sg]
        bag = gather_fm_backward(self._ctx["table"], self.embedding_dims, False, ids, dflat=dx, flat_col0=col0)
        return [self._table_grad(bag)]


# ---------------------------------------------------------------------------
class Trainer:
    """The reference train step (2.FM/ModelManager.py:171-181): BCE on the
    layer's 'output', backward, Adam apply.  ``apply_mode='rowwise'`` touches
    only the unique rows of the batch; ``'keras_dense'`` restates Keras 2.8
    ``Adam._resource_apply_sparse`` (every row of the table is decayed/updated).

    ``graph=True`` stages the inputs of every step into static device buffers
    (one async H2D copy per host column) and replays ONE captured CUDA graph of
    the whole step; the optimizer clock lives on the device for that reason."""

    def __init__(self, layer, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, apply_mode="rowwise", graph=False,
                 poll_every: int = 64, plan_ahead: Optional[bool] = None):
        """``poll_every``: every that many steps the device error word (out-of-range id, mailbox / touched-list
        overflow, peer-barrier timeout) is copied to pinned host memory WITHOUT stalling the loop and checked as soon
        as the copy has landed (at the latest ``poll_every`` steps later, and always in ``state_dict()``); 0 = never.
        The kernels clamp and carry on, so without this a bad id would train silently on the wrong row."""
        assert apply_mode in ("rowwise", "keras_dense")
        self.poll_every, self._steps_since_poll = int(poll_every), 0
        # 5.DIN/ModelManager.py:175-190: L2 on the rows the batch used, factor * tf.nn.l2_loss(gather(embed, unique(ids)));
        # set ``trainer.used_rows_l2 = factor`` (the table gradients are then materialised: no fused apply)
        self.used_rows_l2 = 0.0
        self.last_l2 = None
        self.layer, self.lr, self.b1, self.b2, self.eps = layer, lr, beta_1, beta_2, epsilon
        self.mode = _lib.ADAM_ROWWISE if apply_mode == "rowwise" else _lib.ADAM_KERAS_DENSE
        self.rt = layer.rt
        self.state = self.rt.zeros((2,))            # [iterations, lr_t] on the device
        self.last_grads: List[SparseGrad] = []
        self.use_graph = graph
        # data parallel over the ranks of a sharded layer: local batches, global-mean loss
        self.dp_world = layer.shard_spec[0] if getattr(layer, "shard_spec", None) else 1
        self.peer = getattr(layer, "peer", None)      # peer-memory sharding: nothing syncs with the host
        assert not (graph and self.dp_world > 1 and self.peer is None), \
            "the all-to-all sharded step syncs split sizes on the host: no CUDA graph (use shard='peer')"
        # static buffer sets: 2 = stage step i+1 while step i runs; 3 for the plan-ahead of peer-sharded layers (step i
        # already reads the ids of step i+1, so those are staged while step i-1 runs)
        # plan-ahead: ``train_step(batch, next_batch=staged)`` sorts the ids of the NEXT batch on the side stream while this
        # step runs, so that no step waits for its own sort (always on for peer-sharded layers, where the unique ids are the
        # first thing the exchange needs; opt-in for unsharded fused FM / DeepFM steps, where it takes the sort off the path
        # in front of the fused backward + apply)
        self.plan_ahead = (self.peer is not None) if plan_ahead is None else (bool(plan_ahead) or self.peer is not None)
        self.depth = (3 if self.plan_ahead else 2) if graph else 1
        if self.peer is not None:
            self.peer.n_req_sets = max(self.peer.n_req_sets, self.depth)      # one request-mailbox set per buffer set
        self._graphs: Dict[tuple, list] = {}
        self._copy_stream = self._d2h_stream = None

    @property
    def iterations(self) -> int:                     # synchronises
        return int(self.state[0].item())

    # ------------------------------------------------------- checkpoint / resume
    # The reference checkpoints the model and the optimizer with tf.train.Checkpoint(model=..., optimizer=...)
    # (2.FM/ModelManager.py:112-119, variable names in 2.FM/ranking_model/checkpoint/ckpt-2.index: bias, embed, w,
    # MLP_layer*/kernel_i, bias_i and the Adam slots m / v).  Here: one dict of host tensors, per rank for a
    # row-sharded layer (the table entries are then this rank's shard).
    def state_dict(self) -> Dict[str, torch.Tensor]:
        torch.cuda.synchronize(self.rt.device)
        self.rt.check_peeked_error(wait=True)
        self.rt.poll_error()                                    # never checkpoint a state trained on clamped ids
        P = self.layer.params
        sd = {"optimizer/state": self.state.cpu().clone(), "dense/value": P.value.cpu().clone(),
              "dense/m": P.m.cpu().clone(), "dense/v": P.v.cpu().clone()}
        for name in P.names():
            sd[f"var/{name}"] = P[name].cpu().clone()           # reference-shaped views, for inspection / export
        for i, t in enumerate(self.layer.sparse_tables()):
            t = getattr(t, "local", t)                          # a peer-sharded table checkpoints its own shard
            sd[f"table{i}/data"] = t.data.cpu().clone()
            sd[f"table{i}/m"] = t.m.cpu().clone()
            sd[f"table{i}/v"] = t.v.cpu().clone()
        return sd

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        P, dev = self.layer.params, self.rt.device
        assert sd["dense/value"].shape == P.value.shape, "checkpoint belongs to a different model configuration"
        self.state.copy_(sd["optimizer/state"].to(dev))
        P.value.copy_(sd["dense/value"].to(dev))
        P.m.copy_(sd["dense/m"].to(dev))
        P.v.copy_(sd["dense/v"].to(dev))
        for i, t in enumerate(self.layer.sparse_tables()):
            t = getattr(t, "local", t)
            assert sd[f"table{i}/data"].shape == t.data.shape, "table shape mismatch (rows / sharding changed?)"
            t.data.copy_(sd[f"table{i}/data"].to(dev))
            t.m.copy_(sd[f"table{i}/m"].to(dev))
            t.v.copy_(sd[f"table{i}/v"].to(dev))
        torch.cuda.synchronize(dev)

    def save(self, path: str) -> None:
        torch.save(self.state_dict(), path)

    def restore(self, path: str) -> None:
        self.load_state_dict(torch.load(path, map_location="cpu"))

    # ------------------------------------------------------------ eager step
    def train_step(self, inputs, labels=None, next_batch=None) -> torch.Tensor:
        """One step; returns the (device-resident, un-synchronised) scalar loss.

        ``next_batch`` (peer-sharded layers; a DeviceBatch from ``stage()`` that is already staged): the sorted plan of
        the NEXT batch is built on the side stream while this step runs, so that the next step can send its requests
        at once -- at N > 1 the plan is otherwise the first thing on the critical path (nothing of the table side can
        start before the unique ids are known)."""
        loss = self._graph_step(inputs, labels, next_batch) if self.use_graph else self._eager_step(inputs, labels, next_batch)
        self._poll()
        return loss

    def _poll(self) -> None:
        if self.poll_every <= 0:
            return
        self.rt.check_peeked_error()                            # non-blocking: only if the last peek has landed
        self._steps_since_poll += 1
        if self._steps_since_poll >= self.poll_every:
            self._steps_since_poll = 0
            self.rt.check_peeked_error(wait=True)               # the previous peek is poll_every steps old: done long ago
            self.rt.peek_error_async()

    def _eager_step(self, inputs, labels=None, next_batch=None) -> torch.Tensor:
        rt = self.rt
        if labels is None and isinstance(inputs, DeviceBatch):
            labels = inputs.labels
        sl = getattr(inputs, "_slot", None)
        if sl is not None and not torch.cuda.is_current_stream_capturing():
            torch.cuda.current_stream(rt.device).wait_event(sl.copy_done)     # staged on the copy stream
        y = rt.to_device(labels, torch.float32).reshape(-1)
        # plan-ahead (peer-sharded): this batch's plan was sorted during the previous step; the next batch's starts now
        pre, nsl = None, getattr(next_batch, "_slot", None)
        if self.plan_ahead and sl is not None and sl.plan_ready:
            pre, sl.plan_ready = sl.plan, False
        if self.plan_ahead and nsl is not None and nsl is not sl:
            rows = self.layer.sparse_tables()[0].rows
            if nsl.plan is None:
                assert not torch.cuda.is_current_stream_capturing()
                nsl.plan = SparsePlan(rt, nsl.batch.ids, rows, overlap=True)
                if getattr(self.layer, "embedding_dims", 0) == 16 and FusedFMGrad.apply_kernel == "tile":
                    nsl.plan.prepare_fm()             # row descriptors / long-run items of the tiled kernel
            else:
                nsl.plan.rebuild(nsl.batch.ids, rows)
            if self.peer is not None and getattr(self.layer, "shard_mode", None) == "peer":
                # ... and its requests go to the owners' mailbox set of that buffer set (the barrier that closes this
                # step orders them before the next step's serve)
                with torch.cuda.stream(rt.side_stream):
                    self.peer.request(nsl.plan, nsl.batch.ids.B, nsl.batch.ids.F, nsl.index, rt.side_stream.cuda_stream)
        fused = None
        if getattr(self.layer, "fused_train_ok", None) and self.layer.fused_train_ok():
            fused = self.layer.train_forward_backward(inputs, y, 1.0 / self.dp_world, plan=pre) if pre is not None else \
                self.layer.train_forward_backward(inputs, y, 1.0 / self.dp_world)
        if fused is not None:
            loss, _, grads = fused
        else:
            out = self.layer(inputs, training=True)["output"]
            loss, dlogit = bce_forward_backward(rt, out.reshape(-1), y)
            if self.dp_world > 1:
                dlogit.mul_(1.0 / self.dp_world)      # the loss is the mean over the GLOBAL batch
            grads = self.layer.backward(dlogit)
        if self.peer is not None:
            # rows -> owners' mailboxes, dense grads -> every peer's slot (on the dense stream, under the row push);
            # ONE device-side barrier; then the owner's row apply with the dense sum + dense Adam beside it
            cur, ds = torch.cuda.current_stream(rt.device), self._dense_stream_get()
            ds.wait_stream(cur)
            with torch.cuda.stream(ds):
                self.peer.allreduce_push(self.layer.params.grad)
            for g in grads:
                g.push()
            cur.wait_stream(ds)
            self.peer.barrier()
            self._apply_peer(grads, cur, ds)
        else:
            if self.dp_world > 1:
                import torch.distributed as dist
                dist.all_reduce(self.layer.params.grad)   # replicated dense variables: sum of the ranks' grads
            self.apply_gradients(grads)
        if self.plan_ahead and nsl is not None and nsl is not sl:
            nsl.plan.join()                           # the look-ahead sort + requests belong to THIS step (timed with it)
            nsl.plan._pending = None
            nsl.plan_ready = True
        if self.peer is not None:
            self.peer.barrier()                       # owners have applied: shards and mailboxes are free again
        if sl is not None and not torch.cuda.is_current_stream_capturing():
            sl.used = True
            sl.compute_done.record(torch.cuda.current_stream(rt.device))
        return loss

    def _dense_stream_get(self) -> "torch.cuda.Stream":
        if getattr(self, "_dense_stream", None) is None:
            self._dense_stream = torch.cuda.Stream(device=self.rt.device)
        return self._dense_stream

    def _apply_peer(self, grads, cur, ds) -> None:
        """after the gradient barrier of a peer-sharded step: the owner's row apply on the main stream, the rank-ordered
        sum of the dense-gradient slots + dense Adam beside it on the dense stream (they share only the step's lr_t)"""
        rt = self.rt
        assert not self.used_rows_l2, "used-rows L2 is not wired into the peer-sharded apply"
        check(rt.lib.etr_adam_step_begin(rt.ctx, self.state.data_ptr(), self.lr, self.b1, self.b2, rt.stream))
        d_lr = self.state[1:]
        ds.wait_stream(cur)
        with torch.cuda.stream(ds):
            self.peer.allreduce_sum(self.layer.params.grad)
            self.layer.params.adam_step(0.0, d_lr, self.b1, self.b2, self.eps)
        for g in grads:
            g.apply(d_lr, self.b1, self.b2, self.eps, self.mode)
        cur.wait_stream(ds)
        self.last_grads = grads

    def apply_gradients(self, grads: List[SparseGrad]) -> None:
        rt = self.rt
        check(rt.lib.etr_adam_step_begin(rt.ctx, self.state.data_ptr(), self.lr, self.b1, self.b2, rt.stream))
        d_lr = self.state[1:]
        self.layer.params.adam_step(0.0, d_lr, self.b1, self.b2, self.eps)
        plans: Dict[tuple, SparsePlan] = {}
        if self.used_rows_l2:
            self.last_l2 = rt.zeros((1,))
        for g in grads:
            if isinstance(g, FusedFMGrad) and self.mode == _lib.ADAM_ROWWISE and not self.used_rows_l2:
                g.apply(d_lr, self.b1, self.b2, self.eps)
                continue
            if hasattr(g, "push"):                    # PeerFMGrad: owner-side mailbox reduce + Adam
                g.apply(d_lr, self.b1, self.b2, self.eps, self.mode)
                continue
            key = (id(g.ids), g.table.rows)
            g.reduce(plans.get(key))
            plans[key] = g.plan
            t = g.table.desc()
            if self.used_rows_l2:
                check(rt.lib.etr_used_rows_l2(rt.ctx, C.byref(t), g.plan.unique_ids.data_ptr(), g.plan.counts.data_ptr(),
                                              g.plan.n_slots, float(self.used_rows_l2), g.unique_grad.data_ptr(),
                                              g.unique_grad.shape[1], self.last_l2.data_ptr(), rt.stream))
            check(rt.lib.etr_sparse_adam_apply(rt.ctx, C.byref(t), g.table.m.data_ptr(), g.table.v.data_ptr(),
                                               g.plan.unique_ids.data_ptr(), g.plan.counts.data_ptr(),
                                               g.plan.n_slots, g.unique_grad.data_ptr(), g.unique_grad.shape[1],
                                               0.0, d_lr.data_ptr(), self.b1, self.b2, self.eps, self.mode,
                                               rt.stream))
        self.last_grads = grads

    # ----------------------------------------------------------- graph step
    # Pipelined: ``depth`` sets of static input buffers, each with its own captured
    # graph.  Host columns are copied on a dedicated copy stream (H2D of step i+1
    # overlaps the compute of step i); the loss is read back on a third stream into
    # pinned memory so that reading step i's loss never waits for step i+1.
    class _Slot:
        def __init__(self):
            self.batch = self.ids_buf = self.cont_buf = self.lab_buf = None
            self.loss = None
            self.graphs, self.eager = {}, {}          # per step variant (plan ready?, look-ahead?): captured graph, eager runs
            self.plan, self.plan_ready = None, False  # plan-ahead (peer-sharded layers): this slot's sorted plan, static buffers
            self.copy_done = torch.cuda.Event()
            self.compute_done = torch.cuda.Event()
            self.used = False
            self.host_loss = torch.zeros(1, dtype=torch.float32).pin_memory()
            self.loss_ready = torch.cuda.Event()

    class LossHandle:
        """Result of ``train_step_async``: ``result()`` waits for that step's loss only."""

        def __init__(self, slot):
            self._slot, self._value = slot, None

        def result(self) -> float:
            if self._value is None:
                self._slot.loss_ready.synchronize()
                self._value = float(self._slot.host_loss[0])
            return self._value

    def _slots_for(self, B: int):
        lay, rt = self.layer, self.rt
        key = (B,)
        if key not in self._graphs:
            names = getattr(lay, "feature_names", None) or lay.categorical_features
            cont_names = list(getattr(lay, "continuous_features", []))
            slots = []
            for q in range(self.depth):
                sl = Trainer._Slot()
                sl.index = q
                sl.ids_buf = rt.empty((len(names), B), torch.int64)
                sl.cont_buf = rt.empty((max(len(cont_names), 1), B), torch.float32)
                sl.lab_buf = rt.empty((B,), torch.float32)
                ids = IdsBatch(rt, sl.ids_buf, B, len(names), 1, 1, B, 1, lay.pad_id, lay.pooling)
                sl.batch = DeviceBatch(ids, sl.cont_buf[: len(cont_names)].t() if cont_names else None, sl.lab_buf)
                sl.batch._slot = sl
                slots.append(sl)
            self._graphs[key] = [slots, 0]
        return self._graphs[key]

    def stage(self, inputs, labels) -> DeviceBatch:
        """Copy one step's host (or device) inputs into the next set of static buffers
        (asynchronously, on the copy stream)."""
        lay, rt = self.layer, self.rt
        names = getattr(lay, "feature_names", None) or lay.categorical_features
        cont_names = list(getattr(lay, "continuous_features", []))
        B = int(inputs[names[0]].shape[0])
        entry = self._slots_for(B)
        slots, nxt = entry
        sl = slots[nxt % self.depth]
        entry[1] = nxt + 1
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=rt.device)
            self._d2h_stream = torch.cuda.Stream(device=rt.device)
        cur = torch.cuda.current_stream(rt.device)
        cols = [inputs[n] for n in names] + [inputs[n] for n in cont_names] + [labels]
        if any(isinstance(c, torch.Tensor) and c.is_cuda for c in cols):
            self._copy_stream.wait_stream(cur)            # device-resident inputs: order after their producer
        sl.plan_ready = False                             # the ids are about to change: a plan sorted for the old ones is stale
        with torch.cuda.stream(self._copy_stream):
            if sl.used:
                self._copy_stream.wait_event(sl.compute_done)      # the buffers' previous consumer has finished
            self._copy_columns(sl.ids_buf, [inputs[n] for n in names], torch.int64)
            if cont_names:
                self._copy_columns(sl.cont_buf, [inputs[n] for n in cont_names], torch.float32)
            sl.lab_buf.copy_(torch.as_tensor(labels).reshape(-1), non_blocking=True)
            sl.copy_done.record(self._copy_stream)
        return sl.batch

    @staticmethod
    def _copy_columns(dst: torch.Tensor, cols, dtype: torch.dtype) -> None:
        """dst [n, B] <- n host/device columns.  Columns that sit back to back in one allocation (a
        column-major block, e.g. a pinned staging arena or the columns of a contiguous [n, B] array) go in
        ONE async copy per run instead of one per column: 40 cudaMemcpyAsync per step were host-bound."""
        ts = [torch.as_tensor(c).reshape(-1) for c in cols]
        i, n = 0, len(ts)
        while i < n:
            j = i + 1
            t0 = ts[i]
            if t0.dtype == dtype and t0.is_contiguous():
                nbytes = t0.numel() * t0.element_size()
                while (j < n and ts[j].dtype == dtype and ts[j].device == t0.device and ts[j].is_contiguous() and
                       ts[j].numel() == t0.numel() and ts[j].data_ptr() == t0.data_ptr() + (j - i) * nbytes and
                       ts[j].untyped_storage().data_ptr() == t0.untyped_storage().data_ptr()):
                    j += 1
            if j - i > 1:
                block = torch.as_strided(t0, (j - i, t0.numel()), (t0.numel(), 1))
                dst[i:j].copy_(block, non_blocking=True)
            else:
                dst[i].copy_(t0, non_blocking=True)
            i = j

    def _graph_step(self, inputs, labels, next_batch=None) -> torch.Tensor:
        """Per buffer set and step variant: calls 1-2 run eagerly on the static buffers (real training
        steps; they size the workspace), call 3 captures the step into a CUDA graph,
        every later call is one graph replay."""
        batch = inputs if isinstance(inputs, DeviceBatch) else self.stage(inputs, labels)
        sl = getattr(batch, "_slot", None)
        assert sl is not None, "graph mode needs the static DeviceBatch returned by stage()"
        cur = torch.cuda.current_stream(self.rt.device)
        cur.wait_event(sl.copy_done)
        nsl = getattr(next_batch, "_slot", None) if self.plan_ahead else None
        if nsl is sl:
            nsl = next_batch = None
        if nsl is not None:
            cur.wait_event(nsl.copy_done)     # its ids are sorted inside this step
        else:
            next_batch = None
        key = (self.plan_ahead and sl.plan_ready, nsl is not None)
        if key not in sl.graphs and sl.eager.get(key, 0) < 2:
            sl.eager[key] = sl.eager.get(key, 0) + 1
            loss = self._eager_step(batch, None, next_batch)
            if sl.loss is None:
                sl.loss = self.rt.empty((1,))
            sl.loss.copy_(loss)
        else:
            if key not in sl.graphs:
                torch.cuda.synchronize(self.rt.device)
                ready = (sl.plan_ready, nsl.plan_ready if nsl is not None else False)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    loss = self._eager_step(batch, None, next_batch)
                    sl.loss.copy_(loss)
                sl.graphs[key] = g            # capture does not execute: the replay below runs the step
                sl.plan_ready = ready[0]      # ... so the flags the captured step flipped are restored first
                if nsl is not None:
                    nsl.plan_ready = ready[1]
                cur = torch.cuda.current_stream(self.rt.device)
            sl.graphs[key].replay()
            if self.plan_ahead:               # what the replayed step did to the plan flags
                sl.plan_ready = False
                if nsl is not None:
                    nsl.plan_ready = True
        sl.used = True
        sl.compute_done.record(cur)
        return sl.loss

    def train_step_async(self, inputs, labels=None, next_batch=None) -> "Trainer.LossHandle":
        """Graph mode: enqueue the step and an asynchronous read-back of its loss;
        returns immediately.  ``handle.result()`` blocks only until THIS step is done."""
        assert self.use_graph, "train_step_async needs graph=True"
        loss = self._graph_step(inputs, labels, next_batch)
        self._poll()
        batch_slot = None
        for slots, _ in self._graphs.values():
            for sl in slots:
                if sl.loss is loss:
                    batch_slot = sl
        with torch.cuda.stream(self._d2h_stream):
            self._d2h_stream.wait_event(batch_slot.compute_done)
            batch_slot.host_loss.copy_(batch_slot.loss, non_blocking=True)
            batch_slot.loss_ready.record(self._d2h_stream)
        return Trainer.LossHandle(batch_slot)
