"""torch.autograd hook over the layer shims ("a custom-gradient hook -- torch autograd.Function
here, tf.custom_gradient when TF exists").

The reference train step is ``with tf.GradientTape(): out = model(x); loss = bce(y, out)`` followed by
``tape.gradient`` and ``opt.apply_gradients`` (2.FM/ModelManager.py:172-179).  The same loop in torch, with the
B200 kernels doing every forward and backward, reads

    model = EtrModule(DeepFMRankingLayer(...))                  # an nn.Module around any layer shim
    opt = EtrAdam(model, lr=1e-3)                               # Keras-Adam semantics, dense + IndexedSlices
    out = model(inputs)["output"]                               # K1 + tower kernels; autograd-aware
    loss = torch.nn.functional.binary_cross_entropy(out, y)     # any torch loss on the probabilities
    loss.backward()                                             # -> layer.backward(dL/dz): our CUDA backward
    opt.step()                                                  # dense Adam + sparse (row-wise | keras_dense) Adam

``model.dense.grad`` holds the flat gradient of every small replicated variable (views: ``model.grad_of(name)``);
the table gradients stay in IndexedSlices form (``model.sparse_grads`` -- ``.indexed_slices()`` exports unique ids +
summed rows), which is what the reference's optimizer receives for its Embedding variables.

The fused ``Trainer`` step remains the fast path (one CUDA graph, fused loss); this module is the general one: any
torch loss, any surrounding torch graph, gradient checks against the oracle in tests/test_gpu_autograd.py.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import check
from .runtime import FusedFMGrad, SparseGrad, SparsePlan


class _ModelFunction(torch.autograd.Function):
    """layer(inputs) -> probabilities [B,1]; backward feeds dL/dz (z = pre-sigmoid logit) to ``layer.backward``."""

    @staticmethod
    def forward(ctx, module: "EtrModule", dense: torch.Tensor, inputs):
        layer = module.layer
        out = layer(inputs, training=True)["output"]
        ctx.module = module
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        module = ctx.module
        layer = module.layer
        (out,) = ctx.saved_tensors
        # every model layer of the path ends in a sigmoid (2.FM/CustomLayers.py:155, 305, 496, 531, 598;
        # 3.DCN/CustomLayers.py:267): dL/dz = dL/dp * p (1 - p)
        dz = (grad_out.to(torch.float32) * out * (1.0 - out)).reshape(-1).contiguous()
        layer.params.grad.zero_()
        module.sparse_grads = layer.backward(dz)
        return None, layer.params.grad.clone(), None


class _OpFunction(torch.autograd.Function):
    """Interaction-only layers whose ``call`` takes a dense tensor and whose ``backward`` takes the output
    gradient: InnerProductNetwork / OuterProductNetwork ([B,F,k] -> [B,P]) and CrossLayer / MatrixCrossLayer
    ([B,D] -> [B,D])."""

    @staticmethod
    def forward(ctx, module: "EtrModule", dense: torch.Tensor, x: torch.Tensor):
        layer = module.layer
        ctx.module = module
        return layer.call(x, training=True)

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        module = ctx.module
        layer = module.layer
        if layer.params.grad is not None:
            layer.params.grad.zero_()
        dx = layer.backward(grad_out.to(torch.float32).contiguous())
        module.kernel_grad = getattr(layer, "kernel_grad", None)
        dense_grad = layer.params.grad.clone() if layer.params.grad is not None else None
        return None, dense_grad, dx


class _PairVecFunction(torch.autograd.Function):
    """FieldAwareInteractionLayer: ids [B,F] -> pair vectors [B,P,k]; the table gradient is sparse."""

    @staticmethod
    def forward(ctx, module: "EtrModule", dense: torch.Tensor, X):
        ctx.module = module
        return module.layer.call(X, training=True)

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        module = ctx.module
        module.sparse_grads = module.layer.backward(grad_out.to(torch.float32).contiguous())
        return None, None, None


class EtrModule(torch.nn.Module):
    """nn.Module around a layer shim.  ``self.dense`` is the layer's flat dense-parameter buffer exposed as an
    ``nn.Parameter`` that SHARES storage with ``layer.params.value`` (so the kernels and torch see one copy);
    embedding tables stay HBM-resident buffers of the layer (``layer.sparse_tables()``)."""

    def __init__(self, layer):
        super().__init__()
        self.layer = layer
        self.dense = torch.nn.Parameter(layer.params.value, requires_grad=True)
        self.sparse_grads: List[SparseGrad] = []
        self.kernel_grad: Optional[torch.Tensor] = None
        kind = type(layer).__name__
        if kind == "FieldAwareInteractionLayer":
            self._fn = _PairVecFunction
        elif hasattr(layer, "feature_names") or hasattr(layer, "categorical_features"):
            self._fn = _ModelFunction
        else:
            self._fn = _OpFunction

    def forward(self, inputs):
        out = self._fn.apply(self, self.dense, inputs)
        return {"output": out} if self._fn is _ModelFunction else out

    def grad_of(self, name: str) -> torch.Tensor:
        """reference-shaped view of the gradient of one dense variable (after ``backward()``)"""
        P = self.layer.params
        o, shape = P._views[name]
        n = 1
        for s_ in shape:
            n *= s_
        return self.dense.grad[o:o + n].view(shape)[P._pad[name]:]

    def variables(self) -> Dict[str, torch.Tensor]:
        return {n: self.layer.params[n] for n in self.layer.params.names()}


class EtrAdam:
    """tf.keras.optimizers.Adam(learning_rate, beta_1, beta_2, epsilon) for an ``EtrModule``: dense variables by one
    fused launch, tables by the sorted-ID segment reduction + sparse Adam (``apply_mode`` 'rowwise' | 'keras_dense',
    the latter restating Keras 2.8 ``_resource_apply_sparse``).  2.FM/ModelManager.py:103-104,178."""

    def __init__(self, module: EtrModule, learning_rate: float = 1e-3, beta_1: float = 0.9, beta_2: float = 0.999,
                 epsilon: float = 1e-7, apply_mode: str = "rowwise", lr: Optional[float] = None):
        from .CustomLayers import Trainer
        self.module = module
        self._tr = Trainer(module.layer, lr=learning_rate if lr is None else lr, beta_1=beta_1, beta_2=beta_2,
                           epsilon=epsilon, apply_mode=apply_mode)

    @property
    def iterations(self) -> int:
        return self._tr.iterations

    def zero_grad(self) -> None:
        if self.module.dense.grad is not None:
            self.module.dense.grad = None
        self.module.sparse_grads = []

    def step(self) -> None:
        m = self.module
        if m.dense.grad is not None:
            m.layer.params.grad.copy_(m.dense.grad)        # whatever torch accumulated (e.g. an extra regulariser)
        else:
            m.layer.params.grad.zero_()
        self._tr.apply_gradients(m.sparse_grads)

    def apply_gradients(self, grads_and_vars=None) -> None:
        """Keras spelling of ``step`` (the gradients live on the module)."""
        self.step()
