"""Keras binding of the B200 layers -- the reference-side adapter.

With TensorFlow installed, a reference ``ModelManager`` switches to the B200 hot path by changing two imports:

    from etr_b200.tf_adapter import (FMRankingLayer, DeepFMRankingLayer, FFMRankingLayer, FFMLayer, FwFMLayer,
                                     PNNRankingLayer, PNNLayer, DeepCrossNetworkLayer)      # instead of CustomLayers
    from etr_b200.tf_adapter import Adam                                  # instead of tf.keras.optimizers.Adam

and everything else runs as written: ``tf.keras.Model(input_dic, layer(input_dic))`` over symbolic
``tf.keras.Input`` columns (2.FM/ModelManager.py:87-96), ``model(inputs, training=True)`` under a ``GradientTape``,
``tape.gradient(loss, model.trainable_variables)`` and ``opt.apply_gradients(zip(grads, vars))`` (:172-179).

How: every class below is a real ``tf.keras.layers.Layer``.  ``call`` wraps the torch-side layer shim
(``CustomLayers.py``) in ``tf.py_function`` (so it also works on the symbolic inputs of the functional API) under
``tf.custom_gradient``; tensors cross with DLPack in both directions (zero copy when TF and torch share the GPU).  The
embedding tables and their Adam slots stay HBM-resident on the B200 side -- they are NOT tf.Variables -- so each layer
exposes ONE tiny trainable variable, ``etr_handle``: its "gradient" is how ``tape.gradient`` reaches the layer (the
backward kernels run there and leave the dense gradients + the IndexedSlices-form table gradients on the layer), and
``Adam.apply_gradients`` recognises handles and runs the fused dense + sparse Adam kernels for them; any other
variable in the list goes to a stock ``tf.keras.optimizers.Adam``.

TensorFlow is not installable in the build image (no wheel, no network), so this file is import-guarded and exercised
only for its guard (tests/test_tf_adapter.py); the torch-side twin with the same structure
(``autograd.py``: EtrModule / EtrAdam) is what the GPU tests run.
"""
from __future__ import annotations

from typing import Dict, List, Optional

try:                                                  # pragma: no cover - TensorFlow is absent from the build image
    import tensorflow as tf
except Exception:                                     # ImportError, or a broken install
    tf = None

_HANDLES: Dict[object, "object"] = {}                 # etr_handle.ref() -> Keras shim that owns it


def require_tf():
    if tf is None:
        raise ImportError("etr_b200.tf_adapter needs TensorFlow (>= 2.8, the version the reference pins in "
                          "*/output/saved_model.pb); use etr_b200.CustomLayers / etr_b200.autograd from torch instead")
    return tf


def _to_torch(t):
    import torch
    return torch.utils.dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(t))


def _to_tf(t):
    import torch
    return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t.contiguous()))


def _make(impl_name: str, input_attrs=("feature_names",)):
    """Keras Layer class for the layer shim ``CustomLayers.<impl_name>`` (same ctor arguments)."""
    require_tf()

    class Shim(tf.keras.layers.Layer):
        def __init__(self, *args, **kwargs):
            keras_kw = {k: kwargs.pop(k) for k in ("name", "dtype", "trainable", "dynamic") if k in kwargs}
            super().__init__(**keras_kw)
            from . import CustomLayers
            self.impl = getattr(CustomLayers, impl_name)(*args, **kwargs)
            self._names: List[str] = []
            for a in input_attrs:
                self._names += list(getattr(self.impl, a, []))
            self.sparse_grads = []
            self._trainer = None

        def build(self, input_shape):
            self.etr_handle = self.add_weight(name="etr_handle", shape=(1,), initializer="zeros", trainable=True)
            _HANDLES[self.etr_handle.ref()] = self
            self.built = True

        # --- eager bodies run by tf.py_function -------------------------------------------------
        def _forward(self, training, *cols):
            inputs = {n: _to_torch(c) for n, c in zip(self._names, cols)}
            out = self.impl(inputs, training=bool(training))["output"]
            return _to_tf(out)

        def _backward(self, dy, out):
            dyt, outt = _to_torch(dy), _to_torch(out)
            dz = (dyt * outt * (1.0 - outt)).reshape(-1).contiguous()       # every model layer ends in a sigmoid
            self.impl.params.grad.zero_()
            self.sparse_grads = self.impl.backward(dz)
            return tf.zeros([1], tf.float32)                                  # the handle carries no value

        def call(self, inputs, training=None):
            cols = [inputs[n] for n in self._names]
            train_flag = tf.constant(bool(training))

            @tf.custom_gradient
            def op(handle, *xs):
                out = tf.py_function(self._forward, [train_flag] + list(xs), Tout=tf.float32)
                out.set_shape([None, 1])

                def grad(dy):
                    g = tf.py_function(self._backward, [dy, out], Tout=tf.float32)
                    g.set_shape([1])
                    return [g] + [None] * len(xs)
                return out, grad

            return {"output": op(self.etr_handle, *cols)}

        # reference-named variables (2.FM/CustomLayers.py:123-135 ...), as DLPack-imported tf tensors
        def etr_variables(self) -> Dict[str, object]:
            return {n: _to_tf(self.impl.params[n]) for n in self.impl.params.names()}

    Shim.__name__ = Shim.__qualname__ = impl_name
    return Shim


class Adam:
    """tf.keras.optimizers.Adam drop-in (same ctor keywords, ``iterations``, ``apply_gradients``) that runs the B200
    dense + sparse Adam kernels for etr layers and a stock Keras Adam for everything else.
    ``apply_mode='keras_dense'`` is Keras 2.8's exact sparse semantics (all rows decay every step -- what the
    reference's checkpoints show); ``'rowwise'`` updates only the rows a batch touched."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, apply_mode="keras_dense", **kw):
        require_tf()
        self._cfg = dict(lr=float(learning_rate), beta_1=beta_1, beta_2=beta_2, epsilon=epsilon, apply_mode=apply_mode)
        self._stock = tf.keras.optimizers.Adam(learning_rate=learning_rate, beta_1=beta_1, beta_2=beta_2, epsilon=epsilon, **kw)

    @property
    def iterations(self):
        return self._stock.iterations

    def apply_gradients(self, grads_and_vars, **kw):
        from .CustomLayers import Trainer
        rest = []
        for g, v in grads_and_vars:
            shim = _HANDLES.get(v.ref())
            if shim is None:
                rest.append((g, v))
                continue
            if shim._trainer is None:
                shim._trainer = Trainer(shim.impl, **self._cfg)
            shim._trainer.apply_gradients(shim.sparse_grads)
        if rest:
            return self._stock.apply_gradients(rest, **kw)
        self._stock.iterations.assign_add(1)
        return self._stock.iterations


_MODEL_LAYERS = {
    "FMRankingLayer": ("feature_names",), "DeepFMRankingLayer": ("feature_names", "continuous_features"),
    "FFMRankingLayer": ("feature_names",), "FFMLayer": ("feature_names",), "FwFMLayer": ("feature_names",),
    "PNNRankingLayer": ("feature_names",), "PNNLayer": ("feature_names",),
    "DeepCrossNetworkLayer": ("categorical_features", "continuous_features"),
}


def __getattr__(name):                                  # classes are built on first use, so importing never needs TF
    if name in _MODEL_LAYERS:
        cls = _make(name, _MODEL_LAYERS[name])
        globals()[name] = cls
        return cls
    raise AttributeError(name)
