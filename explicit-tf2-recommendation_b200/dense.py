"""Dense (replicated) parameters and the MLP tower -- host side of K7.

``MLPLayer`` mirrors the reference class of the same name
(2.FM/CustomLayers.py:15-84; copy in 3.DCN/CustomLayers.py:20-90): same ctor
arguments, ``kernels`` / ``biases`` lists named ``kernel_i`` / ``bias_i``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import check
from .runtime import Runtime, cast_bf16, gemm_bf16_tn, gemm_bf16_tn_accumulate, gemm_bf16_wgrad, gemm_f32, _p


class DenseParams:
    """All small replicated variables of one model in ONE flat fp32 buffer (with
    matching flat grad / Adam m / v buffers) so the dense Adam apply is a single
    launch.  Variables are views into the flat buffers."""

    def __init__(self, rt: Runtime):
        self.rt = rt
        self._specs: List[Tuple[str, Tuple[int, ...], torch.Tensor]] = []
        self.value = self.grad = self.m = self.v = None
        self._views: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        self._pad: Dict[str, int] = {}

    def add(self, name: str, init: torch.Tensor, front_pad_rows: int = 0) -> str:
        """``front_pad_rows`` zero rows are stored in front of a 2-D variable (the
        operand it multiplies has that many zero padding columns in front so that
        its embedding part starts 16-byte aligned); ``self[name]`` is the
        reference-shaped view behind them, ``self.full(name)`` the padded one."""
        assert self.value is None, "DenseParams already finalized"
        assert name not in [s[0] for s in self._specs], name
        init = init.detach().to(torch.float32)
        if front_pad_rows:
            init = torch.cat([torch.zeros((front_pad_rows,) + tuple(init.shape[1:]), device=init.device), init], 0)
        self._pad[name] = int(front_pad_rows)
        self._specs.append((name, tuple(init.shape), init.reshape(-1)))
        return name

    def finalize(self):
        # every variable starts on a 16-byte boundary
        off = 0
        for name, shape, _ in self._specs:
            self._views[name] = (off, shape)
            off += (math.prod(shape) + 3) // 4 * 4
        n = max(off, 4)
        self.value = self.rt.zeros((n,))
        self.grad = self.rt.zeros((n,))
        self.m = self.rt.zeros((n,))
        self.v = self.rt.zeros((n,))
        for name, shape, init in self._specs:
            o, _ = self._views[name]
            self.value[o:o + init.numel()] = init.to(self.rt.device)
        return self

    def full(self, name: str) -> torch.Tensor:
        o, shape = self._views[name]
        return self.value[o:o + math.prod(shape)].view(shape)

    def gfull(self, name: str) -> torch.Tensor:
        o, shape = self._views[name]
        return self.grad[o:o + math.prod(shape)].view(shape)

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.full(name)[self._pad[name]:]

    def g(self, name: str) -> torch.Tensor:
        return self.gfull(name)[self._pad[name]:]

    def names(self) -> List[str]:
        return [s[0] for s in self._specs]

    def set(self, name: str, value) -> None:
        self[name].copy_(torch.as_tensor(value, dtype=torch.float32).reshape(self[name].shape))

    def adam_step(self, lr_t: float, d_lr_t: Optional[torch.Tensor], b1: float, b2: float, eps: float) -> None:
        rt = self.rt
        check(rt.lib.etr_dense_adam_apply(rt.ctx, self.value.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                          self.grad.data_ptr(), self.value.numel(), lr_t, _p(d_lr_t), b1, b2, eps,
                                          rt.stream))


# ------------------------------------------------------------ initializers
def glorot_uniform(shape, gen: torch.Generator, device) -> torch.Tensor:
    """Keras glorot_uniform: limit = sqrt(6/(fan_in+fan_out)); for rank>2 the
    leading dims are the receptive field (Keras ``_compute_fans``)."""
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    elif len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rec = math.prod(shape[:-2])
        fan_in, fan_out = shape[-2] * rec, shape[-1] * rec
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return torch.empty(shape, device=device).uniform_(-lim, lim, generator=gen)


def random_normal(shape, gen: torch.Generator, device, std=0.05) -> torch.Tensor:
    return torch.empty(shape, device=device).normal_(0.0, std, generator=gen)


_INIT = {
    "glorot_uniform": glorot_uniform,
    "random_normal": random_normal,
    "zeros": lambda shape, gen, device: torch.zeros(shape, device=device),
}


class MLPLayer:
    """per layer: act(x @ kernel_i + bias_i) -- MatMul, BiasAdd, activation
    (2.FM/CustomLayers.py:72-84).  Batch norm is not on the hot path (never
    enabled by any hot-path caller) and raises; dropout is keyed on a private
    ``is_train`` flag Keras never passes (:72,82) so it is never active in the
    reference either and is ignored here."""

    # layers whose input is at least this wide run on the tensor cores when precision == "bf16"
    TC_MIN_IN = 129      # narrower layers take the exact fp32 thread-per-row kernels (K <= 128)

    def __init__(self, units, activation=None, use_bias=True, is_batch_norm=False, is_dropput=0,
                 kernel_initializer="glorot_uniform", bias_initializer="zeros", name="mlp", precision="fp32",
                 **kwargs):
        self.units = [units] if not isinstance(units, list) else list(units)
        if len(self.units) <= 0:
            raise ValueError(f"Received an invalid value for `units`, expected a positive integer, got {units}.")
        if is_batch_norm:
            raise NotImplementedError("is_batch_norm=True is outside the B200 hot path")
        if activation not in _lib.ACT:
            raise ValueError(f"unsupported activation {activation!r}")
        self.activation, self.use_bias, self.is_dropout = activation, use_bias, is_dropput
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer
        self.name = name
        assert precision in ("fp32", "bf16")
        self.precision = precision      # "bf16": wide layers use tcgen05 (bf16 operands, fp32 accumulate)
        self.fused_skinny = bool(kwargs.pop("fused_skinny", True))   # K7b one-pass backward of a [wide -> 32] layer
        self.params: Optional[DenseParams] = None
        self.in_dim: Optional[int] = None
        self._saved = None
        self._owns_params = False
        self.front_pad = 0

    # -- build ---------------------------------------------------------------
    def build(self, in_dim: int, params: DenseParams, gen: torch.Generator, front_pad: int = 0):
        """``front_pad``: the input operand carries that many zero columns in front
        (kernel_0 gets as many zero rows; the variable is the view behind them)."""
        self.in_dim, self.params, self.front_pad = int(in_dim), params, int(front_pad)
        dims = [self.in_dim] + self.units
        for i in range(len(dims) - 1):
            params.add(f"{self.name}/kernel_{i}", _INIT[self.kernel_initializer]((dims[i], dims[i + 1]), gen,
                                                                                 params.rt.device),
                       front_pad_rows=self.front_pad if i == 0 else 0)
            if self.use_bias:
                params.add(f"{self.name}/bias_{i}", _INIT[self.bias_initializer]((dims[i + 1],), gen,
                                                                               params.rt.device))
        return self

    @property
    def kernels(self) -> List[torch.Tensor]:
        return [self.params[f"{self.name}/kernel_{i}"] for i in range(len(self.units))]

    @property
    def biases(self) -> List[torch.Tensor]:
        return [self.params[f"{self.name}/bias_{i}"] for i in range(len(self.units))] if self.use_bias else []

    # -- forward / backward ----------------------------------------------------
    def __call__(self, inputs, training: bool = False, is_train: bool = False) -> torch.Tensor:
        if self.params is None:                       # stand-alone use, as in the reference docstring
            rt = Runtime.get(inputs.device if isinstance(inputs, torch.Tensor) and inputs.is_cuda else None)
            gen = torch.Generator(device=rt.device)
            gen.manual_seed(0)
            self.build(inputs.shape[-1], DenseParams(rt), gen)
            self.params.finalize()
            self._owns_params = True
        rt = self.params.rt
        x = inputs if (isinstance(inputs, torch.Tensor) and inputs.dtype == torch.bfloat16 and
                       inputs.device == rt.device) else rt.to_device(inputs, torch.float32)
        assert x.dim() == 2 and x.shape[1] == self.front_pad + self.in_dim and x.stride(1) == 1
        acts = [x]          # layer outputs (fp32; acts[0] = the input): what the activation backward reads
        ops = []            # the operand each layer's GEMM actually consumed (bf16 copy on the tensor-core path)
        for i, n_out in enumerate(self.units):
            k = self.params.full(f"{self.name}/kernel_{i}")
            b = self.params[f"{self.name}/bias_{i}"] if self.use_bias else None
            y = rt.empty((x.shape[0], n_out))
            if self._tc(i, x):
                # tensor cores: Y = X K  ==  X [B,in] x (K^T)[out,in]^T ; K^T is a small bf16 copy made per call
                xb = x if x.dtype == torch.bfloat16 else cast_bf16(rt, x)
                ops.append(xb)
                kt = cast_bf16(rt, k, transpose=True)
                gemm_bf16_tn(rt, xb, kt, y, x.shape[0], n_out, x.shape[1], bias=b, act=self.activation)
            else:
                assert x.dtype == torch.float32
                ops.append(x)
                gemm_f32(rt, x, k, y, x.shape[0], n_out, x.shape[1], x.stride(0), n_out, n_out, bias=b,
                         act=self.activation)
            acts.append(y)
            x = y
        if training:
            self._saved = (acts, ops)
        return x

    def _tc(self, i: int, x: torch.Tensor) -> bool:
        return self.precision == "bf16" and x.shape[1] >= self.TC_MIN_IN and x.stride(0) % 8 == 0

    def backward(self, dy: torch.Tensor, need_input_grad: bool = True, dy_is_preact: bool = False,
                 accumulate_into: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """dy = dL/d(output) (post-activation; pre-activation of the LAST layer when
        ``dy_is_preact``), contiguous [B, units[-1]]; it is consumed (modified in
        place).  Fills the kernel/bias grads in ``params.grad``; returns
        dL/d(input) [B, front_pad+in_dim] (added into ``accumulate_into`` if given)."""
        assert self._saved is not None, "call with training=True first"
        rt = self.params.rt
        acts, ops = self._saved
        d = dy
        assert d.is_contiguous()
        last = len(self.units) - 1
        for i in reversed(range(len(self.units))):
            x, y = ops[i], acts[i + 1]
            B, n_in, n_out = x.shape[0], x.shape[1], self.units[i]
            if not (dy_is_preact and i == last):
                check(rt.lib.etr_act_backward(rt.ctx, d.data_ptr(), y.data_ptr(), d.numel(),
                                              _lib.ACT[self.activation], rt.stream))
            gk = self.params.gfull(f"{self.name}/kernel_{i}")
            tc = x.dtype == torch.bfloat16
            if (tc and self.fused_skinny and i == 0 and accumulate_into is None and n_out == 32 and n_in % 16 == 0 and
                    n_in <= 512 and x.stride(0) % 8 == 0):
                # wide-in / narrow-out first layer: dX, dK and db in one pass over the batch (K7b)
                dx = rt.empty((B, n_in), torch.bfloat16) if need_input_grad else None
                gb = self.params.g(f"{self.name}/bias_{i}") if self.use_bias else None
                check(rt.lib.etr_mlp_skinny_backward(rt.ctx, x.data_ptr(), x.stride(0), d.data_ptr(),
                                                     self.params.full(f"{self.name}/kernel_{i}").data_ptr(), B, n_in,
                                                     n_out, _p(dx), n_in, gk.data_ptr(), _p(gb), rt.stream))
                d = dx
                continue
            if tc:
                # dK = X^T d : X [B,in] and d [B,out] as stored (MN-major UMMA operands); split-K over the batch
                gemm_bf16_wgrad(rt, x, cast_bf16(rt, d), gk, n_in, n_out, B)
            else:
                # dK = x^T d : A stored [K=B, M=n_in] -> trans_a
                gemm_f32(rt, x, d, gk, n_in, n_out, B, x.stride(0), n_out, n_out, trans_a=True)
            if self.use_bias:
                gb = self.params.g(f"{self.name}/bias_{i}")
                check(rt.lib.etr_colsum_f32(rt.ctx, d.data_ptr(), B, n_out, n_out, gb.data_ptr(), rt.stream))
            if i > 0 or need_input_grad:
                k = self.params.full(f"{self.name}/kernel_{i}")
                acc = accumulate_into if (i == 0 and accumulate_into is not None) else None
                if tc and acc is None and n_out == 1 and i == 0 and k.is_contiguous():
                    # Dense(1): dx = dz k^T is a rank-1 product -- one elementwise pass instead of a K = 1 GEMM
                    dx = rt.empty((B, n_in), torch.bfloat16)
                    check(rt.lib.etr_outer_bf16(rt.ctx, d.data_ptr(), k.data_ptr(), B, n_in, dx.data_ptr(), n_in, rt.stream))
                elif tc and acc is None:
                    # dx = d K^T : A = d (bf16) [B,out], B operand = K [in,out] as stored; bf16 result
                    db_ = cast_bf16(rt, d)
                    kb = cast_bf16(rt, k)
                    # the model input's gradient goes to the gather backward (reads bf16); an
                    # intermediate one feeds the next activation backward and stays fp32
                    dx = rt.empty((B, n_in), torch.bfloat16 if i == 0 else torch.float32)
                    gemm_bf16_tn(rt, db_, kb, dx, B, n_in, n_out)
                elif tc and acc.dtype == torch.float32 and acc.stride(1) == 1:
                    # the same product accumulated into the gradient another branch already wrote for this input
                    gemm_bf16_tn_accumulate(rt, cast_bf16(rt, d), cast_bf16(rt, k), acc, B, n_in, n_out)
                    dx = acc
                else:
                    dx = acc if acc is not None else rt.empty((B, n_in))
                    # dx = d K^T : B(k,n) = K[n,k] -> trans_b
                    gemm_f32(rt, d, k, dx, B, n_in, n_out, n_out, n_out, dx.stride(0), trans_b=True,
                             beta=1.0 if acc is not None else 0.0)
                d = dx
            else:
                d = None
        self._saved = None
        return d


DenseLayer = MLPLayer      # 3.DCN/CustomLayers.py:153-167: a chain of Dense(x, activation)
