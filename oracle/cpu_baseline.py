"""CPU baseline: the reference DeepFM / FM train step restated on torch-CPU.

TEST / BENCH INFRASTRUCTURE ONLY (see ``oracle/__init__.py``); "CPU restatement
of the reference (TensorFlow unavailable)" -- never call it "TF2".

One step = what 2.FM/ModelManager.py:171-181 does around
``DeepFMRankingLayer.call`` (2.FM/CustomLayers.py:279-308): gather ->
FM terms + MLP -> sigmoid -> Keras BCE -> tape.gradient (the table gradients are
IndexedSlices: per-occurrence rows + ids, deduplicated with unique +
segment-sum) -> Adam (row-wise on the unique rows, or the exact-Keras dense
pass).  The gathered rows are autograd leaves, so no dense [V,k] gradient is
ever materialised -- the same thing TF's IndexedSlices achieves.
"""
from __future__ import annotations

import math
import os
import time
from typing import Optional

import numpy as np
import torch

from . import reference_layers as R


class DeepFMCpuStep:
    def __init__(self, V: int, F: int, k: int, C: int, mlp_dims=(32, 8), lr=1e-3, seed=0, mode="rowwise",
                 threads: Optional[int] = None, with_mlp: bool = True):
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        g = torch.Generator().manual_seed(seed)
        self.V, self.F, self.k, self.C, self.mode, self.lr = V, F, k, C, mode, lr
        self.embed = torch.empty(V, k).uniform_(-0.05, 0.05, generator=g)
        self.w = torch.empty(V, 1).uniform_(-0.05, 0.05, generator=g)
        self.m_e, self.v_e = torch.zeros(V, k), torch.zeros(V, k)
        self.m_w, self.v_w = torch.zeros(V, 1), torch.zeros(V, 1)
        self.bias = torch.zeros(1, requires_grad=True)
        self.with_mlp = with_mlp
        self.dense = [self.bias]
        if with_mlp:
            rng = np.random.default_rng(seed)
            self.mlp1 = R.MLPLayer(list(mlp_dims), "relu").init_weights(C + F * k, rng)
            self.mlp2 = R.MLPLayer([1]).init_weights(mlp_dims[-1], rng)
            self.dense += self.mlp1.variables() + self.mlp2.variables()
        self.dense_state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in self.dense]
        self.t = 0
        self.b1, self.b2, self.eps = 0.9, 0.999, 1e-7

    def step(self, X: torch.Tensor, Xc: Optional[torch.Tensor], y: torch.Tensor) -> float:
        B = X.shape[0]
        emb = self.embed[X].requires_grad_(True)                 # [B,F,k]  ResourceGather
        wv = self.w[X].requires_grad_(True)                      # [B,F,1]
        first, second = R.fm_terms_from_rows(emb, wv)
        z = first + self.bias + second
        if self.with_mlp:
            flat = emb.reshape(B, -1)
            if Xc is not None:
                flat = torch.cat([Xc, flat], dim=1)
            z = z + self.mlp2(self.mlp1(flat))
        loss = R.keras_bce(y.reshape(-1, 1), torch.sigmoid(z))
        grads = torch.autograd.grad(loss, [emb, wv] + self.dense)
        # ---- IndexedSlices dedup: unique + segment-sum
        flat_ids = X.reshape(-1)
        uniq, inv = torch.unique(flat_ids, return_inverse=True)
        ge = torch.zeros(uniq.numel(), self.k).index_add_(0, inv, grads[0].reshape(-1, self.k))
        gw = torch.zeros(uniq.numel(), 1).index_add_(0, inv, grads[1].reshape(-1, 1))
        # ---- Adam
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        with torch.no_grad():
            for var, m, v, g in ((self.embed, self.m_e, self.v_e, ge), (self.w, self.m_w, self.v_w, gw)):
                if self.mode == "keras_dense":
                    m.mul_(self.b1); m[uniq] += (1 - self.b1) * g
                    v.mul_(self.b2); v[uniq] += (1 - self.b2) * g * g
                    var.sub_(lr_t * m / (v.sqrt() + self.eps))
                else:
                    mr = m[uniq] * self.b1 + (1 - self.b1) * g
                    vr = v[uniq] * self.b2 + (1 - self.b2) * g * g
                    m[uniq] = mr; v[uniq] = vr
                    var[uniq] -= lr_t * mr / (vr.sqrt() + self.eps)
            for p, (m, v), g in zip(self.dense, self.dense_state, grads[2:]):
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                p.sub_(lr_t * m / (v.sqrt() + self.eps))
        return float(loss.detach())


def time_cpu_steps(stepper: DeepFMCpuStep, batches, budget_s: float = 15.0, warmup: int = 1, max_steps: int = 1000):
    """Runs whole batches until ~budget_s of CPU time is spent; returns
    (samples_per_s, steps, seconds)."""
    for i in range(warmup):
        stepper.step(*batches[i % len(batches)])
    n, t0 = 0, time.perf_counter()
    samples = 0
    while n < max_steps:
        X, Xc, y = batches[n % len(batches)]
        stepper.step(X, Xc, y)
        samples += X.shape[0]
        n += 1
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    return samples / dt, n, dt
