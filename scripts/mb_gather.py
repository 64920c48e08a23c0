"""Micro-benchmark of the streaming gather kernel (run on the GPU box): what bounds it?
Varies table footprint, row alignment, stores, CTAs/SM.  Prints one line per config."""
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import etr_b200  # noqa: F401,E402
from etr_b200.runtime import EmbeddingTable, IdsBatch, Runtime, gather_fm_forward  # noqa: E402

rt = Runtime.get()
dev = rt.device
B, F = 65536, 26
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def run(tag, V, k, has_w, flat, dist, nb=6, dtype=torch.float32, do_flush=True, align=16, impl="stream"):
    os.environ["ETR_GATHER"] = impl
    tab = EmbeddingTable(rt, V, k + (1 if has_w else 0), dtype, row_align=align)
    tab.data.uniform_(-0.05, 0.05) if dtype == torch.float32 else tab.data.copy_(torch.empty_like(tab.data, dtype=torch.float32).uniform_(-0.05, 0.05))
    rng = np.random.default_rng(1)
    batches = []
    for _ in range(nb):
        u = rng.random((F, B))
        X = np.minimum((V * (u ** 3 if dist == "zipf" else u)).astype(np.int64), V - 1)
        batches.append(IdsBatch(rt, torch.from_numpy(X).to(dev), B, F, 1, 1, B, 1))
    logit = rt.empty((B,))
    fl = None
    if flat:
        fl = rt.empty((B, 16 + F * k), torch.bfloat16 if flat == "bf16" else torch.float32)
    ts = []
    G = 4
    for i in range(10):
        if do_flush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for j in range(G):
            gather_fm_forward(tab, k, has_w, batches[(i * G + j) % nb], logit=logit, flat=fl, flat_col0=16,
                              cont=None)
        b.record()
        ts.append((a, b))
    torch.cuda.synchronize()
    us = statistics.median(a.elapsed_time(b) for a, b in ts[2:]) / G * 1e3
    esz = 4 if dtype == torch.float32 else 2
    alg = B * (F * (k * esz + (esz if has_w else 0) + 8) + 4) + (B * F * k * (2 if flat == "bf16" else 4) if flat else 0)
    print(f"{tag:58s} {us:7.1f} us  {alg / us / 1e3:7.0f} GB/s alg  ({alg / us / 1e3 / 6549.4:.2f})", flush=True)
    del tab


VBIG, VSMALL = 33762577, 700000
# (the r01 sweeps over the staged variant, CTAs per SM and the L1 policy are recorded in profiles/r01_mb_gather.md;
# those knobs were removed from the library afterwards)
for impl in ("generic", "stream"):
    for align in (16, 128):
        run(f"{impl} fp32 k=16 +w align={align} flat=bf16 zipf", VBIG, 16, True, "bf16", "zipf", align=align, impl=impl)
        run(f"{impl} fp32 k=16 +w align={align} flat=bf16 uniform", VBIG, 16, True, "bf16", "uniform", align=align, impl=impl)
for k_, nbytes in ((8, 32), (16, 64), (32, 128), (64, 256)):
    run(f"stream k={k_} no w ({nbytes}B rows) no flat uniform", VBIG, k_, False, None, "uniform", impl="stream")
for V in (1 << 20, 1 << 22, 1 << 24, 1 << 26):
    run(f"stream k=16 no w 64B rows no flat uniform V=2^{V.bit_length() - 1} ({V * 64 >> 20} MB)", V, 16, False, None, "uniform",
        impl="stream")
