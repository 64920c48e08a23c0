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


def run(tag, V, k, has_w, flat, dist, cps=None, nb=6, dtype=torch.float32, do_flush=True, align=16, impl="staged", wpb=None):
    os.environ["ETR_GATHER"] = impl
    if wpb:
        os.environ["ETR_STAGED_WPB"] = str(wpb)
    else:
        os.environ.pop("ETR_STAGED_WPB", None)
    if cps:
        os.environ["ETR_STREAM_CPS"] = str(cps)
    else:
        os.environ.pop("ETR_STREAM_CPS", None)
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
for align in (16, 128):
    run(f"staged fp32 k=16 +w align={align} flat=bf16 zipf", VBIG, 16, True, "bf16", "zipf", align=align)
    run(f"staged fp32 k=16 +w align={align} flat=bf16 uniform", VBIG, 16, True, "bf16", "uniform", align=align)
    run(f"staged fp32 k=16 +w align={align} no flat uniform", VBIG, 16, True, None, "uniform", align=align)
for wpb in (2, 3, 4, 5):
    run(f"staged fp32 k=16 +w align=128 flat=bf16 uniform wpb={wpb}", VBIG, 16, True, "bf16", "uniform", align=128, wpb=wpb)
run("staged 64B rows no w no flat uniform", VBIG, 16, False, None, "uniform")
run("staged 64B rows no w flat=bf16 uniform", VBIG, 16, False, "bf16", "uniform")
run("stream 64B rows no w flat=bf16 uniform", VBIG, 16, False, "bf16", "uniform", impl="stream")
for align in (16, 64):
    run(f"staged bf16 k=16 +w align={align} flat=bf16 zipf", VBIG, 16, True, "bf16", "zipf", dtype=torch.bfloat16, align=align)
    run(f"staged bf16 k=16 +w align={align} flat=bf16 uniform", VBIG, 16, True, "bf16", "uniform", dtype=torch.bfloat16, align=align)
run("staged k=64 256B rows no w flat=bf16 zipf (c3 shape)", VBIG, 64, False, "bf16", "zipf")
run("stream k=64 256B rows no w flat=bf16 zipf (c3 shape)", VBIG, 64, False, "bf16", "zipf", impl="stream")
run("staged k=64 256B rows no w no flat uniform", VBIG, 64, False, None, "uniform")
