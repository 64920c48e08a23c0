"""K1 variants on the REAL bench workload (c2: Criteo cardinalities, per-field Zipf / uniform ids,
13 dense columns in front, bf16 flattened operand): which kernel / cache policy / row layout wins?"""
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import etr_b200  # noqa: F401,E402
from etr_b200.runtime import EmbeddingTable, IdsBatch, Runtime, gather_fm_forward  # noqa: E402

rt = Runtime.get()
dev = rt.device
B, F, K = 65536, 26, 16
V = int(sum(bench.CRITEO_CARDS))
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
tables = {}
for align in (16, 128):
    t = EmbeddingTable(rt, V, K + 1, torch.float32, row_align=align)
    t.data.uniform_(-0.05, 0.05)
    tables[align] = t
ids = {}
for dist in ("zipf", "uniform"):
    host = bench.make_batches(6, B, dist, seed=bench.SEED + 1)
    ids[dist] = [(IdsBatch(rt, torch.from_numpy(np.ascontiguousarray(X.T)).to(dev), B, F, 1, 1, B, 1),
                  torch.from_numpy(np.ascontiguousarray(Xc.T)).to(dev)) for X, Xc, y in host]
logit = rt.empty((B,))
x = rt.empty((B, 16 + F * K), torch.bfloat16)
alg = B * (F * (K * 4 + 4 + 8) + 4) + B * F * K * 2


def run(impl, l1, align, dist, extra=None):
    # (the L1-policy / CTAs-per-SM knobs and the cp.async-staged variant of the r01 experiments were removed from the
    # library after the measurements in profiles/r01_mb_gather.md; ETR_GATHER=stream|generic remains)
    os.environ["ETR_GATHER"] = impl
    tab = tables[align]
    ts = []
    G = 4
    for i in range(12):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for j in range(G):
            idb, xc = ids[dist][(i * G + j) % 6]
            gather_fm_forward(tab, K, True, idb, logit=logit, flat=x, flat_col0=16, cont=xc.t())
        b.record()
        ts.append((a, b))
    torch.cuda.synchronize()
    us = statistics.median(a.elapsed_time(b) for a, b in ts[2:]) / G * 1e3
    print(f"{impl:8s} L1={l1} align={align:3d} {dist:8s} {str(extra or ''):28s} {us:7.1f} us  {alg / us / 1e3:6.0f} GB/s ({alg / us / 1e3 / 6549.4:.3f})",
          flush=True)


for dist in ("zipf", "uniform"):
    for impl in ("generic", "stream"):
        for align in (16, 128):
            run(impl, 0, align, dist)
