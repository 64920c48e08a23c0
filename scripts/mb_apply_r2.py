"""Round 2: time the fused FM backward + segment reduction + Adam alone on the c2 workload (plan precomputed,
L2 flushed before every call), plain three-array layout vs the 256-byte record layout; the record kernel's
variant comes from ETR_FUSED_REC (off | D<depth>S<0|1>) -- run once per variant (the choice is read once per process).
    python scripts/mb_apply_r2.py [record|plain] [zipf|uniform]"""
import os, sys, statistics
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import etr_b200  # noqa
from etr_b200.runtime import EmbeddingTable, IdsBatch, Runtime, SparsePlan, FusedFMGrad
layout = sys.argv[1] if len(sys.argv) > 1 else "record"
dists = sys.argv[2:] or ["zipf", "uniform"]
rt = Runtime.get(); dev = rt.device
B, F, K = 65536, 26, 16
V = int(sum(bench.CRITEO_CARDS))
tab = EmbeddingTable(rt, V, K + 1, record=(layout == "record"))
tab.data[:, :17].uniform_(-0.05, 0.05); tab.m; tab.v
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
# ETR_MB_CLEAN=1: after the write flush, READ 256 MiB so that the L2 holds clean lines (otherwise the timed kernel pays the
# write-back of up to 126 MB of dirty flush lines)
clean = torch.ones(64 << 20, dtype=torch.float32, device=dev) if os.environ.get("ETR_MB_CLEAN") == "1" else None
for dist in dists:
    host = bench.make_batches(4, B, dist, seed=bench.SEED + 1)
    ids = [IdsBatch(rt, torch.from_numpy(np.ascontiguousarray(X.T)).to(dev), B, F, 1, 1, B, 1) for X, _, _ in host]
    plans = [SparsePlan(rt, i, V) for i in ids]
    if FusedFMGrad.apply_kernel == "tile" and layout == "record" and os.environ.get("ETR_MB_PREP", "1") == "1":
        for p_ in plans:
            p_.prepare_fm()
    dl = torch.randn(B, device=dev) * 1e-5; sumv = torch.randn(B, K, device=dev) * 0.1
    dx = (torch.randn(B, 16 + F * K, device=dev) * 1e-5).to(torch.bfloat16)
    lr = torch.tensor([1e-3], device=dev)
    ts = []
    for i in range(int(os.environ.get("ETR_MB_ITERS", "14"))):
        g = FusedFMGrad(tab, ids[i % 4], K, dl, sumv, dx, 16, plan=plans[i % 4])
        flush.zero_()
        if clean is not None:
            clean.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.apply(lr, 0.9, 0.999, 1e-7); b.record(); ts.append((a, b))
    torch.cuda.synchronize()
    nu = plans[0].n_unique
    us = statistics.median(a.elapsed_time(b) for a, b in ts[min(2, len(ts) - 1):]) * 1e3
    print(f"fused apply layout={layout} kernel={os.environ.get('ETR_FUSED_APPLY', 'tile')} rec={os.environ.get('ETR_FUSED_REC', 'default')} ({dist}): {us:.1f} us; "
          f"unique rows {nu}; {nu * 408 / us / 1e3:.0f} GB/s algorithmic (408 B per unique row)", flush=True)
