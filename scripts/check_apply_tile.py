"""Tiled fused apply (csrc/fm_fused_tile.cu) vs the occurrence-parallel kernel (csrc/fm_fused_flat.cu) on the c2
workload: the same plan, the same operands, one Adam step from identical tables -> max |difference| of every touched
record, and a count of untouched rows that changed (must be 0).  python scripts/check_apply_tile.py [zipf|uniform]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import etr_b200  # noqa
from etr_b200.runtime import EmbeddingTable, IdsBatch, Runtime, SparsePlan, FusedFMGrad
rt = Runtime.get(); dev = rt.device
B, F, K = 65536, 26, 16
V = int(sum(bench.CRITEO_CARDS))
for dist in (sys.argv[1:] or ["zipf", "uniform"]):
    X, _, _ = bench.make_batches(1, B, dist, seed=bench.SEED + 1)[0]
    ids = IdsBatch(rt, torch.from_numpy(np.ascontiguousarray(X.T)).to(dev), B, F, 1, 1, B, 1)
    plan = SparsePlan(rt, ids, V)
    g = torch.Generator(device=dev); g.manual_seed(7)
    dl = torch.randn(B, device=dev, generator=g) * 1e-2
    sumv = torch.randn(B, K, device=dev, generator=g) * 0.1
    dx = (torch.randn(B, 16 + F * K, device=dev, generator=g) * 1e-2).to(torch.bfloat16)
    lr = torch.tensor([1e-3], device=dev)
    out = {}
    for kern in ("tile", "flat"):
        tab = EmbeddingTable(rt, V, K + 1, record=True)
        gg = torch.Generator(device=dev); gg.manual_seed(11)
        tab.rec[:, :60].uniform_(0.01, 0.05, generator=gg)       # var, m, v all non-trivial (v >= 0)
        FusedFMGrad.apply_kernel = kern
        FusedFMGrad(tab, ids, K, dl, sumv, dx, 16, plan=plan).apply(lr, 0.9, 0.999, 1e-7)
        torch.cuda.synchronize()
        out[kern] = tab.rec
        if kern == "tile":
            before = torch.empty_like(tab.rec)
            gg.manual_seed(11); before[:, :60].uniform_(0.01, 0.05, generator=gg); before[:, 60:] = 0
    nu = plan.n_unique
    uid = plan.unique_ids[:nu]
    d = (out["tile"][uid] - out["flat"][uid]).abs()
    ref = out["flat"][uid].abs().clamp_min(1e-6)
    moved = (out["tile"][uid, :17] - before[uid, :17]).abs().amax(dim=1)
    mask = torch.ones(V, dtype=torch.bool, device=dev); mask[uid] = False
    stray = int((out["tile"][mask] != before[mask]).any(dim=1).sum())
    print(f"{dist}: unique {nu}  max|tile-flat| = {float(d.max()):.3e}  max rel = {float((d / ref).max()):.3e}  "
          f"rows not moved by tile = {int((moved == 0).sum())}  stray rows changed = {stray}", flush=True)
    assert stray == 0 and float(d.max()) < 2e-6 and int((moved == 0).sum()) == 0
print("check_apply_tile OK")
