"""Debug: the two extra arithmetic modes of the c2 bench (fp32 tower; keras_dense apply), a few eager steps each, for a launch list."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import etr_b200  # noqa
from etr_b200 import CustomLayers as L
mode = sys.argv[1]
dev = torch.device("cuda", 0)
B, F = 65536, 26
V = int(sum(bench.CRITEO_CARDS))
names, cont = [f"C{i + 1}" for i in range(F)], [f"I{i + 1}" for i in range(13)]
host = bench.make_batches(3, B, "zipf", seed=bench.SEED + 1)
devb = [(torch.from_numpy(np.ascontiguousarray(X.T)).to(dev), torch.from_numpy(np.ascontiguousarray(Xc.T)).to(dev), torch.from_numpy(y).to(dev)) for X, Xc, y in host]
def dd(i):
    ids, xc, y = devb[i % 3]
    return {**{n: ids[f] for f, n in enumerate(names)}, **{n: xc[c] for c, n in enumerate(cont)}}, y
lay = L.DeepFMRankingLayer(names, feature_dims=V, embedding_dims=16, continuous_features=cont, seed=1, check_ids=False,
                           mlp_precision="fp32" if mode == "fp32" else "bf16")
graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
tr = L.Trainer(lay, lr=1e-3, apply_mode="keras_dense" if mode == "keras" else "rowwise", graph=graph, plan_ahead=graph and mode != "keras")
for i in range(8 if graph else 3):
    d, y = dd(i); tr.train_step(tr.stage(d, y))
torch.cuda.synchronize()
if len(sys.argv) > 2 and sys.argv[2] == "denorm":      # every Adam m of the table a subnormal float, as ~700 steps after a row's last touch
    lay.table.m.fill_(1e-40)
    torch.cuda.synchronize()
ts = []
for i in range(4):
    d, y = dd(i); b = tr.stage(d, y); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record(); tr.train_step(b); e.record(); torch.cuda.synchronize(); ts.append((a.elapsed_time(e), (time.perf_counter() - t0) * 1e3))
print(mode, "graph" if graph else "eager", "ms per step (device, wall):", [f"{x:.2f}/{w:.2f}" for x, w in ts], flush=True)
