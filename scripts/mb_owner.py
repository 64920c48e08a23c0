"""Owner side of the peer-sharded c2 step alone, on ONE GPU: the mailbox of owner 0 at world = 2 is synthesised from two
c2 batches (the unique ids of each batch that owner 0 holds = one source region), then etr_shard_owner_prep and
etr_shard_owner_apply are timed with the L2 flushed, next to the stamped accumulator + touched-row Adam they replace.
    python scripts/mb_owner.py [world]          knobs: ETR_OWNER_U, ETR_OWNER_BPS"""
import ctypes as C, os, sys, statistics
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import etr_b200  # noqa
from etr_b200._lib import check
from etr_b200.runtime import EmbeddingTable, Runtime
W = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rt = Runtime.get(); dev = rt.device
B, F, K, ld = 65536, 26, 16, 20
V = int(sum(bench.CRITEO_CARDS))
rows = (V + W - 1) // W
cap = (3 * B * F // (2 * W) + 1024 + 63) // 64 * 64
host = bench.make_batches(W, B, "zipf", seed=bench.SEED + 1)
req = torch.zeros((W, cap), dtype=torch.int64); counts = torch.zeros(W, dtype=torch.int32)
for s, (X, _, _) in enumerate(host):
    u = np.unique(X); u = u[u % W == 0] // W
    u = u[np.random.default_rng(s).permutation(len(u))] if os.environ.get("ETR_MB_SHUFFLE") == "1" else u
    req[s, : len(u)] = torch.from_numpy(u); counts[s] = len(u)
n_ent = int(counts.sum()); n_rows = len(np.unique(np.concatenate([req[s, : int(counts[s])].numpy() for s in range(W)])))
req_d, counts_d = req.to(dev), counts.to(dev)
grads = torch.randn((W * cap, ld), device=dev) * 1e-4; grads[:, 17:] = 0
lr = torch.tensor([1e-3], device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn, pre=None, iters=10):
    ts = []
    for _ in range(iters):
        if pre: pre()
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ts.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ts[2:]) * 1e3


for layout in ("record", "plain"):
    tab = EmbeddingTable(rt, rows, K + 1, record=(layout == "record"))
    tab.data[:, :17].uniform_(-0.05, 0.05); tab.m; tab.v
    t = tab.desc()
    own = {"map": rt.zeros((rows,), torch.int64), "step": rt.zeros((1,), torch.int32),
           "mask": rt.zeros((W * cap,), torch.int32), "others": rt.empty((W * cap, W), torch.int32)}
    prep = lambda: check(rt.lib.etr_shard_owner_prep(rt.ctx, req_d.data_ptr(), counts_d.data_ptr(), W, cap, rows, own["map"].data_ptr(),
                                                     own["step"].data_ptr(), own["mask"].data_ptr(), own["others"].data_ptr(), rt.stream))
    apply = lambda: check(rt.lib.etr_shard_owner_apply(rt.ctx, C.byref(t), tab.m.data_ptr(), tab.v.data_ptr(), req_d.data_ptr(),
                                                       counts_d.data_ptr(), grads.data_ptr(), W, cap, ld, own["mask"].data_ptr(),
                                                       own["others"].data_ptr(), K, lr.data_ptr(), 0.9, 0.999, 1e-7, rt.stream))
    us_apply = timed(apply, pre=prep)
    us_prep = timed(prep, pre=lambda: own["mask"].zero_())
    own["mask"].zero_()
    byts = n_ent * (8 + 4 + 80) + n_rows * 480
    print(f"owner W={W} layout={layout} U={os.environ.get('ETR_OWNER_U', '1')} bps={os.environ.get('ETR_OWNER_BPS', 'def')}: "
          f"prep {us_prep:.1f} us, apply {us_apply:.1f} us; entries {n_ent}, rows {n_rows}; "
          f"{byts / us_apply / 1e3:.0f} GB/s (92 B per entry + 480 B per row)", flush=True)
    if layout == "plain" or os.environ.get("ETR_MB_OLD") == "1":
        gacc, epoch = rt.zeros((rows, ld)), rt.zeros((1,), torch.int32)
        touched, n_t = rt.empty((W * cap,), torch.int32), rt.zeros((1,), torch.int32)

        def old():
            epoch.add_(1)
            check(rt.lib.etr_shard_mailbox_accumulate(rt.ctx, req_d.data_ptr(), grads.data_ptr(), counts_d.data_ptr(), W, cap, ld,
                                                      gacc.data_ptr(), epoch.data_ptr(), touched.data_ptr(), n_t.data_ptr(), W * cap, rt.stream))
            check(rt.lib.etr_shard_touched_adam(rt.ctx, C.byref(t), tab.m.data_ptr(), tab.v.data_ptr(), gacc.data_ptr(), ld, touched.data_ptr(),
                                                n_t.data_ptr(), W * cap, K, lr.data_ptr(), 0.9, 0.999, 1e-7, rt.stream))
        print(f"   accumulate + touched-row adam ({layout}): {timed(old):.1f} us", flush=True)
        del gacc
    del tab, own
rt.poll_error()
