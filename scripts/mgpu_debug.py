import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import etr_b200  # noqa
    from etr_b200 import CustomLayers as L
    F, k, V, C, B = 26, 16, 100003, 13, 4096
    names, cont = [f"f{i}" for i in range(F)], [f"c{i}" for i in range(C)]
    for mode, prec, graph, evalfirst in (("peer", "bf16", False, True), ("peer-pull", "fp32", True, True), ("peer", "bf16", False, True), ("peer", "bf16", False, False)):
        sharded = L.DeepFMRankingLayer(names, V, k, continuous_features=cont, seed=3, shard=mode, check_ids=False, mlp_precision=prec)
        full = L.DeepFMRankingLayer(names, V, k, continuous_features=cont, seed=3, check_ids=False, mlp_precision=prec)
        g = torch.Generator(device="cuda").manual_seed(1234)
        table = torch.empty(V, k + 1, device="cuda").uniform_(-0.05, 0.05, generator=g)
        full.table.data[:, : k + 1] = table
        sharded.peer.load_global(table)
        sharded.params.value.copy_(full.params.value)
        torch.cuda.synchronize(); dist.barrier()
        rng = np.random.default_rng(10 + rank)
        tr_s, tr_f = L.Trainer(sharded, lr=1e-2, graph=graph), L.Trainer(full, lr=1e-2)
        if evalfirst:
            X = (rng.random((B, F)) ** 3 * V).astype(np.int64)
            Xc = rng.normal(size=(B, C)).astype(np.float32)
            y = (rng.random(B) < 0.3)
            d = {n: torch.tensor(X[:, i]).cuda() for i, n in enumerate(names)}
            d.update({n: torch.tensor(Xc[:, i]).cuda() for i, n in enumerate(cont)})
            a, b = sharded(d)["output"], full(d)["output"]
            if rank == 0: print("world", world, mode, "eval equal:", torch.equal(a, b), flush=True)
        for step in range(6 if graph else 3):
            X = (rng.random((B, F)) ** 3 * V).astype(np.int64)
            Xc = rng.normal(size=(B, C)).astype(np.float32)
            y = torch.tensor((rng.random(B) < 0.3).astype(np.float32)).cuda()
            d = {n: torch.tensor(X[:, i]).cuda() for i, n in enumerate(names)}
            d.update({n: torch.tensor(Xc[:, i]).cuda() for i, n in enumerate(cont)})
            before_s, before_f = sharded.table.data.clone(), full.table.data.clone()
            ls = tr_s.train_step(d, y)
            gd = {}
            for n, t in d.items():
                parts = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(parts, t); gd[n] = torch.cat(parts)
            ys = [torch.empty_like(y) for _ in range(world)]
            dist.all_gather(ys, y)
            lf = tr_f.train_step(gd, torch.cat(ys))
            torch.cuda.synchronize()
            gs, gf = sharded.params.grad, full.params.grad
            mine = full.table.data[rank::world]
            diff = (sharded.table.data[: mine.shape[0]] - mine).abs()
            upd_f = (full.table.data - before_f).abs()[rank::world]
            nbad = int((diff > 1e-4).sum().item())
            if rank == 0:
                print(f"world {world} {mode}/{prec}/graph={graph}/eval={evalfirst} step {step}: dense grad maxdiff {(gs - gf).abs().max().item():.3e} (scale {gf.abs().max().item():.3e}) "
                      f"dense value diff {(sharded.params.value - full.params.value).abs().max().item():.3e} "
                      f"table maxdiff {diff.max().item():.3e} entries>1e-4: {nbad} max update {upd_f.max().item():.3e}", flush=True)
        sharded.rt.poll_error()
        dist.barrier()
    dist.destroy_process_group()

main()
