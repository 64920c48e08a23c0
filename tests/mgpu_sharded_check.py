"""Multi-GPU check of the row-sharded tables (run under torchrun, one rank per GPU):

    gpurun --gpus 2 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_sharded_check.py'

1. forward: sharded lookup + fused kernel == the unsharded layer on the same local batch, BIT-EXACT;
2. training: data-parallel sharded steps == an unsharded trainer fed the all-gathered global batch
   (up to fp32 summation order).
The unsharded layer is our own CUDA path (already parity-checked against the oracle on one GPU).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import etr_b200  # noqa: F401
    from etr_b200 import CustomLayers as L

    F, k, V, C, B = 26, 16, 100003, 13, 4096
    names, cont = [f"f{i}" for i in range(F)], [f"c{i}" for i in range(C)]
    msgs = []
    # ahead: Trainer next_batch= (the plan of step i+1 is sorted on the side stream during step i; 3 buffer sets)
    for mode, graph, prec, ahead in (("a2a", False, "fp32", False), ("peer-pull", False, "fp32", False),
                                     ("peer-pull", True, "fp32", False), ("peer", False, "bf16", False),
                                     ("peer", True, "bf16", False), ("peer", True, "bf16", True)):
        # "peer" + bf16 tower takes the fused train step with the de-duplicated request/serve exchange
        sharded = L.DeepFMRankingLayer(names, V, k, continuous_features=cont, seed=3, shard=mode, check_ids=False,
                                       mlp_precision=prec)
        full = L.DeepFMRankingLayer(names, V, k, continuous_features=cont, seed=3, check_ids=False, mlp_precision=prec)
        # identical global weights on every rank
        g = torch.Generator(device="cuda").manual_seed(1234)
        table = torch.empty(V, k + 1, device="cuda").uniform_(-0.05, 0.05, generator=g)
        full.table.data[:, : k + 1] = table
        (sharded.peer if mode.startswith("peer") else sharded.shard).load_global(table)
        sharded.params.value.copy_(full.params.value)
        torch.cuda.synchronize()
        dist.barrier()                                  # every shard is loaded before anyone reads it

        rng = np.random.default_rng(10 + rank)

        def batch():
            X = (rng.random((B, F)) ** 3 * V).astype(np.int64)
            Xc = rng.normal(size=(B, C)).astype(np.float32)
            y = (rng.random(B) < 0.3).astype(np.float32)
            d = {n: torch.tensor(X[:, i]).cuda() for i, n in enumerate(names)}
            d.update({n: torch.tensor(Xc[:, i]).cuda() for i, n in enumerate(cont)})
            return d, torch.tensor(y).cuda()

        d, y = batch()
        a = sharded(d)["output"]
        b = full(d)["output"]
        assert torch.equal(a, b), f"rank {rank} {mode}: sharded forward is not bit-exact ({(a - b).abs().max().item()})"

        tr_s = L.Trainer(sharded, lr=1e-2, graph=graph)
        tr_f = L.Trainer(full, lr=1e-2)
        n_steps = (13 if ahead else 6) if graph else 3  # graph mode: 2 (3) buffer sets x (2 eager + capture/replay)
        feed = [batch() for _ in range(n_steps + 1)]
        staged = tr_s.stage(*feed[0]) if ahead and graph else None
        for step in range(n_steps):
            d, y = feed[step]
            if ahead and graph:
                nxt = tr_s.stage(*feed[step + 1])
                ls = tr_s.train_step(staged, None, nxt).clone()
                staged = nxt
            else:
                ls = tr_s.train_step(d, y).clone()
            # the unsharded reference sees the global batch
            gd = {}
            for n, t in d.items():
                parts = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(parts, t)
                gd[n] = torch.cat(parts)
            ys = [torch.empty_like(y) for _ in range(world)]
            dist.all_gather(ys, y)
            lf = tr_f.train_step(gd, torch.cat(ys))
            lsum = ls.clone()
            dist.all_reduce(lsum)
            assert abs(float(lsum.item()) / world - float(lf.item())) < (1e-5 if prec == "fp32" else 1e-4), \
                (mode, graph, step, float(lsum.item()) / world, float(lf.item()))
            if step == 0:
                # one step from identical weights: the sharded and the unsharded update agree to fp32 rounding in
                # EVERY mode (later steps of the bf16 tower do not: a 1e-7 difference in an fp32 weight can flip its
                # bf16 rounding, a 0.4 % change of that operand -- measured drift 1e-5 .. 6e-4 after 3-6 steps)
                torch.cuda.synchronize()
                mine0 = full.table.data[rank::world]
                e_t = (sharded.table.data[: mine0.shape[0]] - mine0).abs().max().item()
                e_d = (sharded.params.value - full.params.value).abs().max().item()
                assert e_t < 2e-6 and e_d < 2e-6, (mode, graph, "first step", rank, e_t, e_d)
        sharded.rt.poll_error()
        mine = full.table.data[rank::world]
        err_t = (sharded.table.data[: mine.shape[0]] - mine).abs().max().item()
        err_d = (sharded.params.value - full.params.value).abs().max().item()
        # lr 1e-2: < 0.2% of one Adam step (fp32).  bf16 tower: the two sides run DIFFERENT kernels on different table
        # layouts (unsharded: 256-byte records, occurrence-parallel apply with MUFU sqrt / rcp; sharded: plain arrays,
        # owner-side apply), so their fp32 weights differ in the last ulp after one step (checked above: < 2e-6); a last-ulp
        # difference can flip the bf16 rounding of an operand, and in its first steps Adam turns a flipped sign of a
        # near-zero gradient into a full +-lr step: entries a step or two of lr apart after 3-6 steps are expected
        tol = 2e-5 if prec == "fp32" else 2.5e-2
        assert err_t < tol and err_d < tol, (mode, graph, rank, err_t, err_d)
        if mode.startswith("peer"):
            # replicated dense variables must stay BIT-identical across ranks (deterministic rank-order sum)
            ref = sharded.params.value.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(ref, sharded.params.value), f"rank {rank}: dense replicas diverged"
        dist.barrier()
        msgs.append(f"{mode}{'+graph' if graph else ''}{'+ahead' if ahead else ''}/{prec}: table_err={err_t:.2e} dense_err={err_d:.2e}")
    # every other model family that gathers from one shared table: all-to-all form (DCN vector / matrix cross, PNN)
    for model in ("dcn_vec", "dcn_matrix", "pnn"):
        Fm, Vm, Bm = 10, 60013, 1024
        nm, cm = [f"f{i}" for i in range(Fm)], [f"c{i}" for i in range(3)]

        def make(shard):
            if model == "pnn":
                return L.PNNRankingLayer(nm, Vm, k, seed=4, shard=shard, check_ids=False)
            return L.DeepCrossNetworkLayer(nm, cm, feature_dims=Vm, embedding_dims=k, type="vec" if model == "dcn_vec" else "matrix",
                                           seed=4, shard=shard, check_ids=False)

        sharded, full = make("a2a"), make(None)
        g = torch.Generator(device="cuda").manual_seed(77)
        table = torch.empty(Vm, k, device="cuda").uniform_(-0.05, 0.05, generator=g)
        full.table.data[:, :k] = table
        sharded.shard.load_global(table)
        sharded.params.value.copy_(full.params.value)
        torch.cuda.synchronize()
        dist.barrier()
        rng = np.random.default_rng(20 + rank)
        tr_s, tr_f = L.Trainer(sharded, lr=1e-2), L.Trainer(full, lr=1e-2)
        for step in range(3):
            X = (rng.random((Bm, Fm)) ** 3 * Vm).astype(np.int64)
            Xc = rng.normal(size=(Bm, 3)).astype(np.float32)
            y = torch.tensor((rng.random(Bm) < 0.3).astype(np.float32)).cuda()
            d = {n: torch.tensor(X[:, i]).cuda() for i, n in enumerate(nm)}
            if model != "pnn":
                d.update({n: torch.tensor(Xc[:, i]).cuda() for i, n in enumerate(cm)})
            if step == 0:
                assert torch.equal(sharded(d)["output"], full(d)["output"]), f"rank {rank} {model}: forward not bit-exact"
            ls = tr_s.train_step(d, y).clone()
            gd = {}
            for n, t in d.items():
                parts = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(parts, t)
                gd[n] = torch.cat(parts)
            ys = [torch.empty_like(y) for _ in range(world)]
            dist.all_gather(ys, y)
            lf = tr_f.train_step(gd, torch.cat(ys))
            dist.all_reduce(ls)
            assert abs(float(ls.item()) / world - float(lf.item())) < 1e-5, (model, step)
        mine = full.table.data[rank::world, :k]
        err_t = (sharded.table.data[: mine.shape[0], :k] - mine).abs().max().item()
        err_d = (sharded.params.value - full.params.value).abs().max().item()
        assert err_t < 2e-5 and err_d < 2e-5, (model, rank, err_t, err_d)
        dist.barrier()
        msgs.append(f"{model}/a2a: table_err={err_t:.2e} dense_err={err_d:.2e}")
    if rank == 0:
        print(f"MGPU_OK world={world} " + " | ".join(msgs))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
